"""Per-kernel timing of the HW3D step pieces with CUDA events (no profiler): Laplace solve, fused RHS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi

def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for dims in ((516, 132, 512, 2, 2), (516, 68, 256, 2, 2), (260, 68, 256, 2, 2), (1028, 1, 1024, 2, 0)):
    g = capi.Grid(*dims)
    dev = "cuda"
    one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device=dev)
    ydep = torch.linspace(1.0, 1.3, g.ny, dtype=torch.float64, device=dev)[None, :].expand(g.nx, g.ny).contiguous()
    f = torch.rand(g.shape3, dtype=torch.float64, device=dev) - 0.5
    out = torch.zeros_like(f)
    for per, block, g11 in ((False, "0", one(1.0)), (False, "1", one(1.0)), (False, "1", ydep), (True, "0", one(1.0))):
        md = dict(dx=one(0.2), g11=g11, g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0), g_22=one(1.0))
        os.environ["BOUT_B200_LAP_BLOCK"] = block
        lap = capi.Laplace(g, md, 0.2 * g.nz, periodic_x=per)
        t = timeit(lambda: lap.solve(f, out=out))
        lap.status()
        npts = g.n_interior
        print(f"{dims} periodic_x={per} block_path={block} y_dependent={g11 is ydep}: solve {t*1e3:8.1f} us  "
              f"{npts/t/1e6:8.2f} Gpts/s  {16*npts/t/1e6:8.1f} GB/s algorithmic (16 B/pt)  {1e3/t:8.1f} solves/s", flush=True)
