"""Timing ablations of the block-path rows kernels (BOUT_B200_LAPB_DBG bits switch phases off; results are wrong
then): which phase costs what, measured in place."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi

def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

dims = (516, 132, 512, 2, 2)
g = capi.Grid(*dims)
one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device="cuda")
md = dict(dx=one(0.2), g11=one(1.0), g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0), g_22=one(1.0))
f = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
out = torch.zeros_like(f)
lap = capi.Laplace(g, md, 0.2 * g.nz)
for dbg in (0, 256, 1, 64, 2, 4, 8, 7, 15):
    os.environ["BOUT_B200_LAPB_DBG"] = str(dbg)
    t = timeit(lambda: lap.solve(f, out=out))
    off = [n for b, n in ((1, "per-mode"), (2, "inverse fft"), (4, "forward fft"), (8, "row loads"), (64, "SOLVE table loads"), (256, "Nyquist systems")) if dbg & b]
    print(f"dbg={dbg:2d} off: {', '.join(off) or '-':45s} solve {t*1e3:8.1f} us", flush=True)
