"""configs[1] microbenchmarks (256x64x256 Field3D) timed three ways, to separate the kernels from the timing method:
  A  one launch between CUDA events after flush.zero_()  (L2 left full of DIRTY lines: their write-back lands
     inside the timed launch);
  B  one launch between CUDA events after a READ of a 320 MB buffer (L2 left full of clean lines);
  C  `nset` distinct operand sets (total > L2) used round-robin, `reps` launches back to back between two events:
     steady-state throughput of the operator as it runs inside an RHS, no per-launch event / launch latency.
Usage: python tools/time_micro.py [nset reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

capi = ge.load_package().capi
PEAK = 6545.3
nset = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
dev = "cuda"
g = capi.Grid(260, 68, 256, 2, 2)
sets = [tuple(torch.rand(g.shape3, dtype=torch.float64, device=dev) for _ in range(3)) + (torch.zeros(g.shape3, dtype=torch.float64, device=dev),)
        for _ in range(nset)]
dx = torch.full(g.shape2, 0.2, dtype=torch.float64, device=dev)
flush = torch.empty(40 * 1024 * 1024, dtype=torch.float64, device=dev)
flush.zero_()
ops = (
    ("DDX_C2", lambda a, b, c, o: capi.stencil(g, "DDX_C2", "X", a, scale1=dx, out=o), 16),
    ("DDX_C4", lambda a, b, c, o: capi.stencil(g, "DDX_C4", "X", a, scale1=dx, out=o), 16),
    ("DDX_W3", lambda a, b, c, o: capi.stencil(g, "DDX_W3", "X", a, scale1=dx, out=o), 16),
    ("DDY_C2", lambda a, b, c, o: capi.stencil(g, "DDX_C2", "Y", a, scale1=dx, out=o), 16),
    ("DDZ_C2", lambda a, b, c, o: capi.stencil(g, "DDX_C2", "Z", a, scale1=dx, out=o), 16),
    ("DDZ_FFT", lambda a, b, c, o: capi.ddz_fft(g, a, out=o), 16),
    ("VDDX_U1", lambda a, b, c, o: capi.stencil(g, "VDDX_U1", "X", a, v=b, scale1=dx, out=o), 24),
    ("bracket_arakawa", lambda a, b, c, o: capi.bracket_arakawa(g, a, b, dx, dx, out=o), 24),
    ("2a+bc", lambda a, b, c, o: capi.elementwise(g, [(2.0, a), (1.0, b, c)], out=o), 32),
    ("torch copy_", lambda a, b, c, o: o.copy_(a), 16 * (260 * 68) / (256 * 64)),
)


def single(fn, how):
    ts = []
    for i in range(12):
        if how == "A":
            flush.zero_()
        else:
            flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(*sets[0]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def rotating(fn):
    for s in sets:
        fn(*s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(*sets[i % nset])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


npts = g.n_interior
print(f"{'op':18s} {'A us':>8s} {'A frac':>7s} {'B us':>8s} {'B frac':>7s} {'C us':>8s} {'C frac':>7s}   ({nset} sets, {reps} launches)")
for name, fn, bpp in ops:
    fn(*sets[0])
    r = []
    for t in (single(fn, "A"), single(fn, "B"), rotating(fn)):
        r += [t * 1e6, bpp * npts / t / 1e9 / PEAK]
    print(f"{name:18s} {r[0]:8.1f} {r[1]:7.1%} {r[2]:8.1f} {r[3]:7.1%} {r[4]:8.1f} {r[5]:7.1%}")
