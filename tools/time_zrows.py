"""z-FFT row operations: register path (default) vs the shared-memory Stockham path (BOUT_B200_ZROWS_REG=0 must be
set before the library is first used, so the two are separate processes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi
PEAK = 6545.3

def timeit(fn, reps=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]

flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
tag = "reg" if os.environ.get("BOUT_B200_ZROWS_REG", "1") != "0" else "smem"
for dims in ((260, 68, 256, 2, 2), (516, 68, 256, 2, 2), (516, 132, 512, 2, 2)):
    g = capi.Grid(*dims)
    f = torch.rand(g.shape3, dtype=torch.float64, device="cuda")
    out = torch.zeros_like(f)
    npts = g.n_interior
    t = timeit(lambda: capi.ddz_fft(g, f, out=out), flush=flush)
    print(f"[{tag}] {dims} DDZ_FFT      {t*1e6:8.1f} us  {16*npts/t/1e9:8.1f} GB/s  {16*npts/t/1e9/PEAK:6.1%}", flush=True)
    t = timeit(lambda: capi.zlowpass(g, f, 16, True, out=out, region="RGN_NOBNDRY"), flush=flush)
    print(f"[{tag}] {dims} lowPass      {t*1e6:8.1f} us  {16*npts/t/1e9:8.1f} GB/s  {16*npts/t/1e9/PEAK:6.1%}", flush=True)
    zs = torch.rand(g.shape2, dtype=torch.float64, device="cuda")
    yup, ydn = [torch.zeros_like(f)], [torch.zeros_like(f)]
    t = timeit(lambda: capi.parallel_slices(g, f, zs, 6.283, nslices=1, yup=yup, ydown=ydn), flush=flush)
    print(f"[{tag}] {dims} slice pair   {t*1e6:8.1f} us  {24*npts/t/1e9:8.1f} GB/s  {24*npts/t/1e9/PEAK:6.1%}", flush=True)
