"""Quick per-kernel timing on the GPU box (CUDA events, inputs larger than L2 or L2 flushed)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

capi = ge.load_package().capi
PEAK = 6552.6


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


def main():
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for name, dims in (("C2 256x64x256", (260, 68, 256, 2, 2)), ("C4 512x128x512", (516, 132, 512, 2, 2))):
        g = capi.Grid(*dims)
        f = torch.rand(g.shape3, dtype=torch.float64, device="cuda")
        h = torch.rand(g.shape3, dtype=torch.float64, device="cuda")
        v = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
        out = torch.zeros_like(f)
        dx = torch.full(g.shape2, 0.2, dtype=torch.float64, device="cuda")
        dz = torch.full(g.shape2, 0.1, dtype=torch.float64, device="cuda")
        npts = g.n_interior
        rows = []
        for op, d, bpp in (("DDX_C2", "X", 16), ("DDX_C4", "X", 16), ("DDX_C2", "Y", 16), ("DDX_C4", "Y", 16),
                           ("DDX_C2", "Z", 16), ("DDX_C4", "Z", 16), ("DDX_W3", "X", 16), ("D2DX2_C2", "X", 16),
                           ("VDDX_U1", "X", 24), ("VDDX_W3", "X", 24), ("VDDX_W3", "Z", 24), ("FDDX_U1", "X", 24)):
            vv = v if not op.startswith(("DDX", "D2DX2", "D4DX4")) else None
            t = timeit(lambda: capi.stencil(g, op, d, f, v=vv, scale1=dx, out=out), flush=flush)
            rows.append((f"{op}/{d}", t, bpp))
        t = timeit(lambda: capi.bracket_arakawa(g, f, h, dx, dz, out=out), flush=flush)
        rows.append(("bracket_arakawa", t, 24))
        t = timeit(lambda: capi.elementwise(g, [(2.0, f), (1.0, h, v)], out=out), flush=flush)
        rows.append(("2a + b*c", t, 32))
        t = timeit(lambda: capi.ddz_fft(g, f, out=out), flush=flush)
        rows.append(("DDZ_FFT", t, 16))
        t = timeit(lambda: out.copy_(f), flush=flush)
        rows.append(("torch copy_", t, 16 * f.numel() / npts))
        print(f"== {name}: {npts/1e6:.2f} M interior points")
        for n, t, bpp in rows:
            gbs = bpp * npts / t / 1e9
            print(f"{n:20s} {t*1e6:9.1f} us  {npts/t/1e9:8.2f} Gpts/s  {gbs:8.1f} GB/s algorithmic  {gbs/PEAK:6.1%} of measured HBM peak")


if __name__ == "__main__":
    main()
