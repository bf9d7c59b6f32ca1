"""Timing experiment (results at the y wrap rows are NOT right here): does running the LaplaceCyclic solve of one half
of the y rows beside the fused HW3D kernel of the other half (FP64-bound next to HBM-bound) shorten an RHS?
V0 whole solve then whole fused kernel; V1 halves back to back on one stream; V2 solve(B) on a second stream beside
fused(A); V3 quarters, pipelined."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import __graft_entry__ as ge
pkg = ge.load_package()
capi, models = pkg.capi, pkg.models
dev = torch.device("cuda", 0)
g = capi.Grid(bench.MXSUB + 2 * bench.MXG, bench.MYSUB + 2 * bench.MYG, bench.MZ, bench.MXG, bench.MYG)
md = bench.hw3d_metric(torch, g, dev)
m = models.HW3D(g, md, device=dev, **bench.HW)
n0, v0 = (torch.rand(g.shape3, dtype=torch.float64, device=dev) - 0.5 for _ in range(2))
m.set_state(n0, v0)
zl = float(m.md["dz"][0, 0].item()) * g.nz


def sub(a, b):
    gg = capi.Grid(g.nx, g.ny, g.nz, g.mxg, g.myg)
    gg.c.ystart, gg.c.yend = a, b
    return gg


def parts(k):
    ny = g.yend - g.ystart + 1
    edges = [g.ystart + ny * i // k for i in range(k + 1)]
    out = []
    for i in range(k):
        gs = sub(edges[i], edges[i + 1] - 1)
        lap = capi.Laplace(gs, m.md, zl, nonuniform=True)
        lap.workspace(dev)
        out.append((gs, lap))
    return out


E = capi.make_epilogue((m.n, m.vort), (m.an, m.av), (m.tn, m.tv), (m.an, m.av), 1e-6, 1e-6)
side = torch.cuda.Stream(device=dev)


def v0_():
    m.lap.solve(m.sv, out=m.phi)
    m._fused(g, m.sn, m.sv, E, False, None)


def seq(ps):
    def f():
        for gs, lap in ps:
            lap.solve(m.sv, out=m.phi)
        for gs, lap in ps:
            m._fused(gs, m.sn, m.sv, E, False, None)
    return f


def pipe(ps):
    def f():
        main = torch.cuda.current_stream()
        evs = []
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for gs, lap in ps:
                lap.solve(m.sv, out=m.phi)
                e = torch.cuda.Event()
                e.record(side)
                evs.append(e)
        for i, (gs, lap) in enumerate(ps):
            main.wait_event(evs[min(i + 1, len(ps) - 1)])   # rows of part i read phi of part i + 1's first row
            m._fused(gs, m.sn, m.sv, E, False, None)
    return f


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


m.sn.copy_(n0)
m.sv.copy_(v0)
print(f"V0 whole solve + whole fused kernel: {timeit(v0_):.1f} us")
for k in (2, 4, 8):
    ps = parts(k)
    print(f"k = {k}: back to back {timeit(seq(ps)):.1f} us, solves on a second stream beside the fused parts {timeit(pipe(ps)):.1f} us")
