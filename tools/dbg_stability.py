import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import bench
dev = torch.device("cuda", 0)
pkg = ge.load_package(); capi, models = pkg.capi, pkg.models
def run(MX, MY, dt, steps, tag):
    bench.MXSUB, bench.MYSUB = MX, MY
    G = capi.Grid(MX + 4, MY + 4, bench.MZ, 2, 2)
    one = lambda v: torch.full(G.shape2, v, dtype=torch.float64, device=dev)
    met = dict(dx=one(0.2), dy=one(1.0), dz=one(0.2), J=one(1.0), g_22=one(1.0), g11=one(1.0), g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0))
    M = models.HW3D(G, met, device=dev, mesh=None, **bench.HW)
    n0, v0 = bench.initial_state(0, 1, torch, dev)
    M.set_state(n0.to(dev), v0.to(dev))
    I = G.interior
    for s in range(steps):
        M.rk4_step(dt)
        n = M.n[I]
        amax = n.abs().amax(dim=(1, 2))
        ix = int(amax.argmax())
        spec = torch.fft.rfft(n[ix], dim=-1).abs().amax(dim=0)
        print(f"{tag} step {s}: max|n| {float(amax.max()):.3e} at x={ix}; max|phi| {float(M.phi[I].abs().max()):.3e}; dominant kz {int(spec.argmax())}", flush=True)
        if not np.isfinite(float(amax.max())) or float(amax.max()) > 1e10: break
run(512, 16, 1e-3, 12, "MX512 dt1e-3")
run(1024, 16, 1e-3, 12, "MX1024 dt1e-3")
run(1024, 16, 2.5e-4, 12, "MX1024 dt2.5e-4")
run(768, 16, 1e-3, 12, "MX768 dt1e-3")
