"""A few LaplaceCyclic solves at the bench sizes (for ncu captures): HW3D grid with a y-independent metric,
elm-pb grid with a y-dependent one, and one emulated X-decomposed reduce / finish pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi
for dims, ydep in (((516, 132, 512, 2, 2), False), ((516, 68, 256, 2, 2), True)):
    g = capi.Grid(*dims)
    one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device="cuda")
    g11 = one(1.0)
    if ydep:
        g11 = torch.linspace(1.0, 1.3, g.ny, dtype=torch.float64, device="cuda")[None, :].expand(g.nx, g.ny).contiguous()
    md = dict(dx=one(0.2), g11=g11, g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0), g_22=one(1.0))
    f = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
    out = torch.zeros_like(f)
    lap = capi.Laplace(g, md, 0.2 * g.nz)
    for _ in range(3):
        lap.solve(f, out=out)
    if len(sys.argv) > 1 and sys.argv[1] == "coupled":
        lap2 = capi.Laplace(g, md, 0.2 * g.nz, first_x=True, last_x=False)
        iface = lap2.reduce(f)
        ia = torch.cat([iface, iface])
        lap2.finish(ia, 2, 0, out)
    torch.cuda.synchronize()
