"""One isolated + one emulated-coupled LaplaceCyclic solve at the bench size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi
g = capi.Grid(516, 132, 512, 2, 2)
one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device="cuda")
md = dict(dx=one(0.2), g11=one(1.0), g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0), g_22=one(1.0))
f = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
out = torch.zeros_like(f)
lap = capi.Laplace(g, md, 0.2 * g.nz)
for _ in range(3):
    lap.solve(f, out=out)
lap2 = capi.Laplace(g, md, 0.2 * g.nz, first_x=True, last_x=False)
iface = lap2.reduce(f)
for _ in range(2):
    lap2.reduce(f, iface=iface)
ia = torch.cat([iface, iface])
lap2.finish(ia, 2, 0, out)
torch.cuda.synchronize()
