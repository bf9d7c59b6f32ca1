"""Phase clocks of a few rows-kernel CTAs (BOUT_B200_LAPB_DBG=32: device printf)."""
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
for _ in range(2):
    lap.solve(f, out=out)
torch.cuda.synchronize()
os.environ["BOUT_B200_LAPB_DBG"] = "32"
lap.solve(f, out=out)
torch.cuda.synchronize()
