"""One LaplaceCyclic solve per ablation setting at the HW3D size (run under an ncu launch list to see each
kernel's duration with the phase switched off).  BOUT_B200_LAPB_DBG is read at every solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi
dims = (516, 132, 512, 2, 2)
g = capi.Grid(*dims)
one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device="cuda")
md = dict(dx=one(0.2), g11=one(1.0), g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0), g_22=one(1.0))
f = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
out = torch.zeros_like(f)
lap = capi.Laplace(g, md, 0.2 * g.nz)
lap.solve(f, out=out)
for dbg in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 15]:
    os.environ["BOUT_B200_LAPB_DBG"] = str(dbg)
    lap.solve(f, out=out)
    lap.solve(f, out=out)
torch.cuda.synchronize()
