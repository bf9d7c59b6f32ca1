import os, sys
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi
g = capi.Grid(516, 132, 512, 2, 2)
one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device="cuda")
md = dict(dx=one(0.2), g11=one(1.0), g33=one(1.0), g13=one(0.0), G1=one(0.0), G3=one(0.0), g_11=one(1.0), g_22=one(1.0))
f = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
o = torch.zeros_like(f)
L = capi.Laplace(g, md, 0.2 * g.nz)
for _ in range(3):
    L.solve(f, out=o)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    L.solve(f, out=o)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("BOUT_B200_LAPB_NT"), os.environ.get("BOUT_B200_LAPB_PD"), f"{e0.elapsed_time(e1)/30*1e3:.1f} us per solve")
