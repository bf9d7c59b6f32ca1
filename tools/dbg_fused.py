import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import __graft_entry__ as ge
capi = ge.load_package().capi; orc = ge.load_oracle(); port = orc.Port()
from util import make_metric, seeded_fields, rel_err
grid=(21,5,128,2,1)
g, og = capi.Grid(*grid), orc.Grid(*grid)
m = make_metric(orc, g.shape2, g.nz)
f, x0 = seeded_fields(g.shape3, 2)
dev=lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
md = {n: dev(getattr(m, n)) for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22")}
lap = capi.Laplace(g, md, float(m.dz[0, 0]) * g.nz)
os.environ["BOUT_B200_LAP_FUSED"]="1"
a = lap.solve(dev(f), dev(x0)).cpu().numpy()
os.environ["BOUT_B200_LAP_FUSED"]="0"
b = lap.solve(dev(f), dev(x0)).cpu().numpy()
want = port.laplace_cyclic(og, m, f, x0=x0)
region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
print("fused vs plain", np.abs(a-b)[region].max(), "fused vs oracle", rel_err(a, want, region), "plain vs oracle", rel_err(b, want, region))
d=np.abs(a-b)[region]; print("per-x max diff", d.max(axis=(1,2)))
