"""elm-pb RHS at BASELINE config C5 size (516 x 64(+4) x 256): ms per RHS with CUDA events.
Usage: python tools/time_elmpb.py [reps] [nx ny nz]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
capi, models, elm_setup = pkg.capi, pkg.models, pkg.elm_setup

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
nx, nyi, nz = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (516, 64, 256)
m = models.ELMpb.from_setup(elm_setup.elmpb_setup(nx, nyi, nz))
m.set_state(*models.elmpb_initial_state(m.g))
for _ in range(2):
    m.rhs()
torch.cuda.synchronize()
n0 = capi.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    m.rhs()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
npts = m.g.n_interior
fin = all(bool(torch.isfinite(f).all()) for f in (m.ddt_U, m.ddt_P, m.ddt_Psi))
print(f"elm-pb {nx}x{nyi}x{nz}: {ms:.3f} ms/RHS  {npts / ms / 1e6:.3f} Gpts/s  launches/RHS {(capi.launch_count() - n0) / reps:.0f}  "
      f"finite={fin}  max|ddt| U {m.ddt_U.abs().max().item():.3e} P {m.ddt_P.abs().max().item():.3e} Psi {m.ddt_Psi.abs().max().item():.3e}")
