"""A few launches of the configs[1] operators at 256x64x256 (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
capi = ge.load_package().capi
g = capi.Grid(260, 68, 256, 2, 2)
sets = [tuple(torch.rand(g.shape3, dtype=torch.float64, device="cuda") for _ in range(3)) + (torch.zeros(g.shape3, dtype=torch.float64, device="cuda"),)
        for _ in range(6)]
dx = torch.full(g.shape2, 0.2, dtype=torch.float64, device="cuda")
for rep in range(2):
    for a, b, c, o in sets:
        capi.stencil(g, "DDX_C2", "X", a, scale1=dx, out=o)
        capi.stencil(g, "DDX_C4", "X", a, scale1=dx, out=o)
        capi.bracket_arakawa(g, a, b, dx, dx, out=o)
        capi.ddz_fft(g, a, out=o)
        capi.elementwise(g, [(2.0, a), (1.0, b, c)], out=o)
torch.cuda.synchronize()
