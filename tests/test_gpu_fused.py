"""GPU parity: fused model right-hand sides (one kernel per RHS, RK stage update folded in) vs the
oracle restatement of the reference models (examples/hasegawa-wakatani-3d/hw.cxx:58-74,
examples/hasegawa-wakatani/hw.cxx:94-118).  Tolerance 1e-12 relative (stencil/bracket/Delp2 class)."""
import numpy as np
import pytest
import torch

from util import make_metric, rel_err, seeded_fields

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_metric(capi, m):
    return capi.make_metric({n: dev(getattr(m, n)) for n in ("dx", "dy", "dz", "J", "g_22", "g11", "g33",
                                                              "g13", "G1", "G3", "d1_dx")})


# nz = 128, 256, 1024 take the TMA-pipelined kernel (k_hw3d_rhs_tma), the others the register-marching one
@pytest.mark.parametrize("grid", [(20, 8, 32, 2, 2), (9, 7, 7, 2, 1), (70, 6, 16, 2, 2), (14, 1, 16, 2, 0),
                                  (12, 6, 128, 2, 2), (75, 5, 256, 2, 1), (9, 3, 1024, 2, 1), (10, 1, 128, 2, 0)])
@pytest.mark.parametrize("trivial", [True, False])
def test_hw3d_rhs(capi, orc, grid, trivial):
    g = capi.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz, trivial=trivial)
    n, vort, phi = seeded_fields(g.shape3, 3)
    alpha, kappa, Dn, Dvort = 1.0, 0.5, 1e-3, 2e-3
    want_n, want_v = orc.hw3d_rhs(m, n, vort, phi, alpha, kappa, Dn, Dvort)
    dn, dv = torch.zeros(g.shape3, dtype=torch.float64, device="cuda"), torch.zeros(g.shape3, dtype=torch.float64, device="cuda")
    capi.hw3d_rhs(g, gpu_metric(capi, m), alpha, kappa, Dn, Dvort, dev(n), dev(vort), dev(phi), dn, dv)
    assert rel_err(dn.cpu().numpy(), want_n, g.interior) < 1e-12
    assert rel_err(dv.cpu().numpy(), want_v, g.interior) < 1e-12


@pytest.mark.parametrize("grid", [(20, 8, 32, 2, 2), (70, 5, 128, 2, 2), (11, 4, 512, 2, 1)])
def test_hw3d_rhs_with_rk_epilogue(capi, orc, grid):
    g = capi.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz)
    n, vort, phi, sn, sv, an, av = seeded_fields(g.shape3, 7)
    want_n, want_v = orc.hw3d_rhs(m, n, vort, phi, 1.0, 0.5, 1e-3, 2e-3)
    a, b = 0.005, 0.01 / 3
    z = lambda: torch.zeros(g.shape3, dtype=torch.float64, device="cuda")
    out_n, out_v, acc_n, acc_v = z(), z(), dev(an), dev(av)
    ep = capi.make_epilogue((dev(sn), dev(sv)), (acc_n, acc_v), (out_n, out_v), (acc_n, acc_v), a, b)
    capi.hw3d_rhs(g, gpu_metric(capi, m), 1.0, 0.5, 1e-3, 2e-3, dev(n), dev(vort), dev(phi), epilogue=ep)
    I = g.interior
    assert rel_err(out_n.cpu().numpy(), sn + a * want_n, I) < 1e-12
    assert rel_err(out_v.cpu().numpy(), sv + a * want_v, I) < 1e-12
    assert rel_err(acc_n.cpu().numpy(), an + b * want_n, I) < 1e-12
    assert rel_err(acc_v.cpu().numpy(), av + b * want_v, I) < 1e-12
    # first stage: acc_in = NULL means acc = start + b*ddt
    acc2 = z()
    ep = capi.make_epilogue((dev(sn), dev(sv)), None, (None, None), (acc2, None), a, b)
    capi.hw3d_rhs(g, gpu_metric(capi, m), 1.0, 0.5, 1e-3, 2e-3, dev(n), dev(vort), dev(phi), epilogue=ep)
    assert rel_err(acc2.cpu().numpy(), sn + b * want_n, I) < 1e-12


@pytest.mark.parametrize("grid", [(20, 1, 32, 2, 0), (36, 4, 16, 2, 1), (9, 1, 7, 2, 0)])
@pytest.mark.parametrize("modified", [True, False])
def test_hw_convective(capi, orc, port, grid, modified):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz, trivial=True)
    n, vort, phi = seeded_fields(g.shape3, 3)
    alpha, kappa = 1.0, 0.5
    zn = zp = None
    nz_n, nz_phi = n, phi
    if modified:  # nonzonal = f - averageY(DC(f)) (hw.cxx:106-110; smoothing.cxx:272-323)
        ys = slice(og.ystart, og.yend + 1)
        zn = np.ascontiguousarray(np.repeat(port.dc(og, n)[:, ys].mean(axis=1, keepdims=True), g.ny, axis=1))
        zp = np.ascontiguousarray(np.repeat(port.dc(og, phi)[:, ys].mean(axis=1, keepdims=True), g.ny, axis=1))
        nz_n, nz_phi = n - zn[:, :, None], phi - zp[:, :, None]
    coupling = alpha * (nz_phi - nz_n)
    want_n = -port.bracket_arakawa(og, phi, n, m.dx, m.dz) + coupling - kappa * port.ddz(og, m, phi, "C2")
    want_v = -port.bracket_arakawa(og, phi, vort, m.dx, m.dz) + coupling
    z = lambda: torch.zeros(g.shape3, dtype=torch.float64, device="cuda")
    dn, dv = z(), z()
    capi.hw_convective(g, dev(m.dx), dev(m.dz), alpha, kappa, dev(n), dev(vort), dev(phi),
                       None if zn is None else dev(zn), None if zp is None else dev(zp), dn, dv)
    assert rel_err(dn.cpu().numpy(), want_n, g.interior) < 1e-12
    assert rel_err(dv.cpu().numpy(), want_v, g.interior) < 1e-12
