"""Set-up-time host logic for the elm-pb workload (NOT on the per-RHS path): evaluates the synthetic
closed-form tokamak grid of BASELINE config C5 and restates what ELMpb::init does with it
(examples/elm-pb/elm_pb.cxx:317-337 grid reads, :770-783 shear handling, :786-910 normalisation,
:975-994 resistivity profile, :1040-1058 metric tensor, :1123-1131 phi0) and the pieces of
Coordinates::geometry the RHS kernels consume (src/mesh/coordinates.cxx:1018-1026 G1, G3).

The grid expressions are those of tests/golden/cases.py:elmpb_ini, evaluated on BOUT++'s cell-centred
global coordinates (symmetricGlobalX/Y = true: x = (0.5 + gx - MXG)/MX, y = 2 pi (0.5 + gy - MYG)/MY,
src/mesh/impls/bout/boutmesh.cxx:3159-3185).  numpy only; tests/test_elm_setup.py pins every output
against the reference's own post-init state (tests/golden/elmpb.npz).
"""
from __future__ import annotations

import numpy as np

MU0 = 4.0e-7 * np.pi
MI = 2.0 * 1.6726e-27          # elm_pb.cxx: const BoutReal Mi = 2.0 * 1.6726e-27 (deuterium)


def _ddx_c4(f):
    """index DDX_C4 of a Field2D on interior x (index_derivs.cxx:75), zero elsewhere."""
    r = np.zeros_like(f)
    r[2:-2] = (8.0 * f[3:-1] - 8.0 * f[1:-3] + f[:-4] - f[4:]) / 12.0
    return r


def _ddy_c2(f):
    r = np.zeros_like(f)
    r[:, 1:-1] = 0.5 * (f[:, 2:] - f[:, :-2])
    return r


def elmpb_setup(nx, ny_interior, nz, mxg=2, myg=2, zperiod=15, density=1e19, vacuum_pressure=0.02,
                vacuum_trans=0.01, vac_lund=1e8, core_lund=1e8, dia_fact=1.0):
    """Returns dict(grid=(nx, ny, nz, mxg, myg), md={Field2D metrics}, eq={J0, P0, phi0, B0, b0xcv_*, eta},
    zShift, zlength, dnorm) as float64 numpy arrays, [nx][ny] with guard cells."""
    mx, ny = nx - 2 * mxg, ny_interior + 2 * myg
    x = ((0.5 + np.arange(nx) - mxg) / mx)[:, None] * np.ones((1, ny))
    y = (2.0 * np.pi * (0.5 + np.arange(ny) - myg) / ny_interior)[None, :] * np.ones((nx, 1))
    pi = np.pi
    # ---- the grid "file" (tests/golden/cases.py:elmpb_ini [mesh] section)
    rminor = 0.4 + 0.3 * x
    qsafe = 1.2 + 2 * x ** 2
    Rxy = 3 + rminor * np.cos(y)
    Btxy = 6 / Rxy
    Bpxy = rminor * Btxy / (qsafe * Rxy)
    B0 = np.sqrt(Btxy ** 2 + Bpxy ** 2)
    hthe = rminor * (1 + 0.05 * np.cos(y))
    sinty = 1.5 * x * (y - pi)
    Bbar, Lbar = 2.0, 3.0
    dx = (0.15 / mx) * (1 + 0.2 * x)
    dy = 2 * pi / ny_interior + 0 * x
    zlength = 2 * pi / zperiod
    dz = zlength / nz + 0 * x
    zShift = qsafe * (y - pi + 0.2 * rminor * np.sin(y))
    J0 = 8e5 * (1 - x ** 2) * (1 + 0.1 * np.cos(y))
    P0 = 2e4 * (1 - np.tanh((x - 0.6) / 0.15)) + 500
    bx = 0.02 * np.sin(y) * (1 + 0.3 * x)
    by = 0.003 * np.cos(y)
    bz = 0.05 * np.cos(y) + 0.01 * x
    # Mesh::get communicates what it reads (src/mesh/mesh.cxx:188-207): the grid is all-core, so y guard
    # cells hold the periodic image, which matters for the one non-periodic expression (sinty)
    def ywrap(f):
        f = f.copy()
        f[:, :myg] = f[:, ny - 2 * myg:ny - myg]
        f[:, ny - myg:] = f[:, myg:2 * myg]
        return f
    sinty = ywrap(sinty)
    # ---- ELMpb::init: Dimits-style shear (IncIntShear off): I folded into b0xcv.z, then dropped
    bz = bz + sinty * bx
    Ishear = 0.0 * x
    # ---- normalisation (elm_pb.cxx:786-910)
    Va = np.sqrt(Bbar * Bbar / (MU0 * density * MI))
    Tbar = Lbar / Va
    dnorm = dia_fact * MI / (2.0 * 1.602e-19 * Bbar * Tbar)
    vac_resist = 1.0 / vac_lund if vac_lund > 0 else 0.0
    core_resist = 1.0 / core_lund if core_lund > 0 else 0.0
    J0 = -MU0 * Lbar * J0 / B0
    P0 = 2.0 * MU0 * P0 / (Bbar * Bbar)
    bx = bx / Bbar
    by = by * Lbar * Lbar
    bz = bz * Lbar * Lbar
    Rxy, Bpxy, Btxy, B0, hthe = Rxy / Lbar, Bpxy / Bbar, Btxy / Bbar, B0 / Bbar, hthe / Lbar
    dx = dx / (Lbar * Lbar * Bbar)
    # ---- resistivity profile (:975-994), uniform in z
    pnorm = P0.max()
    vp, vt = vacuum_pressure * pnorm, vacuum_trans * pnorm
    vac_mask = (1.0 - np.tanh((P0 - vp) / vt)) / 2.0
    eta2 = core_resist + (vac_resist - core_resist) * vac_mask
    # ---- metric tensor (:1040-1058)
    g11 = (Rxy * Bpxy) ** 2
    g22 = 1.0 / hthe ** 2
    g33 = Ishear ** 2 * g11 + B0 ** 2 / g11
    g12 = 0.0 * x
    g13 = -Ishear * g11
    g23 = -Btxy / (hthe * Bpxy * Rxy)
    J = hthe / Bpxy
    g_11 = 1.0 / g11 + (Ishear * Rxy) ** 2
    g_22 = (B0 * hthe / Bpxy) ** 2
    g_33 = Rxy * Rxy
    g_12 = Btxy * hthe * Ishear * Rxy / Bpxy
    g_13 = Ishear * Rxy * Rxy
    g_23 = Btxy * hthe * Rxy / Bpxy
    # ---- Coordinates::geometry: G1, G3 (coordinates.cxx:1018-1026; Field2D => DDZ terms vanish)
    G1 = (_ddx_c4(J * g11) / dx + _ddy_c2(J * g12) / dy) / J
    G3 = (_ddx_c4(J * g13) / dx + _ddy_c2(J * g23) / dy) / J
    phi0 = -0.5 * dnorm * P0 / B0
    md = dict(dx=dx, dy=dy, dz=dz, J=J, Bxy=B0.copy(), g11=g11, g22=g22, g33=g33, g12=g12, g13=g13, g23=g23,
              g_11=g_11, g_22=g_22, g_33=g_33, g_12=g_12, g_13=g_13, g_23=g_23, G1=G1, G3=G3)
    eq = dict(J0=J0, P0=P0, phi0=phi0, B0=B0, b0xcv_x=bx, b0xcv_y=by, b0xcv_z=bz,
              eta=np.ascontiguousarray(np.broadcast_to(eta2[:, :, None], (nx, ny, nz))))
    md = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in md.items()}
    eq = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in eq.items()}
    return dict(grid=(nx, ny, nz, mxg, myg), md=md, eq=eq, zShift=np.ascontiguousarray(zShift),
                zlength=zlength, dnorm=dnorm, hyperresist=1e-4)
