"""numpy restatement of the elm-pb right-hand side and of the shifted-metric operators it is built from
(TEST INFRASTRUCTURE ONLY -- see oracle/oracle.py; nothing in bout-dev_b200/ imports this).

Every function cites the reference file:line it follows.  The tridiagonal pieces (LaplaceCyclic, Delp2) come
from the plain-C port (oracle/bout_oracle.c through oracle.Port); everything else is numpy.  Pinned on CPU by
tests/test_elm_oracle.py against outputs of the reference itself: tests/golden/ops_shifted.npz (operators) and
tests/golden/elmpb.npz (one Solver::run_rhs of the unchanged examples/elm-pb/elm_pb.cxx).

Arrays are [nx][ny][nz] / [nx][ny] with guard cells, as everywhere in this repository.
"""
from __future__ import annotations

import numpy as np

WENO_SMALL = 1.0e-8  # include/bout/index_derivs.hxx


# ---- z-FFT row operations -------------------------------------------------------------------------------------
def shift_z(f, angle, zlength):
    """shiftZ of every (x,y) row by angle[x,y] (src/field/field3d.cxx:715-741, ShiftedMetric::shiftZ
    src/mesh/parallel/shiftedmetric.cxx:200-224): mode k >= 1 times exp(-i k 2 pi / zlength * angle); c2r drops
    the imaginary part of the Nyquist mode (numpy's irfft does the same)."""
    nz = f.shape[2]
    F = np.fft.rfft(f, axis=2)
    k = np.arange(nz // 2 + 1) * 2.0 * np.pi / zlength
    ph = np.exp(-1j * k[None, None, :] * np.asarray(angle)[:, :, None])
    ph[:, :, 0] = 1.0
    return np.fft.irfft(F * ph, n=nz, axis=2)


def to_field_aligned(f, zshift, zlength):      # shiftedmetric.cxx:130-140
    return shift_z(f, -zshift, zlength)


def from_field_aligned(f, zshift, zlength):    # shiftedmetric.cxx:142-152
    return shift_z(f, zshift, zlength)


def parallel_slices(f, zshift, zlength):
    """ShiftedMetric::calcParallelSlices for offsets +-1 (shiftedmetric.cxx:226-246, phases :109-123): yup holds at
    row y+1 the field's row y+1 shifted by zShift(y) - zShift(y+1); ydown likewise at row y-1."""
    yup, ydown = np.zeros_like(f), np.zeros_like(f)
    ang_up = np.zeros_like(zshift)
    ang_up[:, 1:] = zshift[:, :-1] - zshift[:, 1:]          # at row y' = y+1: zs(y'-1) - zs(y')
    ang_dn = np.zeros_like(zshift)
    ang_dn[:, :-1] = zshift[:, 1:] - zshift[:, :-1]         # at row y' = y-1: zs(y'+1) - zs(y')
    yup[:, 1:] = shift_z(f, ang_up, zlength)[:, 1:]
    ydown[:, :-1] = shift_z(f, ang_dn, zlength)[:, :-1]
    return yup, ydown


def z_filter(f, n0):                            # field3d.cxx:620-661
    F = np.fft.rfft(f, axis=2)
    G = np.zeros_like(F)
    G[:, :, n0] = F[:, :, n0]
    return np.fft.irfft(G, n=f.shape[2], axis=2)


def z_lowpass(f, zmax, keep_zonal):             # field3d.cxx:664-710
    nz = f.shape[2]
    if (zmax >= nz // 2 or zmax < 0) and keep_zonal:
        return f
    F = np.fft.rfft(f, axis=2)
    F[:, :, zmax + 1:] = 0.0
    if not keep_zonal:
        F[:, :, 0] = 0.0
    return np.fft.irfft(F, n=nz, axis=2)


# ---- index stencils (src/mesh/index_derivs.cxx) -----------------------------------------------------------------
def _sh(f, axis, d):
    """f shifted so that result[i] = f[i + d] along axis (edges wrap: only interior results are used; z is periodic)."""
    return np.roll(f, -d, axis=axis)


def ddx_c4(f, axis):                            # index_derivs.cxx:75
    return (8.0 * _sh(f, axis, 1) - 8.0 * _sh(f, axis, -1) + _sh(f, axis, -2) - _sh(f, axis, 2)) / 12.0


def weno3(v, f, axis):
    """VDDX_WENO3 (index_derivs.cxx:190-216) along axis for velocity v at the point."""
    fm, fmm, fp, fpp = _sh(f, axis, -1), _sh(f, axis, -2), _sh(f, axis, 1), _sh(f, axis, 2)
    pos = v > 0.0
    r_pos = (WENO_SMALL + (f - 2.0 * fm + fmm) ** 2) / (WENO_SMALL + (fp - 2.0 * f + fm) ** 2)
    r_neg = (WENO_SMALL + (fpp - 2.0 * fp + f) ** 2) / (WENO_SMALL + (fp - 2.0 * f + fm) ** 2)
    d_pos = -fmm + 3.0 * fm - 3.0 * f + fp
    d_neg = -fm + 3.0 * f - 3.0 * fp + fpp
    r = np.where(pos, r_pos, r_neg)
    d = np.where(pos, d_pos, d_neg)
    w = 1.0 / (1.0 + 2.0 * r * r)
    return v * (0.5 * ((fp - fm) - w * d))


def b0xgrad_dot_grad(md, phi, A, phi_slices=None, a_slices=None, zshift=None, zlength=None):
    """b0xGrad_dot_Grad (src/mesh/difops.cxx:457-601) with the elm-pb methods (first C4 / C2 / C4, upwind W3 / C2 /
    W3).  phi, A: Field3D or Field2D arrays; *_slices = (yup, ydown) for Field3D operands that carry parallel
    slices.  A Field2D A carries none, so its y-advection term is evaluated in field-aligned coordinates
    (include/bout/index_derivs_interface.hxx:360-390): toFieldAligned -> VDDY -> fromFieldAligned."""
    b3 = lambda a: a[:, :, None] if a.ndim == 2 else a
    rdx, rdy, rdz = 1.0 / b3(md["dx"]), 1.0 / b3(md["dy"]), 1.0 / b3(md["dz"])
    g_22, g_23, g_12 = b3(md["g_22"]), b3(md["g_23"]), b3(md["g_12"])

    def ddy(f, sl):
        if f.ndim == 2:
            return 0.5 * (_sh(f, 1, 1) - _sh(f, 1, -1))[:, :, None]
        up, dn = sl
        return 0.5 * (_sh(up, 1, 1) - _sh(dn, 1, -1))

    dpdx = ddx_c4(b3(phi), 0) * rdx
    dpdy = ddy(phi, phi_slices) * rdy
    dpdz = 0.0 if phi.ndim == 2 else ddx_c4(phi, 2) * rdz
    vx = g_22 * dpdz - g_23 * dpdy
    vy = g_23 * dpdx - g_12 * dpdz
    vz = g_12 * dpdy - g_22 * dpdx
    shape = phi.shape if phi.ndim == 3 else A.shape
    vx, vy, vz = (np.broadcast_to(v, shape) for v in (vx, vy, vz))
    A3 = np.broadcast_to(b3(A), shape)
    res = weno3(vx, A3, 0) * rdx
    if A.ndim == 2:
        yterm = vy * (0.5 * (_sh(A, 1, 1) - _sh(A, 1, -1)))[:, :, None] * rdy      # VDDY_C2
        yterm = from_field_aligned(to_field_aligned(yterm, zshift, zlength), zshift, zlength)
        res = res + yterm
    else:
        up, dn = a_slices
        res = res + vy * 0.5 * (_sh(up, 1, 1) - _sh(dn, 1, -1)) * rdy
        res = res + weno3(vz, A3, 2) * rdz
    return res / (b3(md["J"]) * np.sqrt(g_22))


# ---- boundaries, communicate, smoothing ----------------------------------------------------------------------
def halo_y_periodic(f, myg):
    """Mesh::communicate on one rank of an all-core grid: y periodic (boutmesh.cxx:1245-1555)."""
    ny = f.shape[1]
    f[:, :myg] = f[:, ny - 2 * myg:ny - myg]
    f[:, ny - myg:] = f[:, myg:2 * myg]
    return f


def bndry_dirichlet_o2(f, mxg, myg, side):
    """BoundaryDirichlet::apply, value 0 (src/mesh/boundary_standard.cxx:527-581): first guard = -f_in, rest 0."""
    ys = slice(myg, f.shape[1] - myg)
    if side == "xin":
        f[mxg - 1, ys] = -f[mxg, ys]
        f[:mxg - 1, ys] = 0.0
    else:
        nx = f.shape[0]
        f[nx - mxg, ys] = -f[nx - mxg - 1, ys]
        f[nx - mxg + 1:, ys] = 0.0
    return f


def bndry_neumann(f, mxg, myg, side):
    """BoundaryNeumann::apply, gradient 0, CELL_CENTRE, width 2 (boundary_standard.cxx:2167-2190): mirror."""
    ys = slice(myg, f.shape[1] - myg)
    nx = f.shape[0]
    if side == "xin":
        f[mxg - 1, ys] = f[mxg, ys]
        f[mxg - 2, ys] = f[mxg + 1, ys]
    else:
        f[nx - mxg, ys] = f[nx - mxg - 1, ys]
        f[nx - mxg + 1, ys] = f[nx - mxg - 2, ys]
    return f


def bndry_zerolaplace(f, md, zlength, mxg, myg, side):
    """BoundaryZeroLaplace::apply(Field3D) (boundary_standard.cxx:2630-2686)."""
    nx, ny, nz = f.shape
    k = np.arange(nz // 2 + 1) * 2.0 * np.pi / zlength
    xin, b = (mxg, -1) if side == "xin" else (nx - mxg - 1, 1)
    for y in range(myg, ny - myg):
        c0 = np.fft.rfft(f[xin, y]) / nz
        c1 = np.fft.rfft(f[xin - b, y]) / nz
        grad = c0[0] - c1[0]
        x = xin + b
        while 0 <= x < nx:
            c0 = c0.copy()
            c0[0] += grad
            coef = -1.0 * np.sqrt(md["g33"][x, y] / md["g11"][x, y]) * md["dx"][x, y]
            c0[1:] *= np.exp(coef * k[1:])
            f[x, y] = np.fft.irfft(c0 * nz, n=nz)
            x += b
    return f


def smooth_x(f):
    """smooth_x (src/physics/smoothing.cxx:46-73), before its communicate."""
    r = f.copy()
    r[1:-1] = 0.5 * f[1:-1] + 0.25 * (f[:-2] + f[2:])
    return r


# ---- the model ---------------------------------------------------------------------------------------------------
def elmpb_rhs(port, og, metric, md, eq, zshift, zlength, dnorm, hyperresist, U, P, Psi, filter_z_mode=1,
              low_pass_z=16, zonal=(False, True, False)):
    """ELMpb::rhs with the shipped [highbeta] switches (examples/elm-pb/elm_pb.cxx:1240-1955) after Solver::pre_rhs
    (solver.cxx:1486-1500).  port/og/metric: oracle.Port, oracle.Grid, oracle.Metric for the tridiagonal pieces;
    md: name -> Field2D (same metric); eq: J0, P0, phi0, B0, b0xcv_x/y/z (Field2D), eta (Field3D).
    Returns dict(ddt_U, ddt_P, ddt_Psi, phi, Jpar, Jpar2, U, P, Psi) (the states after boundary conditions)."""
    mxg, myg = og.mxg, og.myg
    U, P, Psi = U.copy(), P.copy(), Psi.copy()
    # Solver::pre_rhs: [all] dirichlet_o2, [P] bndry_core = neumann, [Psi] zerolaplace both sides
    bndry_dirichlet_o2(U, mxg, myg, "xin"), bndry_dirichlet_o2(U, mxg, myg, "xout")
    bndry_neumann(P, mxg, myg, "xin"), bndry_dirichlet_o2(P, mxg, myg, "xout")
    bndry_zerolaplace(Psi, md, zlength, mxg, myg, "xin"), bndry_zerolaplace(Psi, md, zlength, mxg, myg, "xout")
    for f in (U, P, Psi):                                                    # mesh->communicate(comms)
        halo_y_periodic(f, myg)
    sU, sP, sPsi = (parallel_slices(f, zshift, zlength) for f in (U, P, Psi))
    B0 = eq["B0"][:, :, None]
    # phi = phiSolver->solve(U) - 0.5 dnorm P / B0                            elm_pb.cxx:1472-1477
    phi = port.laplace_cyclic(og, metric, U, flags=(0, 4, 5), all_terms=True, nonuniform=True)
    phi = phi - 0.5 * dnorm * P / B0
    halo_y_periodic(phi, myg)
    sphi = parallel_slices(phi, zshift, zlength)
    phi[0], phi[-1] = phi[1], phi[-2]                                        # :1498-1516 (MXG = 2)

    def bc_j(f):                                                             # [J] bndry_core = neumann
        bndry_neumann(f, mxg, myg, "xin"), bndry_dirichlet_o2(f, mxg, myg, "xout")
        return f
    # Jpar = Delp2(Psi); BC; communicate; smooth_x; BC; communicate          :1518-1563
    Jpar = bc_j(port.delp2_fft(og, metric, Psi, all_terms=True, nonuniform=True))
    halo_y_periodic(Jpar, myg)
    Jpar = halo_y_periodic(smooth_x(Jpar), myg)
    halo_y_periodic(bc_j(Jpar), myg)
    sJ = parallel_slices(Jpar, zshift, zlength)
    d2J = port.delp2_fft(og, metric, Jpar, all_terms=True, nonuniform=True)
    Jpar2 = halo_y_periodic(bc_j(d2J.copy()), myg)                           # :1565-1572
    # ddt(Psi) = -Grad_par(phi B0)/B0 + eta Jpar - bracket(phi0, Psi, STD) B0 - eta hyperresist Delp2(Jpar)
    inv_sg = 1.0 / np.sqrt(md["g_22"])[:, :, None]
    rdy = 1.0 / md["dy"][:, :, None]
    fa = to_field_aligned(phi * B0, zshift, zlength)        # no slices on phi*B0: index_derivs_interface.hxx:205-211
    gp = from_field_aligned(0.5 * (_sh(fa, 1, 1) - _sh(fa, 1, -1)), zshift, zlength) * rdy * inv_sg
    bg = b0xgrad_dot_grad(md, eq["phi0"], Psi, None, sPsi) / md["Bxy"][:, :, None]     # difops.cxx:805-808
    ddt_Psi = -gp / B0 + eq["eta"] * Jpar - bg * B0 - eq["eta"] * hyperresist * d2J
    # ddt(U) = B0^2 b0xGrad_dot_Grad(Psi_loc, J0) + b0xcv . Grad(P) - B0^2 Grad_par(Jpar) - b0xGrad_dot_Grad(phi0, U)
    t1 = b0xgrad_dot_grad(md, Psi, eq["J0"], sPsi, None, zshift, zlength)
    dPx = ddx_c4(P, 0) / md["dx"][:, :, None]
    dPy = 0.5 * (_sh(sP[0], 1, 1) - _sh(sP[1], 1, -1)) * rdy
    dPz = ddx_c4(P, 2) / md["dz"][:, :, None]
    gpJ = 0.5 * (_sh(sJ[0], 1, 1) - _sh(sJ[1], 1, -1)) * rdy * inv_sg
    t4 = b0xgrad_dot_grad(md, eq["phi0"], U, None, sU)
    ddt_U = (B0 ** 2 * t1 + eq["b0xcv_x"][:, :, None] * dPx + eq["b0xcv_y"][:, :, None] * dPy
             + eq["b0xcv_z"][:, :, None] * dPz - B0 ** 2 * gpJ - t4)
    # ddt(P) = -b0xGrad_dot_Grad(phi, P0) - b0xGrad_dot_Grad(phi0, P)
    ddt_P = (-b0xgrad_dot_grad(md, phi, eq["P0"], sphi, None, zshift, zlength)
             - b0xgrad_dot_grad(md, eq["phi0"], P, None, sP))
    out = {}
    for name, f, zon in (("ddt_Psi", ddt_Psi, zonal[2]), ("ddt_U", ddt_U, zonal[0]), ("ddt_P", ddt_P, zonal[1])):
        f = z_filter(f, filter_z_mode)                                       # :1931-1955
        if low_pass_z > 0:
            f = z_lowpass(f, low_pass_z, zon)
        out[name] = f
    out.update(phi=phi, Jpar=Jpar, Jpar2=Jpar2, U=U, P=P, Psi=Psi)
    return out


# ---- Y operators on parallel slices (the remaining elm-pb / 6-field vocabulary) ----------------------------------
def _b3(a):
    return a[:, :, None] if a.ndim == 2 else a


def ddy_slices(md, f, sl):
    """DDY on a field with parallel slices: index DDY_C2 (YOrthogonal, index_derivs_interface.hxx:199-203) / dy
    (coordinates.cxx:1488-1499)."""
    up, dn = sl
    return 0.5 * (_sh(up, 1, 1) - _sh(dn, 1, -1)) / _b3(md["dy"])


def d2dy2_slices(md, f, sl):
    """D2DY2 (src/sys/derivs.cxx:194-213): index D2DY2_C2 / dy^2 + d1_dy * index DDY_C2 / dy (non-uniform mesh)."""
    up, dn = sl
    dy = _b3(md["dy"])
    second = (_sh(up, 1, 1) + _sh(dn, 1, -1) - 2.0 * f) / (dy * dy)
    return second + _b3(md["d1_dy"]) * (0.5 * (_sh(up, 1, 1) - _sh(dn, 1, -1))) / dy


def grad_par(md, f, sl):                          # coordinates.cxx:1529-1535
    return ddy_slices(md, f, sl) / np.sqrt(_b3(md["g_22"]))


def vddy_slices(md, v, f, sl):                    # VDDY_C2 (index_derivs.cxx:148) / dy
    up, dn = sl
    return v * 0.5 * (_sh(up, 1, 1) - _sh(dn, 1, -1)) / _b3(md["dy"])


def vpar_grad_par(md, v, f, sl):                  # coordinates.cxx:1541-1566
    return vddy_slices(md, v, f, sl) / np.sqrt(_b3(md["g_22"]))


def div_par(md, f, sl):
    """Div_par (coordinates.cxx:1571-1594): Bxy * Grad_par(f / Bxy), the slices divided by Bxy at their own rows."""
    B = _b3(md["Bxy"])
    up, dn = sl
    return B * grad_par(md, f / B, (up / B, dn / B))


def grad2_par2(md, f, sl):
    """Grad2_par2 (coordinates.cxx:1611-1628, coefficient :1889-1905):
    DDY(1/sqrt(g_22)) / sqrt(g_22) * DDY(f) + D2DY2(f) / g_22."""
    g_22 = md["g_22"]
    sg = 1.0 / np.sqrt(g_22)
    dsg = np.zeros_like(sg)
    dsg[:, 1:-1] = 0.5 * (sg[:, 2:] - sg[:, :-2])
    coef = dsg / md["dy"] * sg
    return _b3(coef) * ddy_slices(md, f, sl) + d2dy2_slices(md, f, sl) / _b3(g_22)


def average_y(f, myg):                            # src/physics/smoothing.cxx:272-323 (one rank in y)
    mean = f[:, myg:f.shape[1] - myg].mean(axis=1, keepdims=True)
    return np.broadcast_to(mean, f.shape).copy()
