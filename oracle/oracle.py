"""Python bindings for the two CPU oracles (TEST INFRASTRUCTURE ONLY).

* ``Port``  -> oracle/_ref/liboracle_port.so, built from oracle/bout_oracle.c, the plain-C
  restatement of the reference algorithms (each C function cites reference file:line).
* ``Ref``   -> oracle/_ref/libboutref.so, the BOUT++ reference's own sources compiled in place from
  /root/reference against the shims in oracle/shim/ (see oracle/Makefile, oracle/ref_driver.cxx).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (bout-dev_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_ref", "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libboutref.so")

# stencil op codes (oracle/bout_oracle.c enum; the CUDA library uses the same numbering)
OPS = {
    "DDX_C2": 0, "DDX_C4": 1, "DDX_W2": 2, "DDX_S2": 3, "DDX_W3": 4,
    "D2DX2_C2": 10, "D2DX2_C4": 11, "D4DX4_C2": 12,
    "VDDX_C2": 20, "VDDX_C4": 21, "VDDX_U1": 22, "VDDX_U2": 23, "VDDX_U3": 24, "VDDX_W3": 25,
    "FDDX_U1": 30, "FDDX_U2": 31, "FDDX_C2": 32, "FDDX_C4": 33,
}
DIRS = {"X": 0, "Y": 1, "Z": 2}

# Laplacian flags (include/bout/invert_laplace.hxx:72-104 of the reference)
INVERT_DC_GRAD, INVERT_AC_GRAD, INVERT_AC_LAP, INVERT_SYM, INVERT_SET, INVERT_RHS = 1, 2, 4, 8, 16, 32
INVERT_DC_LAP, INVERT_BNDRY_ONE, INVERT_DC_GRADPAR, INVERT_DC_GRADPARINV = 64, 128, 256, 512
INVERT_IN_CYLINDER = 1024
INVERT_ZERO_DC, INVERT_START_NEW, INVERT_BOTH_BNDRY_ONE = 1, 2, 4


class _Grid(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "ny", "nz", "xstart", "xend", "ystart", "yend")]


@dataclass
class Grid:
    """Local grid including guards (reference: src/mesh/impls/bout/boutmesh.cxx:391-404)."""
    nx: int
    ny: int
    nz: int
    mxg: int
    myg: int

    def __post_init__(self):
        for n in ("nx", "ny", "nz", "mxg", "myg"):
            setattr(self, n, int(getattr(self, n)))

    @property
    def xstart(self): return self.mxg
    @property
    def xend(self): return self.nx - self.mxg - 1
    @property
    def ystart(self): return self.myg
    @property
    def yend(self): return self.ny - self.myg - 1
    @property
    def shape3(self): return (self.nx, self.ny, self.nz)
    @property
    def shape2(self): return (self.nx, self.ny)
    @property
    def interior(self):
        return (slice(self.xstart, self.xend + 1), slice(self.ystart, self.yend + 1), slice(None))
    @property
    def n_interior(self):
        return (self.xend - self.xstart + 1) * (self.yend - self.ystart + 1) * self.nz

    def c(self):
        return _Grid(self.nx, self.ny, self.nz, self.xstart, self.xend, self.ystart, self.yend)


def _p(a):
    if a is None:
        return None
    assert a.dtype in (np.float64, np.complex128) and a.flags.c_contiguous
    return a.ctypes.data_as(C.c_void_p)


class _LaplMetric(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dx", "dz", "g11", "g33", "g13", "g_11", "g_22", "G1",
                                          "G3", "IntShiftTorsion")] + \
               [("zlength", C.c_double)] + \
               [(n, C.c_int) for n in ("all_terms", "nonuniform", "inc_int_shear", "periodic_x",
                                       "first_x", "last_x")]


class _LaplFlags(C.Structure):
    _fields_ = [("global_", C.c_int), ("inner", C.c_int), ("outer", C.c_int)]


@dataclass
class Metric:
    """Field2D metric bundle (reference: include/bout/coordinates.hxx:83-116)."""
    dx: np.ndarray
    dy: np.ndarray
    dz: np.ndarray
    g11: np.ndarray
    g22: np.ndarray
    g33: np.ndarray
    g12: np.ndarray
    g13: np.ndarray
    g23: np.ndarray
    g_11: np.ndarray
    g_22: np.ndarray
    g_33: np.ndarray
    g_12: np.ndarray
    g_13: np.ndarray
    g_23: np.ndarray
    J: np.ndarray
    Bxy: np.ndarray
    G1: np.ndarray
    G2: np.ndarray
    G3: np.ndarray
    d1_dx: np.ndarray
    d1_dy: np.ndarray
    non_uniform: bool = True

    NAMES = ("dx", "dy", "dz", "g11", "g22", "g33", "g12", "g13", "g23", "g_11", "g_22", "g_33",
             "g_12", "g_13", "g_23", "J", "Bxy", "G1", "G2", "G3", "d1_dx", "d1_dy")


class Port:
    """ctypes view of oracle/bout_oracle.c."""

    def __init__(self):
        self.lib = C.CDLL(PORT_SO)
        self.lib.bo_rk4_stage.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        self.lib.bo_rk4_final.argtypes = [C.c_size_t, C.c_void_p, C.c_double] + [C.c_void_p] * 5
        self.lib.bo_rk3ssp_stage2.argtypes = [C.c_size_t] + [C.c_void_p] * 3 + [C.c_double, C.c_void_p]
        self.lib.bo_rk3ssp_stage3.argtypes = [C.c_size_t] + [C.c_void_p] * 3 + [C.c_double, C.c_void_p]

    # -- stencils
    def index_deriv(self, g: Grid, op: str, direction: str, f, v=None, yup=None, ydown=None, out=None):
        out = np.zeros(g.shape3) if out is None else out
        gc = g.c()
        self.lib.bo_index_deriv(C.byref(gc), OPS[op], DIRS[direction], _p(v), _p(f), _p(yup),
                                _p(ydown), _p(out))
        return out

    def index_ddz_fft(self, g: Grid, order: int, f, fft_filter=0.0):
        out = np.zeros(g.shape3)
        gc = g.c()
        self.lib.bo_index_ddz_fft.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        self.lib.bo_index_ddz_fft(C.byref(gc), order, fft_filter, _p(f), _p(out))
        return out

    def rfft(self, x):
        n = len(x)
        out = np.zeros(n // 2 + 1, dtype=np.complex128)
        self.lib.bo_rfft(_p(np.ascontiguousarray(x)), n, _p(out))
        return out

    def irfft(self, X, n):
        out = np.zeros(n)
        self.lib.bo_irfft(_p(np.ascontiguousarray(X)), n, _p(out))
        return out

    # -- brackets
    def bracket_arakawa(self, g: Grid, f, gg, dx, dz):
        out = np.zeros(g.shape3)
        gc = g.c()
        self.lib.bo_bracket_arakawa(C.byref(gc), _p(f), _p(gg), _p(dx), _p(dz), _p(out))
        return out

    def bracket_arakawa_3d2d(self, g: Grid, f, g2d, dx, dz):
        out = np.zeros(g.shape3)
        gc = g.c()
        self.lib.bo_bracket_arakawa_3d2d(C.byref(gc), _p(f), _p(g2d), _p(dx), _p(dz), _p(out))
        return out

    # -- Laplacian
    def _lm(self, g: Grid, m: Metric, all_terms=True, nonuniform=None, periodic_x=False,
            first_x=True, last_x=True):
        lm = _LaplMetric()
        self._keep = [np.ascontiguousarray(getattr(m, n)) for n in
                      ("dx", "dz", "g11", "g33", "g13", "g_11", "g_22", "G1", "G3")]
        for n, a in zip(("dx", "dz", "g11", "g33", "g13", "g_11", "g_22", "G1", "G3"), self._keep):
            setattr(lm, n, a.ctypes.data)
        lm.IntShiftTorsion = None
        lm.zlength = float(m.dz[0, 0]) * g.nz
        lm.all_terms = int(all_terms)
        lm.nonuniform = int(m.non_uniform if nonuniform is None else nonuniform)
        lm.inc_int_shear = 0
        lm.periodic_x = int(periodic_x)
        lm.first_x = int(first_x)
        lm.last_x = int(last_x)
        return lm

    def tridag_coefs(self, g, m, jx, jy, kwave, c1=None, c2=None, d=None, **kw):
        lm = self._lm(g, m, **kw)
        gc = g.c()
        abc = np.zeros(3, dtype=np.complex128)
        self.lib.bo_tridag_coefs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        self.lib.bo_tridag_coefs(C.byref(gc), C.byref(lm), jx, jy, kwave, _p(c1), _p(c2), _p(d),
                                 abc[0:].ctypes.data, abc[1:].ctypes.data, abc[2:].ctypes.data)
        return abc

    def cyclic_solve(self, a, b, c, r, periodic=False):
        a, b, c, r = (np.ascontiguousarray(v, dtype=np.complex128) for v in (a, b, c, r))
        x = np.zeros_like(r)
        self.lib.bo_cyclic_solve.restype = C.c_int
        st = self.lib.bo_cyclic_solve(len(r), int(periodic), _p(a), _p(b), _p(c), _p(r), _p(x))
        if st:
            raise ZeroDivisionError("Zero pivot in CyclicReduce::reduce")
        return x

    def laplace_cyclic(self, g, m, rhs, x0=None, flags=(0, 0, 0), maxmode=None, A=None, C1=None,
                       C2=None, D=None, **kw):
        lm = self._lm(g, m, **kw)
        fl = _LaplFlags(*flags)
        gc = g.c()
        if maxmode is None:
            maxmode = g.nz // 2
        ones = np.ones(g.shape2)
        A = np.zeros(g.shape2) if A is None else A
        C1 = ones if C1 is None else C1
        C2 = ones if C2 is None else C2
        D = ones if D is None else D
        x0 = np.zeros(g.shape3) if x0 is None else x0
        out = np.zeros(g.shape3)
        self.lib.bo_laplace_cyclic.restype = C.c_int
        st = self.lib.bo_laplace_cyclic(C.byref(gc), C.byref(lm), C.byref(fl), int(maxmode), _p(A),
                                        _p(C1), _p(C2), _p(D), _p(rhs), _p(x0), _p(out))
        if st:
            raise ZeroDivisionError("Zero pivot in CyclicReduce::reduce")
        return out

    def delp2_fft(self, g, m, f, **kw):
        lm = self._lm(g, m, **kw)
        gc = g.c()
        out = np.zeros(g.shape3)
        self.lib.bo_delp2_fft(C.byref(gc), C.byref(lm), _p(f), _p(out))
        return out

    # -- boundaries
    def bndry_x(self, g, kind, side, f, dx, val=0.0):
        gc = g.c()
        self.lib.bo_bndry_x.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        self.lib.bo_bndry_x(C.byref(gc), {"dirichlet": 0, "neumann": 1, "free_o2": 2, "free_o3": 3}[kind], side, val, _p(dx), _p(f))
        return f

    def dc(self, g, f):
        out = np.zeros(g.shape2)
        gc = g.c()
        self.lib.bo_dc(C.byref(gc), _p(f), _p(out))
        return out

    # -- numpy-level compositions of the C kernels ------------------------------------------
    # Coordinate derivatives: index derivative times reciprocal spacing.
    # Field3D / Field2D multiplies by the reciprocal (generated_fieldops.cxx:253-274).
    def ddx(self, g, m, f, method="C2"):           # coordinates.cxx:1467-1478 (IncIntShear off)
        return self.index_deriv(g, "DDX_" + method, "X", f) * (1.0 / m.dx)[:, :, None]

    def ddy(self, g, m, f, method="C2"):           # derivs.cxx:69-73
        return self.index_deriv(g, "DDX_" + method, "Y", f) * (1.0 / m.dy)[:, :, None]

    def ddz(self, g, m, f, method="C2"):           # derivs.cxx:82-86
        if method == "FFT":
            return self.index_ddz_fft(g, 1, f) * (1.0 / m.dz)[:, :, None]
        return self.index_deriv(g, "DDX_" + method, "Z", f) * (1.0 / m.dz)[:, :, None]

    def d2dx2(self, g, m, f, method="C2", first="C2"):  # derivs.cxx:149-166
        r = self.index_deriv(g, "D2DX2_" + method, "X", f) * (1.0 / m.dx ** 2)[:, :, None]
        if m.non_uniform:
            r = r + (m.d1_dx[:, :, None] * self.index_deriv(g, "DDX_" + first, "X", f)) \
                * (1.0 / m.dx)[:, :, None]
        return r

    def d2dz2(self, g, m, f, method="C2"):         # derivs.cxx:241-255
        if method == "FFT":
            return self.index_ddz_fft(g, 2, f) * (1.0 / m.dz ** 2)[:, :, None]
        return self.index_deriv(g, "D2DX2_" + method, "Z", f) * (1.0 / m.dz ** 2)[:, :, None]

    def zrows(self, g, f, modify, region=None):
        """rfft every (x,y) row in the region, let `modify(spectrum, x, y)` edit it, irfft back
        (the common shape of filter/lowPass/shiftZ, src/field/field3d.cxx:620-741)."""
        out = np.zeros(g.shape3)
        xs, xe, ys, ye = region if region is not None else (0, g.nx - 1, 0, g.ny - 1)
        for x in range(xs, xe + 1):
            for y in range(ys, ye + 1):
                spec = self.rfft(f[x, y])
                modify(spec, x, y)
                out[x, y] = self.irfft(spec, g.nz)
        return out

    def zfilter(self, g, f, n0):                       # field3d.cxx:620-661
        def mod(s, x, y):
            keep = s[n0] if 0 <= n0 < len(s) else 0.0
            s[:] = 0.0
            if 0 <= n0 < len(s):
                s[n0] = keep
        return self.zrows(g, f, mod)

    def zlowpass(self, g, f, zmax, keep_zonal=True):   # field3d.cxx:664-710
        if (zmax >= g.nz // 2 or zmax < 0) and keep_zonal:
            return f.copy()
        def mod(s, x, y):
            s[zmax + 1:] = 0.0
            if not keep_zonal:
                s[0] = 0.0
        return self.zrows(g, f, mod)

    def zshift(self, g, f, angle, zlength):            # field3d.cxx:715-741
        angle = np.broadcast_to(angle, g.shape2)
        zlength = np.broadcast_to(zlength, g.shape2)
        def mod(s, x, y):
            for jz in range(1, g.nz // 2 + 1):
                kwave = jz * 2.0 * np.pi / zlength[x, y]
                s[jz] *= complex(np.cos(kwave * angle[x, y]), -np.sin(kwave * angle[x, y]))
        if g.nz == 1:
            return f.copy()
        return self.zrows(g, f, mod)

    def rk4_final(self, start, dt, k1, k2, k3, k4):
        out = np.zeros_like(start)
        self.lib.bo_rk4_final(start.size, _p(start), dt, _p(k1), _p(k2), _p(k3), _p(k4), _p(out))
        return out

    def rk4_stage(self, start, k, coef_dt):
        out = np.zeros_like(start)
        self.lib.bo_rk4_stage(start.size, _p(start), _p(k), coef_dt, _p(out))
        return out


class Ref:
    """ctypes view of the reference's own sources (oracle/ref_driver.cxx).  One live mesh at a time."""

    _live = None

    def __init__(self, ini_text: str):
        if Ref._live is not None:
            Ref._live.close()
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.boutref_last_error.restype = C.c_char_p
        L.boutref_op.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p]
        L.boutref_time_op.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_double, C.c_int]
        L.boutref_time_op.restype = C.c_double
        L.boutref_time_laplace_solve.restype = C.c_double
        L.boutref_time_laplace_solve.argtypes = [C.c_void_p, C.c_int]
        L.boutref_laplace_set_coef.argtypes = [C.c_char, C.c_void_p]
        L.boutref_tridag_coefs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(L.boutref_init(ini_text.encode()))
        Ref._live = self
        v = (C.c_int * 12)()
        L.boutref_mesh_info(v)
        self.info = list(v)
        self.grid = Grid(v[0], v[1], v[2], v[3], v[5])
        self.periodic_x = bool(v[10])

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.boutref_last_error().decode(errors="replace")[-2000:])

    def close(self):
        if Ref._live is self:
            self.lib.boutref_finalise()
            Ref._live = None

    def metric(self) -> Metric:
        vals = {}
        for n in Metric.NAMES:
            a = np.zeros(self.grid.shape2)
            self._check(self.lib.boutref_get_metric(n.encode(), _p(a)))
            vals[n] = a
        return Metric(**vals)

    def set_metric(self, name, a):
        self._check(self.lib.boutref_set_metric(name.encode(), _p(np.ascontiguousarray(a))))

    def geometry(self):
        self._check(self.lib.boutref_geometry())

    def expr3d(self, expr):
        out = np.zeros(self.grid.shape3)
        self._check(self.lib.boutref_expr3d(expr.encode(), _p(out)))
        return out

    def op(self, name, *inputs, method="", scalar=0.0):
        ins = [np.ascontiguousarray(a) for a in inputs]
        arr = (C.c_void_p * max(3, len(ins)))(*[a.ctypes.data for a in ins])
        out = np.zeros(self.grid.shape3)
        self._check(self.lib.boutref_op(name.encode(), method.encode(), len(ins), arr, scalar, _p(out)))
        return out

    def parallel_slice(self, f, offset):
        """Field3D::ynext(offset) after Mesh::communicate (ParallelTransform::calcParallelSlices)."""
        out = np.zeros(self.grid.shape3)
        self._check(self.lib.boutref_parallel_slice(_p(np.ascontiguousarray(f)), int(offset), _p(out)))
        return out

    def zshift(self):
        out = np.zeros(self.grid.shape2)
        zl = C.c_double(0.0)
        self._check(self.lib.boutref_zshift(_p(out), C.byref(zl)))
        return out, zl.value

    def time_op(self, name, *inputs, method="", scalar=0.0, reps=5):
        ins = [np.ascontiguousarray(a) for a in inputs]
        arr = (C.c_void_p * max(3, len(ins)))(*[a.ctypes.data for a in ins])
        t = self.lib.boutref_time_op(name.encode(), method.encode(), len(ins), arr, scalar, reps)
        if t < 0:
            raise RuntimeError(self.lib.boutref_last_error().decode(errors="replace")[-2000:])
        return t

    def rfft(self, x):
        out = np.zeros(len(x) // 2 + 1, dtype=np.complex128)
        self._check(self.lib.boutref_rfft(_p(np.ascontiguousarray(x)), len(x), _p(out)))
        return out

    def irfft(self, X, n):
        out = np.zeros(n)
        self._check(self.lib.boutref_irfft(_p(np.ascontiguousarray(X)), n, _p(out)))
        return out

    def cyclic_reduce(self, a, b, c, r, periodic=False):
        a, b, c, r = (np.ascontiguousarray(v, dtype=np.complex128) for v in (a, b, c, r))
        nsys, n = r.shape
        x = np.zeros_like(r)
        self._check(self.lib.boutref_cyclic_reduce(nsys, n, int(periodic), _p(a), _p(b), _p(c), _p(r), _p(x)))
        return x

    def tridag_coefs(self, jx, jy, jz, ccoef=None, d=None):
        abc = np.zeros(3, dtype=np.complex128)
        self._check(self.lib.boutref_tridag_coefs(jx, jy, jz, _p(ccoef), _p(d), _p(abc)))
        return abc

    def laplace_create(self, section="laplace"):
        self._check(self.lib.boutref_laplace_create(section.encode()))

    def laplace_set_coef(self, which, a):
        self._check(self.lib.boutref_laplace_set_coef(which.encode(), _p(np.ascontiguousarray(a))))

    def laplace_set_flags(self, global_=0, inner=0, outer=0):
        self._check(self.lib.boutref_laplace_set_flags(global_, inner, outer))

    def laplacexz_solve(self, section, A2d, B2d, b, x0):
        """LaplaceXZ::create(mesh, &options[section]) -> setCoefs(A, B) -> solve(b, x0); unwritten cells are 0."""
        out = np.zeros_like(b)
        self.lib.boutref_laplacexz_solve.argtypes = [C.c_char_p] + [C.c_void_p] * 5
        self._check(self.lib.boutref_laplacexz_solve(section.encode(), _p(np.ascontiguousarray(A2d)),
                                                     _p(np.ascontiguousarray(B2d)), _p(np.ascontiguousarray(b)),
                                                     _p(np.ascontiguousarray(x0)), _p(out)))
        return out

    def invert_par_solve(self, section, f, A=None, B=None, Cc=None, D=None, E=None):
        """InvertPar::create(&options[section]) -> setCoefA..E -> solve(f); unwritten cells are 0."""
        out = np.zeros_like(f)
        keep = [None if a is None else np.ascontiguousarray(a) for a in (A, B, Cc, D, E)]
        self.lib.boutref_invert_par_solve.argtypes = [C.c_char_p] + [C.c_void_p] * 7
        self._check(self.lib.boutref_invert_par_solve(section.encode(), *[None if a is None else _p(a) for a in keep],
                                                      _p(np.ascontiguousarray(f)), _p(out)))
        return out

    def invert_par_solve_all_y(self, section, f, A=None, B=None, Cc=None, D=None, E=None):
        """As invert_par_solve, returning every y of the interior x range (open field lines: the solver writes the y
        guard cells too)."""
        out = np.zeros_like(f)
        keep = [None if a is None else np.ascontiguousarray(a) for a in (A, B, Cc, D, E)]
        self.lib.boutref_invert_par_solve_all_y.argtypes = [C.c_char_p] + [C.c_void_p] * 7
        self._check(self.lib.boutref_invert_par_solve_all_y(section.encode(), *[None if a is None else _p(a) for a in keep],
                                                            _p(np.ascontiguousarray(f)), _p(out)))
        return out

    def laplace_solve_perp(self, jy, b, x0):
        """Laplacian::solve(FieldPerp, FieldPerp) on the y = jy slices of two Field3D arrays -> [nx][nz]."""
        out = np.zeros((b.shape[0], b.shape[2]))
        self.lib.boutref_laplace_solve_perp.argtypes = [C.c_int] + [C.c_void_p] * 3
        self._check(self.lib.boutref_laplace_solve_perp(jy, _p(np.ascontiguousarray(b)), _p(np.ascontiguousarray(x0)), _p(out)))
        return out

    def laplace_solve(self, b, x0=None):
        out = np.zeros(self.grid.shape3)
        self._check(self.lib.boutref_laplace_solve(_p(np.ascontiguousarray(b)),
                                                   _p(None if x0 is None else np.ascontiguousarray(x0)),
                                                   _p(out)))
        return out

    def time_laplace_solve(self, b, reps=3):
        t = self.lib.boutref_time_laplace_solve(_p(np.ascontiguousarray(b)), reps)
        if t < 0:
            raise RuntimeError(self.lib.boutref_last_error().decode(errors="replace")[-2000:])
        return t

    def communicate(self, *fields):
        fs = [np.ascontiguousarray(f).copy() for f in fields]
        arr = (C.c_void_p * len(fs))(*[f.ctypes.data for f in fs])
        self._check(self.lib.boutref_communicate(len(fs), arr))
        return fs

    def apply_boundary(self, f, condition):
        f = np.ascontiguousarray(f).copy()
        self._check(self.lib.boutref_apply_boundary(_p(f), condition.encode()))
        return f


    # -- unchanged reference example models (examples/*/) through the reference's RK solvers
    def model_create(self, name, solver="rk4"):
        self._check(self.lib.boutref_model_create(name.encode(), solver.encode()))
        self.lib.boutref_model_varname.restype = C.c_char_p
        self.model_vars = [self.lib.boutref_model_varname(i).decode()
                           for i in range(self.lib.boutref_model_nvars())]
        return self.model_vars

    def model_get(self, name):
        out = np.zeros(self.grid.shape3)
        self._check(self.lib.boutref_model_get(self.model_vars.index(name), _p(out)))
        return out

    def model_set(self, name, a):
        self._check(self.lib.boutref_model_set(self.model_vars.index(name), _p(np.ascontiguousarray(a))))

    def model_ddt(self, name):
        out = np.zeros(self.grid.shape3)
        self._check(self.lib.boutref_model_get_ddt(self.model_vars.index(name), _p(out)))
        return out

    def model_rhs(self, t=0.0):
        self.lib.boutref_model_rhs.argtypes = [C.c_double]
        self._check(self.lib.boutref_model_rhs(t))

    def model_time_rhs(self, reps=3):
        self.lib.boutref_model_time_rhs.restype = C.c_double
        t = self.lib.boutref_model_time_rhs(reps)
        if t < 0:
            raise RuntimeError(self.lib.boutref_last_error().decode(errors="replace")[-2000:])
        return t

    def elmpb_get(self, name):
        """examples/elm-pb only: a post-init equilibrium field / scalar or a per-RHS auxiliary by name
        (oracle/elmpb_export.cxx).  Returns a float for scalars."""
        k = C.c_int(0)
        self.lib.boutref_elmpb_get.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int)]
        if self.lib.boutref_elmpb_get(name.encode(), None, C.byref(k)) != 0:
            raise KeyError(name)
        if k.value == 0:
            out = np.zeros(1)
        else:
            out = np.zeros(self.grid.shape2 if k.value == 2 else self.grid.shape3)
        self._check(self.lib.boutref_elmpb_get(name.encode(), _p(out), C.byref(k)))
        return float(out[0]) if k.value == 0 else out

    def omp_threads(self):
        """OpenMP threads the reference build uses in this process."""
        return int(self.lib.boutref_omp_threads())

    def model_rhs_calls(self):
        """Solver::resetRHSCounter (src/solver/solver.cxx:946): run_rhs + run_convective calls since the last query."""
        return int(self.lib.boutref_model_rhs_calls())

    def model_solve(self, nout, output_step):
        self.lib.boutref_model_solve.argtypes = [C.c_int, C.c_double]
        self._check(self.lib.boutref_model_solve(nout, output_step))


# ---- numpy restatements of whole-model right-hand sides -------------------------------------------
def _sh(f, dx=0, dy=0, dz=0):
    """f shifted so that result[x,y,z] = f[x+dx, y+dy, z+dz] (z periodic; x,y edges wrap and are
    never used: only interior points are compared)."""
    return np.roll(f, (-dx, -dy, -dz), axis=(0, 1, 2))


def si_bracket(f, g, dx, dz):
    """bracket(FieldAccessor f, FieldAccessor g, i) -- include/bout/single_index_ops.hxx:93-128."""
    Jpp = ((_sh(f, dz=1) - _sh(f, dz=-1)) * (_sh(g, dx=1) - _sh(g, dx=-1))
           - (_sh(f, dx=1) - _sh(f, dx=-1)) * (_sh(g, dz=1) - _sh(g, dz=-1)))
    Jpx = (_sh(g, dx=1) * (_sh(f, 1, 0, 1) - _sh(f, 1, 0, -1)) - _sh(g, dx=-1) * (_sh(f, -1, 0, 1) - _sh(f, -1, 0, -1))
           - _sh(g, dz=1) * (_sh(f, 1, 0, 1) - _sh(f, -1, 0, 1)) + _sh(g, dz=-1) * (_sh(f, 1, 0, -1) - _sh(f, -1, 0, -1)))
    Jxp = (_sh(g, 1, 0, 1) * (_sh(f, dz=1) - _sh(f, dx=1)) - _sh(g, -1, 0, -1) * (_sh(f, dx=-1) - _sh(f, dz=-1))
           - _sh(g, -1, 0, 1) * (_sh(f, dz=1) - _sh(f, dx=-1)) + _sh(g, 1, 0, -1) * (_sh(f, dx=1) - _sh(f, dz=-1)))
    return (Jpp + Jpx + Jxp) / (12 * dx[:, :, None] * dz[:, :, None])


def si_ddz(f, dz):
    """single_index_ops.hxx:201-209."""
    return 0.5 * (_sh(f, dz=1) - _sh(f, dz=-1)) / dz[:, :, None]


def si_delp2(f, m):
    """single_index_ops.hxx:222-246 (non-uniform correction always included)."""
    b = lambda a: a[:, :, None]
    dx, dz = b(m.dx), b(m.dz)
    return ((b(m.G1) + b(m.d1_dx) * b(m.g11)) * (_sh(f, dx=1) - _sh(f, dx=-1)) / (2.0 * dx)
            + b(m.G3) * (_sh(f, dz=1) - _sh(f, dz=-1)) / (2.0 * dz)
            + b(m.g11) * (_sh(f, dx=1) - 2.0 * f + _sh(f, dx=-1)) / dx ** 2
            + b(m.g33) * (_sh(f, dz=1) - 2.0 * f + _sh(f, dz=-1)) / dz ** 2
            + 2 * b(m.g13) * ((_sh(f, 1, 0, 1) - _sh(f, -1, 0, 1)) - (_sh(f, 1, 0, -1) - _sh(f, -1, 0, -1)))
            / (4. * dz * dx))


def si_div_par_grad_par(f, m):
    """single_index_ops.hxx:255-278 with identity parallel transform (yup/ydown = the field)."""
    b = lambda a: a[:, :, None]
    s2 = lambda a, d: np.roll(a, -d, axis=1)
    dy, J, g_22 = m.dy, m.J, m.g_22
    gu = 2. * (_sh(f, dy=1) - f) / b(dy + s2(dy, 1))
    fu = gu * b(J + s2(J, 1)) / b(g_22 + s2(g_22, 1))
    gl = 2. * (f - _sh(f, dy=-1)) / b(dy + s2(dy, -1))
    fl = gl * b(J + s2(J, -1)) / b(g_22 + s2(g_22, -1))
    return (fu - fl) / b(dy * J)


def hw3d_rhs(m, n, vort, phi, alpha, kappa, Dn, Dvort):
    """examples/hasegawa-wakatani-3d/hw.cxx:65-74; valid on interior points."""
    div_current = alpha * si_div_par_grad_par(phi - n, m) if n.shape[1] > 1 else 0.0
    ddt_n = -si_bracket(phi, n, m.dx, m.dz) - div_current - kappa * si_ddz(phi, m.dz) + Dn * si_delp2(n, m)
    ddt_v = -si_bracket(phi, vort, m.dx, m.dz) - div_current + Dvort * si_delp2(vort, m)
    return ddt_n, ddt_v


def make_ini(nx, ny, nz, mxg=2, myg=2, dx="0.2", dy="1.0", dz="0.2", periodic_x=False, extra="",
             mesh_extra=""):
    """BOUT.inp text for a mesh-from-options (reference: src/mesh/data/gridfromoptions.cxx)."""
    return f"""
MXG = {mxg}
MYG = {myg}
periodicX = {'true' if periodic_x else 'false'}
[mesh]
nx = {nx}
ny = {ny}
nz = {nz}
dx = {dx}
dy = {dy}
dz = {dz}
{mesh_extra}
{extra}
"""


def halo_self_reference(g: Grid, f, periodic_x=False, periodic_y=True):
    """What Mesh::communicate leaves in the guard cells when the rank is its own neighbour
    (BoutMesh::sendY/sendX/wait, src/mesh/impls/bout/boutmesh.cxx:1245-1555): Y first, then X with
    the y-guard corners (src/mesh/mesh.cxx:351-362).  Pinned against the reference build in
    tests/test_oracle_reference.py."""
    f = f.copy()
    if periodic_y and g.myg > 0:
        ny_int = g.yend - g.ystart + 1
        f[:, :g.ystart] = f[:, ny_int:ny_int + g.ystart]
        f[:, g.yend + 1:] = f[:, g.ystart:g.ystart + (g.ny - g.yend - 1)]
    if periodic_x and g.mxg > 0:
        nx_int = g.xend - g.xstart + 1
        f[:g.xstart] = f[nx_int:nx_int + g.xstart]
        f[g.xend + 1:] = f[g.xstart:g.xstart + (g.nx - g.xend - 1)]
    return f
