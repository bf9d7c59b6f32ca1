"""ctypes binding of include/bout_b200.h for the Python harness (tests/, bench.py).

Torch is used for device memory, streams and torch.distributed only; every numerical operation
below is a call into libbout_b200.so.  If the library is missing this module raises at import --
there is deliberately no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BOUT_B200_LIB") or os.path.join(_HERE, "csrc", "libbout_b200.so")  # env: A/B builds
if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(bout-dev_b200 has no CPU fallback)")
lib = C.CDLL(LIB_PATH)
lib.bout_b200_last_error.restype = C.c_char_p
lib.bout_b200_launch_count.restype = C.c_ulonglong

OPS = {
    "DDX_C2": 0, "DDX_C4": 1, "DDX_W2": 2, "DDX_S2": 3, "DDX_W3": 4,
    "D2DX2_C2": 10, "D2DX2_C4": 11, "D4DX4_C2": 12,
    "VDDX_C2": 20, "VDDX_C4": 21, "VDDX_U1": 22, "VDDX_U2": 23, "VDDX_U3": 24, "VDDX_W3": 25,
    "FDDX_U1": 30, "FDDX_U2": 31, "FDDX_C2": 32, "FDDX_C4": 33,
}
DIRS = {"X": 0, "Y": 1, "Z": 2}
REGIONS = {"RGN_NOBNDRY": 0, "RGN_ALL": 1, "RGN_NOX": 2, "RGN_NOY": 3}


class BoutB200Error(RuntimeError):
    """Raised when a C-ABI call returns a negative status (the host layer throws BoutException)."""

    def __init__(self, status):
        self.status = status
        super().__init__(f"bout_b200 status {status}: {lib.bout_b200_last_error().decode()}")


class CGrid(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "ny", "nz", "xstart", "xend", "ystart", "yend")]


class Grid:
    """Local grid, guard cells included (reference boutmesh.cxx:391-404)."""

    def __init__(self, nx, ny, nz, mxg, myg):
        self.nx, self.ny, self.nz, self.mxg, self.myg = nx, ny, nz, mxg, myg
        self.xstart, self.xend = mxg, nx - mxg - 1
        self.ystart, self.yend = myg, ny - myg - 1
        self.c = CGrid(nx, ny, nz, self.xstart, self.xend, self.ystart, self.yend)
        self.shape3 = (nx, ny, nz)
        self.shape2 = (nx, ny)
        self.n_interior = (self.xend - self.xstart + 1) * (self.yend - self.ystart + 1) * nz
        self.interior = (slice(self.xstart, self.xend + 1), slice(self.ystart, self.yend + 1),
                         slice(None))


def _check(rc):
    if rc != 0:
        raise BoutB200Error(rc)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype in (torch.float64, torch.complex128) and t.is_contiguous(), \
        "expected a contiguous fp64 CUDA tensor"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib.bout_b200_launch_count())


def stencil(g: Grid, op: str, direction: str, f, v=None, yup=None, ydown=None, scale1=None,
            op2=None, scale2=None, out=None, region="RGN_NOBNDRY"):
    out = torch.zeros_like(f) if out is None else out
    _check(lib.bout_b200_stencil(C.byref(g.c), OPS[op], DIRS[direction], REGIONS[region], _ptr(v),
                                 _ptr(f), _ptr(yup), _ptr(ydown), _ptr(scale1),
                                 -1 if op2 is None else OPS[op2], _ptr(scale2), _ptr(out), _stream()))
    return out


def bracket_arakawa(g: Grid, f, gg, dx, dz, out=None):
    out = torch.zeros_like(f) if out is None else out
    fn = lib.bout_b200_bracket_arakawa if gg.dim() == 3 else lib.bout_b200_bracket_arakawa_3d2d
    _check(fn(C.byref(g.c), _ptr(f), _ptr(gg), _ptr(dx), _ptr(dz), _ptr(out), _stream()))
    return out


def ddz_fft(g: Grid, f, order=1, fft_filter=0.0, scale=None, out=None, region="RGN_NOBNDRY"):
    out = torch.zeros_like(f) if out is None else out
    _check(lib.bout_b200_ddz_fft(C.byref(g.c), order, C.c_double(fft_filter), REGIONS[region],
                                 _ptr(f), _ptr(scale), _ptr(out), _stream()))
    return out


def zfilter(g: Grid, f, n0, out=None, region="RGN_ALL"):
    out = torch.zeros_like(f) if out is None else out
    _check(lib.bout_b200_zfilter(C.byref(g.c), REGIONS[region], n0, _ptr(f), _ptr(out), _stream()))
    return out


def zlowpass(g: Grid, f, zmax, keep_zonal=True, out=None, region="RGN_ALL"):
    out = torch.zeros_like(f) if out is None else out
    _check(lib.bout_b200_zlowpass(C.byref(g.c), REGIONS[region], zmax, int(keep_zonal), _ptr(f),
                                  _ptr(out), _stream()))
    return out


def zfilter_lowpass(g: Grid, f, n0, zmax, keep_zonal=True, out=None, region="RGN_ALL"):
    """filter(f, n0) then lowPass(., zmax, keep_zonal) in one pass."""
    out = torch.zeros_like(f) if out is None else out
    _check(lib.bout_b200_zfilter_lowpass(C.byref(g.c), REGIONS[region], n0, zmax, int(keep_zonal), _ptr(f),
                                         _ptr(out), _stream()))
    return out


def zshift(g: Grid, f, angle, zlength, out=None, region="RGN_ALL"):
    """angle / zlength: python floats or Field2D tensors."""
    out = torch.zeros_like(f) if out is None else out
    a2 = angle if torch.is_tensor(angle) else None
    z2 = zlength if torch.is_tensor(zlength) else None
    _check(lib.bout_b200_zshift(C.byref(g.c), REGIONS[region], _ptr(f), _ptr(a2),
                                C.c_double(0.0 if a2 is not None else angle), _ptr(z2),
                                C.c_double(0.0 if z2 is not None else zlength), _ptr(out), _stream()))
    return out


def dc(g: Grid, f, out=None):
    out = torch.zeros(g.shape2, dtype=torch.float64, device=f.device) if out is None else out
    _check(lib.bout_b200_dc(C.byref(g.c), _ptr(f), _ptr(out), _stream()))
    return out


class CLaplace(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22",
                                          "A", "C1", "C2", "D")] + \
               [("zlength", C.c_double)] + \
               [(n, C.c_int) for n in ("all_terms", "nonuniform", "global_flags",
                                       "inner_boundary_flags", "outer_boundary_flags", "maxmode",
                                       "periodic_x", "first_x", "last_x", "include_yguards_lower",
                                       "include_yguards_upper", "wrap_yguards")] + [("ypeer", C.c_void_p)]


def copy_xz(g, jy, field3d, perp, to_perp=True):
    """sliceXZ (to_perp) / Field3D = FieldPerp (not to_perp): plane jy of a Field3D <-> an [nx][nz] array."""
    _check(lib.bout_b200_copy_xz(C.byref(g.c), int(jy), _ptr(field3d), _ptr(perp), int(bool(to_perp)), _stream()))
    return perp if to_perp else field3d


class CLaplaceXZ(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("A", "B", "J", "g11", "g33", "dx")] + [
        ("zlength", C.c_double), ("inner_boundary_flags", C.c_int), ("outer_boundary_flags", C.c_int)]


class CInvertPar(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("A", "B", "C", "D", "E", "sg", "g_22", "dy", "d1_dy", "twist_shift")] + [
        ("zlength", C.c_double), ("closed", C.c_int)]


class LaplaceXZ:
    """LaplaceXZ "cyclic" (reference src/invert/laplacexz/impls/cyclic/laplacexz-cyclic.cxx): setCoefs(A, B), solve."""

    def __init__(self, g: Grid, metric: dict, zlength, A, B, inner_boundary_flags=0, outer_boundary_flags=0):
        self.g = g
        self.keep = dict(A=A, B=B, J=metric["J"], g11=metric["g11"], g33=metric["g33"], dx=metric["dx"])
        c = CLaplaceXZ()
        for n, t in self.keep.items():
            setattr(c, n, t.data_ptr())
        c.zlength = zlength
        c.inner_boundary_flags, c.outer_boundary_flags = inner_boundary_flags, outer_boundary_flags
        self.c, self.ws = c, None

    def solve(self, rhs, x0=None, out=None):
        out = torch.zeros_like(rhs) if out is None else out
        lib.bout_b200_laplacexz_workspace.restype = C.c_size_t
        n = lib.bout_b200_laplacexz_workspace(C.byref(self.g.c))
        if self.ws is None or self.ws.numel() < n:
            self.ws = torch.empty(n, dtype=torch.uint8, device=rhs.device)
        _check(lib.bout_b200_laplacexz_solve(C.byref(self.g.c), C.byref(self.c), _ptr(rhs), _ptr(x0), _ptr(out),
                                             C.c_void_p(self.ws.data_ptr()), C.c_size_t(self.ws.numel()), _stream()))
        return out


class InvertPar:
    """InvertPar "cyclic" (reference src/invert/parderiv/impls/cyclic/cyclic.cxx) on a field-aligned input.  sg is
    formed as the reference's constructor does, with the 2nd-order central DDY of a Field2D."""

    def __init__(self, g: Grid, metric: dict, zlength, A=None, B=None, Cc=None, D=None, E=None, closed=True,
                 nonuniform=True, twist_shift=None):
        self.g = g
        g22, dy = metric["g_22"], metric["dy"]
        r = 1.0 / torch.sqrt(g22)
        ddy = torch.zeros_like(r)
        if closed:   # guard rows hold the periodic images
            ddy[:, 1:-1] = 0.5 * (r[:, 2:] - r[:, :-2]) / dy[:, 1:-1]
        else:
            ddy[:, 1:-1] = 0.5 * (r[:, 2:] - r[:, :-2]) / dy[:, 1:-1]
        sg = (ddy / torch.sqrt(g22)).contiguous()
        self.keep = dict(A=A, B=B, C=Cc, D=D, E=E, sg=sg, g_22=g22, dy=dy,
                         d1_dy=metric["d1_dy"] if nonuniform else None, twist_shift=twist_shift)
        c = CInvertPar()
        for n, t in self.keep.items():
            setattr(c, n, None if t is None else t.data_ptr())
        c.zlength, c.closed = zlength, int(closed)
        self.c, self.ws = c, None

    def solve(self, f, out=None):
        out = torch.zeros_like(f) if out is None else out
        lib.bout_b200_invert_par_workspace.restype = C.c_size_t
        n = lib.bout_b200_invert_par_workspace(C.byref(self.g.c))
        if self.ws is None or self.ws.numel() < n:
            self.ws = torch.empty(n, dtype=torch.uint8, device=f.device)
        _check(lib.bout_b200_invert_par_solve(C.byref(self.g.c), C.byref(self.c), _ptr(f), _ptr(out),
                                              C.c_void_p(self.ws.data_ptr()), C.c_size_t(self.ws.numel()), _stream()))
        return out


class Laplace:
    """Host-side mirror of the Laplacian settings (reference invert_laplace.cxx:68-138)."""

    def __init__(self, g: Grid, metric: dict, zlength, all_terms=True, nonuniform=True, flags=(0, 0, 0),
                 maxmode=None, periodic_x=False, A=None, C1=None, C2=None, D=None, first_x=True, last_x=True,
                 wrap_yguards=False):
        self.g = g
        self.keep = dict(metric)
        self.keep.update(A=A, C1=C1, C2=C2, D=D)
        c = CLaplace()
        for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22", "A", "C1", "C2", "D"):
            t = self.keep.get(n)
            setattr(c, n, None if t is None else t.data_ptr())
        c.zlength = zlength
        c.all_terms, c.nonuniform = int(all_terms), int(nonuniform)
        c.global_flags, c.inner_boundary_flags, c.outer_boundary_flags = flags
        c.maxmode = g.nz // 2 if maxmode is None else maxmode
        c.periodic_x, c.first_x, c.last_x = int(periodic_x), int(first_x), int(last_x)
        c.wrap_yguards = int(wrap_yguards)   # single rank, periodic y: the solve also writes the y guard rows
        self.c = c
        self.ws = None

    def set_ypeer(self, ypeer):
        """bout_b200_laplace.ypeer for the next solves (a CYPeer, or None)."""
        self._ypeer = ypeer
        self.c.ypeer = None if ypeer is None else C.addressof(ypeer)

    def workspace(self, device):
        lib.bout_b200_laplace_cyclic_workspace.restype = C.c_size_t
        n = lib.bout_b200_laplace_cyclic_workspace(C.byref(self.g.c), C.byref(self.c))
        if self.ws is None or self.ws.numel() < n:
            self.ws = torch.empty(n, dtype=torch.uint8, device=device)
        return self.ws

    def solve(self, rhs, x0=None, out=None):
        out = torch.zeros_like(rhs) if out is None else out
        ws = self.workspace(rhs.device)
        _check(lib.bout_b200_laplace_cyclic_solve(C.byref(self.g.c), C.byref(self.c), _ptr(rhs),
                                                  _ptr(x0), _ptr(out), C.c_void_p(ws.data_ptr()),
                                                  C.c_size_t(ws.numel()), _stream()))
        return out

    def solve_perp(self, jy, rhs, x0=None, out=None):
        """LaplaceCyclic::solve(FieldPerp, FieldPerp): rhs, x0, out are [nx][nz] planes belonging to y index jy."""
        out = torch.zeros_like(rhs) if out is None else out
        lib.bout_b200_laplace_cyclic_workspace_perp.restype = C.c_size_t
        n = lib.bout_b200_laplace_cyclic_workspace_perp(C.byref(self.g.c), C.byref(self.c))
        if getattr(self, "ws_perp", None) is None or self.ws_perp.numel() < n:
            self.ws_perp = torch.empty(n, dtype=torch.uint8, device=rhs.device)
        _check(lib.bout_b200_laplace_cyclic_solve_perp(C.byref(self.g.c), C.byref(self.c), int(jy), _ptr(rhs), _ptr(x0),
                                                       _ptr(out), C.c_void_p(self.ws_perp.data_ptr()),
                                                       C.c_size_t(self.ws_perp.numel()), _stream()))
        return out

    def factor(self, device="cuda"):
        """Compute the coefficient table and pivot reciprocals once; later solves with unchanged
        coefficients reuse them.  Call again after changing the contents of any metric/coefficient array."""
        ws = self.workspace(device)
        _check(lib.bout_b200_laplace_cyclic_factor(C.byref(self.g.c), C.byref(self.c), C.c_void_p(ws.data_ptr()),
                                                   C.c_size_t(ws.numel()), _stream()))
        return self

    def invalidate(self):
        if self.ws is not None:
            _check(lib.bout_b200_laplace_cyclic_invalidate(C.c_void_p(self.ws.data_ptr())))

    def __del__(self):
        try:
            self.invalidate()   # the workspace address may be handed to another solver by the allocator
        except Exception:
            pass

    # -- X-decomposed solve: reduce -> (all-gather of the interface records) -> finish
    def iface_doubles(self):
        lib.bout_b200_laplace_cyclic_iface_doubles.restype = C.c_size_t
        return int(lib.bout_b200_laplace_cyclic_iface_doubles(C.byref(self.g.c), C.byref(self.c)))

    def reduce(self, rhs, x0=None, iface=None):
        ws = self.workspace(rhs.device)
        if iface is None:
            iface = torch.empty(self.iface_doubles(), dtype=torch.float64, device=rhs.device)
        _check(lib.bout_b200_laplace_cyclic_reduce(C.byref(self.g.c), C.byref(self.c), _ptr(rhs), _ptr(x0),
                                                   C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                                   _ptr(iface), _stream()))
        return iface

    def finish(self, iface_all, nxpe, pe_xind, out):
        ws = self.ws
        _check(lib.bout_b200_laplace_cyclic_finish(C.byref(self.g.c), C.byref(self.c), _ptr(iface_all), nxpe,
                                                   pe_xind, _ptr(out), C.c_void_p(ws.data_ptr()),
                                                   C.c_size_t(ws.numel()), _stream()))
        return out

    def status(self):
        _check(lib.bout_b200_laplace_cyclic_status(C.c_void_p(self.ws.data_ptr()), _stream()))

    def delp2(self, f, out=None):
        out = torch.zeros_like(f) if out is None else out
        _check(lib.bout_b200_delp2_fft(C.byref(self.g.c), C.byref(self.c), _ptr(f), _ptr(out), _stream()))
        return out


def delp2_fd(g: Grid, f, dx, dz, g11, g33, g13, G1, G3, d1_dx=None, out=None):
    out = torch.zeros_like(f) if out is None else out
    _check(lib.bout_b200_delp2_fd(C.byref(g.c), _ptr(f), _ptr(dx), _ptr(dz), _ptr(g11), _ptr(g33),
                                  _ptr(g13), _ptr(G1), _ptr(G3), _ptr(d1_dx), _ptr(out), _stream()))
    return out


class CEwTerm(C.Structure):
    _fields_ = [("coef", C.c_double), ("f1", C.c_void_p), ("f2", C.c_void_p), ("div", C.c_void_p),
                ("f1_is2d", C.c_int), ("f2_is2d", C.c_int), ("div_is2d", C.c_int)]


def elementwise(g: Grid, terms, out=None, region="RGN_ALL", like=None):
    """terms: list of (coef, f1, f2, div) with tensors or None; 2-D tensors broadcast over z."""
    arr = (CEwTerm * len(terms))()
    ref = like
    for i, t in enumerate(terms):
        coef, f1, f2, dv = (list(t) + [None, None, None])[:4]
        arr[i].coef = coef
        for name, f in (("f1", f1), ("f2", f2), ("div", dv)):
            setattr(arr[i], name, None if f is None else f.data_ptr())
            setattr(arr[i], name + "_is2d", int(f is not None and f.dim() == 2))
            if f is not None and f.dim() == 3 and ref is None:
                ref = f
    out = torch.zeros_like(ref) if out is None else out
    _check(lib.bout_b200_elementwise(C.byref(g.c), REGIONS[region], len(terms), arr, _ptr(out), _stream()))
    return out


UNARY = {"neg": 0, "sqrt": 1, "abs": 2, "exp": 3, "log": 4, "sin": 5, "cos": 6, "tanh": 7, "sq": 8,
         "pow": 9, "copy": 10}


def unary(g: Grid, op, f, arg=0.0, out=None):
    out = torch.empty_like(f) if out is None else out
    _check(lib.bout_b200_unary(C.byref(g.c), UNARY[op], _ptr(f), C.c_double(arg), _ptr(out), _stream()))
    return out


def where(g: Grid, test, gt0, le0, out=None):
    """where(test, gt0, le0): test > 0 ? gt0 : le0; test 3-D or 2-D, gt0 / le0 tensors or numbers."""
    ref = gt0 if torch.is_tensor(gt0) else (le0 if torch.is_tensor(le0) else None)
    out = torch.empty(g.shape3, dtype=torch.float64, device=test.device) if out is None else out
    a, av = (gt0, 0.0) if torch.is_tensor(gt0) else (None, float(gt0))
    b, bv = (le0, 0.0) if torch.is_tensor(le0) else (None, float(le0))
    _check(lib.bout_b200_where(C.byref(g.c), _ptr(test), int(test.dim() == 2), _ptr(a), C.c_double(av), _ptr(b),
                               C.c_double(bv), _ptr(out), _stream()))
    return out


def boundary(g: Grid, kind, side, f, value=0.0, spacing=None, lo=-1, hi=-1):
    """In place.  kind: dirichlet | neumann | zero | extend | free_o2 | free_o3; side: xin | xout | ydown | yup."""
    k = {"dirichlet": 0, "dirichlet_o2": 0, "neumann": 1, "zero": 2, "extend": 3, "free_o2": 4, "free_o3": 5}[kind]
    sd = {"xin": 0, "xout": 1, "ydown": 2, "yup": 3}[side]
    _check(lib.bout_b200_boundary(C.byref(g.c), k, sd, C.c_double(value), _ptr(spacing), lo, hi, _ptr(f),
                                  _stream()))
    return f


class CBndrySpec(C.Structure):
    _fields_ = [("f", C.c_void_p), ("kind", C.c_int), ("side", C.c_int), ("value", C.c_double), ("spacing2d", C.c_void_p),
                ("lo", C.c_int), ("hi", C.c_int)]


def boundary_batch(g: Grid, entries):
    """Several boundary conditions in one launch: entries = [(kind, side, f[, value[, spacing]]), ...]; entries must
    not depend on each other (bout_b200_boundary_batch)."""
    kinds = {"dirichlet": 0, "dirichlet_o2": 0, "neumann": 1, "zero": 2, "extend": 3, "free_o2": 4, "free_o3": 5}
    sides = {"xin": 0, "xout": 1, "ydown": 2, "yup": 3}
    arr = (CBndrySpec * len(entries))()
    for i, e in enumerate(entries):
        kind, side, f = e[0], e[1], e[2]
        value = e[3] if len(e) > 3 else 0.0
        spacing = e[4] if len(e) > 4 else None
        arr[i].f = f.data_ptr()
        arr[i].kind, arr[i].side, arr[i].value = kinds[kind], sides[side], value
        arr[i].spacing2d = None if spacing is None else spacing.data_ptr()
        arr[i].lo = arr[i].hi = -1
    _check(lib.bout_b200_boundary_batch(C.byref(g.c), len(entries), arr, _stream()))


def boundary_zerolaplace(g: Grid, side, f, dx, g11, g33, zlength, lo=-1, hi=-1):
    """In place.  BoundaryZeroLaplace on the xin / xout guard cells."""
    sd = {"xin": 0, "xout": 1, "both": 2}[side]
    _check(lib.bout_b200_boundary_zerolaplace(C.byref(g.c), sd, _ptr(dx), _ptr(g11), _ptr(g33),
                                              C.c_double(zlength), lo, hi, _ptr(f), _stream()))
    return f


def _ptr_array(fields):
    return (C.c_void_p * len(fields))(*[f.data_ptr() for f in fields])


def halo_self(g: Grid, fields, periodic_x=False, periodic_y=True):
    _check(lib.bout_b200_halo_self(C.byref(g.c), len(fields), _ptr_array(fields), int(periodic_x),
                                   int(periodic_y), _stream()))


def halo_pack_y(g: Grid, fields, y0, nrows, xlo, xhi, buf):
    _check(lib.bout_b200_halo_pack_y(C.byref(g.c), len(fields), _ptr_array(fields), y0, nrows, xlo, xhi,
                                     _ptr(buf), _stream()))


def halo_unpack_y(g: Grid, fields, y0, nrows, xlo, xhi, buf):
    _check(lib.bout_b200_halo_unpack_y(C.byref(g.c), len(fields), _ptr_array(fields), y0, nrows, xlo, xhi,
                                       _ptr(buf), _stream()))


# ---- Y exchange over NVLink peer memory (raw device addresses as python ints) -------------------------------
class _RawCuda:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


def tensor_from_ptr(ptr, shape, device):
    """A torch view (fp64, contiguous) of device memory this library allocated (bout_b200_peer_alloc)."""
    return torch.as_tensor(_RawCuda(ptr, shape), device=device)


def peer_alloc(nbytes):
    p = C.c_void_p()
    _check(lib.bout_b200_peer_alloc(C.c_size_t(nbytes), C.byref(p)))
    return p.value


def peer_export(ptr):
    h = (C.c_ubyte * 64)()
    _check(lib.bout_b200_peer_export(C.c_void_p(ptr), h))
    return bytes(h)


def peer_open(handle: bytes):
    p = C.c_void_p()
    h = (C.c_ubyte * 64).from_buffer_copy(handle)
    _check(lib.bout_b200_peer_open(h, C.byref(p)))
    return p.value


def halo_push_y(g: Grid, fields, y0, nrows, xlo, xhi, peer_buf, peer_flag, value, counter):
    _check(lib.bout_b200_halo_push_y(C.byref(g.c), len(fields), _ptr_array(fields), y0, nrows, xlo, xhi,
                                     C.c_void_p(peer_buf), C.c_void_p(peer_flag), C.c_ulonglong(value),
                                     C.c_void_p(counter), _stream()))


def _xchg_args(y0, bufs, flags):
    return ((C.c_int * 2)(*y0), (C.c_void_p * 2)(*[b or None for b in bufs]),
            (C.c_void_p * 2)(*[f or None for f in flags]))


def halo_push_y2(g: Grid, fields, nrows, y0, peer_bufs, peer_flags, value, counters):
    a, b, f = _xchg_args(y0, peer_bufs, peer_flags)
    _check(lib.bout_b200_halo_push_y2(C.byref(g.c), len(fields), _ptr_array(fields), nrows, a, b, f,
                                      C.c_ulonglong(value), C.c_void_p(counters), _stream()))


def halo_wait_unpack_y2(g: Grid, fields, nrows, y0, local_bufs, local_flags, value, status):
    a, b, f = _xchg_args(y0, local_bufs, local_flags)
    _check(lib.bout_b200_halo_wait_unpack_y2(C.byref(g.c), len(fields), _ptr_array(fields), nrows, a, b, f,
                                             C.c_ulonglong(value), C.c_void_p(status), _stream()))


def halo_wait(flag, value, status):
    _check(lib.bout_b200_halo_wait(C.c_void_p(flag), C.c_ulonglong(value), C.c_void_p(status), _stream()))


def halo_unpack_y_raw(g: Grid, fields, y0, nrows, xlo, xhi, buf):
    _check(lib.bout_b200_halo_unpack_y(C.byref(g.c), len(fields), _ptr_array(fields), y0, nrows, xlo, xhi,
                                       C.c_void_p(buf), _stream()))


def rk_stage(g: Grid, mode, dt, start, k1, k2=None, k3=None, k4=None, c=1.0, out=None,
             region="RGN_NOBNDRY"):
    m = {"axpy": 0, "rk4_final": 1, "rk3_stage2": 2, "rk3_stage3": 3}[mode]
    out = torch.zeros_like(start) if out is None else out
    _check(lib.bout_b200_rk_stage(C.byref(g.c), REGIONS[region], m, C.c_double(dt), C.c_double(c),
                                  _ptr(start), _ptr(k1), _ptr(k2), _ptr(k3), _ptr(k4), _ptr(out), _stream()))
    return out


def lincomb(g: Grid, coefs, fields, out, region="RGN_NOBNDRY"):
    """out[region] = sum_j coefs[j] * fields[j] (1..8 terms; out must not alias a field): the state / output
    loops of euler, rkgeneric and splitrk (bout_b200_lincomb)."""
    n = len(coefs)
    assert n == len(fields)
    cs = (C.c_double * n)(*[float(c) for c in coefs])
    for f in fields:
        _ptr(f)
    _check(lib.bout_b200_lincomb(C.byref(g.c), REGIONS[region], n, cs, _ptr_array(fields), _ptr(out), _stream()))
    return out


lib.bout_b200_err_sum_scratch_doubles.restype = C.c_size_t


def err_sum(g: Grid, a, b, atol, accumulate_onto=None, region="RGN_NOBNDRY"):
    """sum over region of |a-b| / (|a|+|b|+atol) (bout_b200_err_sum; rkscheme.cxx:150-156), returned as a float.
    accumulate_onto: a previous sum to add to on the device (the per-field accumulation of a state)."""
    n = int(lib.bout_b200_err_sum_scratch_doubles())
    scratch = torch.empty(n + 1, dtype=torch.float64, device=a.device)
    result = scratch[n:]
    if accumulate_onto is not None:
        result.fill_(float(accumulate_onto))
    _check(lib.bout_b200_err_sum(C.byref(g.c), REGIONS[region], _ptr(a), _ptr(b), C.c_double(atol), _ptr(scratch),
                                 0 if accumulate_onto is None else 1, C.c_void_p(result.data_ptr()), _stream()))
    return float(result.item())


def err_max(g: Grid, a, b, accumulate_onto=None, region="RGN_NOBNDRY"):
    """max over region of |a-b| (bout_b200_err_max; adams_bashforth.cxx:229-253), returned as a float."""
    n = int(lib.bout_b200_err_sum_scratch_doubles())
    scratch = torch.empty(n + 1, dtype=torch.float64, device=a.device)
    result = scratch[n:]
    if accumulate_onto is not None:
        result.fill_(float(accumulate_onto))
    _check(lib.bout_b200_err_max(C.byref(g.c), REGIONS[region], _ptr(a), _ptr(b), _ptr(scratch),
                                 0 if accumulate_onto is None else 1, C.c_void_p(result.data_ptr()), _stream()))
    return float(result.item())


class CMetric(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dx", "dy", "dz", "J", "g_22", "g11", "g33", "g13", "G1", "G3",
                                          "d1_dx")]


class CRkEpilogue(C.Structure):
    _fields_ = [("start", C.c_void_p * 2), ("acc_in", C.c_void_p * 2), ("state_out", C.c_void_p * 2),
                ("acc_out", C.c_void_p * 2), ("a", C.c_double), ("b", C.c_double)]


def make_metric(md: dict):
    """md: name -> Field2D tensor (missing names become NULL).  Keeps the tensors alive."""
    c = CMetric()
    for n, _ in CMetric._fields_:
        t = md.get(n)
        setattr(c, n, None if t is None else t.data_ptr())
    c._keep = md
    return c


def make_epilogue(start, acc_in, state_out, acc_out, a, b):
    e = CRkEpilogue()
    for name, pair in (("start", start), ("acc_in", acc_in), ("state_out", state_out), ("acc_out", acc_out)):
        arr = getattr(e, name)
        for i in range(2):
            arr[i] = None if pair is None or pair[i] is None else pair[i].data_ptr()
    e.a, e.b = a, b
    e._keep = (start, acc_in, state_out, acc_out)
    return e


def fused_coef_workspace(g: Grid, device):
    lib.bout_b200_fused_coef_doubles.restype = C.c_size_t
    return torch.empty(lib.bout_b200_fused_coef_doubles(C.byref(g.c)), dtype=torch.float64, device=device)


def fused_coefs(g: Grid, metric: CMetric, coef_ws):
    _check(lib.bout_b200_fused_coefs(C.byref(g.c), C.byref(metric), _ptr(coef_ws), _stream()))
    return coef_ws


class CYPeer(C.Structure):
    """bout_b200_ypeer: Y guard rows delivered by the producing kernel into the neighbours' mapped arenas."""
    _fields_ = [("delta", C.c_longlong * 2), ("peer_flag", C.c_void_p * 2), ("signal", C.c_ulonglong),
                ("counters", C.c_void_p), ("wait_flag", (C.c_void_p * 2) * 2), ("wait_value", (C.c_ulonglong * 2) * 2),
                ("status", C.c_void_p)]


def make_ypeer(delta, peer_flag, signal, counters, status, waits=None):
    """delta / peer_flag: (lower, upper) neighbour; waits[d] = up to two (flag address, value) the CTAs next to side d
    wait for.  None entries = nothing on that side."""
    y = CYPeer()
    for d in range(2):
        y.delta[d] = int(delta[d] or 0)
        y.peer_flag[d] = peer_flag[d]
        for k, w in enumerate((waits[d] if waits is not None else ())[:2]):
            if w is not None:
                y.wait_flag[d][k], y.wait_value[d][k] = w
    y.signal, y.counters, y.status = int(signal), counters, status
    return y


class CGuardFill(C.Structure):
    _fields_ = [("value", C.c_double), ("x_inner", C.c_int), ("x_outer", C.c_int), ("wrap_y", C.c_int),
                ("ystart", C.c_int), ("yend", C.c_int), ("ypeer", C.c_void_p)]


def make_guard_fill(g: Grid, value=0.0, x_inner=True, x_outer=True, wrap_y=True, ypeer=None):
    """Guard cells written by the kernel that produces the next stage state (bout_b200_guard_fill): constant-value
    Dirichlet on the x sides, periodic y copies on a single rank or delivery to the Y neighbours (ypeer: a CYPeer);
    g is the MESH's grid (its ystart / yend)."""
    gf = CGuardFill(float(value), int(x_inner), int(x_outer), int(wrap_y), g.c.ystart, g.c.yend,
                    None if ypeer is None else C.addressof(ypeer))
    gf._keep = ypeer
    return gf


def hw3d_rhs(g: Grid, metric, alpha, kappa, Dn, Dvort, n, vort, phi, ddt_n=None, ddt_vort=None,
             epilogue=None, coef_ws=None, guards=None):
    coef_ws = fused_coef_workspace(g, n.device) if coef_ws is None else coef_ws
    _check(lib.bout_b200_hw3d_rhs_guards(C.byref(g.c), None if metric is None else C.byref(metric),
                                         C.c_double(alpha), C.c_double(kappa),
                                         C.c_double(Dn), C.c_double(Dvort), _ptr(n), _ptr(vort), _ptr(phi),
                                         _ptr(ddt_n), _ptr(ddt_vort),
                                         None if epilogue is None else C.byref(epilogue),
                                         None if guards is None else C.byref(guards), _ptr(coef_ws), _stream()))


def hw_convective(g: Grid, dx, dz, alpha, kappa, n, vort, phi, zonal_n=None, zonal_phi=None, ddt_n=None,
                  ddt_vort=None, epilogue=None):
    _check(lib.bout_b200_hw_convective(C.byref(g.c), _ptr(dx), _ptr(dz), C.c_double(alpha),
                                       C.c_double(kappa), _ptr(n), _ptr(vort), _ptr(phi), _ptr(zonal_n),
                                       _ptr(zonal_phi), _ptr(ddt_n), _ptr(ddt_vort),
                                       None if epilogue is None else C.byref(epilogue), _stream()))


# ---- shifted-metric / elm-pb operator set -----------------------------------------------------------
def stencil_yslices(g: Grid, op: str, f, yup, ydown, v=None, scale1=None, op2=None, scale2=None, out=None,
                    region="RGN_NOBNDRY"):
    """Y stencil whose neighbours come from parallel slices: yup / ydown are lists of 1 or 2 fields."""
    out = torch.zeros_like(f) if out is None else out
    _check(lib.bout_b200_stencil_yslices(C.byref(g.c), OPS[op], REGIONS[region], _ptr(v), _ptr(f), len(yup),
                                         _ptr_array(yup), _ptr_array(ydown), _ptr(scale1),
                                         -1 if op2 is None else OPS[op2], _ptr(scale2), _ptr(out), _stream()))
    return out


def parallel_slices(g: Grid, f, zshift, zlength, nslices=None, yup=None, ydown=None):
    """ShiftedMetric::calcParallelSlices: returns (yup, ydown), lists of nslices fields (offsets +-1..)."""
    n = g.myg if nslices is None else nslices
    yup = [torch.zeros_like(f) for _ in range(n)] if yup is None else yup
    ydown = [torch.zeros_like(f) for _ in range(n)] if ydown is None else ydown
    _check(lib.bout_b200_parallel_slices(C.byref(g.c), _ptr(f), _ptr(zshift), C.c_double(zlength), n,
                                         _ptr_array(yup), _ptr_array(ydown), _stream()))
    return yup, ydown


class CB0xGrad(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dx", "dy", "dz", "J", "g_22", "g_12", "g_23", "int_shift_torsion")] + \
               [(n, C.c_int) for n in ("ddx_op", "ddy_op", "ddz_op", "vddx_op", "vddy_op", "vddz_op")]


def make_b0xgrad(md: dict, first=("C4", "C2", "C4"), upwind=("W3", "C2", "W3"), int_shift_torsion=None):
    """md: name -> Field2D tensor; first / upwind: (x, y, z) method names (mesh:dd?:first / :upwind)."""
    c = CB0xGrad()
    for n in ("dx", "dy", "dz", "J", "g_22", "g_12", "g_23"):
        setattr(c, n, md[n].data_ptr())
    c.int_shift_torsion = None if int_shift_torsion is None else int_shift_torsion.data_ptr()
    c.ddx_op, c.ddy_op, c.ddz_op = (OPS["DDX_" + m] for m in first)
    c.vddx_op, c.vddy_op, c.vddz_op = (OPS["VDDX_" + m] for m in upwind)
    c._keep = (md, int_shift_torsion)
    return c


def b0xgrad_dot_grad(g: Grid, cfg, phi, A, phi_slices=None, a_slices=None, out=None, shifted=None):
    """phi / A: Field3D or Field2D tensors; *_slices = (yup list, ydown list) or None (field itself).
    shifted = (zShift, zlength) on a shifted-metric mesh: with a Field2D A the y-advection term is taken
    through toFieldAligned / fromFieldAligned as the reference does."""
    ref = phi if phi.dim() == 3 else A
    out = torch.zeros_like(ref) if out is None else out
    yterm = torch.zeros_like(ref) if (shifted is not None and A.dim() == 2) else None
    pu, pd = phi_slices if phi_slices is not None else ([], [])
    au, ad = a_slices if a_slices is not None else ([], [])
    _check(lib.bout_b200_b0xgrad_dot_grad(
        C.byref(g.c), C.byref(cfg), _ptr(phi), int(phi.dim() == 2), len(pu),
        _ptr_array(pu) if pu else None, _ptr_array(pd) if pd else None, _ptr(A), int(A.dim() == 2), len(au),
        _ptr_array(au) if au else None, _ptr_array(ad) if ad else None, _ptr(out), _ptr(yterm), _stream()))
    if yterm is not None:
        zs, zl = shifted
        aligned = zshift(g, yterm, (-zs).contiguous(), zl, region="RGN_NOBNDRY")
        back = zshift(g, aligned, zs, zl, region="RGN_NOBNDRY")
        elementwise(g, [(1.0, out), (1.0, back)], out=out, region="RGN_NOBNDRY")
    return out


class CElmpbRhs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "dx", "dy", "dz", "J", "g_22", "g_12", "g_23", "Bxy", "B0", "J0", "P0", "phi0", "b0xcv_x", "b0xcv_y",
        "b0xcv_z", "U", "P", "Psi", "phi", "Jpar", "d2Jpar", "eta", "U_up", "U_dn", "P_up", "P_dn", "Psi_up",
        "Psi_dn", "phi_up", "phi_dn", "J_up", "J_dn", "U_mode", "P_mode", "J_mode", "zShift",
        "g11", "g33", "g13", "G1", "G3")] + \
        [("delp2_all_terms", C.c_int), ("delp2_nonuniform", C.c_int), ("rowc_ws", C.c_void_p), ("zlength", C.c_double),
         ("hyperresist", C.c_double), ("keep_U", C.c_double), ("keep_P", C.c_double), ("keep_Psi", C.c_double),
         ("eta_is2d", C.c_int), ("filter_n0", C.c_int)] + \
        [(n, C.c_void_p) for n in ("ddt_U", "ddt_P", "ddt_Psi", "rk_start_U", "rk_start_P", "rk_start_Psi",
                                   "rk_stage_U", "rk_stage_P", "rk_stage_Psi")] + [("rk_cdt", C.c_double)]


def row_mode(g: Grid, fields, n0, out=None):
    """One Fourier coefficient per (x,y) row: list of complex [nx][ny] tensors (stored as float64 [nx][ny][2])."""
    if out is None:
        out = [torch.zeros(g.shape2 + (2,), dtype=torch.float64, device=fields[0].device) for _ in fields]
    _check(lib.bout_b200_row_mode(C.byref(g.c), len(fields), _ptr_array(fields), n0, _ptr_array(out), _stream()))
    return out


def elmpb_rhs(g: Grid, fields: dict, hyperresist, filter_n0, keep=(1.0, 1.0, 1.0), rk=None, zlength=0.0):
    """fields: name -> tensor for every pointer member of bout_b200_elmpb_rhs_args (eta may be 2-D).
    rk = (start fields (U, P, Psi), stage-state fields (U, P, Psi), c * dt) or None."""
    c = CElmpbRhs()
    for n, _t in CElmpbRhs._fields_:
        if n in fields:
            setattr(c, n, fields[n].data_ptr())
    if rk is not None:
        start, stage, cdt = rk
        for nm, a, b in zip(("U", "P", "Psi"), start, stage):
            setattr(c, "rk_start_" + nm, a.data_ptr())
            setattr(c, "rk_stage_" + nm, b.data_ptr())
        c.rk_cdt = cdt
    c.hyperresist = hyperresist
    c.zlength = zlength
    c.delp2_all_terms, c.delp2_nonuniform = 1, 1     # the default Laplacian instance Delp2 uses
    c.keep_U, c.keep_P, c.keep_Psi = keep
    c.eta_is2d = int(fields["eta"].dim() == 2)
    c.filter_n0 = filter_n0
    _check(lib.bout_b200_elmpb_rhs(C.byref(g.c), C.byref(c), _stream()))


def average_y(g: Grid, f, out=None, y_group=None, nype=1):
    """averageY(Field3D): local mean -> (SUM all-reduce over the Y communicator when NYPE > 1) -> broadcast."""
    xz = torch.empty((g.nx, g.nz), dtype=torch.float64, device=f.device)
    _check(lib.bout_b200_ymean(C.byref(g.c), _ptr(f), _ptr(xz), _stream()))
    scale = 1.0
    if nype > 1:
        import torch.distributed as dist
        dist.all_reduce(xz, group=y_group)
        scale = 1.0 / nype
    out = torch.empty_like(f) if out is None else out
    _check(lib.bout_b200_yspread(C.byref(g.c), _ptr(xz), C.c_double(scale), _ptr(out), _stream()))
    return out


def smooth_x(g: Grid, f, out=None):
    out = torch.empty_like(f) if out is None else out
    _check(lib.bout_b200_smooth_x(C.byref(g.c), _ptr(f), _ptr(out), _stream()))
    return out
