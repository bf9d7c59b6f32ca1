"""CPU: index algebra of the block-partitioned LaplaceCyclic path and of the register-resident z-FFT.

tests/emul/lapb_emul.cu steps simulated threads through the SAME per-thread phase functions the kernels run
(bout-dev_b200/csrc/laplace_block.cuh, fft_reg.cuh), with barriers where the kernels have them, on host memory.
It is test infrastructure (built into a temp directory here, never shipped): it pins the transform's index maps,
the separator / reduced-system algebra, every tridagMatrix flag row and the on-device change detection without a
GPU.  The kernels themselves are checked on the device by tests/test_gpu_laplace.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import make_metric, rel_err, seeded_fields

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = "/usr/local/cuda/bin/nvcc"


class CGrid(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "ny", "nz", "xstart", "xend", "ystart", "yend")]


class CLaplace(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22",
                                          "A", "C1", "C2", "D")] + \
               [("zlength", C.c_double)] + \
               [(n, C.c_int) for n in ("all_terms", "nonuniform", "global_flags",
                                       "inner_boundary_flags", "outer_boundary_flags", "maxmode",
                                       "periodic_x", "first_x", "last_x", "include_yguards_lower",
                                       "include_yguards_upper", "wrap_yguards")] + [("ypeer", C.c_void_p)]


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    so = str(tmp_path_factory.mktemp("emul") / "liblapb_emul.so")
    subprocess.run([NVCC, "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets",
                    "-diag-suppress", "128", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", so,
                    os.path.join(HERE, "emul", "lapb_emul.cu")], check=True)
    lib = C.CDLL(so)
    lib.lapb_emul_solve.restype = C.c_int
    lib.regfft_emul_roundtrip.restype = C.c_int
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def emul_solve(lib, grid, m, f, x0=None, flags=(0, 0, 0), maxmode=None, all_terms=True, nonuniform=True,
               fresh=True, force=False, keep=None, nt512=256, **coef):
    nx, ny, nz, mxg, myg = grid
    g = CGrid(nx, ny, nz, mxg, nx - mxg - 1, myg, ny - myg - 1)
    L = CLaplace()
    arrays = keep if keep is not None else {}
    for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22"):
        arrays.setdefault(n, np.ascontiguousarray(getattr(m, n)))
    for n in ("A", "C1", "C2", "D"):
        if coef.get(n) is not None:
            arrays.setdefault(n, np.ascontiguousarray(coef[n]))
    for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22", "A", "C1", "C2", "D"):
        setattr(L, n, ptr(arrays.get(n)))
    L.zlength = float(m.dz[0, 0]) * nz
    L.all_terms, L.nonuniform = int(all_terms), int(nonuniform)
    L.global_flags, L.inner_boundary_flags, L.outer_boundary_flags = flags
    L.maxmode = nz // 2 if maxmode is None else maxmode
    L.periodic_x, L.first_x, L.last_x = 0, 1, 1
    out = np.zeros_like(f)
    rc = lib.lapb_emul_solve(C.byref(g), C.byref(L), ptr(f), ptr(x0), ptr(out), int(fresh), int(force), int(nt512))
    assert rc >= 0
    return out, rc, arrays


@pytest.mark.parametrize("nz", [128, 256, 512])
def test_register_fft_round_trip(emul, nz):
    """bout::fft::rfft / irfft conventions (fft_fftw.cxx:116-240): forward scaled by 1/nz, backward unscaled."""
    rng = np.random.default_rng(nz)
    row = rng.uniform(-1, 1, nz)
    spec = np.zeros(nz + 2)
    back = np.zeros(nz)
    assert emul.regfft_emul_roundtrip(nz, ptr(row), ptr(spec), ptr(back)) == 0
    want = np.fft.rfft(row) / nz
    got = spec[0::2] + 1j * spec[1::2]
    assert np.abs(got - want).max() < 5e-16
    assert np.abs(back - row).max() < 2e-15


# (global, inner, outer): the flag rows of tests/test_gpu_laplace.py
FLAGSETS = [(0, 0, 0), (0, 3, 0), (0, 0, 3), (0, 2, 2), (0, 0, 1), (1, 4, 64), (0, 16, 16), (0, 32 | 3, 32),
            (4, 0, 0), (0, 128, 0), (0, 3 | 16, 32), (0, 1024, 0)]
# row counts: one block only, a trailing separator-only block (n = 16 k + 1), short last blocks, several CTAs
GRIDS = [(12, 3, 128, 2, 1), (17, 4, 128, 2, 1), (40, 4, 128, 2, 1), (33, 3, 256, 2, 1), (70, 4, 256, 2, 1),
         (20, 3, 512, 2, 1), (49, 4, 512, 2, 1), (36, 4, 512, 1, 1)]


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("flags", FLAGSETS)
@pytest.mark.parametrize("nt512", [256, 128])
def test_block_path_matches_oracle(emul, orc, port, grid, flags, nt512):
    if nt512 == 128 and grid[2] != 512:
        pytest.skip("the 128-thread rows kernels exist for nz = 512 only")
    og = orc.Grid(*grid)
    m = make_metric(orc, (grid[0], grid[1]), grid[2])
    f, x0 = seeded_fields(grid[:3], 2)
    want = port.laplace_cyclic(og, m, f, x0=x0, flags=flags)
    got, rc, _ = emul_solve(emul, grid, m, f, x0, flags=flags, nt512=nt512)
    assert rc & 1 == 0 and rc & 2                 # no zero pivot; factored
    assert bool(rc & 4) == (og.yend > og.ystart)  # this metric varies along y: per-y tables when several y are solved
    region = (slice(None), slice(og.ystart, og.yend + 1), slice(None))
    assert rel_err(got, want, region) < 1e-10


@pytest.mark.parametrize("grid", [(40, 4, 128, 2, 1), (70, 3, 256, 2, 1), (49, 3, 512, 2, 1)])
def test_block_path_coefficients_filter_and_shared_tables(emul, orc, port, grid):
    """A / C / D coefficients, truncated spectra (laplace:filter), a y-independent metric (one table for all y)."""
    og = orc.Grid(*grid)
    nx, ny, nz = grid[:3]
    f, x0 = seeded_fields(grid[:3], 2)
    X = (np.arange(nx)[:, None] + 0.5) / nx * np.ones((nx, ny))
    A = np.ascontiguousarray(-0.1 - 0.05 * X)
    D = np.ascontiguousarray(1 + 0.2 * X)
    Cc = np.ascontiguousarray(1 + 0.3 * np.sin(3 * X))
    region = (slice(None), slice(og.ystart, og.yend + 1), slice(None))
    m = make_metric(orc, (nx, ny), nz)
    for kw in (dict(A=A), dict(A=A, D=D), dict(C1=Cc, C2=Cc), dict(A=A, C1=Cc, C2=Cc, D=D)):
        for maxmode in (None, nz // 4):
            want = port.laplace_cyclic(og, m, f, x0=x0, maxmode=maxmode, **kw)
            for nt512 in ((256, 128) if nz == 512 else (256,)):
                got, rc, _ = emul_solve(emul, grid, m, f, x0, maxmode=maxmode, nt512=nt512, **kw)
                assert rc & 1 == 0
                assert rel_err(got, want, region) < 1e-10
    mt = make_metric(orc, (nx, ny), nz, trivial=True)
    want = port.laplace_cyclic(og, mt, f, x0=x0, flags=(0, 3, 0), A=A, nonuniform=False)
    got, rc, _ = emul_solve(emul, grid, mt, f, x0, flags=(0, 3, 0), A=A, nonuniform=False)
    assert rc & 4 == 0 and rc & 1 == 0            # tables shared by every y
    assert rel_err(got, want, region) < 1e-10


def test_block_path_refactors_only_when_an_input_changed(emul, orc, port):
    grid = (40, 4, 128, 2, 1)
    og = orc.Grid(*grid)
    m = make_metric(orc, grid[:2], grid[2])
    f, g2 = seeded_fields(grid[:3], 2)
    A = np.full(grid[:2], -0.1)
    region = (slice(None), slice(og.ystart, og.yend + 1), slice(None))
    got, rc, keep = emul_solve(emul, grid, m, f, A=A)
    assert rc & 2
    assert rel_err(got, port.laplace_cyclic(og, m, f, A=A), region) < 1e-10
    got, rc, keep = emul_solve(emul, grid, m, g2, fresh=False, keep=keep)      # new right-hand side only
    assert rc & 2 == 0
    assert rel_err(got, port.laplace_cyclic(og, m, g2, A=A), region) < 1e-10
    keep["A"][5, 2] = -0.7                                                      # same array, one new value
    got, rc, keep = emul_solve(emul, grid, m, g2, fresh=False, keep=keep)
    assert rc & 2
    A2 = A.copy()
    A2[5, 2] = -0.7
    assert rel_err(got, port.laplace_cyclic(og, m, g2, A=A2), region) < 1e-10


def test_block_path_reports_zero_pivot(emul, orc):
    grid = (20, 3, 128, 2, 1)
    m = make_metric(orc, grid[:2], grid[2], trivial=True)
    m.g11[:] = 0.0
    m.g33[:] = 0.0
    f = np.ones(grid[:3])
    _, rc, _ = emul_solve(emul, grid, m, f, flags=(0, 1, 1), all_terms=False, nonuniform=False)
    assert rc & 1
