"""CPU: pins the plain-C oracle (oracle/bout_oracle.c) to the reference.

 * the reference's own known-answer tests for this path (tests/unit/include/test_cyclic_reduction.cxx:41-110,
   tests/unit/invert/test_fft.cxx:20-120, tests/integrated/test-cyclic/test_cyclic.cxx:45-131);
 * the reference's golden file tests/integrated/test-laplace/data/benchmark.0.nc (1e-6, float32) and
 * outputs of the reference's own sources compiled in place (tests/golden/*.npz, made by
   tests/golden/make_golden.py) at the north_star tolerances: 1e-12 stencils / bracket / Delp2,
   1e-10 LaplaceCyclic.
"""
import os

import numpy as np
import pytest

from util import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ops():
    return dict(np.load(os.path.join(GOLD, "ops_small.npz")))


@pytest.fixture(scope="module")
def lapb():
    return dict(np.load(os.path.join(GOLD, "laplace_benchmark.npz")))


def metric_from(orc, d, prefix="metric_"):
    m = orc.Metric(**{n: np.ascontiguousarray(d[prefix + n]) for n in orc.Metric.NAMES})
    m.non_uniform = True
    return m


# ---- known-answer vectors of the reference's unit tests ---------------------------------------------
def test_cyclic_reduction_known_answers(port):
    x = port.cyclic_solve([0, 1, 1, 1, 1], [5, 4, 3, 2, 1], [2, 2, 2, 2, 0], [0, 1, 2, 2, 3])
    assert np.allclose(x, [-1, 2.5, -4, 5.75, -2.75], atol=1e-14, rtol=0)
    x = port.cyclic_solve([0, -2, -2, -2, -2], [1, 1, 1, 1, 1], [2, 2, 2, 2, 0], [5, 4, 5, 4, 5])
    assert np.allclose(x, [3.4, 0.8, 5, 0.8, 6.6], atol=1e-14, rtol=0)


@pytest.mark.parametrize("n", [8, 9])
def test_fft_known_spectrum(port, n):
    i = np.arange(n)
    sig = np.sin(i * 2 * np.pi / n) + np.cos(2 * i * 2 * np.pi / n)
    want = np.zeros(n // 2 + 1, complex)
    want[1], want[2] = -0.5j, 0.5
    assert np.abs(port.rfft(sig) - want).max() < 1e-12       # forward normalised by 1/n (fft_fftw.cxx:170-176)
    assert np.abs(port.irfft(want, n) - sig).max() < 1e-12
    assert np.abs(port.irfft(port.rfft(sig), n) - sig).max() < 1e-12


@pytest.mark.parametrize("n", [4, 5, 16, 33])
@pytest.mark.parametrize("periodic", [False, True])
def test_cyclic_known_integer_solution(port, n, periodic):
    """tests/integrated/test-cyclic/test_cyclic.cxx:45-131: x_i = (i mod 4) - 2, random diagonally
    dominant complex coefficients, tolerance 1e-10."""
    rng = np.random.default_rng(n)
    for _ in range(5):
        x = ((np.arange(n) % 4) - 2).astype(complex)
        a = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
        c = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
        b = 4.0 + rng.uniform(0, 1, n) + 1j * rng.uniform(-1, 1, n)
        if periodic:
            r = a * np.roll(x, 1) + b * x + c * np.roll(x, -1)
        else:
            r = b * x
            r[1:] += a[1:] * x[:-1]
            r[:-1] += c[:-1] * x[1:]
        assert np.abs(port.cyclic_solve(a, b, c, r, periodic) - x).max() < 1e-10


# ---- the reference's golden LaplaceCyclic file --------------------------------------------------------
LAP_CASES = {"": ("zeros", "ones", "ones"), "a": ("a", "ones", "ones"), "ac": ("a", "c", "ones"),
             "ad": ("a", "ones", "d")}


@pytest.mark.parametrize("suffix", list(LAP_CASES))
@pytest.mark.parametrize("kind", ["flag0", "flag3", "flagis", "flagos"])
def test_laplace_benchmark_file(port, orc, lapb, suffix, kind):
    g = orc.Grid(*lapb["grid"])
    m = metric_from(orc, lapb)
    m.non_uniform = False
    coef = {"zeros": np.zeros(g.shape2), "ones": np.ones(g.shape2), "a": lapb["a"], "c": lapb["c"], "d": lapb["d"]}
    A, Cc, D = (np.ascontiguousarray(coef[k]) for k in LAP_CASES[suffix])
    flags = {"flag0": (0, 0, 0), "flag3": (0, 3, 0), "flagis": (0, 16, 0), "flagos": (0, 0, 16)}[kind]
    x0 = np.ascontiguousarray(np.repeat(lapb["set_to"][:, :, None], g.nz, axis=2))
    maxmode = int(round((1.0 - 0.2) * g.nz / 2.0))   # laplace:filter = 0.2 (invert_laplace.cxx:89-98)
    got = port.laplace_cyclic(g, m, lapb["input"], x0=x0, flags=flags, maxmode=maxmode, A=A, C1=Cc, C2=Cc, D=D,
                              all_terms=False, nonuniform=False)
    name = kind + suffix
    ys = slice(g.ystart, g.yend + 1)
    # the reference's own tolerance against its float32 golden file (test-laplace/runtest:34)
    assert np.allclose(got[:, ys], lapb["bench32_" + name], atol=1e-6, rtol=1e-6)
    # and the north_star tolerance against the float64 output of the reference build
    assert rel_err(got, lapb["ref64_" + name], (slice(None), ys, slice(None))) < 1e-10


# ---- reference outputs for every operator on one small grid ---------------------------------------------
def _grid(orc, ops):
    return orc.Grid(*ops["grid"])


INDEX_CASES = [(b, d, m) for b, ms in (("DDX", ["C2", "C4", "W2", "S2", "W3"]), ("D2DX2", ["C2", "C4"]),
                                       ("D4DX4", ["C2"]), ("VDDX", ["C2", "C4", "U1", "U2", "U3", "W3"]),
                                       ("FDDX", ["U1", "U2", "C2", "C4"])) for d in "XYZ" for m in ms]


@pytest.mark.parametrize("base,d,meth", INDEX_CASES)
def test_index_derivatives_match_reference(port, orc, ops, base, d, meth):
    g = _grid(orc, ops)
    v = ops["v"] if base in ("VDDX", "FDDX") else None
    got = port.index_deriv(g, f"{base}_{meth}", d, ops["f"], v=v)
    want = ops[f"index_{base.replace('X', d)}_{meth}"]
    assert rel_err(got, want, g.interior) < 1e-12


def test_fft_derivatives_match_reference(port, orc, ops):
    g = _grid(orc, ops)
    assert rel_err(port.index_ddz_fft(g, 1, ops["f"]), ops["index_DDZ_FFT"], g.interior) < 1e-12
    assert rel_err(port.index_ddz_fft(g, 2, ops["f"]), ops["index_D2DZ2_FFT"], g.interior) < 1e-12


def test_coordinate_derivatives_match_reference(port, orc, ops):
    g, m, f = _grid(orc, ops), metric_from(orc, ops), ops["f"]
    for meth in ("C2", "C4"):
        assert rel_err(port.ddx(g, m, f, meth), ops["DDX_" + meth], g.interior) < 1e-12
        assert rel_err(port.ddy(g, m, f, meth), ops["DDY_" + meth], g.interior) < 1e-12
        assert rel_err(port.ddz(g, m, f, meth), ops["DDZ_" + meth], g.interior) < 1e-12
        assert rel_err(port.d2dx2(g, m, f, meth, first="C2"), ops["D2DX2_" + meth], g.interior) < 1e-12
        assert rel_err(port.d2dz2(g, m, f, meth), ops["D2DZ2_" + meth], g.interior) < 1e-12


def test_brackets_match_reference(port, orc, ops):
    g, m = _grid(orc, ops), metric_from(orc, ops)
    got = port.bracket_arakawa(g, ops["f"], ops["h"], m.dx, m.dz)
    assert rel_err(got, ops["bracket_arakawa"], g.interior) < 1e-12
    got = port.bracket_arakawa_3d2d(g, ops["f"], np.ascontiguousarray(ops["f2"]), m.dx, m.dz)
    assert rel_err(got, ops["bracket_arakawa_3d2d"], g.interior) < 1e-12
    # bracket(Field2D, Field3D) = -bracket(Field3D, Field2D) (difops.cxx:796-799)
    got = -port.bracket_arakawa_3d2d(g, ops["h"], np.ascontiguousarray(ops["f2"]), m.dx, m.dz)
    assert rel_err(got, ops["bracket_arakawa_2d3d"], g.interior) < 1e-12


def test_single_index_forms_match_field_operators(port, orc, ops):
    """tests/unit/include/bout/test_single_index_ops.cxx:37-290: the single-index closed forms equal
    the field operators (there at 1e-15 on a unit metric; here relative on a varying metric)."""
    g, m = _grid(orc, ops), metric_from(orc, ops)
    assert rel_err(orc.si_bracket(ops["f"], ops["h"], m.dx, m.dz), ops["bracket_arakawa"], g.interior) < 1e-12
    assert rel_err(orc.si_ddz(ops["f"], m.dz), ops["DDZ_C2"], g.interior) < 1e-12
    assert rel_err(orc.si_delp2(ops["f"], m), ops["Delp2_fd"], g.interior) < 1e-12


def test_delp2_matches_reference(port, orc, ops):
    g, m = _grid(orc, ops), metric_from(orc, ops)
    got = port.delp2_fft(g, m, ops["f"], all_terms=True, nonuniform=True)
    assert rel_err(got, ops["Delp2_fft"], g.interior) < 1e-12


def test_zfft_ops_match_reference(port, orc, ops):
    g, m, f = _grid(orc, ops), metric_from(orc, ops), ops["f"]
    allr = (slice(None),) * 3
    assert rel_err(port.zfilter(g, f, 3), ops["filter_3"], allr) < 1e-12
    assert rel_err(port.zlowpass(g, f, 4), ops["lowPass_4"], allr) < 1e-12
    assert rel_err(port.zshift(g, f, 0.37, float(m.dz[0, 0]) * g.nz), ops["shiftZ_0.37"], allr) < 1e-12
    dc = port.dc(g, f)
    assert rel_err(np.repeat(dc[:, :, None], g.nz, 2), ops["DC3d"], allr) < 1e-13


@pytest.mark.parametrize("ci", [0, 1])
def test_laplace_flags_match_reference(port, orc, ops, ci):
    g, m = _grid(orc, ops), metric_from(orc, ops)
    kw = {} if ci == 0 else dict(A=ops["lap_A"], C1=ops["lap_C"], C2=ops["lap_C"], D=ops["lap_D"])
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    for fi, fl in enumerate(ops["lap_flagsets"]):
        got = port.laplace_cyclic(g, m, ops["f"], x0=ops["h"], flags=tuple(int(v) for v in fl), **kw)
        assert rel_err(got, ops[f"lap_c{ci}_f{fi}"], region) < 1e-10, fl


def test_boundaries_and_communicate_match_reference(port, orc, ops):
    g, m, f = _grid(orc, ops), metric_from(orc, ops), ops["f"]
    for cond, kind, val in (("dirichlet", "dirichlet", 0.0), ("dirichlet_o2", "dirichlet", 0.0),
                            ("neumann", "neumann", 0.0), ("dirichlet(0.7)", "dirichlet", 0.7),
                            ("neumann(0.7)", "neumann", 0.7)):
        got = f.copy()
        port.bndry_x(g, kind, -1, got, m.dx, val)
        port.bndry_x(g, kind, +1, got, m.dx, val)
        want = ops["bndry_" + cond]
        # x boundaries only here (the y direction is periodic on this grid: ixseps = nx)
        assert np.abs(got - want).max() < 1e-15, cond
    assert np.array_equal(orc.halo_self_reference(g, f, periodic_x=False, periodic_y=True), ops["communicate"])


def test_free_boundaries_match_reference(port, orc, ops):
    """BoundaryFree_O2 / BoundaryFree_O3 (src/mesh/boundary_standard.cxx:3281-3293, :3545-3558) against the
    reference build's output on the same field (tests/golden/bndry_free.npz)."""
    gold = np.load(os.path.join(GOLD, "bndry_free.npz"))
    g, m, f = _grid(orc, ops), metric_from(orc, ops), ops["f"]
    for kind in ("free_o2", "free_o3"):
        got = f.copy()
        port.bndry_x(g, kind, -1, got, m.dx)
        port.bndry_x(g, kind, +1, got, m.dx)
        # 5e-15 (a few ulp, amplified x3 into the second guard cell): the reference build contracts the three-term sum into FMAs (-march=x86-64-v3), the port does not
        assert np.abs(got - gold["bndry_" + kind]).max() <= 5e-15, kind


def test_field_algebra_matches_reference(ops):
    """generated_fieldops.cxx arithmetic restated in numpy with the reference's operation order."""
    f, h, v, f2 = ops["f"], ops["h"], ops["v"], ops["f2"]
    assert np.array_equal(2. * f + h * v, ops["arith_2a_bc"])
    assert np.array_equal(f * h, ops["mul"])
    assert np.array_equal(f * (1.0 / f2)[:, :, None], ops["div_2d"])  # multiplies by the reciprocal (:253-274)
    assert np.array_equal(0.25 - f, ops["s_sub"])
