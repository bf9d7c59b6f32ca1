"""GPU parity against the committed golden fixtures (outputs of the reference's own sources, see
tests/golden/make_golden.py) -- no oracle in the loop.  Tolerances (BASELINE.json north_star):
stencils / bracket / Delp2 1e-12 relative, LaplaceCyclic 1e-10 relative; the reference's float32
benchmark.0.nc at its own 1e-6."""
import os

import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops():
    return dict(np.load(os.path.join(GOLD, "ops_small.npz")))


@pytest.fixture(scope="module")
def lapb():
    return dict(np.load(os.path.join(GOLD, "laplace_benchmark.npz")))


def laplace(capi, g, d, **kw):
    md = {n: dev(d["metric_" + n]) for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22")}
    for k in ("A", "C1", "C2", "D"):
        if kw.get(k) is not None:
            kw[k] = dev(kw[k])
    return capi.Laplace(g, md, float(d["metric_dz"][0, 0]) * g.nz, **kw)


INDEX_CASES = [(b, d, m) for b, ms in (("DDX", ["C2", "C4", "W2", "S2", "W3"]), ("D2DX2", ["C2", "C4"]),
                                       ("D4DX4", ["C2"]), ("VDDX", ["C2", "C4", "U1", "U2", "U3", "W3"]),
                                       ("FDDX", ["U1", "U2", "C2", "C4"])) for d in "XYZ" for m in ms]


@pytest.mark.parametrize("base,d,meth", INDEX_CASES)
def test_index_derivatives(capi, ops, base, d, meth):
    g = capi.Grid(*[int(v) for v in ops["grid"]])
    v = dev(ops["v"]) if base in ("VDDX", "FDDX") else None
    got = capi.stencil(g, f"{base}_{meth}", d, dev(ops["f"]), v=v).cpu().numpy()
    assert rel_err(got, ops[f"index_{base.replace('X', d)}_{meth}"], g.interior) < 1e-12


def test_coordinate_derivatives_fused_scaling(capi, ops):
    """DDX = index DDX / dx etc. in ONE kernel (coordinates.cxx:1467-1514, derivs.cxx:149-166)."""
    g = capi.Grid(*[int(v) for v in ops["grid"]])
    f = dev(ops["f"])
    dx, dy, dz, d1 = (ops["metric_" + n] for n in ("dx", "dy", "dz", "d1_dx"))
    for meth in ("C2", "C4"):
        got = capi.stencil(g, "DDX_" + meth, "X", f, scale1=dev(1 / dx)).cpu().numpy()
        assert rel_err(got, ops["DDX_" + meth], g.interior) < 1e-12
        got = capi.stencil(g, "DDX_" + meth, "Y", f, scale1=dev(1 / dy)).cpu().numpy()
        assert rel_err(got, ops["DDY_" + meth], g.interior) < 1e-12
        got = capi.stencil(g, "DDX_" + meth, "Z", f, scale1=dev(1 / dz)).cpu().numpy()
        assert rel_err(got, ops["DDZ_" + meth], g.interior) < 1e-12
        got = capi.stencil(g, "D2DX2_" + meth, "X", f, scale1=dev(1 / dx ** 2), op2="DDX_C2",
                           scale2=dev(d1 / dx)).cpu().numpy()
        assert rel_err(got, ops["D2DX2_" + meth], g.interior) < 1e-12
    got = capi.ddz_fft(g, f, order=1).cpu().numpy()
    assert rel_err(got, ops["index_DDZ_FFT"], g.interior) < 1e-12
    got = capi.ddz_fft(g, f, order=2).cpu().numpy()
    assert rel_err(got, ops["index_D2DZ2_FFT"], g.interior) < 1e-12


def test_brackets_delp2_zfft(capi, ops):
    g = capi.Grid(*[int(v) for v in ops["grid"]])
    f, h, f2 = dev(ops["f"]), dev(ops["h"]), dev(ops["f2"])
    dx, dz = dev(ops["metric_dx"]), dev(ops["metric_dz"])
    assert rel_err(capi.bracket_arakawa(g, f, h, dx, dz).cpu().numpy(), ops["bracket_arakawa"], g.interior) < 1e-12
    assert rel_err(capi.bracket_arakawa(g, f, f2, dx, dz).cpu().numpy(), ops["bracket_arakawa_3d2d"],
                   g.interior) < 1e-12
    lap = laplace(capi, g, ops)
    assert rel_err(lap.delp2(f).cpu().numpy(), ops["Delp2_fft"], g.interior) < 1e-12
    md = {n: dev(ops["metric_" + n]) for n in ("g11", "g33", "g13", "G1", "G3", "d1_dx")}
    got = capi.delp2_fd(g, f, dx, dz, md["g11"], md["g33"], md["g13"], md["G1"], md["G3"], md["d1_dx"])
    assert rel_err(got.cpu().numpy(), ops["Delp2_fd"], g.interior) < 1e-12
    allr = (slice(None),) * 3
    assert rel_err(capi.zfilter(g, f, 3).cpu().numpy(), ops["filter_3"], allr) < 1e-12
    assert rel_err(capi.zlowpass(g, f, 4).cpu().numpy(), ops["lowPass_4"], allr) < 1e-12
    zl = float(ops["metric_dz"][0, 0]) * g.nz
    assert rel_err(capi.zshift(g, f, 0.37, zl).cpu().numpy(), ops["shiftZ_0.37"], allr) < 1e-12
    dc = capi.dc(g, f).cpu().numpy()
    assert rel_err(np.repeat(dc[:, :, None], g.nz, 2), ops["DC3d"], allr) < 1e-13


@pytest.mark.parametrize("ci", [0, 1])
def test_laplace_flags(capi, ops, ci):
    g = capi.Grid(*[int(v) for v in ops["grid"]])
    kw = {} if ci == 0 else dict(A=ops["lap_A"], C1=ops["lap_C"], C2=ops["lap_C"], D=ops["lap_D"])
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    for fi, fl in enumerate(ops["lap_flagsets"]):
        lap = laplace(capi, g, ops, flags=tuple(int(v) for v in fl), **dict(kw))
        got = lap.solve(dev(ops["f"]), dev(ops["h"])).cpu().numpy()
        lap.status()
        assert rel_err(got, ops[f"lap_c{ci}_f{fi}"], region) < 1e-10, fl


LAP_CASES = {"": ("zeros", "ones", "ones"), "a": ("a", "ones", "ones"), "ac": ("a", "c", "ones"),
             "ad": ("a", "ones", "d")}


@pytest.mark.parametrize("suffix", list(LAP_CASES))
@pytest.mark.parametrize("kind", ["flag0", "flag3", "flagis", "flagos"])
def test_laplace_benchmark_file(capi, lapb, suffix, kind):
    """The reference's golden file tests/integrated/test-laplace/data/benchmark.0.nc."""
    g = capi.Grid(*[int(v) for v in lapb["grid"]])
    coef = {"zeros": None, "ones": None, "a": lapb["a"], "c": lapb["c"], "d": lapb["d"]}
    A, Cc, D = (coef[k] for k in LAP_CASES[suffix])
    flags = {"flag0": (0, 0, 0), "flag3": (0, 3, 0), "flagis": (0, 16, 0), "flagos": (0, 0, 16)}[kind]
    x0 = np.ascontiguousarray(np.repeat(lapb["set_to"][:, :, None], g.nz, axis=2))
    maxmode = int(round((1.0 - 0.2) * g.nz / 2.0))
    lap = laplace(capi, g, lapb, flags=flags, maxmode=maxmode, A=A, C1=Cc, C2=Cc, D=D, all_terms=False,
                  nonuniform=False)
    got = lap.solve(dev(lapb["input"]), dev(x0)).cpu().numpy()
    lap.status()
    name = kind + suffix
    ys = slice(g.ystart, g.yend + 1)
    assert np.allclose(got[:, ys], lapb["bench32_" + name], atol=1e-6, rtol=1e-6)
    assert rel_err(got, lapb["ref64_" + name], (slice(None), ys, slice(None))) < 1e-10


def test_boundaries_and_communicate(capi, ops):
    g = capi.Grid(*[int(v) for v in ops["grid"]])
    dx = dev(ops["metric_dx"])
    for cond, kind, val in (("dirichlet", "dirichlet", 0.0), ("dirichlet_o2", "dirichlet", 0.0),
                            ("neumann", "neumann", 0.0), ("dirichlet(0.7)", "dirichlet", 0.7),
                            ("neumann(0.7)", "neumann", 0.7)):
        F = dev(ops["f"])
        capi.boundary(g, kind, "xin", F, val, dx)
        capi.boundary(g, kind, "xout", F, val, dx)
        assert np.abs(F.cpu().numpy() - ops["bndry_" + cond]).max() < 1e-15, cond
    F = dev(ops["f"])
    capi.halo_self(g, [F], periodic_x=False, periodic_y=True)
    assert np.array_equal(F.cpu().numpy(), ops["communicate"])


def test_field_algebra(capi, ops):
    g = capi.Grid(*[int(v) for v in ops["grid"]])
    f, h, v, f2 = (dev(ops[k]) for k in ("f", "h", "v", "f2"))
    allr = (slice(None),) * 3
    assert rel_err(capi.elementwise(g, [(2.0, f), (1.0, h, v)], like=f).cpu().numpy(), ops["arith_2a_bc"], allr) < 1e-15
    assert np.array_equal(capi.elementwise(g, [(1.0, f, h)], like=f).cpu().numpy(), ops["mul"])
    assert np.array_equal(capi.elementwise(g, [(1.0, f, None, f2)], like=f).cpu().numpy(), ops["div_2d"])
    assert np.array_equal(capi.elementwise(g, [(0.25,), (-1.0, f)], like=f).cpu().numpy(), ops["s_sub"])
