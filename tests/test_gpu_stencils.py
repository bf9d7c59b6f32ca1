"""GPU parity: index-derivative stencils and Arakawa brackets through the C-ABI vs the oracle.

Tolerance: 1e-12 relative to max|reference| (BASELINE.json north_star: "stencil, bracket and Delp2
outputs within 1e-12 relative").  The reference tests these operators on random fields at 1e-15
(tests/unit/include/bout/test_single_index_ops.cxx:37-290) and on a 3x5x7 mesh
(tests/unit/fake_mesh_fixture.hxx) -- odd nz and tiny grids are covered here too.
"""
import numpy as np
import pytest
import torch

from util import metric_2d, rel_err, seeded_fields

pytestmark = pytest.mark.gpu
TOL = 1e-12

STD_OPS = ["DDX_C2", "DDX_C4", "DDX_W2", "DDX_S2", "DDX_W3", "D2DX2_C2", "D2DX2_C4", "D4DX4_C2"]
UPWIND_OPS = ["VDDX_C2", "VDDX_C4", "VDDX_U1", "VDDX_U2", "VDDX_U3", "VDDX_W3"]
FLUX_OPS = ["FDDX_U1", "FDDX_U2", "FDDX_C2", "FDDX_C4"]
GRIDS = [  # nx, ny, nz, mxg, myg
    (20, 12, 64, 2, 2),
    (7, 9, 7, 2, 2),      # odd nz -> scalar path (FakeMeshFixture-like 3x5x7 interior)
    (37, 6, 2, 2, 2),     # nz = 2: every z neighbour wraps
    (70, 5, 16, 2, 2),    # several marching chunks in x
]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("direction", ["X", "Y", "Z"])
@pytest.mark.parametrize("op", STD_OPS + UPWIND_OPS + FLUX_OPS)
def test_index_stencil_matches_oracle(capi, orc, port, op, direction, grid):
    g = capi.Grid(*grid)
    og = orc.Grid(*grid)
    f, v = seeded_fields(g.shape3, 2)
    v = v - 0.05  # both signs, exercises the upwind branches
    needs_v = not op.startswith(("DDX", "D2DX2", "D4DX4"))
    want = port.index_deriv(og, op, direction, f, v=v if needs_v else None)
    got = capi.stencil(g, op, direction, dev(f), v=dev(v) if needs_v else None).cpu().numpy()
    # floor: e.g. W3 along z with nz=2 is analytically zero -> compare against |f| instead
    assert rel_err(got, want, g.interior, floor=1e-3 * np.abs(f).max()) < TOL
    # cells outside the region are untouched (zeros from the harness allocation)
    mask = np.ones(g.shape3, bool)
    mask[g.interior] = False
    assert np.all(got[mask] == 0.0)


@pytest.mark.parametrize("grid", GRIDS[:2])
def test_coordinate_derivatives_fused_scaling(capi, orc, port, grid):
    """DDX = index/dx, D2DX2 = index/dx^2 + d1_dx*DDX/dx in one pass (derivs.cxx:149-166)."""
    g = capi.Grid(*grid)
    og = orc.Grid(*grid)
    (f,) = seeded_fields(g.shape3, 1)
    dx, dy, dz = metric_2d(g.shape2)
    d1_dx = 0.2 * np.cos(np.arange(g.nx))[:, None] * np.ones(g.shape2)
    m = orc.Metric(**{n: np.ones(g.shape2) for n in orc.Metric.NAMES})
    m.dx, m.dy, m.dz, m.d1_dx, m.non_uniform = dx, dy, dz, d1_dx, True
    for meth in ("C2", "C4"):
        want = port.ddx(og, m, f, meth)
        got = capi.stencil(g, "DDX_" + meth, "X", dev(f), scale1=dev(1 / dx)).cpu().numpy()
        assert rel_err(got, want, g.interior) < TOL
        want = port.ddz(og, m, f, meth)
        got = capi.stencil(g, "DDX_" + meth, "Z", dev(f), scale1=dev(1 / dz)).cpu().numpy()
        assert rel_err(got, want, g.interior) < TOL
        for first in ("C2", "C4"):
            want = port.d2dx2(og, m, f, meth, first)
            got = capi.stencil(g, "D2DX2_" + meth, "X", dev(f), scale1=dev(1 / dx ** 2),
                               op2="DDX_" + first, scale2=dev(d1_dx / dx)).cpu().numpy()
            assert rel_err(got, want, g.interior) < TOL


def test_y_stencil_with_parallel_slices(capi, orc, port):
    """DIRECTION::YOrthogonal reads yup/ydown (stencils.hxx:40-119)."""
    grid = (12, 10, 32, 2, 2)
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    f, up, down = seeded_fields(g.shape3, 3)
    for op in ("DDX_C2", "D2DX2_C2"):
        want = port.index_deriv(og, op, "Y", f, yup=up, ydown=down)
        got = capi.stencil(g, op, "Y", dev(f), yup=dev(up), ydown=dev(down)).cpu().numpy()
        assert rel_err(got, want, g.interior) < TOL


def test_single_point_direction_gives_zero(capi):
    """nPoint == 1 -> zeroFrom(f) (index_derivs_interface.hxx:152-156)."""
    g = capi.Grid(12, 1, 8, 2, 0)
    f = torch.rand(g.shape3, dtype=torch.float64, device="cuda")
    out = torch.full_like(f, 7.0)
    capi.stencil(g, "DDX_C2", "Y", f, out=out)
    assert torch.all(out == 0)


def test_guard_cell_requirement_is_an_error(capi):
    """ASSERT2(getNguard(direction) >= nGuards) (index_derivs.hxx:89) -> status, not UB."""
    g = capi.Grid(12, 6, 8, 1, 1)
    f = torch.rand(g.shape3, dtype=torch.float64, device="cuda")
    with pytest.raises(capi.BoutB200Error) as e:
        capi.stencil(g, "DDX_C4", "X", f)
    assert e.value.status == -3
    capi.stencil(g, "DDX_C2", "X", f)  # one guard is enough for C2
    with pytest.raises(capi.BoutB200Error):
        capi.stencil(g, "DDX_C2", "X", f, region="RGN_ALL")


@pytest.mark.parametrize("grid", GRIDS + [(12, 1, 16, 2, 0)])
def test_arakawa_bracket(capi, orc, port, grid):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    f, h = seeded_fields(g.shape3, 2)
    dx, dy, dz = metric_2d(g.shape2)
    want = port.bracket_arakawa(og, f, h, dx, dz)
    got = capi.bracket_arakawa(g, dev(f), dev(h), dev(dx), dev(dz)).cpu().numpy()
    floor = 1e-3 * np.abs(f).max() * np.abs(h).max() / (dx.min() * dz.min())  # nz=2: analytically 0
    assert rel_err(got, want, g.interior, floor=floor) < TOL
    h2 = np.ascontiguousarray(h[:, :, 0])
    want = port.bracket_arakawa_3d2d(og, f, h2, dx, dz)
    got = capi.bracket_arakawa(g, dev(f), dev(h2), dev(dx), dev(dz)).cpu().numpy()
    assert rel_err(got, want, g.interior, floor=floor) < TOL


def test_arakawa_antisymmetry_full_size(capi):
    """Size-independent property at the BASELINE C2 size: [f,g] = -[g,f] exactly term-by-term is
    not guaranteed in floating point, but to round-off it is; and [f,f] = 0."""
    g = capi.Grid(260, 68, 256, 2, 2)
    gen = torch.Generator(device="cuda").manual_seed(5)
    f = torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen)
    h = torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen)
    dx = torch.full(g.shape2, 0.2, dtype=torch.float64, device="cuda")
    dz = torch.full(g.shape2, 0.1, dtype=torch.float64, device="cuda")
    a = capi.bracket_arakawa(g, f, h, dx, dz)
    b = capi.bracket_arakawa(g, h, f, dx, dz)
    scale = a.abs().max()
    assert ((a + b).abs().max() / scale).item() < 1e-13
    c = capi.bracket_arakawa(g, f, f, dx, dz)
    assert (c.abs().max() / scale).item() < 1e-13
