"""GPU parity: field algebra, boundary conditions, self halo exchange, RK stage updates."""
import numpy as np
import pytest
import torch

from util import metric_2d, rel_err, seeded_fields

pytestmark = pytest.mark.gpu
GRIDS = [(12, 8, 16, 2, 2), (7, 6, 7, 2, 1), (20, 1, 32, 2, 0)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("grid", GRIDS)
def test_elementwise_matches_generated_fieldops(capi, grid):
    """Bit-exact against the arithmetic of src/field/generated_fieldops.cxx (same operation order;
    division by Field2D / scalar multiplies by the reciprocal, :253-274, :543-558)."""
    g = capi.Grid(*grid)
    a, b, c = seeded_fields(g.shape3, 3)
    b = b + 2.0  # keep divisors away from zero
    f2 = metric_2d(g.shape2)[0]
    A, B, Cc, F2 = dev(a), dev(b), dev(c), dev(f2)
    b3 = lambda x: x[:, :, None]
    cases = [
        ([(1.0, A), (1.0, B)], a + b),
        ([(1.0, A), (-1.0, B)], a - b),
        ([(1.0, A, B)], a * b),
        ([(1.0, A, None, B)], a / b),
        ([(1.0, A, F2)], a * b3(f2)),
        ([(1.0, A, None, F2)], a * b3(1.0 / f2)),
        ([(1.0, A), (1.0, F2)], a + b3(f2)),
        ([(3.5, A)], a * 3.5),
        ([(1.0 / 3.5, A)], a * (1.0 / 3.5)),
        ([(1.0, A), (0.25,)], a + 0.25),
        ([(0.25,), (-1.0, A)], 0.25 - a),
        ([(2.0, None, None, B)], 2.0 / b),
        ([(2.0, A), (1.0, B, Cc)], 2. * a + b * c),   # examples/performance/arithmetic
    ]
    for terms, want in cases:
        got = capi.elementwise(g, terms, like=A).cpu().numpy()
        assert np.array_equal(got, want), terms
    got = capi.elementwise(g, [(2.0, A), (1.0, B, Cc)], region="RGN_NOBNDRY", like=A).cpu().numpy()
    assert np.array_equal(got[g.interior], (2. * a + b * c)[g.interior])
    mask = np.ones(g.shape3, bool)
    mask[g.interior] = False
    assert np.all(got[mask] == 0)


def test_unary(capi):
    g = capi.Grid(8, 6, 16, 2, 2)
    (a,) = seeded_fields(g.shape3, 1)
    A = dev(a)
    assert np.array_equal(capi.unary(g, "neg", A).cpu().numpy(), -1.0 * a)
    assert np.array_equal(capi.unary(g, "abs", A).cpu().numpy(), np.abs(a))
    assert np.array_equal(capi.unary(g, "sq", A).cpu().numpy(), a * a)
    assert np.array_equal(capi.unary(g, "sqrt", dev(np.abs(a))).cpu().numpy(), np.sqrt(np.abs(a)))
    assert np.allclose(capi.unary(g, "exp", A).cpu().numpy(), np.exp(a), rtol=1e-15)
    assert np.allclose(capi.unary(g, "tanh", A).cpu().numpy(), np.tanh(a), rtol=1e-14, atol=1e-16)


@pytest.mark.parametrize("grid", [(12, 8, 16, 2, 2), (9, 6, 7, 1, 1)])
@pytest.mark.parametrize("kind", ["dirichlet", "neumann"])
def test_x_boundaries(capi, orc, port, grid, kind):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    (f,) = seeded_fields(g.shape3, 1)
    dx = metric_2d(g.shape2)[0]
    for val in (0.0, 0.7):
        want = f.copy()
        port.bndry_x(og, kind, -1, want, dx, val)
        port.bndry_x(og, kind, +1, want, dx, val)
        F = dev(f)
        capi.boundary(g, kind, "xin", F, val, dev(dx))
        capi.boundary(g, kind, "xout", F, val, dev(dx))
        if kind == "dirichlet":  # 2*val - f is exact on both sides
            assert np.array_equal(F.cpu().numpy(), want)
        else:  # f + delta*val: the GPU contracts to an FMA, gcc does not -> 1 ulp
            assert rel_err(F.cpu().numpy(), want, (slice(None),) * 3) < 1e-15


def test_zero_boundary_apply_ddt(capi):
    g = capi.Grid(12, 8, 16, 2, 2)
    F = torch.ones(g.shape3, dtype=torch.float64, device="cuda")
    capi.boundary(g, "zero", "xin", F)
    capi.boundary(g, "zero", "xout", F)
    f = F.cpu().numpy()
    assert np.all(f[:2, 2:6] == 0) and np.all(f[-2:, 2:6] == 0) and np.all(f[2:-2] == 1)
    assert np.all(f[:2, :2] == 1)  # y guards are not part of the x boundary region


@pytest.mark.parametrize("grid", [(12, 8, 16, 2, 2), (10, 6, 7, 2, 1), (12, 1, 8, 2, 0)])
@pytest.mark.parametrize("px,py", [(False, True), (True, False), (True, True)])
def test_halo_self_exchange(capi, orc, grid, px, py):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    fields = seeded_fields(g.shape3, 3)
    devs = [dev(f) for f in fields]
    capi.halo_self(g, devs, periodic_x=px, periodic_y=py)
    for f, d in zip(fields, devs):
        assert np.array_equal(d.cpu().numpy(), orc.halo_self_reference(og, f, px, py))


def test_pack_unpack_y_roundtrip(capi):
    g = capi.Grid(12, 8, 16, 2, 2)
    fields = [dev(f) for f in seeded_fields(g.shape3, 2)]
    buf = torch.zeros(2 * g.nx * 2 * g.nz, dtype=torch.float64, device="cuda")
    capi.halo_pack_y(g, fields, g.yend - 1, 2, 0, g.nx - 1, buf)
    want = torch.stack([f[:, g.yend - 1:g.yend + 1, :] for f in fields]).reshape(-1)
    assert torch.equal(buf, want)  # [field][x][y][z] as pack_data lays it out
    capi.halo_unpack_y(g, fields, 0, 2, 0, g.nx - 1, buf)
    for f in fields:
        assert torch.equal(f[:, 0:2, :], f[:, g.yend - 1:g.yend + 1, :])


def test_rk_stage_updates(capi, orc, port):
    grid = (12, 8, 16, 2, 2)
    g = capi.Grid(*grid)
    s, k1, k2, k3, k4 = seeded_fields(g.shape3, 5)
    dt = 0.0371
    I = g.interior
    got = capi.rk_stage(g, "axpy", dt, dev(s), dev(k1), c=0.5).cpu().numpy()
    assert rel_err(got, port.rk4_stage(s, k1, 0.5 * dt), I) < 1e-15  # FMA vs mul+add
    got = capi.rk_stage(g, "rk4_final", dt, dev(s), dev(k1), dev(k2), dev(k3), dev(k4)).cpu().numpy()
    assert rel_err(got, port.rk4_final(s, dt, k1, k2, k3, k4), I) < 1e-15
    got = capi.rk_stage(g, "rk3_stage2", dt, dev(s), dev(k1), dev(k2)).cpu().numpy()
    assert rel_err(got, 0.75 * s + 0.25 * k1 + 0.25 * dt * k2, I) < 1e-15
    got = capi.rk_stage(g, "rk3_stage3", dt, dev(s), dev(k1), dev(k2)).cpu().numpy()
    assert rel_err(got, (1. / 3) * s + (2. / 3.) * (k1 + dt * k2), I) < 1e-15


def test_boundary_batch_equals_single_calls(capi):
    """bout_b200_boundary_batch: several (field, side, condition) entries in one launch give bitwise what one
    bout_b200_boundary call per entry gives (x sides in one batch, y sides -- whose corner cells read what the x sides
    wrote -- in a second one, as Solver::pre_rhs of the host layer issues them); more entries than one launch holds."""
    g = capi.Grid(11, 9, 12, 2, 2)
    gen = torch.Generator(device="cuda").manual_seed(9)
    fields = [torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen) for _ in range(5)]
    dx = torch.rand(g.shape2, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    dy = torch.rand(g.shape2, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    kinds = ["dirichlet", "neumann", "free_o2", "free_o3", "extend"]
    want = [f.clone() for f in fields]
    for f, k in zip(want, kinds):
        for side in ("xin", "xout"):
            capi.boundary(g, k, side, f, value=0.3, spacing=dx)
        for side in ("ydown", "yup"):
            capi.boundary(g, k, side, f, value=0.3, spacing=dy)
    got = [f.clone() for f in fields]
    capi.boundary_batch(g, [(k, side, f, 0.3, dx) for f, k in zip(got, kinds) for side in ("xin", "xout")])     # 10 entries
    capi.boundary_batch(g, [(k, side, f, 0.3, dy) for f, k in zip(got, kinds) for side in ("ydown", "yup")])
    for a, b, k in zip(got, want, kinds):
        assert torch.equal(a, b), k
