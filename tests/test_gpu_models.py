"""GPU parity of whole-model right-hand sides and short RK trajectories against the reference's
UNCHANGED example models run through the reference's own Solver (tests/golden/models.npz, made by
tests/golden/make_golden.py).  Tolerances (north_star): RHS 1e-12 relative except where a
LaplaceCyclic solve is inside the RHS (1e-10); short-horizon trajectories 1e-8."""
import os

import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLD, "models.npz")))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def metric(gold, tag):
    return {n: dev(gold[f"{tag}_metric_{n}"]) for n in ("dx", "dy", "dz", "J", "g_22", "g11", "g33", "g13",
                                                          "G1", "G3", "d1_dx", "g_11")}


def test_hw3d_rhs_and_rk4_trajectory(capi, gold):
    import __graft_entry__ as ge
    models = ge.load_package().models
    g = capi.Grid(*[int(v) for v in gold["hw3d_grid"]])
    m = models.HW3D(g, metric(gold, "hw3d"), alpha=1.0, kappa=0.5, Dn=1e-3, Dvort=1e-3)
    m.set_state(dev(gold["hw3d_init_n"]), dev(gold["hw3d_init_vort"]))
    m.rhs()
    I = g.interior
    assert rel_err(m.ddt_n.cpu().numpy(), gold["hw3d_ddt_n"], I) < 1e-10
    assert rel_err(m.ddt_vort.cpu().numpy(), gold["hw3d_ddt_vort"], I) < 1e-10
    dt, nsteps = gold["hw3d_dt_nsteps"]
    m.set_state(dev(gold["hw3d_init_n"]), dev(gold["hw3d_init_vort"]))
    for _ in range(int(nsteps)):
        m.rk4_step(float(dt))
    assert rel_err(m.n.cpu().numpy(), gold["hw3d_final_n"], I) < 1e-8
    assert rel_err(m.vort.cpu().numpy(), gold["hw3d_final_vort"], I) < 1e-8


def test_blob2d_harness_rhs_and_rk4_trajectory(capi, gold):
    """models.Blob2D (the operator sequence of examples/blob2d/blob2d.cxx:165-213 through the C-ABI; the model bench.py
    decomposes in X for configs[2]) against the reference build's run of the unchanged source: initial profile,
    one Solver::run_rhs (1e-10) and the 5-step RK4 trajectory (1e-8)."""
    import __graft_entry__ as ge
    models = ge.load_package().models
    g = capi.Grid(*[int(v) for v in gold["blob2d_grid"]])
    m = models.Blob2D(g, metric(gold, "blob2d"), Te0=5.0, B0=0.35, D_n=1e-6, D_vort=1e-6, R_c=1.5, L_par=10.0)
    I = (slice(g.xstart, g.xend + 1), slice(g.ystart, g.yend + 1), slice(None))
    n0, o0 = models.blob2d_initial_state(g, height=0.5, width=0.1)
    assert rel_err(n0.cpu().numpy(), gold["blob2d_init_n"], I) < 1e-13
    m.set_state(dev(gold["blob2d_init_n"]), dev(gold["blob2d_init_omega"]))
    m.rhs()
    assert rel_err(m.ddt_n.cpu().numpy(), gold["blob2d_ddt_n"], I) < 1e-10
    assert rel_err(m.ddt_omega.cpu().numpy(), gold["blob2d_ddt_omega"], I) < 1e-10
    dt, nsteps = gold["blob2d_dt_nsteps"]
    m.set_state(dev(gold["blob2d_init_n"]), dev(gold["blob2d_init_omega"]))
    for _ in range(int(nsteps)):
        m.rk4_step(float(dt))
    assert rel_err(m.n.cpu().numpy(), gold["blob2d_final_n"], I) < 1e-8
    assert rel_err(m.omega.cpu().numpy(), gold["blob2d_final_omega"], I) < 1e-8


@pytest.mark.parametrize("nz", [16, 128])
def test_hw3d_guard_cells_written_by_the_producing_kernels(capi, gold, nz, monkeypatch):
    """bout_b200_hw3d_rhs_guards / bout_b200_laplace.wrap_yguards: the x guard cells (Dirichlet of Solver::pre_rhs) and
    periodic y guard rows (single-rank Mesh::communicate) written by the fused kernel's epilogue and by the solve are
    bitwise what the separate boundary / exchange launches give (nz = 16: register-marching kernel + three-kernel
    solve, guard cells by the fall-back launches; nz = 128: bulk-copy kernel + block solve, in-kernel)."""
    import __graft_entry__ as ge
    models = ge.load_package().models
    g = capi.Grid(20, 9, nz, 2, 2)
    one = lambda v: torch.full(g.shape2, v, dtype=torch.float64, device="cuda")
    md = dict(dx=one(0.3), dy=one(0.7), dz=one(0.2), J=one(1.0), g_22=one(1.0), g11=one(1.0), g33=one(1.0), g13=one(0.0),
              G1=one(0.0), G3=one(0.0), d1_dx=one(0.0), g_11=one(1.0))
    gen = torch.Generator(device="cuda").manual_seed(5)
    n0 = torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    v0 = torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    res = {}
    for fill in ("0", "1"):
        monkeypatch.setenv("BOUT_B200_GUARD_FILL", fill)
        m = models.HW3D(g, md, alpha=1.0, kappa=0.5, Dn=1e-3, Dvort=1e-3)
        assert m.guard_fill == (fill == "1")
        m.set_state(n0, v0)
        launches = capi.launch_count()
        for _ in range(3):
            m.rk4_step(1e-3)
        launches = capi.launch_count() - launches
        phi = m.phi.clone()
        n, v = m.n.clone(), m.vort.clone()
        if fill == "0":   # what the next right-hand side's pre_rhs + communicate would make of them
            m._pre_rhs(n, v)
            capi.halo_self(g, [n, v], periodic_x=False, periodic_y=True)
            capi.halo_self(g, [phi], periodic_x=False, periodic_y=True)
        res[fill] = (n, v, phi, launches)
    for a, b, name in zip(res["0"][:3], res["1"][:3], ("n", "vort", "phi")):
        assert torch.equal(a, b), name
    assert torch.isfinite(res["1"][0]).all() and float(res["1"][0].abs().max()) > 0.1
    if nz == 128:
        # no boundary / exchange launches between the stages: 12 right-hand sides of 4 solve launches + the fused kernel
        # as three (the edge rows apart from the rest), plus the boundary / exchange launches of the very first state
        assert res["1"][3] <= 12 * 7 + 4


# ---- UNCHANGED reference models compiled against the C++ host layer -----------------------------------
@pytest.fixture(scope="module")
def hostmod():
    import __graft_entry__ as ge
    return ge.load_package().host


def _cases():
    import sys
    sys.path.insert(0, GOLD)
    import cases
    return cases


@pytest.mark.parametrize("tag,model,solver", [("hw", "hw", "rk4"), ("hw_rk3", "hw", "rk3ssp"), ("blob2d", "blob2d", "rk4"),
                                              ("hw3d", "hw3d", "rk4")])
def test_unchanged_model_through_host_layer(hostmod, gold, tag, model, solver):
    """examples/hasegawa-wakatani/hw.cxx, examples/blob2d/blob2d.cxx and examples/hasegawa-wakatani-3d/hw.cxx (the
    headline workload: its single-index BOUT_FOR_RAJA loop is recorded and dispatched to bout_b200_hw3d_rhs), compiled
    unchanged against
    bout-dev_b200/host/include/bout/*.hxx, run device-resident: initial profiles, one Solver::run_rhs
    and a 5-step RK trajectory against the reference build."""
    cases = _cases()
    ini = {"hw": cases.HW_INI, "hw_rk3": cases.HW_INI, "blob2d": cases.BLOB_INI, "hw3d": cases.HW3D_INI}[tag]
    h = hostmod.Host(ini)
    try:
        names = h.model_create(model, solver)
        nx, ny, nz, mxg, myg = [int(v) for v in gold[tag + "_grid"]]
        assert h.shape3 == (nx, ny, nz)
        I = (slice(mxg, nx - mxg), slice(myg, ny - myg), slice(None))
        for v in names:   # initial profiles (expression evaluator, GlobalX/Y/Z) match the reference's FieldFactory
            assert rel_err(h.model_get(v), gold[f"{tag}_init_{v}"], I, floor=1e-300) < 1e-13, v
        for n in ("dx", "dz", "g11", "g33", "J", "g_22"):
            assert np.allclose(h.metric(n), gold[f"{tag}_metric_{n}"], rtol=1e-14, atol=0)
        h.model_rhs(0.0)
        for v in names:
            assert rel_err(h.model_ddt(v), gold[f"{tag}_ddt_{v}"], I) < 1e-10, v
        allocs = None
        dt, nsteps = gold[tag + "_dt_nsteps"]
        for v in names:
            h.model_set(v, gold[f"{tag}_init_{v}"])
        h.model_solve(1, float(dt) * int(nsteps))
        for v in names:
            assert rel_err(h.model_get(v), gold[f"{tag}_final_{v}"], I) < 1e-8, v
        allocs = h.pool_allocations()
        assert allocs > 0
    finally:
        h.close()


def test_host_layer_pool_reaches_steady_state(hostmod, gold):
    """Array pool (reference include/bout/array.hxx:423-472): no device allocation in steady state."""
    cases = _cases()
    h = hostmod.Host(cases.HW_INI)
    try:
        h.model_create("hw", "rk4")
        h.model_rhs(0.0)
        h.model_rhs(0.0)
        a0 = h.pool_allocations()
        for _ in range(5):
            h.model_rhs(0.0)
        assert h.pool_allocations() == a0
    finally:
        h.close()


def test_single_index_loop_generic_dispatch():
    """tests/models/si_loop.cxx: a BOUT_FOR_RAJA loop outside the fused hw3d pattern, dispatched operator by operator
    on the device, against the same expression with whole-field operators (tests/si_loop_check.py)."""
    import json, os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, os.path.join(here, "si_loop_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


@pytest.mark.parametrize("lazy", ["1", "0"])
def test_deferred_fused_field_algebra_on_device(lazy):
    """tests/models/lazy_ops.cxx on the device (tests/lazy_ops_check.py): deferred, fused Field3D algebra of the host layer
    against the element-wise formulas; with fusion the 21 operators of the model's rhs() are at most 10 launches."""
    import json, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, os.path.join(here, "lazy_ops_check.py")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, BOUT_B200_LAZY=lazy))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["ok"]
    if lazy == "1":
        assert 0 < res["launches"] <= 10, res


def test_vectors_grad_where_through_the_host_layer():
    """tests/models/vec_ops.cxx on the device (tests/vec_ops_check.py): Vector2D / Vector3D, Grad, dot products,
    toContravariant, where (bout_b200_where), interp_to against the reference build's run of the same source (1e-12)."""
    import json, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, os.path.join(here, "vec_ops_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_parallel_operators_and_flux_derivatives_through_the_host_layer():
    """tests/models/par_ops.cxx on the device (tests/par_ops_check.py): Grad_par, Vpar_Grad_par, Div_par, Grad2_par2,
    VDDY, FDDX / FDDY / FDDZ of the host layer against the reference build's run of the same source (1e-12)."""
    import json, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    p = subprocess.run([sys.executable, os.path.join(here, "par_ops_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_where_kernel(capi):
    """bout_b200_where against numpy (reference include/bout/where.hxx: test > 0 ? gt0 : le0), 3-D and 2-D tests,
    field and constant branches; bit-exact (a selection)."""
    g = capi.Grid(9, 6, 10, 2, 1)
    gen = torch.Generator(device="cuda").manual_seed(2)
    t3 = torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    t3[2, 3, 4] = 0.0                      # the boundary case takes the le0 branch
    t2 = torch.rand(g.shape2, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    a = torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen)
    b = torch.rand(g.shape3, dtype=torch.float64, device="cuda", generator=gen)
    assert torch.equal(capi.where(g, t3, a, b), torch.where(t3 > 0, a, b))
    assert torch.equal(capi.where(g, t3, 2.5, b), torch.where(t3 > 0, torch.full_like(b, 2.5), b))
    assert torch.equal(capi.where(g, t3, a, -1.0), torch.where(t3 > 0, a, torch.full_like(a, -1.0)))
    assert torch.equal(capi.where(g, t2, a, b), torch.where(t2[:, :, None] > 0, a, b))
