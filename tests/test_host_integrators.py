"""CPU: host logic of the explicit integrators in the C++ host layer (bout-dev_b200/host: rk4 fixed/adaptive, rk3ssp,
euler, rkgeneric with rkf45 / cashkarp / rk4 / rkf34, splitrk) against the reference's own solvers on
tests/models/ode_pair.cxx (tests/golden/integrators.npz, made by tests/golden/make_golden.py integrators).

There is no CPU compute path in the product, so the child process LD_PRELOADs tests/fake_device/fake_b200.c, a
host-memory test double of the few C-ABI entry points the integrators call.  What this pins: Solver::create, the
accept / reject control flow (same number of RHS evaluations as the reference), the stage and output formulas, the
copy-on-write state handling.  The same cases run through the real kernels in tests/test_zz_gpu_integrators.py."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases import ELEM_CASES, HW_INTEGRATOR_CASES, INTEGRATOR_CASES  # noqa: E402

ALL_CASES = sorted(INTEGRATOR_CASES) + sorted(ELEM_CASES)

HOST_LIB = os.path.join(ROOT, "bout-dev_b200", "host", "libbout_b200_host.so")
MODEL = os.path.join(ROOT, "bout-dev_b200", "host", "models", "ode_pair.so")
MODEL2 = os.path.join(ROOT, "bout-dev_b200", "host", "models", "elem_access.so")


@pytest.fixture(scope="module")
def fake_device(tmp_path_factory):
    for p in (HOST_LIB, MODEL, MODEL2):
        if not os.path.exists(p):
            pytest.fail(f"{p} missing: run python -c 'import __graft_entry__ as g; g.build()'")
    so = str(tmp_path_factory.mktemp("fake") / "libfake_b200.so")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                    "-o", so, os.path.join(HERE, "fake_device", "fake_b200.c"), "-lm", "-ldl"], check=True)
    return so


@pytest.fixture(scope="module")
def results(fake_device):
    """All cases in ONE child process (one torch import)."""
    env = dict(os.environ, LD_PRELOAD=fake_device)
    p = subprocess.run([sys.executable, os.path.join(HERE, "integrator_check.py")] + ALL_CASES, env=env,
                       capture_output=True, text=True, timeout=600)
    out = {}
    for line in p.stdout.splitlines():
        if line.startswith("{"):
            r = json.loads(line)
            out[r["tag"]] = r
    assert out, p.stdout[-2000:] + p.stderr[-2000:]
    return out


@pytest.mark.parametrize("tag", ALL_CASES)
def test_integrator_matches_reference_solver(results, tag):
    """INTEGRATOR_CASES: every solver / scheme / controller and the boundary-condition names on ode_pair.cxx;
    ELEM_CASES: host-side element access (f(x,y,z), f[i], BOUT_FOR, Ind3D::zp) mixed with device operations."""
    res = results[tag]
    assert "error" not in res, res
    assert res["rhs_calls"] == res["ref_rhs_calls"], res
    assert res["ok"] and res["tol"] <= 1e-10, res


def test_without_the_test_double_the_host_layer_fails_loudly():
    """No GPU and no test double: the first device allocation must raise, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([sys.executable, os.path.join(HERE, "integrator_check.py"), "euler"], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode != 0
    assert "HostError" in p.stdout and "error" in p.stdout, p.stdout[-1500:] + p.stderr[-1500:]


def test_unknown_solver_and_scheme_are_rejected(fake_device):
    code = (
        "import sys, os\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(HERE, 'golden')!r})\n"
        "import __graft_entry__ as ge\n"
        "from cases import ode_ini\n"
        "hostmod = ge.load_package().host\n"
        "for solver, opts in (('cvode', ''), ('rkgeneric', 'scheme = dopri5')):\n"
        "    h = hostmod.Host(ode_ini(opts, False, 1, 0.1))\n"
        "    try:\n"
        "        h.model_create('ode_pair', solver)\n"
        "    except hostmod.HostError as e:\n"
        "        print('rejected:', e)\n"
        "    else:\n"
        "        sys.exit(3)\n"
        "    finally:\n"
        "        h.close()\n")
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, LD_PRELOAD=fake_device), capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("rejected:") == 2


def test_output_step_writes_a_readable_dump_file(fake_device, tmp_path):
    """[output] path: NetCDF classic dump file (dims, scalars, t_array, iteration, one record per output incl. the
    initial state); record k of a 3-output run == final state of a k-output run (tests/output_check.py)."""
    p = subprocess.run([sys.executable, os.path.join(HERE, "output_check.py"), str(tmp_path)],
                       env=dict(os.environ, LD_PRELOAD=fake_device), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["ok"] and res["records"] == 4, res


def test_output_monitor_can_stop_the_run(fake_device, tmp_path):
    """PhysicsModel::outputMonitor returning non-zero ends Solver::solve with the reference's exception text
    (Solver::call_monitors, src/solver/solver.cxx:893-896); the records written so far stay in the dump file."""
    path = str(tmp_path / "BOUT.dmp.0.nc")
    code = (
        "import sys, os\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(HERE, 'golden')!r})\n"
        "import __graft_entry__ as ge\n"
        "from cases import ode_ini\n"
        "hostmod = ge.load_package().host\n"
        f"ini = ode_ini('timestep = 0.0625', False, 4, 0.25).replace('[ode]', '[ode]\\nquit_at = 1') + '\\n[output]\\npath = {path}\\n'\n"
        "h = hostmod.Host(ini)\n"
        "h.model_create('ode_pair', 'rk4')\n"
        "try:\n"
        "    h.model_solve(4, 0.25)\n"
        "except hostmod.HostError as e:\n"
        "    print('stopped:', e)\n"
        "else:\n"
        "    sys.exit(3)\n"
        "finally:\n"
        "    h.close()\n")
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, LD_PRELOAD=fake_device), capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Monitor signalled to quit (return code 7)" in p.stdout
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as nc:
        assert list(nc.variables["iteration"][:]) == [-1, 0, 1]


def test_unchanged_imex_example_with_its_shipped_input(fake_device):
    """examples/IMEX/advection-diffusion: imex.cxx and data/BOUT.inp as shipped (solver type = splitrk, nstages 9,
    adapt_period 5; `zs = z / (2π)`; W3 upwinding, C4 second derivative; nout reduced to 2) through the host layer:
    Options -> Mesh -> Solver::create -> initial profile -> the whole trajectory, with the stencils and the halo
    served by the oracle's C restatement behind the test double.  Final state within 1e-8 of the reference's own run
    (observed 3e-12, the reference's run-to-run spread) and the same 465 RHS evaluations: this pins the host layer's
    derivative dispatch (method names -> operators, 1/dx factors, communicate) and the splitrk controller."""
    port = os.path.join(ROOT, "oracle", "_ref", "liboracle_port.so")
    for need in (os.path.join(ROOT, "bout-dev_b200", "host", "models", "imex.so"), port):
        if not os.path.exists(need):
            pytest.skip(f"{need} not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "integrator_check.py"), "imex_shipped"],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=port), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["ok"] and res["rhs_calls"] == res["ref_rhs_calls"] and res["err_U"] < 1e-8, res


def test_unchanged_eigen_box_example_with_its_shipped_input(fake_device):
    """examples/eigen-box: eigen-box.cxx and data/BOUT.inp as shipped (mesh sizes as expressions of MXG, nz = 1,
    MXG = 1, a quoted string value, per-variable bndry_all incl. `none`), stepped with rk4 instead of the shipped slepc
    eigen-solver, against the reference's own run: final state at rounding level, same RHS count."""
    port = os.path.join(ROOT, "oracle", "_ref", "liboracle_port.so")
    for need in (os.path.join(ROOT, "bout-dev_b200", "host", "models", "eigenbox.so"), port):
        if not os.path.exists(need):
            pytest.skip(f"{need} not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "integrator_check.py"), "eigenbox_shipped"],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=port), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["ok"] and res["rhs_calls"] == res["ref_rhs_calls"], res


def test_unchanged_advection_reaction_example(fake_device):
    """examples/IMEX/advection-reaction/split_operator.cxx unchanged (bracket(phi, U, BRACKET_SIMPLE), split operator,
    initial_profile(), SAVE_ONCE, nz from the root option MZ) with the shipped methods / solver (splitrk) and profiles
    on a mesh given as options (the shipped grid file is NetCDF): 1e-8 of the reference's run, same RHS count."""
    port = os.path.join(ROOT, "oracle", "_ref", "liboracle_port.so")
    for need in (os.path.join(ROOT, "bout-dev_b200", "host", "models", "advreact.so"), port):
        if not os.path.exists(need):
            pytest.skip(f"{need} not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "integrator_check.py"), "advreact"],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=port), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert res["ok"] and res["rhs_calls"] == res["ref_rhs_calls"], res


def test_expression_evaluator_and_model_setup_match_reference(fake_device):
    """tests/host_parser_check.py: 41 expressions against the reference's FieldFactory (precedence quirks, implicit
    product, UTF-8 symbols, all generators but ballooning) and the set-up of the unchanged hw.cxx / blob2d.cxx
    (initial profiles, metrics) against the reference build."""
    p = subprocess.run([sys.executable, os.path.join(HERE, "host_parser_check.py")], env=dict(os.environ, LD_PRELOAD=fake_device),
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


PORT = os.path.join(ROOT, "oracle", "_ref", "liboracle_port.so")


@pytest.fixture(scope="module")
def hw_results(fake_device):
    """cases.HW_INTEGRATOR_CASES in one child process; stencils, brackets, Delp2, LaplaceCyclic served by the oracle
    port behind the test double."""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "integrator_check.py")] + sorted(HW_INTEGRATOR_CASES),
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT), capture_output=True, text=True,
                       timeout=900)
    out = {}
    for line in p.stdout.splitlines():
        if line.startswith("{"):
            r = json.loads(line)
            out[r["tag"]] = r
    assert out, p.stdout[-2000:] + p.stderr[-2000:]
    return out


@pytest.mark.parametrize("tag", sorted(HW_INTEGRATOR_CASES))
def test_unchanged_hw_model_with_the_other_integrators_host_logic(hw_results, tag):
    """examples/hasegawa-wakatani/hw.cxx (split-operator, Arakawa brackets, Delp2, LaplaceCyclic per RHS) stepped by
    euler / rkgeneric / splitrk of the host layer: final state vs the reference's own solver run, same RHS count."""
    res = hw_results[tag]
    assert "error" not in res, res
    assert res["ok"] and res["rhs_calls"] == res["ref_rhs_calls"], res


def test_unchanged_hw_and_blob2d_through_the_host_layer_host_logic(fake_device):
    """tests/host_models_check.py: the checks of tests/test_gpu_models.py (one run_rhs at 1e-10, 5-step RK4 / RK3-SSP
    trajectories at 1e-8 against tests/golden/models.npz) on CPU -- options, mesh, Laplacian flags, boundary
    conditions, method dispatch and split-operator handling of the host layer with the device replaced by the
    oracle port."""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "host_models_check.py")],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT), capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_single_index_loop_generic_dispatch_host_logic(fake_device):
    """tests/si_loop_check.py on CPU: the recorder behind BOUT_FOR_RAJA / FieldAccessor / single_index_ops.hxx
    (NaN-boxed intermediates, += assignments, integer constants) and its one-kernel-per-operator dispatch give the
    same ddt as the whole-field operators."""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "si_loop_check.py")],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


@pytest.mark.parametrize("lazy", ["1", "0"])
def test_deferred_fused_field_algebra_host_logic(fake_device, lazy):
    """tests/lazy_ops_check.py on CPU: the host layer's deferred, fused Field3D algebra (absorbed combinations, an operand
    written in place after an expression that reads it was built, pending multipliers / divisors, more than eight
    terms) and the one-launch-per-operator mode give the element-wise formulas of the statements, in statement order."""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "lazy_ops_check.py")],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT, BOUT_B200_LAZY=lazy),
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_vectors_grad_where_through_the_host_layer_host_logic(fake_device):
    """tests/vec_ops_check.py on CPU: the host layer's Vector2D / Vector3D, Grad, dot products (components / through the
    metric), toContravariant, where and interp_to against the same model source run by the reference build on a
    non-orthogonal metric."""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "vec_ops_check.py")],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_parallel_operators_and_flux_derivatives_host_logic(fake_device):
    """tests/par_ops_check.py on CPU: the host layer's Grad_par, Vpar_Grad_par, Div_par, Grad2_par2 (identity parallel
    transform), VDDY and FDDX / FDDY / FDDZ against the same model source run by the reference build on a mesh whose
    g_22, dy and Bxy vary.  (On the device these are bout_b200_stencil calls in the Y direction and element-wise passes,
    both covered by the -m gpu operator tests.)"""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "par_ops_check.py")],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_average_y_through_the_host_layer_host_logic(fake_device):
    """tests/avg_ops_check.py on CPU: the host layer's averageY(Field3D) (bout_b200_ymean + bout_b200_yspread),
    averageY(Field2D), DC and smooth_x against the same model source run by the reference build (tests/golden/avg_ops.npz)."""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "avg_ops_check.py")],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_fieldperp_through_the_host_layer_host_logic(fake_device):
    """tests/perp_ops_check.py on CPU: the host layer's FieldPerp (sliceXZ, algebra, element access,
    Laplacian::solve(FieldPerp), Field3D = FieldPerp) against the same model source run by the reference build."""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    p = subprocess.run([sys.executable, os.path.join(HERE, "perp_ops_check.py")],
                       env=dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


@pytest.mark.parametrize("direction,ranks", [("y", 2), ("y", 4), ("x", 2), ("x", 4), ("xy", 4)])
def test_host_layer_mesh_decomposed_over_gloo_ranks(fake_device, direction, ranks):
    """tests/host_ydecomp_check.py, world_size 2 and 4 over gloo: the C++ Mesh with NYPE, NXPE or 2 x 2 ranks (global index in
    profiles and metric, guard cells of fields and of the metric from the neighbouring rank through
    bout::b200::YDecomposition, x boundary conditions on the end ranks only, the X-decomposed LaplaceCyclic solve as
    reduce -> all-gather -> finish, error norms of the adaptive integrators reduced over the ranks) -- unchanged
    hasegawa-wakatani-3d/hw.cxx, conduction.cxx (y) and blob2d.cxx (x) against the reference build's single-rank run,
    slab by slab.  (In the test double the distributed elimination is emulated: see fake_b200.c.)"""
    if not os.path.exists(PORT):
        pytest.skip("oracle/_ref/liboracle_port.so not built")
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(ranks):
        env = dict(os.environ, LD_PRELOAD=fake_device, B200_FAKE_ORACLE_PORT=PORT, RANK=str(r), WORLD_SIZE=str(ranks),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "host_ydecomp_check.py"), direction], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=900))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o[-3000:] + e[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["ok"] and res["ranks"] == ranks and res["direction"] == direction, res
    assert res["adaptive_rhs_calls"] == [res["adaptive_rhs_calls_single"]] * ranks, res
    if direction == "x":
        assert res["blob2d_gathers_per_rhs"] == 1 and res["hw3d_gathers_per_rhs"] == 1, res
