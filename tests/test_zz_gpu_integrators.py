"""GPU: the explicit integrators of the C++ host layer through the real kernels (bout_b200_lincomb = k_elementwise,
bout_b200_err_sum = k_err_partial + k_err_final, bout_b200_rk_stage) against the reference's own solvers
(tests/golden/integrators.npz), plus the error-sum kernel against numpy.  Tolerance 1e-10 relative on the final
state; identical RHS-evaluation counts."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases import ELEM_CASES, HW_INTEGRATOR_CASES, INTEGRATOR_CASES  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def results(capi):
    env = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}
    p = subprocess.run([sys.executable, os.path.join(HERE, "integrator_check.py")] + sorted(INTEGRATOR_CASES)
                       + sorted(ELEM_CASES) + sorted(HW_INTEGRATOR_CASES) + ["imex_shipped", "advreact"], env=env,
                       capture_output=True, text=True, timeout=900)
    out = {}
    for line in p.stdout.splitlines():
        if line.startswith("{"):
            r = json.loads(line)
            out[r["tag"]] = r
    assert out, p.stdout[-2000:] + p.stderr[-2000:]
    return out


@pytest.mark.parametrize("tag", sorted(INTEGRATOR_CASES) + sorted(ELEM_CASES))
def test_integrator_on_device_matches_reference_solver(results, tag):
    res = results[tag]
    assert "error" not in res, res
    assert res["rhs_calls"] == res["ref_rhs_calls"], res
    assert res["ok"] and res["tol"] <= 1e-10, res


@pytest.mark.parametrize("tag", sorted(HW_INTEGRATOR_CASES))
def test_unchanged_hw_model_with_the_other_integrators(results, tag):
    """examples/hasegawa-wakatani/hw.cxx (split-operator model, Arakawa brackets + LaplaceCyclic per RHS) stepped by
    euler / rkgeneric / splitrk of the host layer against the reference's solvers: 1e-8 on the final state."""
    res = results[tag]
    assert "error" not in res, res
    assert res["rhs_calls"] == res["ref_rhs_calls"], res
    assert res["ok"] and res["tol"] <= 1e-8, res


def test_unchanged_imex_example_with_its_shipped_input(results):
    """examples/IMEX/advection-diffusion: imex.cxx and data/BOUT.inp as shipped (splitrk, nstages 9, adapt_period 5,
    W3 upwinding, C4 second derivative; nout reduced to 2) against the reference's own run: 1e-8 on the final state."""
    res = results["imex_shipped"]
    assert "error" not in res, res
    assert res["ok"] and res["tol"] <= 1e-8, res


def test_unchanged_advection_reaction_example_on_device(results):
    """examples/IMEX/advection-reaction/split_operator.cxx (bracket SIMPLE = VDDX(DDZ f, g) + VDDZ(-DDX f, g), W3 / C4,
    splitrk) on the device against the reference's run: 1e-8."""
    res = results["advreact"]
    assert "error" not in res, res
    assert res["ok"] and res["tol"] <= 1e-8, res


@pytest.mark.parametrize("shape,guards", [((12, 1, 8), (2, 0)), ((20, 9, 16), (2, 2)), ((68, 36, 130), (2, 2)),
                                          ((260, 68, 256), (2, 2)), ((7, 5, 3), (1, 1))])
def test_err_sum_and_lincomb_kernels(capi, shape, guards):
    """sum |a-b| / (|a|+|b|+atol) over RGN_NOBNDRY (rkscheme.cxx:150-156) and out = sum_j c_j x_j."""
    import torch
    nx, ny, nz = shape
    mxg, myg = guards
    g = capi.Grid(nx, ny, nz, mxg, myg)
    rng = np.random.default_rng(nx * 1000 + nz)
    a, b = rng.standard_normal(shape), rng.standard_normal(shape)
    a[0, 0, 0] = np.nan   # guard cells must not enter the sum (myg = 0: x guard)
    I = (slice(mxg, nx - mxg), slice(myg, ny - myg), slice(None))
    atol = 1e-3
    want = float((np.abs(a[I] - b[I]) / (np.abs(a[I]) + np.abs(b[I]) + atol)).sum())
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    got = capi.err_sum(g, da, db, atol)
    assert abs(got - want) <= 1e-12 * abs(want)
    assert capi.err_sum(g, da, db, atol) == got                      # fixed summation order
    assert abs(capi.err_sum(g, da, db, atol, accumulate_onto=got) - 2 * want) <= 1e-12 * abs(want)
    assert capi.err_max(g, da, db) == float(np.abs(a[I] - b[I]).max())   # max |a-b| (adams_bashforth.cxx:229-253): exact
    assert capi.err_max(g, da, db, accumulate_onto=1e9) == 1e9
    assert capi.err_max(g, da, da.clone()) == 0.0 or np.isnan(a[I]).any()
    xs = [torch.from_numpy(rng.standard_normal(shape)).cuda() for _ in range(8)]
    cs = [float(c) for c in rng.standard_normal(8)]
    for n in (1, 3, 8):
        out = torch.full(shape, 7.0, dtype=torch.float64, device="cuda")
        capi.lincomb(g, cs[:n], xs[:n], out)
        ref = np.zeros(shape)
        for c, x in zip(cs[:n], xs[:n]):
            ref = ref + c * x.cpu().numpy() if ref.any() else c * x.cpu().numpy()
        o = out.cpu().numpy()
        assert np.abs(o[I] - ref[I]).max() <= 2e-15 * n * max(1.0, np.abs(ref[I]).max())
        mask = np.ones(shape, bool)
        mask[I] = False
        assert (o[mask] == 7.0).all()                                # guards untouched
    with pytest.raises(capi.BoutB200Error):
        capi.lincomb(g, [1.0, 2.0], [xs[0], xs[1]], xs[1])           # out must not alias an operand


@pytest.mark.parametrize("mode", ["ode_pair", "hw"])
def test_output_step_on_device(capi, tmp_path, mode):
    """Dump snapshots leave the device on a second stream (page-locked staging) while the solver keeps stepping:
    every record must equal the state at its output time (tests/output_check.py)."""
    env = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}
    args = [sys.executable, os.path.join(HERE, "output_check.py"), str(tmp_path)] + (["hw"] if mode == "hw" else [])
    p = subprocess.run(args, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


def test_free_boundaries_on_device(capi):
    """free_o2 / free_o3 (kinds 4, 5 of bout_b200_boundary) on the x sides against the reference build's output
    (tests/golden/bndry_free.npz) and on the y sides against the closed form (boundary_standard.cxx:3281-3293,
    :3545-3558: every guard cell, outwards, from the 2 / 3 cells inside it)."""
    import torch
    ops = np.load(os.path.join(HERE, "golden", "ops_small.npz"))
    gold = np.load(os.path.join(HERE, "golden", "bndry_free.npz"))
    nx, ny, nz, mxg, myg = [int(v) for v in ops["grid"]]
    g = capi.Grid(nx, ny, nz, mxg, myg)
    for kind in ("free_o2", "free_o3"):
        F = torch.from_numpy(np.ascontiguousarray(ops["f"])).cuda()
        capi.boundary(g, kind, "xin", F)
        capi.boundary(g, kind, "xout", F)
        assert np.abs(F.cpu().numpy() - gold["bndry_" + kind]).max() <= 5e-15, kind
        F = torch.from_numpy(np.ascontiguousarray(ops["f"])).cuda()
        capi.boundary(g, kind, "ydown", F)
        capi.boundary(g, kind, "yup", F)
        w = ops["f"].copy()
        for i in range(myg):
            lo, hi = myg - 1 - i, ny - myg + i
            if kind == "free_o2":
                w[:, lo] = 2 * w[:, lo + 1] - w[:, lo + 2]
                w[:, hi] = 2 * w[:, hi - 1] - w[:, hi - 2]
            else:
                w[:, lo] = 3.0 * w[:, lo + 1] - 3.0 * w[:, lo + 2] + w[:, lo + 3]
                w[:, hi] = 3.0 * w[:, hi - 1] - 3.0 * w[:, hi - 2] + w[:, hi - 3]
        assert np.abs(F.cpu().numpy() - w).max() <= 5e-15, kind + " y"
    tiny = capi.Grid(6, 6, 4, 2, 2)   # 2 interior cells: free_o3 needs 3
    with pytest.raises(capi.BoutB200Error):
        capi.boundary(tiny, "free_o3", "xin", torch.zeros(6, 6, 4, dtype=torch.float64, device="cuda"))


def test_average_y_through_the_host_layer():
    """tests/models/avg_ops.cxx on the device (tests/avg_ops_check.py): averageY(Field3D) = bout_b200_ymean +
    bout_b200_yspread, averageY(Field2D), DC and smooth_x of the host layer against the reference build's run of the same source
    (1e-12)."""
    p = subprocess.run([sys.executable, os.path.join(HERE, "avg_ops_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert json.loads(p.stdout.strip().splitlines()[-1])["ok"]


@pytest.mark.gpu
@pytest.mark.parametrize("direction,ranks", [("y", 2), ("y", 4), ("x", 2), ("x", 4), ("xy", 4)])
def test_host_layer_mesh_decomposed_over_nccl_ranks(direction, ranks):
    """tests/host_ydecomp_check.py on N GPUs (NCCL P2P between the pack and unpack steps of Mesh::communicate; NXPE: one
    all-gather between bout_b200_laplace_cyclic_reduce and _finish): the unchanged hasegawa-wakatani-3d/hw.cxx,
    conduction.cxx and blob2d.cxx through the C++ host layer with NYPE = N, NXPE = N or NXPE x NYPE = 2 x 2 against the reference build's
    single-rank run.  Skipped on a one-GPU box; the gloo run of tests/test_host_integrators.py covers the host logic on
    CPU."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr",
           "127.0.0.1", "--master-port", str(29800 + ranks + {"y": 0, "x": 10, "xy": 20}[direction]),
           os.path.join(HERE, "host_ydecomp_check.py"), direction]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
