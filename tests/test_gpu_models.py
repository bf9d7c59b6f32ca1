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
