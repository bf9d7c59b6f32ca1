"""CPU: the set-up-time restatement of ELMpb::init + Coordinates::geometry (bout-dev_b200/elm_setup.py)
against the reference's own post-init state recorded in tests/golden/elmpb.npz (generated from the
unchanged examples/elm-pb/elm_pb.cxx by tests/golden/make_golden.py)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load_setup():
    spec = importlib.util.spec_from_file_location("elm_setup", os.path.join(ROOT, "bout-dev_b200", "elm_setup.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_setup_matches_reference_init():
    gold = dict(np.load(os.path.join(HERE, "golden", "elmpb.npz")))
    nx, ny, nz, mxg, myg = [int(v) for v in gold["grid"]]
    s = load_setup().elmpb_setup(nx, ny - 2 * myg, nz, mxg, myg)
    assert s["grid"] == (nx, ny, nz, mxg, myg)

    def rel(a, b, region):
        sc = np.abs(b[region]).max()
        d = np.abs(a[region] - b[region]).max()
        return d / sc if sc > 0 else d

    alls = (slice(None), slice(None))
    interior = (slice(mxg, nx - mxg), slice(myg, ny - myg))
    for k, v in s["md"].items():
        # G1, G3 guard cells are extrapolated by the reference (coordinates.cxx:1066-1068) and never read
        # by the RHS kernels: compared on the interior, everything else on all cells
        region = interior if k in ("G1", "G3") else alls
        assert rel(v, gold["metric_" + k], region) < 1e-13, k
    for k, v in s["eq"].items():
        assert rel(v, gold["eq_" + k], alls if v.ndim == 2 else (slice(None),) * 3) < 1e-13, k
    assert rel(s["zShift"], gold["zShift"], alls) < 1e-14
    assert abs(s["zlength"] - float(gold["zlength"])) < 1e-15
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from elmpb_names import ELMPB_SCALARS
    sc = dict(zip(ELMPB_SCALARS, gold["scalars"]))
    assert abs(s["dnorm"] / sc["dnorm"] - 1) < 1e-14
    assert s["hyperresist"] == sc["hyperresist"]


def test_golden_psi_loc_is_psi():
    """elm_pb.cxx:1752-1753: Psi_loc = interp_to(Psi, CELL_CENTRE) + the same zerolaplace condition again
    is bitwise Psi -- the harness model relies on it."""
    gold = dict(np.load(os.path.join(HERE, "golden", "elmpb.npz")))
    assert np.array_equal(gold["rhs_Psi_loc"], gold["rhs_bc_Psi"])
