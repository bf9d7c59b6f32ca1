"""One integrator case of tests/golden/cases.INTEGRATOR_CASES through the C++ host layer (bout-dev_b200/host):
tests/models/ode_pair.cxx, compiled against host/include/bout, stepped by Solver::create(type) and compared with the
reference's own solver on the same model (tests/golden/integrators.npz).  Run as a child process:

    python tests/integrator_check.py <tag> [<tag> ...]          real device (GPU tests)
    LD_PRELOAD=tests/fake_device/libfake_b200.so python ...     host-memory test double (CPU: host logic only)

Prints one JSON line per tag {"tag", "err_f", "err_g", "rhs_calls", "ref_rhs_calls"} (or {"tag", "error"});
exit code 0 iff every case is within tolerance.
Tolerance: 1e-10 relative on the final state of the algebra-only model, 1e-8 (north_star: short trajectories) for
cases.HW_INTEGRATOR_CASES (unchanged hasegawa-wakatani model, LaplaceCyclic inside the RHS), and the SAME number of
RHS evaluations, i.e. the adaptive controllers took the same accept / reject decisions."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))
TOL = 1e-10


def run(tag):
    import __graft_entry__ as ge
    from cases import ELEM_CASES, HW_INTEGRATOR_CASES, INTEGRATOR_CASES, hw_integrator_ini, ode_ini
    hostmod = ge.load_package().host
    gold = np.load(os.path.join(HERE, "golden", "integrators.npz"))
    if tag in HW_INTEGRATOR_CASES:   # unchanged examples/hasegawa-wakatani/hw.cxx (needs the real device)
        solver, opts, nout, step = HW_INTEGRATOR_CASES[tag]
        model, ini, tol = "hw", hw_integrator_ini(opts, nout, step), 1e-8
    elif tag == "imex_shipped":     # unchanged examples/IMEX/advection-diffusion/imex.cxx with its shipped input
        from cases import IMEX_INI, IMEX_NOUT, IMEX_STEP
        solver, nout, step = "splitrk", IMEX_NOUT, IMEX_STEP
        model, ini, tol = "imex", IMEX_INI, 1e-8
    elif tag == "advreact":          # unchanged examples/IMEX/advection-reaction/split_operator.cxx (BRACKET_SIMPLE)
        from cases import ADVREACT_INI
        solver, nout, step = "splitrk", 1, 0.01
        model, ini, tol = "advreact", ADVREACT_INI, 1e-8
    elif tag == "eigenbox_shipped":  # unchanged examples/eigen-box/eigen-box.cxx with its shipped input, rk4
        from cases import EIGENBOX_INI
        solver, nout, step = "rk4", 1, 0.2
        model, ini, tol = "eigenbox", EIGENBOX_INI, 1e-10
    elif tag in ELEM_CASES:         # tests/models/elem_access.cxx: host-side element access
        solver, opts, nout, step = ELEM_CASES[tag]
        model, ini, tol = "elem_access", ode_ini(opts, False, nout, step), TOL
    else:
        solver, opts, split, nout, step, *bndry = INTEGRATOR_CASES[tag]
        model, ini, tol = "ode_pair", ode_ini(opts, split, nout, step, *bndry), TOL
    h = hostmod.Host(ini)
    try:
        names = h.model_create(model, solver)
        nx, ny, nz = h.shape3
        mxg = h.info[3]
        I = (slice(mxg, nx - mxg), slice(None), slice(None))
        if tag.startswith("bc_"):   # X boundaries: guard cells are set by the boundary conditions and compared too
            I = (slice(None), slice(None), slice(None))
        res = {"tag": tag}
        for v in names:
            assert np.abs(h.model_get(v)[I] - gold[f"{tag}_init_{v}"][I]).max() < 1e-13, "initial profile " + v
        if tag == "imex_shipped":   # mesh expressions of the shipped input: dx = 1 / nx, dz = 1 / (nz-4)
            assert np.allclose(h.metric("dx"), 1.0 / 132, rtol=1e-15) and np.allclose(h.metric("dz"), 1.0 / 124, rtol=1e-15)
        if os.environ.get("B200_CHECK_INIT_ONLY") == "1":   # CPU: options, mesh and initial profile only
            res.update(ok=True, init_only=True, tol=tol)
            print(json.dumps(res), flush=True)
            return 0
        h.model_solve(nout, step)
        for v in names:
            a, b = h.model_get(v)[I], gold[f"{tag}_final_{v}"][I]
            res["err_" + v] = float(np.abs(a - b).max() / np.abs(b).max())
        res["rhs_calls"] = h.rhs_calls()
        res["ref_rhs_calls"] = int(gold[f"{tag}_rhs_calls"])
    finally:
        h.close()
    res["tol"] = tol
    ok = all(res["err_" + v] < tol for v in names) and res["rhs_calls"] == res["ref_rhs_calls"]
    res["ok"] = bool(ok)
    print(json.dumps(res), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    rc = 0
    for tag in sys.argv[1:]:
        try:
            rc |= run(tag)
        except Exception as e:  # noqa: BLE001 -- reported per case, the caller asserts on it
            print(json.dumps({"tag": tag, "error": f"{type(e).__name__}: {e}"}), flush=True)
            rc |= 1
    sys.exit(rc)
