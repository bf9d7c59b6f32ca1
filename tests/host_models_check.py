"""The unchanged examples/hasegawa-wakatani/hw.cxx, examples/blob2d/blob2d.cxx and examples/hasegawa-wakatani-3d/hw.cxx
(its BOUT_FOR_RAJA single-index loop recorded and dispatched to the fused kernel) through the C++ host layer against
the reference build (tests/golden/models.npz): initial profiles, one Solver::run_rhs, a 5-step RK trajectory -- the
checks of tests/test_gpu_models.py::test_unchanged_model_through_host_layer, runnable on CPU with the test double
(LD_PRELOAD=<libfake_b200.so> B200_FAKE_ORACLE_PORT=<liboracle_port.so>), where they pin the HOST logic: options,
mesh, Laplacian flags and coefficients, boundary conditions, derivative-method dispatch, split-operator RHS, RK4 /
RK3-SSP.  Prints one JSON line; exit code 0 iff all pass."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))


def rel(a, b, I):
    return float(np.abs(a[I] - b[I]).max() / max(np.abs(b[I]).max(), 1e-300))


def main():
    import __graft_entry__ as ge
    import cases
    hostmod = ge.load_package().host
    gold = np.load(os.path.join(HERE, "golden", "models.npz"))
    res, bad = {}, []
    for tag, model, solver, ini in (("hw", "hw", "rk4", cases.HW_INI), ("hw_rk3", "hw", "rk3ssp", cases.HW_INI),
                                    ("blob2d", "blob2d", "rk4", cases.BLOB_INI),
                                    ("hw3d", "hw3d", "rk4", cases.HW3D_INI),
                                    ("conduction", "conduction", "rk4", cases.CONDUCTION_INI),
                                    ("blob2d_outer", "blob2d_outer", "rk4", cases.BLOB_INI)):
        h = hostmod.Host(ini)
        try:
            names = h.model_create(model, solver)
            nx, ny, nz, mxg, myg = [int(v) for v in gold[tag + "_grid"]]
            I = (slice(mxg, nx - mxg), slice(myg, ny - myg), slice(None))
            h.model_rhs(0.0)
            for v in names:
                res[f"{tag}_ddt_{v}"] = rel(h.model_ddt(v), gold[f"{tag}_ddt_{v}"], I)
                if res[f"{tag}_ddt_{v}"] > 1e-10:
                    bad.append(f"{tag} ddt {v}")
            dt, nsteps = gold[tag + "_dt_nsteps"]
            for v in names:
                h.model_set(v, gold[f"{tag}_init_{v}"])
            h.model_solve(1, float(dt) * int(nsteps))
            for v in names:
                res[f"{tag}_final_{v}"] = rel(h.model_get(v), gold[f"{tag}_final_{v}"], I)
                if res[f"{tag}_final_{v}"] > 1e-8:
                    bad.append(f"{tag} final {v}")
        finally:
            h.close()
    res.update(ok=not bad, bad=bad)
    print(json.dumps(res), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
