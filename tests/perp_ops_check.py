"""tests/models/perp_ops.cxx through the C++ host layer (FieldPerp: sliceXZ, algebra with FieldPerp / BoutReal / Field3D
operands, compound assignment, host-side element access, Laplacian::solve(FieldPerp, FieldPerp), Field3D = FieldPerp)
against the SAME source run by the reference build (tests/golden/invert.npz: host_perp_ddt_f / _g, made by
tests/golden/make_golden_invert.py).  Runs on the GPU, or on CPU with the test double
(LD_PRELOAD=<libfake_b200.so> B200_FAKE_ORACLE_PORT=<liboracle_port.so>).  Prints one JSON line; exit code 0 iff ok."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))


def main():
    import __graft_entry__ as ge
    import cases
    gold = dict(np.load(os.path.join(HERE, "golden", "invert.npz")))
    hostmod = ge.load_package().host
    h = hostmod.Host(cases.INVERT_LAP_INI)
    try:
        h.model_create("perp_ops", "rk4")
        h.model_set("f", gold["lap_b"])
        h.model_set("g", gold["lap_x0"])
        h.model_rhs(0.0)
        df, dg = h.model_ddt("f"), h.model_ddt("g")
    finally:
        h.close()
    wf, wg = gold["host_perp_ddt_f"], gold["host_perp_ddt_g"]
    jy = 3
    # the algebra is the same sequence of IEEE operations on both sides; the solve is an FFT + tridiagonal solve
    e_alg = float(np.abs(df[:, jy] - wf[:, jy]).max() / np.abs(wf[:, jy]).max())
    e_sol = float(np.abs(dg[:, jy] - wg[:, jy]).max() / np.abs(wg[:, jy]).max())
    others = float(max(np.abs(np.delete(df, jy, axis=1)).max(), np.abs(np.delete(dg, jy, axis=1)).max()))
    ok = e_alg < 1e-14 and e_sol < 1e-10 and others == 0.0 and float(np.abs(wf).max()) > 1.0
    print(json.dumps({"algebra_rel_err": e_alg, "solve_rel_err": e_sol, "other_planes_max": others, "ok": ok}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
