"""tests/models/avg_ops.cxx through the C++ host layer (averageY of Field3D and Field2D, DC, smooth_x: smoothing.cxx:46-73, 238-323)
against the SAME source run by the reference build (tests/golden/avg_ops.npz, made by tests/golden/make_golden_avg.py)
on the mesh of cases.OPS_MESH_INI with the input fields of vec_ops.npz.  Runs on the GPU, or on CPU with the test double
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
    gold = dict(np.load(os.path.join(HERE, "golden", "vec_ops.npz")))
    gold.update(np.load(os.path.join(HERE, "golden", "avg_ops.npz")))
    h = ge.load_package().host.Host(cases.OPS_MESH_INI)
    try:
        h.model_create("avg_ops", "rk4")
        h.model_set("f", gold["f0"])
        h.model_set("g", gold["g0"])
        h.model_rhs(0.0)
        df, dg = h.model_ddt("f"), h.model_ddt("g")
    finally:
        h.close()
    nx, ny, nz = df.shape
    I = (slice(2, nx - 2), slice(2, ny - 2), slice(None))
    ef = float(np.abs(df[I] - gold["avg_ddt_f"][I]).max() / np.abs(gold["avg_ddt_f"][I]).max())
    eg = float(np.abs(dg[I] - gold["avg_ddt_g"][I]).max() / np.abs(gold["avg_ddt_g"][I]).max())
    ok = ef < 1e-12 and eg < 1e-12 and float(np.abs(gold["avg_ddt_f"][I]).max()) > 0.1
    print(json.dumps({"rel_err_ddt_f": ef, "rel_err_ddt_g": eg, "ok": bool(ok)}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
