"""tests/models/si_loop.cxx through the C++ host layer: a BOUT_FOR_RAJA single-index loop that is not the fused
hasegawa-wakatani-3d pattern (boxed intermediate, `+=`, integer constants, a direct accessor read) against the same
expression written with whole-field operators in the same model.  Runs on the GPU, or on CPU with the test double
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
    hostmod = ge.load_package().host
    h = hostmod.Host(cases.HW3D_INI)
    try:
        h.model_create("si_loop", "rk4")
        rng = np.random.default_rng(3)
        h.model_set("a", rng.uniform(-1, 1, h.shape3))
        h.model_set("b", np.zeros(h.shape3))
        h.model_rhs(0.0)
        da, db = h.model_ddt("a"), h.model_ddt("b")
        nx, ny, nz = h.shape3
        I = (slice(2, nx - 2), slice(2, ny - 2), slice(None))
        err = float(np.abs(da[I] - db[I]).max() / np.abs(db[I]).max())
    finally:
        h.close()
    ok = err < 1e-13 and float(np.abs(db[I]).max()) > 1.0
    print(json.dumps({"rel_err": err, "ok": ok}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
