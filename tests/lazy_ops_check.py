"""tests/models/lazy_ops.cxx through the C++ host layer against the element-wise formulas in numpy: the deferred,
fused Field3D algebra (and, with BOUT_B200_LAZY=0, one launch per operator) must give what the statements say, in
statement order.  Runs on the GPU, or on CPU with the test double (LD_PRELOAD=<libfake_b200.so>).  Prints one JSON line;
exit code 0 iff ok."""
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
    pkg = ge.load_package()
    h = pkg.host.Host(cases.HW3D_INI)
    try:
        h.model_create("lazy_ops", "rk4")
        rng = np.random.default_rng(11)
        a0, b0 = rng.uniform(-1, 1, h.shape3), rng.uniform(-1, 1, h.shape3)
        h.model_set("a", a0)
        h.model_set("b", b0)
        l0 = pkg.capi.launch_count() if hasattr(pkg.capi, "launch_count") else 0
        h.model_rhs(0.0)
        da, db = h.model_ddt("a"), h.model_ddt("b")
        launches = (pkg.capi.launch_count() - l0) if hasattr(pkg.capi, "launch_count") else -1
        a_after = h.model_get("a")
    finally:
        h.close()
    nx, ny, nz = a0.shape
    x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    c = 0.1 * x + 0.2 * y + 0.05 * z + 1.0
    m2 = (1.0 + 0.01 * x + 0.02 * y)
    a1 = a0.copy()
    a1[3, 2, 1] = 7.5
    t = a0 * b0
    u = t + 2.0 * a0 - b0 / m2 + 1.5
    v = u * c / (a1 + 4.0)
    w = v - u * u
    s = a1 + b0 + c + t + u + v + w + a1 * c + b0 * c + 2.0
    wa = s / 3.0 - (1.0 - w)
    wb = -wa + m2 * u
    # interior only: Solver::run_rhs applies the boundary conditions to the guard cells of a and b before rhs()
    I = (slice(2, nx - 2), slice(2, ny - 2), slice(None))
    ea = float(np.abs(da[I] - wa[I]).max() / np.abs(wa[I]).max())
    eb = float(np.abs(db[I] - wb[I]).max() / np.abs(wb[I]).max())
    ok = ea < 1e-13 and eb < 1e-13 and a_after[3, 2, 1] == 7.5 and float(np.abs(a_after[I] - a1[I]).max()) == 0.0
    print(json.dumps({"rel_err_ddt_a": ea, "rel_err_ddt_b": eb, "launches": launches, "ok": bool(ok),
                      "lazy": os.environ.get("BOUT_B200_LAZY", "1")}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
