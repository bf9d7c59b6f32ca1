"""Host-side set-up logic of the C++ host layer against the reference (run with LD_PRELOAD=<test double> on CPU):
 * the expression evaluator on cases.EXPRESSIONS against tests/golden/expressions.npz (reference FieldFactory);
 * the options-file reader (comments, multi-line bracketed values, bare flags, case-sensitive names) through option
   references against tests/golden/options_probe.json;
 * Options -> Mesh -> model init() -> Solver::add of the unchanged hw.cxx / blob2d.cxx: initial profiles and metrics
   against tests/golden/models.npz.
Prints one JSON line; exit code 0 iff everything matches."""
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
    bad = []
    ref = np.load(os.path.join(HERE, "golden", "expressions.npz"))
    h = hostmod.Host(cases.EXPR_INI)
    try:
        for i, e in enumerate(cases.EXPRESSIONS):
            try:
                a, b = h.expr3d(e), ref[str(i)]
                ok = np.array_equal(np.isnan(a), np.isnan(b)) and np.nanmax(np.abs(a - b)) <= 1e-13 * max(1.0, np.nanmax(np.abs(b)))
            except Exception as ex:  # noqa: BLE001
                ok = False
                e = f"{e}: {ex}"
            if not ok:
                bad.append(e)
    finally:
        h.close()
    probe = json.load(open(os.path.join(HERE, "golden", "options_probe.json")))
    h = hostmod.Host(cases.OPTIONS_PROBE_INI)
    try:
        for e in cases.OPTIONS_PROBE_EXPRS:
            try:
                v = float(h.expr3d(e)[2, 2, 1])
            except hostmod.HostError:
                v = "ERR"
            if v != probe[e]:
                bad.append(f"option reference {e}: {v} (reference {probe[e]})")
    finally:
        h.close()
    gold = np.load(os.path.join(HERE, "golden", "models.npz"))
    for tag, model, ini in (("hw", "hw", cases.HW_INI), ("blob2d", "blob2d", cases.BLOB_INI)):
        h = hostmod.Host(ini)
        try:
            names = h.model_create(model, "rk4")
            nx, ny, nz, mxg, myg = [int(v) for v in gold[tag + "_grid"]]
            I = (slice(mxg, nx - mxg), slice(myg, ny - myg), slice(None))
            for v in names:
                if np.abs(h.model_get(v)[I] - gold[f"{tag}_init_{v}"][I]).max() > 1e-13:
                    bad.append(f"{tag} initial {v}")
            for n in ("dx", "dz", "g11", "g33", "J", "g_22"):
                if not np.allclose(h.metric(n), gold[f"{tag}_metric_{n}"], rtol=1e-14, atol=0):
                    bad.append(f"{tag} metric {n}")
        finally:
            h.close()
    print(json.dumps({"ok": not bad, "bad": bad, "expressions": len(cases.EXPRESSIONS)}), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
