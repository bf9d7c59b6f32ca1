"""Host-side set-up logic of the C++ host layer against the reference (run with LD_PRELOAD=<test double> on CPU):
 * the expression evaluator on cases.EXPRESSIONS against tests/golden/expressions.npz (reference FieldFactory);
 * the options-file reader (comments, multi-line bracketed values, bare flags, case-sensitive names) through option
   references against tests/golden/options_probe.json;
 * Coordinates (covariant metric, J, Bxy, G1..G3, d1_dx, guard cells) of a general metric against ops_small.npz;
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
    # Coordinates of a non-orthogonal, x- and y-dependent metric against the reference's (tests/golden/ops_small.npz)
    ops = np.load(os.path.join(HERE, "golden", "ops_small.npz"))
    h = hostmod.Host(cases.OPS_MESH_INI)
    try:
        nx, ny, nz, mxg, myg = [int(v) for v in ops["grid"]]
        nocorners = np.ones((nx, ny), bool)
        for xs in (slice(0, mxg), slice(nx - mxg, nx)):
            for ys in (slice(0, myg), slice(ny - myg, ny)):
                nocorners[xs, ys] = False   # corner cells of derived quantities are never written by the reference
        for n in ("dx", "dy", "dz", "g11", "g22", "g33", "g12", "g13", "g23", "g_11", "g_22", "g_33", "g_12", "g_13", "g_23",
                  "J", "Bxy", "G1", "G2", "G3", "d1_dx", "d1_dy"):
            a, b = h.metric(n), ops["metric_" + n]
            if np.abs(a - b)[nocorners].max() > 1e-13 * max(1.0, np.abs(b).max()):
                bad.append(f"metric {n}: {np.abs(a - b)[nocorners].max():.3e}")
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
