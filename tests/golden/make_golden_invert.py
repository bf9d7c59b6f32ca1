"""Golden vectors for SURVEY 8(f)4, produced by the reference's own solvers (oracle/_ref/libboutref.so):
Laplacian types pcr and pcr_thomas next to cyclic, LaplaceXZ cyclic, InvertPar cyclic.  Writes tests/golden/invert.npz.
Run in the build container (needs /root/reference to have been compiled by oracle/Makefile)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import __graft_entry__ as ge  # noqa: E402
import cases  # noqa: E402

orc = ge.load_oracle()
out = {}
rng = np.random.default_rng(20261018)

# ---- Laplacian: pcr / pcr_thomas (src/invert/laplace/impls/pcr/pcr.cxx, impls/pcr_thomas/pcr_thomas.cxx) -------------
NX, NYI, NZ = cases.INVERT_GRID
r = orc.Ref(cases.INVERT_LAP_INI)
nx, ny, nz = NX, NYI + 4, NZ
b = rng.uniform(-1, 1, (nx, ny, nz))
x0 = rng.uniform(-1, 1, (nx, ny, nz))
A = 0.3 + 0.1 * rng.uniform(size=(nx, ny))
Cc = 1.0 + 0.2 * rng.uniform(size=(nx, ny))
out["lap_b"], out["lap_x0"], out["lap_A"], out["lap_C"] = b, x0, A, Cc
met = r.metric()
for name in orc.Metric.NAMES:
    out["metric_" + name] = getattr(met, name)
for i, flags in enumerate(cases.INVERT_LAP_FLAGS):
    for sec in ("lap_cyclic", "lap_pcr", "lap_pcr_thomas"):
        r.laplace_create(sec)
        r.laplace_set_flags(*flags)
        r.laplace_set_coef("A", A)
        r.laplace_set_coef("C", Cc)
        out[f"{sec}_{i}"] = r.laplace_solve(b, x0)

# LaplaceCyclic::solve(FieldPerp) (cyclic_laplace.cxx:120-185) on two y slices, Dirichlet and INVERT_SET + Neumann flags
for i, flags in enumerate(((0, 0, 0), (0, 16, 3))):
    r.laplace_create("lap_cyclic")
    r.laplace_set_flags(*flags)
    r.laplace_set_coef("A", A)
    r.laplace_set_coef("C", Cc)
    for jy in (3, 6):
        out[f"perp_{i}_{jy}"] = r.laplace_solve_perp(jy, b, x0)

# ---- LaplaceXZ cyclic (src/invert/laplacexz/impls/cyclic/laplacexz-cyclic.cxx) ------------------------------------
A2 = 1.0 + 0.3 * rng.uniform(size=(nx, ny))
B2 = -0.5 - 0.2 * rng.uniform(size=(nx, ny))
out["xz_A"], out["xz_B"] = A2, B2
for i, (inner, outer) in enumerate(cases.INVERT_XZ_FLAGS):
    out[f"xz_{i}"] = r.laplacexz_solve(f"xz{i}", A2, B2, b, x0)
r.close()

# ---- InvertPar cyclic (src/invert/parderiv/impls/cyclic/cyclic.cxx): closed field lines (y periodic) ----------------
r = orc.Ref(cases.INVERT_LAP_INI)
pa = 1.0 + 0.1 * rng.uniform(size=(nx, ny))
pb = -0.4 - 0.1 * rng.uniform(size=(nx, ny))
pc = 0.05 * rng.uniform(size=(nx, ny))
pd = -0.02 * rng.uniform(size=(nx, ny))
pe = 0.03 * rng.uniform(size=(nx, ny))
out["par_A"], out["par_B"], out["par_C"], out["par_D"], out["par_E"] = pa, pb, pc, pd, pe
out["par_closed_AB"] = r.invert_par_solve("parderiv", b, A=pa, B=pb)
out["par_closed_all"] = r.invert_par_solve("parderiv", b, A=pa, B=pb, Cc=pc, D=pd, E=pe)

# the calls of tests/models/invert_ops.cxx (coefficients from the metric), for the host-layer test
out["host_par"] = r.invert_par_solve("parderiv", b, A=met.g33, B=-0.4 * met.g11)
out["host_xz"] = r.laplacexz_solve("xz1", met.g11, -met.g33, b, x0)
r.laplace_create("lap_pcr")
r.laplace_set_flags(0, 2, 0)
r.laplace_set_coef("A", 0.3 * met.g33)
out["host_pcr"] = r.laplace_solve(b, x0)
# tests/models/perp_ops.cxx (FieldPerp algebra, element access, LaplaceCyclic::solve(FieldPerp), Field3D = FieldPerp)
# run by the reference build: one right-hand side on (f, g) = (b, x0)
r.model_create("perp_ops", "rk4")
r.model_set("f", b)
r.model_set("g", x0)
r.model_rhs(0.0)
out["host_perp_ddt_f"], out["host_perp_ddt_g"] = r.model_ddt("f"), r.model_ddt("g")
r.close()

# ---- the same on open field lines (boundary rows b = 1, c = -1 / a = -1, b = 1 in the y guard cells) ----------------
r = orc.Ref(cases.INVERT_OPEN_INI)
out["par_open_all"] = r.invert_par_solve_all_y("parderiv", b, A=pa, B=pb, Cc=pc, D=pd, E=pe)
met = r.metric()
out["open_metric_g_22"], out["open_metric_dy"], out["open_metric_d1_dy"] = met.g_22, met.dy, met.d1_dy
r.close()

# Cells a reference solver leaves unwritten hold whatever its emptyFrom() allocation held: arrays already in the committed
# file are kept as they are (the tests only read the written cells), new keys are added.
path = os.path.join(HERE, "invert.npz")
if os.path.exists(path):
    old = dict(np.load(path))
    out = {**out, **{k: v for k, v in old.items() if k in out and v.shape == out[k].shape}}
np.savez_compressed(path, **out)
print("wrote invert.npz:", sorted(out)[:8], "...", len(out), "arrays")
