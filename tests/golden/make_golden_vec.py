"""Golden values of tests/models/vec_ops.cxx (Vector2D / Vector3D, Grad, dot products, toContravariant, where,
interp_to) from the REFERENCE build (oracle/_ref/libboutref.so: the same source compiled against the reference's own
headers and sources) on the non-orthogonal, x- and y-dependent metric of cases.OPS_MESH_INI.  Writes vec_ops.npz.
Run in the build container (needs /root/reference for oracle/_ref); the GPU box only reads the .npz."""
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
r = orc.Ref(cases.OPS_MESH_INI)
r.model_create("vec_ops", "rk4")
rng = np.random.default_rng(23)
shape = r.model_get("f").shape
nx, ny, nz = shape
x, y, z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
# smooth fields with some zero crossings (the where() test on f) plus noise
f0 = np.sin(0.7 * x + 0.3 * y) * np.cos(2 * np.pi * z / nz) + 0.1 * rng.uniform(-1, 1, shape)
g0 = np.cos(0.4 * x - 0.5 * y) + 0.5 * np.sin(4 * np.pi * z / nz + 0.2 * x) + 0.1 * rng.uniform(-1, 1, shape)
r.model_set("f", f0)
r.model_set("g", g0)
r.model_rhs(0.0)
out = {"f0": f0, "g0": g0, "ddt_f": r.model_ddt("f"), "ddt_g": r.model_ddt("g")}
r.close()
# tests/models/par_ops.cxx (parallel operators, VDDY, flux derivatives) on the same mesh and fields
r = orc.Ref(cases.OPS_MESH_INI)
r.model_create("par_ops", "rk4")
r.model_set("f", f0)
r.model_set("g", g0)
r.model_rhs(0.0)
out["par_ddt_f"], out["par_ddt_g"] = r.model_ddt("f"), r.model_ddt("g")
r.close()
# the reference leaves the guard cells of a time derivative as its allocation found them: keep the fixture deterministic
for k in ("ddt_f", "ddt_g", "par_ddt_f", "par_ddt_g"):
    a = np.zeros_like(out[k])
    a[2:-2, 2:-2, :] = out[k][2:-2, 2:-2, :]
    out[k] = a
np.savez_compressed(os.path.join(HERE, "vec_ops.npz"), **out)
print("wrote vec_ops.npz", {k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
