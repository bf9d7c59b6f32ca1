"""Golden values of tests/models/avg_ops.cxx (averageY of Field3D / Field2D, DC, smooth_x) from the REFERENCE build
(oracle/_ref/libboutref.so: the same source compiled against the reference's own headers and sources) on the mesh of
cases.OPS_MESH_INI, with the input fields of vec_ops.npz.  Writes avg_ops.npz.  Run in the build container (needs
/root/reference for oracle/_ref); the GPU box only reads the .npz."""
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
src = np.load(os.path.join(HERE, "vec_ops.npz"))
r = orc.Ref(cases.OPS_MESH_INI)
r.model_create("avg_ops", "rk4")
r.model_set("f", src["f0"])
r.model_set("g", src["g0"])
r.model_rhs(0.0)
out = {}
for k in ("f", "g"):   # guard cells of a time derivative are whatever its allocation held: keep the fixture deterministic
    a = np.zeros_like(src["f0"])
    a[2:-2, 2:-2, :] = r.model_ddt(k)[2:-2, 2:-2, :]
    out["avg_ddt_" + k] = a
r.close()
np.savez_compressed(os.path.join(HERE, "avg_ops.npz"), **out)
print("wrote avg_ops.npz", {k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
