"""ctypes binding of the C++ host layer's harness (host/src/host_driver.cxx -> libbout_b200_host.so).

The host layer is the product's C++17 side: BOUT++'s public API (Field3D, DDX.., bracket, Delp2,
Laplacian, Mesh::communicate, PhysicsModel, the explicit solvers, dump / restart files) implemented on the C-ABI.  This module lets
tests/ and bench.py run UNCHANGED reference models (host/models/*.so, compiled from the reference's
example sources) with raw arrays in and out, the same way the test harness drives the reference build.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "host", "libbout_b200_host.so")
MODEL_DIR = os.path.join(_HERE, "host", "models")


class HostError(RuntimeError):
    pass


class Host:
    """One live mesh/model at a time (like BOUT++ itself: global mesh, global options)."""

    def __init__(self, ini_text: str):
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built (python -c 'import __graft_entry__ as g; g.build()')")
        self.lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L = self.lib
        L.b200host_last_error.restype = C.c_char_p
        L.b200host_model_varname.restype = C.c_char_p
        L.b200host_pool_allocations.restype = C.c_ulonglong
        L.b200host_cow_copies.restype = C.c_ulonglong
        L.b200host_model_rhs.argtypes = [C.c_double]
        L.b200host_model_solve.argtypes = [C.c_int, C.c_double]
        self._check(L.b200host_init(ini_text.encode()))
        v = (C.c_int * 11)()
        self._check(L.b200host_mesh_info(v))
        self.info = list(v)
        self.shape3 = tuple(v[0:3])
        self.shape2 = tuple(v[0:2])
        self.model_vars = []
        self._open = True
        atexit.register(self.close)   # the library's static objects must not outlive the mesh at interpreter exit

    def _check(self, rc):
        if rc != 0:
            raise HostError(self.lib.b200host_last_error().decode(errors="replace"))

    def close(self):
        if self._open:
            self._open = False
            self.lib.b200host_finalise()

    def metric(self, name):
        a = np.zeros(self.shape2)
        self._check(self.lib.b200host_get_metric(name.encode(), a.ctypes.data_as(C.c_void_p)))
        return a

    def set_metric(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self._check(self.lib.b200host_set_metric(name.encode(), a.ctypes.data_as(C.c_void_p)))

    def expr3d(self, expr):
        a = np.zeros(self.shape3)
        self._check(self.lib.b200host_expr3d(expr.encode(), a.ctypes.data_as(C.c_void_p)))
        return a

    def model_create(self, name, solver="rk4"):
        path = name if os.path.isabs(name) else os.path.join(MODEL_DIR, name + ".so")
        self._check(self.lib.b200host_model_create(path.encode(), solver.encode()))
        self.model_vars = [self.lib.b200host_model_varname(i).decode() for i in range(self.lib.b200host_model_nvars())]
        return self.model_vars

    def _arr(self, fn, *args):
        a = np.zeros(self.shape3)
        self._check(fn(*args, a.ctypes.data_as(C.c_void_p)))
        return a

    def model_get(self, name):
        return self._arr(self.lib.b200host_model_get, self.model_vars.index(name))

    def model_ddt(self, name):
        return self._arr(self.lib.b200host_model_get_ddt, self.model_vars.index(name))

    def model_output(self, name):
        return self._arr(self.lib.b200host_model_output, name.encode())

    def model_set(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self._check(self.lib.b200host_model_set(self.model_vars.index(name), a.ctypes.data_as(C.c_void_p)))

    def model_rhs(self, t=0.0):
        self._check(self.lib.b200host_model_rhs(t))

    def model_solve(self, nout, output_step):
        self._check(self.lib.b200host_model_solve(nout, output_step))

    def stream_sync(self):
        self._check(self.lib.b200host_stream_sync())

    def rhs_calls(self):
        return int(self.lib.b200host_rhs_calls())

    def pool_allocations(self):
        return int(self.lib.b200host_pool_allocations())

    def cow_copies(self):
        """Whole-block device copies made by copy-on-write so far (each is a pass over a field that no kernel list shows)."""
        return int(self.lib.b200host_cow_copies())
