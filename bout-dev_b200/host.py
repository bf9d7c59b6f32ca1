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


_EXCHANGE_T = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_int, C.c_void_p)
_ALLGATHER_T = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_void_p)
_ALLREDUCE_T = C.CFUNCTYPE(C.c_double, C.c_double, C.c_int, C.c_void_p)
_ALLREDUCE_Y_T = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_int, C.c_void_p)


class YRanks:
    """The communicator side of the host layer's decomposed Mesh (bout::b200::YDecomposition, b200_host.hxx): one
    process per GPU, `torch.distributed` already initialised (NCCL on GPUs; gloo with the CPU test double, whose
    "device" memory is host memory).  direction "y": NYPE = the group's size (closed field lines, periodic);
    direction "x": NXPE = the group's size (no wrap).  Owns the exchange buffers and the callbacks the C++ Mesh calls:
    exchange(count, dir) = sendY / sendX / wait of boutmesh.cxx:1245-1555, allgather(count) for the interface rows
    of an X-decomposed LaplaceCyclic solve (PE_XIND order), allreduce(value, op) for the adaptive integrators' error
    norms.  torch's collectives are ordered after the work already on the current (= default) stream and their
    wait() orders later work on it after the receive, which is what the C++ side's pack -> exchange -> unpack needs."""

    def __init__(self, capacity_doubles: int, device=None, group=None, direction: str = "y", nxpe: "int | None" = None):
        import torch
        import torch.distributed as dist
        if direction not in ("x", "y", "xy"):
            raise ValueError("direction must be 'x', 'y' or 'xy'")
        self.dist, self.torch, self.group, self.direction = dist, torch, group, direction
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        # BoutMesh's layout: rank = PE_YIND * NXPE + PE_XIND (boutmesh.cxx:255-341)
        self.nxpe = {"x": self.size, "y": 1}.get(direction, nxpe)
        if not self.nxpe or self.size % self.nxpe != 0:
            raise ValueError(f"direction 'xy' needs nxpe dividing the {self.size} ranks")
        self.nype = self.size // self.nxpe
        self.xind, self.yind = self.rank % self.nxpe, self.rank // self.nxpe
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.capacity = int(capacity_doubles)
        self.buf = torch.zeros(4, self.capacity, dtype=torch.float64, device=device)
        self.gather_send = torch.zeros(self.capacity, dtype=torch.float64, device=device)
        self.gather_recv = torch.zeros(self.nxpe * self.capacity if self.nxpe > 1 else 1, dtype=torch.float64, device=device)
        self.exchanges = 0
        self.gathers = 0
        self.error = None
        grank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        at = lambda ix, iy: grank((iy % self.nype) * self.nxpe + ix)
        # [dir][below, above]; None = no neighbour (the ends of an X row)
        self.nbr = [(at(self.xind, self.yind - 1), at(self.xind, self.yind + 1)),
                    (at(self.xind - 1, self.yind) if self.xind > 0 else None,
                     at(self.xind + 1, self.yind) if self.xind < self.nxpe - 1 else None)]
        # the ranks of one X row all-gather the interface rows of a LaplaceCyclic solve: one sub-group per row (every
        # rank takes part in creating every group)
        self.row_group = group
        if self.nxpe > 1 and self.nype > 1:
            for iy in range(self.nype):
                gr = dist.new_group([at(ix, iy) for ix in range(self.nxpe)])
                if iy == self.yind:
                    self.row_group = gr
        # ... and the ranks of one Y column all-reduce averageY's partial sums
        self.col_group = group
        if self.nxpe > 1 and self.nype > 1:
            for ix in range(self.nxpe):
                gr = dist.new_group([at(ix, iy) for iy in range(self.nype)])
                if ix == self.xind:
                    self.col_group = gr
        self._exchange = _EXCHANGE_T(self._do_exchange)
        self._allgather = _ALLGATHER_T(self._do_allgather)
        self._allreduce_y = _ALLREDUCE_Y_T(self._do_allreduce_y)
        self._allreduce = _ALLREDUCE_T(self._do_allreduce)

    def _do_exchange(self, count, direction, _user):
        try:
            dist = self.dist
            send_lo, send_hi, recv_lo, recv_hi = (self.buf[i, :count] for i in range(4))
            below, above = self.nbr[1 if direction == 1 else 0]
            up, down = above is not None, below is not None
            # with two ranks in y both neighbours are the same process: its first send (upwards) pairs with the first
            # receive posted here (from below), its second with the second
            ops = []
            if up:
                ops.append(dist.P2POp(dist.isend, send_hi, above, self.group))
            if down:
                ops.append(dist.P2POp(dist.irecv, recv_lo, below, self.group))
                ops.append(dist.P2POp(dist.isend, send_lo, below, self.group))
            if up:
                ops.append(dist.P2POp(dist.irecv, recv_hi, above, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            self.exchanges += 1
            return 0
        except Exception as e:   # an exception must not unwind through the C++ frames
            self.error = e
            return 1

    def _do_allgather(self, count, _user):
        try:
            self.dist.all_gather_into_tensor(self.gather_recv[:self.nxpe * count], self.gather_send[:count], group=self.row_group)
            self.gathers += 1
            return 0
        except Exception as e:
            self.error = e
            return 1

    def _do_allreduce_y(self, count, op, _user):
        try:
            dist = self.dist
            dist.all_reduce(self.gather_send[:count], op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=self.col_group)
            return 0
        except Exception as e:
            self.error = e
            return 1

    def _do_allreduce(self, value, op, _user):
        dist, torch = self.dist, self.torch
        t = torch.tensor([value], dtype=torch.float64, device=self.buf.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=self.group)
        return float(t.item())


class Host:
    """One live mesh/model at a time (like BOUT++ itself: global mesh, global options).  `yranks`: a YRanks for a mesh
    decomposed in y or x over the ranks of its process group (this process holds slab yranks.rank of yranks.size)."""

    def __init__(self, ini_text: str, yranks: "YRanks | None" = None):
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
        self.yranks = yranks
        L.b200host_set_decomposition.argtypes = [C.c_int] * 4 + [_EXCHANGE_T, _ALLGATHER_T, _ALLREDUCE_Y_T, _ALLREDUCE_T, C.c_void_p] \
            + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p, C.c_void_p]
        if yranks is not None and yranks.size > 1:
            b = yranks.buf
            self._check(L.b200host_set_decomposition(yranks.nxpe, yranks.xind, yranks.nype, yranks.yind, yranks._exchange, yranks._allgather, yranks._allreduce_y, yranks._allreduce, None,
                                                     b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(), b[3].data_ptr(),
                                                     yranks.capacity, yranks.gather_send.data_ptr(), yranks.gather_recv.data_ptr()))
        else:
            self._check(L.b200host_set_decomposition(1, 0, 1, 0, _EXCHANGE_T(), _ALLGATHER_T(), _ALLREDUCE_Y_T(), _ALLREDUCE_T(), None, None, None,
                                                     None, None, 0, None, None))
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
