"""BoutMesh decomposition and guard-cell exchange across GPUs (harness side of Mesh::communicate).

Mirrors the reference's BoutMesh (src/mesh/impls/bout/boutmesh.cxx): NXPE x NYPE ranks with
rank = PE_YIND * NXPE + PE_XIND (:1041), equal slabs MXSUB = MX / NXPE, MYSUB = MY / NYPE (:368-404),
neighbours with periodic wrap (:1012, :1588-1654), and Mesh::communicate = Y exchange first, then X
with the y-guard corners included (src/mesh/mesh.cxx:351-362, boutmesh.cxx:1245-1555).

One process per GPU; torch.distributed (NCCL on GPUs) is the transport.  Because x is the slowest
index, the X guard planes of a Field3D are contiguous slices of the field itself and are sent and
received in place -- no pack_data / unpack_data (boutmesh.cxx:2400-2466) on that path; Y guard rows are
strided and are packed by the CUDA kernels bout_b200_halo_pack_y / _unpack_y.  All sends and receives of
one communicate() are posted as one batch (one NCCL group).  Rectangular decomposition only (no
branch cuts / twist-shift).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class CudaHaloOps:
    """Local copies done by the CUDA library (the only implementation the product has)."""

    def __init__(self, capi):
        self.capi = capi

    def self_exchange(self, g, fields, periodic_x, periodic_y):
        self.capi.halo_self(g, fields, periodic_x=periodic_x, periodic_y=periodic_y)

    def pack_y(self, g, fields, y0, nrows, buf):
        self.capi.halo_pack_y(g, fields, y0, nrows, 0, g.nx - 1, buf)

    def unpack_y(self, g, fields, y0, nrows, buf):
        self.capi.halo_unpack_y(g, fields, y0, nrows, 0, g.nx - 1, buf)


class PeerYChannel:
    """Y guard-row exchange through the neighbours' memory (bout_b200_halo_push_y / _wait): this rank owns
    two double-buffered receive areas (rows arriving from below / from above), two arrival flags, two block
    counters and a status word in ONE device allocation whose CUDA IPC handle every neighbour opens once."""

    def __init__(self, mesh, nf, device):
        from . import capi
        self.capi, self.mesh, self.nf = capi, mesh, nf
        g = mesh.grid
        self.cap = nf * g.nx * mesh.myg * g.nz                      # doubles per message
        self.flag_off = 4 * self.cap * 8                            # recv[parity][side][cap], then the flags
        total = self.flag_off + 2 * 8 + 2 * 4 + 8
        self.base = capi.peer_alloc(total)
        mine = torch.tensor(list(capi.peer_export(self.base)), dtype=torch.uint8, device=device)
        allh = torch.empty(dist.get_world_size() * 64, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu().numpy().tobytes()
        self.peer = {}
        for r in {mesh.yup, mesh.ydown} - {None, mesh.rank}:
            self.peer[r] = capi.peer_open(allh[64 * r:64 * r + 64])
        self.count = 0

    def _recv(self, base, parity, side):
        return base + ((parity * 2 + side) * self.cap) * 8

    def exchange(self, fields):
        capi, m, g = self.capi, self.mesh, self.mesh.grid
        self.count += 1
        c, par = self.count, self.count & 1
        flags, counters = self.base + self.flag_off, self.base + self.flag_off + 16
        status = self.base + self.flag_off + 24
        up, dn = (self.peer[m.yup] if m.yup is not None else 0), (self.peer[m.ydown] if m.ydown is not None else 0)
        # one launch: my top interior rows arrive "from below" (side 0) at the upper neighbour, my bottom
        # interior rows "from above" (side 1) at the lower neighbour, each followed by that neighbour's flag
        capi.halo_push_y2(g, fields, m.myg, (g.yend - m.myg + 1, g.ystart),
                          (self._recv(up, par, 0) if up else 0, self._recv(dn, par, 1) if dn else 0),
                          (up + self.flag_off if up else 0, dn + self.flag_off + 8 if dn else 0), c, counters)
        # one launch: wait for my two flags, then unpack the lower / upper guard rows
        capi.halo_wait_unpack_y2(g, fields, m.myg, (0, g.yend + 1),
                                 (self._recv(self.base, par, 0) if dn else 0, self._recv(self.base, par, 1) if up else 0),
                                 (flags, flags + 8), c, status)

    def timed_out(self):
        """True if a wait gave up (reads the status word after synchronising the device)."""
        import ctypes as C
        torch.cuda.synchronize()
        v = C.c_int(0)
        self.capi._check(self.capi.lib.bout_b200_memcpy_d2h(C.byref(v), C.c_void_p(self.base + self.flag_off + 24),
                                                            C.c_size_t(4), None))
        return v.value != 0


class PeerFieldArena:
    """Fields whose Y guard rows are written by the NEIGHBOURS' kernels (bout_b200_ypeer): all of them live at the same
    offsets of one device allocation per rank, which the two Y neighbours map with CUDA IPC, followed by the arrival
    flags [channel][side] (side 0: rows from the lower neighbour have arrived, side 1: from the upper one), two block
    counters per channel and a status word."""

    NCH = 2  # channels: 0 = evolving-field stage states, 1 = phi

    def __init__(self, mesh, nfields, device):
        from . import capi
        self.capi, self.mesh = capi, mesh
        g = mesh.grid
        self.fbytes = (g.nx * g.ny * g.nz * 8 + 255) // 256 * 256
        self.flag_off = nfields * self.fbytes
        self.base = capi.peer_alloc(self.flag_off + 256)
        mine = torch.tensor(list(capi.peer_export(self.base)), dtype=torch.uint8, device=device)
        allh = torch.empty(dist.get_world_size() * 64, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu().numpy().tobytes()
        opened = {r: capi.peer_open(allh[64 * r:64 * r + 64]) for r in {mesh.yup, mesh.ydown} - {None, mesh.rank}}
        self.peer = [opened.get(mesh.ydown), opened.get(mesh.yup)]          # lower, upper neighbour's arena (or None)
        self.fields = [capi.tensor_from_ptr(self.base + i * self.fbytes, g.shape3, device) for i in range(nfields)]
        self.ptrs = {f.data_ptr() for f in self.fields}
        self.counters = [self.base + self.flag_off + 8 * 2 * self.NCH + 8 * ch for ch in range(self.NCH)]
        self.status = self.base + self.flag_off + 8 * 2 * self.NCH + 8 * self.NCH

    def flag(self, ch, side, base=None):
        return (self.base if base is None else base) + self.flag_off + 8 * (2 * ch + side)

    def ypeer(self, ch, signal, deliver=True, waits=None):
        """A CYPeer for a launch on channel ch that raises the neighbours' flags to `signal`: this rank is the UPPER
        neighbour of its lower neighbour (raises that one's side-1 flag) and the lower neighbour of its upper one."""
        pf = [None, None]
        delta = [0, 0]
        if deliver:
            for d in range(2):
                if self.peer[d] is not None:
                    pf[d] = self.flag(ch, 1 - d, self.peer[d])
                    assert (self.peer[d] - self.base) % 8 == 0
                    delta[d] = (self.peer[d] - self.base) // 8
        return self.capi.make_ypeer(delta, pf, signal, self.counters[ch], self.status, waits)

    def timed_out(self):
        import ctypes as C
        torch.cuda.synchronize()
        v = C.c_int(0)
        self.capi._check(self.capi.lib.bout_b200_memcpy_d2h(C.byref(v), C.c_void_p(self.status), C.c_size_t(4), None))
        return v.value != 0


def decompose(rank, nxpe, nype):
    """(PE_XIND, PE_YIND) of a rank (boutmesh.cxx:1041: MYPE = PE_YIND * NXPE + PE_XIND)."""
    if not 0 <= rank < nxpe * nype:
        raise ValueError(f"rank {rank} outside NXPE*NYPE = {nxpe * nype}")
    return rank % nxpe, rank // nxpe


def neighbours(pe_xind, pe_yind, nxpe, nype, periodic_x, periodic_y):
    """Ranks of the inner/outer X and lower/upper Y neighbours (None = physical boundary)."""
    def r(ix, iy):
        return iy * nxpe + ix
    xin = r(pe_xind - 1, pe_yind) if pe_xind > 0 else (r(nxpe - 1, pe_yind) if periodic_x else None)
    xout = r(pe_xind + 1, pe_yind) if pe_xind < nxpe - 1 else (r(0, pe_yind) if periodic_x else None)
    ydown = r(pe_xind, pe_yind - 1) if pe_yind > 0 else (r(pe_xind, nype - 1) if periodic_y else None)
    yup = r(pe_xind, pe_yind + 1) if pe_yind < nype - 1 else (r(pe_xind, 0) if periodic_y else None)
    return xin, xout, ydown, yup


class BoutMesh:
    def __init__(self, nxpe, nype, rank, mxsub, mysub, mz, mxg, myg, periodic_x=False, periodic_y=True,
                 device=None, ops=None, make_grid=None):
        if dist.is_initialized() and dist.get_world_size() != nxpe * nype:
            raise ValueError("NXPE * NYPE must equal the number of ranks (boutmesh.cxx:255-295)")
        self.nxpe, self.nype, self.rank = nxpe, nype, rank
        self.pe_xind, self.pe_yind = decompose(rank, nxpe, nype)
        self.mxsub, self.mysub, self.mz, self.mxg, self.myg = mxsub, mysub, mz, mxg, myg
        self.periodic_x, self.periodic_y = periodic_x, periodic_y
        self.first_x, self.last_x = self.pe_xind == 0, self.pe_xind == nxpe - 1
        self.first_y, self.last_y = self.pe_yind == 0, self.pe_yind == nype - 1
        self.xin, self.xout, self.ydown, self.yup = neighbours(self.pe_xind, self.pe_yind, nxpe, nype,
                                                               periodic_x, periodic_y)
        self.device = device
        if ops is None:
            from . import capi
            ops = CudaHaloOps(capi)
            make_grid = capi.Grid
        self.ops = ops
        self.grid = make_grid(mxsub + 2 * mxg, mysub + 2 * myg, mz, mxg, myg)
        self._ybuf = {}
        # Y guard rows go through NVLink peer memory when the ranks are CUDA processes of one node
        # (BOUT_B200_HALO=nccl keeps them on NCCL send/recv); X guard planes use NCCL
        import os
        self._peer_ok = (type(ops).__name__ == "CudaHaloOps" and dist.is_initialized() and nype > 1
                         and os.environ.get("BOUT_B200_HALO", "peer") == "peer")
        self._peer = {}
        # sub-communicator of this rank's X row (CyclicReduce's comm, boutmesh.cxx:668-760); every rank
        # must create every group
        self.x_group = None
        if dist.is_initialized() and nxpe > 1:
            for iy in range(nype):
                grp = dist.new_group([iy * nxpe + ix for ix in range(nxpe)])
                if iy == self.pe_yind:
                    self.x_group = grp

    # global index of local x (incl. guards): boutmesh.cxx XGLOBAL
    def x_global(self, jx):
        return jx + self.pe_xind * self.mxsub

    def _buf(self, key, n, like):
        b = self._ybuf.get(key)
        if b is None or b.numel() != n or b.device != like.device:
            b = torch.empty(n, dtype=like.dtype, device=like.device)
            self._ybuf[key] = b
        return b

    def communicate(self, fields):
        """Mesh::communicate(FieldGroup&): Y guards, then X guards including the y-guard corners."""
        g = self.grid
        nf = len(fields)
        # ---- Y --------------------------------------------------------------------------------
        if self.myg > 0:
            if self.nype == 1:
                if self.periodic_y:
                    self.ops.self_exchange(g, fields, False, True)
            elif self._peer_ok and fields[0].is_cuda and self.rank not in (self.yup, self.ydown):
                key = f"{nf}@{torch.cuda.current_stream().cuda_stream}"
                ch = self._peer.get(key)
                if ch is None:
                    ch = self._peer[key] = PeerYChannel(self, nf, fields[0].device)
                ch.exchange(fields)
            else:
                n = nf * g.nx * self.myg * g.nz
                ops, keep = [], []
                tag = f"{nf}@{torch.cuda.current_stream().cuda_stream}" if fields[0].is_cuda else f"{nf}"
                # Post order: send up, receive from below, send down, receive from above -- so that two
                # messages between the same pair of ranks (NYPE = 2, periodic) match in order.
                if self.yup is not None:
                    sb = self._buf("s_up" + tag, n, fields[0])
                    self.ops.pack_y(g, fields, g.yend - self.myg + 1, self.myg, sb)
                    ops.append(dist.P2POp(dist.isend, sb, self.yup))
                if self.ydown is not None:
                    rb = self._buf("r_dn" + tag, n, fields[0])
                    ops.append(dist.P2POp(dist.irecv, rb, self.ydown))
                    keep.append((rb, 0))
                    sb = self._buf("s_dn" + tag, n, fields[0])
                    self.ops.pack_y(g, fields, g.ystart, self.myg, sb)
                    ops.append(dist.P2POp(dist.isend, sb, self.ydown))
                if self.yup is not None:
                    rb = self._buf("r_up" + tag, n, fields[0])
                    ops.append(dist.P2POp(dist.irecv, rb, self.yup))
                    keep.append((rb, g.yend + 1))
                if ops:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                for rb, y0 in keep:
                    self.ops.unpack_y(g, fields, y0, self.myg, rb)
        # ---- X (planes are contiguous: sent and received in place) ---------------------------------
        if self.mxg > 0:
            if self.nxpe == 1:
                if self.periodic_x:
                    self.ops.self_exchange(g, fields, True, False)
            else:
                ops = []
                for f in fields:
                    # send outward, receive from inside, send inward, receive from outside: two messages
                    # between the same pair of ranks (NXPE = 2, periodicX) then match in order
                    if self.xout is not None:   # boutmesh.cxx:1289
                        ops.append(dist.P2POp(dist.isend, f[g.xend - self.mxg + 1:g.xend + 1], self.xout))
                    if self.xin is not None:    # boutmesh.cxx:1458
                        ops.append(dist.P2POp(dist.irecv, f[:g.xstart], self.xin))
                        ops.append(dist.P2POp(dist.isend, f[g.xstart:g.xstart + self.mxg], self.xin))  # :1273
                    if self.xout is not None:   # boutmesh.cxx:1464
                        ops.append(dist.P2POp(dist.irecv, f[g.xend + 1:], self.xout))
                if ops:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()

    def halo_timed_out(self):
        """Peer-memory exchange only: did any wait on a neighbour's flag give up?  (Synchronises.)"""
        return any(ch.timed_out() for ch in self._peer.values())

    def all_gather_x(self, local, out):
        """All-gather along this rank's X row in PE_XIND order (CyclicReduce interface exchange)."""
        dist.all_gather_into_tensor(out, local, group=self.x_group)
        return out
