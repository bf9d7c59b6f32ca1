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

    def all_gather_x(self, local, out):
        """All-gather along this rank's X row in PE_XIND order (CyclicReduce interface exchange)."""
        dist.all_gather_into_tensor(out, local, group=self.x_group)
        return out
