"""CPU, world_size 2 and 4 over gloo: the host logic of the multi-GPU path -- BoutMesh decomposition,
neighbour tables, the Mesh::communicate message schedule (Y first, then X with corners; periodic
wrap; two messages between the same pair of ranks) and the X-row all-gather order of the
CyclicReduce interface exchange.

Method of the reference's tests/integrated/test-communications: every cell of a global field carries
a unique integer tag; after communicate() every guard cell must hold the tag of the global cell it
mirrors.  The CUDA copy kernels cannot run here, so the test injects index-copy stand-ins for the
three local copy operations (test infrastructure; the product has only the CUDA ones)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Grid:
    def __init__(self, nx, ny, nz, mxg, myg):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.xstart, self.xend, self.ystart, self.yend = mxg, nx - mxg - 1, myg, ny - myg - 1


class IndexCopyOps:
    """Stand-ins for bout_b200_halo_self / _pack_y / _unpack_y with the same semantics."""

    def self_exchange(self, g, fields, periodic_x, periodic_y):
        for f in fields:
            if periodic_y and g.ystart > 0:
                n = g.yend - g.ystart + 1
                f[:, :g.ystart] = f[:, n:n + g.ystart].clone()
                f[:, g.yend + 1:] = f[:, g.ystart:g.ystart + (g.ny - g.yend - 1)].clone()
            if periodic_x and g.xstart > 0:
                n = g.xend - g.xstart + 1
                f[:g.xstart] = f[n:n + g.xstart].clone()
                f[g.xend + 1:] = f[g.xstart:g.xstart + (g.nx - g.xend - 1)].clone()

    def pack_y(self, g, fields, y0, nrows, buf):
        buf.copy_(torch.stack([f[:, y0:y0 + nrows, :] for f in fields]).reshape(-1))

    def unpack_y(self, g, fields, y0, nrows, buf):
        b = buf.reshape(len(fields), g.nx, nrows, g.nz)
        for i, f in enumerate(fields):
            f[:, y0:y0 + nrows, :] = b[i]


def global_tags(gnx, gny, nz, nf):
    return [np.arange(gnx * gny * nz, dtype=np.float64).reshape(gnx, gny, nz) + 1e6 * (i + 1) for i in range(nf)]


def expected_local(G, pe_x, pe_y, mxsub, mysub, mxg, myg, px, py, MX, MY):
    """Local array incl. guards that communicate() must produce: interior = own slab; guards = the
    periodic image of the global INTERIOR (or untouched = -1 at physical boundaries)."""
    nz = G.shape[2]
    out = -np.ones((mxsub + 2 * mxg, mysub + 2 * myg, nz))
    for jx in range(mxsub + 2 * mxg):
        gx = pe_x * mxsub + jx - mxg          # global interior x index
        if not 0 <= gx < MX:
            if not px:
                continue
            gx %= MX
        for jy in range(mysub + 2 * myg):
            gy = pe_y * mysub + jy - myg
            if not 0 <= gy < MY:
                if not py:
                    continue
                gy %= MY
            out[jx, jy] = G[gx, gy]
    return out


def worker(rank, world, port, nxpe, nype, px, py, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    meshmod = ge.load_package().mesh
    mxsub, mysub, nz, mxg, myg, nf = 4, 3, 5, 2, 2, 2
    MX, MY = mxsub * nxpe, mysub * nype
    mesh = meshmod.BoutMesh(nxpe, nype, rank, mxsub, mysub, nz, mxg, myg, periodic_x=px, periodic_y=py,
                            ops=IndexCopyOps(), make_grid=Grid)
    tags = global_tags(MX, MY, nz, nf)   # tags of the global interior
    fields = []
    for G in tags:
        loc = -np.ones((mxsub + 2 * mxg, mysub + 2 * myg, nz))
        loc[mxg:-mxg, myg:-myg] = G[mesh.pe_xind * mxsub:(mesh.pe_xind + 1) * mxsub,
                                    mesh.pe_yind * mysub:(mesh.pe_yind + 1) * mysub]
        fields.append(torch.from_numpy(loc))
    mesh.communicate(fields)
    ok = True
    for G, f in zip(tags, fields):
        want = expected_local(G, mesh.pe_xind, mesh.pe_yind, mxsub, mysub, mxg, myg, px, py, MX, MY)
        ok = ok and np.array_equal(f.numpy(), want)
    # X-row all-gather order
    if nxpe > 1:
        local = torch.full((3,), float(rank))
        out = torch.empty(3 * nxpe)
        mesh.all_gather_x(local, out)
        want = np.repeat([mesh.pe_yind * nxpe + ix for ix in range(nxpe)], 3).astype(float)
        ok = ok and np.array_equal(out.numpy(), want)
    results[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


CASES = [(2, 1, False, True), (2, 1, True, True), (1, 2, False, True), (1, 2, True, False),
         (2, 2, False, True), (2, 2, True, True), (4, 1, False, True)]


@pytest.mark.parametrize("nxpe,nype,px,py", CASES)
def test_communicate_tags(nxpe, nype, px, py):
    world = nxpe * nype
    port = 29600 + (os.getpid() + nxpe * 7 + nype * 3 + px * 2 + py) % 300
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(worker, args=(world, port, nxpe, nype, px, py, results), nprocs=world, join=True)
        assert dict(results) == {r: True for r in range(world)}


def test_decomposition_rules():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    m = ge.load_package().mesh
    assert m.decompose(5, 4, 2) == (1, 1)          # MYPE = PE_YIND*NXPE + PE_XIND (boutmesh.cxx:1041)
    assert m.neighbours(0, 0, 4, 2, False, True) == (None, 1, 4, 4)
    assert m.neighbours(3, 1, 4, 2, True, True) == (6, 4, 3, 3)
    assert m.neighbours(0, 0, 1, 1, False, False) == (None, None, None, None)
    with pytest.raises(ValueError):
        m.decompose(8, 4, 2)
