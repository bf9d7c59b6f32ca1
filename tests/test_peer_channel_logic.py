"""CPU: host logic of the NVLink peer-memory Y exchange (bout-dev_b200/mesh.py:PeerYChannel) with a recording
stand-in for the C-ABI: which rows go where, which flag is raised, how the receive areas alternate.  (The data
path itself is covered on GPUs by tests/multigpu_check.py.)"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_mesh_module():
    # mesh.py only needs torch / torch.distributed at import; the capi import inside PeerYChannel is bypassed
    pkg = types.ModuleType("bout_dev_b200_cpu")
    pkg.__path__ = [os.path.join(ROOT, "bout-dev_b200")]
    sys.modules["bout_dev_b200_cpu"] = pkg
    spec = importlib.util.spec_from_file_location("bout_dev_b200_cpu.mesh", os.path.join(ROOT, "bout-dev_b200", "mesh.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class Recorder:
    def __init__(self):
        self.calls = []

    def halo_push_y2(self, g, fields, nrows, y0, bufs, flags, value, counters):
        self.calls.append(("push", nrows, tuple(y0), tuple(bufs), tuple(flags), value, counters))

    def halo_wait_unpack_y2(self, g, fields, nrows, y0, bufs, flags, value, status):
        self.calls.append(("recv", nrows, tuple(y0), tuple(bufs), tuple(flags), value, status))


def make_channel(meshmod, yup, ydown, rank=1):
    g = types.SimpleNamespace(nx=8, ny=10, nz=4, ystart=2, yend=7)
    mesh = types.SimpleNamespace(grid=g, myg=2, yup=yup, ydown=ydown, rank=rank)
    ch = object.__new__(meshmod.PeerYChannel)
    ch.capi, ch.mesh, ch.nf = Recorder(), mesh, 3
    ch.cap = 3 * g.nx * 2 * g.nz
    ch.flag_off = 4 * ch.cap * 8
    ch.base = 1 << 20
    ch.peer = {r: (r + 2) << 24 for r in {yup, ydown} - {None}}
    ch.count = 0
    return ch, g


def test_exchange_rows_flags_and_double_buffering():
    meshmod = load_mesh_module()
    ch, g = make_channel(meshmod, yup=2, ydown=0)
    ch.exchange(["f"] * 3)
    ch.exchange(["f"] * 3)
    (p1, r1, p2, r2) = ch.capi.calls
    up, dn, cap8 = ch.peer[2], ch.peer[0], ch.cap * 8
    # exchange 1 (parity 1): top interior rows -> upper neighbour's "from below" area, bottom rows -> lower neighbour's
    # "from above" area; flags 0 / 1 of the respective neighbour; value = exchange counter
    assert p1[0] == "push" and p1[1] == 2 and p1[2] == (g.yend - 1, g.ystart)
    assert p1[3] == (up + (1 * 2 + 0) * cap8, dn + (1 * 2 + 1) * cap8)
    assert p1[4] == (up + ch.flag_off, dn + ch.flag_off + 8) and p1[5] == 1
    # receive: lower guard rows from my "from below" area, upper guard rows from my "from above" area, my own flags
    assert r1[0] == "recv" and r1[2] == (0, g.yend + 1)
    assert r1[3] == (ch.base + (1 * 2 + 0) * cap8, ch.base + (1 * 2 + 1) * cap8)
    assert r1[4] == (ch.base + ch.flag_off, ch.base + ch.flag_off + 8) and r1[5] == 1
    # exchange 2 uses the other parity of every receive area and the next flag value
    assert p2[3] == (up + 0 * cap8, dn + 1 * cap8) and p2[5] == 2
    assert r2[3] == (ch.base + 0 * cap8, ch.base + 1 * cap8) and r2[5] == 2
    # the status word and the block counters live behind the two flags
    assert p1[6] == ch.base + ch.flag_off + 16 and r1[6] == ch.base + ch.flag_off + 24


def test_physical_boundary_sides_are_skipped():
    meshmod = load_mesh_module()
    ch, g = make_channel(meshmod, yup=None, ydown=0)       # no upper neighbour (non-periodic edge)
    ch.exchange(["f"])
    push, recv = ch.capi.calls
    assert push[3][0] == 0 and push[4][0] == 0             # nothing sent upwards
    assert recv[3][1] == 0                                 # nothing received from above
    assert push[3][1] != 0 and recv[3][0] != 0
