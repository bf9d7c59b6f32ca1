"""GPU parity: z-FFT row operations (FFT derivatives, filter, lowPass, shiftZ, DC) vs the oracle.
Tolerance 1e-12 relative (north_star).  Sizes include non-power-of-two nz (direct-DFT path) and
the n = 8, 9 lengths the reference's own FFT test uses (tests/unit/invert/test_fft.cxx:50-120)."""
import numpy as np
import pytest
import torch

from util import rel_err, seeded_fields

pytestmark = pytest.mark.gpu
TOL = 1e-12
GRIDS = [(8, 6, 64, 2, 2), (6, 5, 8, 1, 1), (6, 5, 9, 1, 1), (5, 4, 7, 1, 1), (6, 3, 256, 2, 1), (6, 4, 128, 2, 1),
         (5, 3, 512, 1, 1), (4, 3, 1024, 1, 1), (5, 3, 12, 1, 1), (5, 3, 16, 1, 1), (5, 3, 32, 1, 1)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("grid", GRIDS)
def test_fft_derivatives(capi, orc, port, grid):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    (f,) = seeded_fields(g.shape3, 1)
    for order in (1, 2):
        want = port.index_ddz_fft(og, order, f)
        got = capi.ddz_fft(g, dev(f), order=order).cpu().numpy()
        assert rel_err(got, want, g.interior) < TOL
    want = port.index_ddz_fft(og, 1, f, fft_filter=0.3)
    got = capi.ddz_fft(g, dev(f), order=1, fft_filter=0.3).cpu().numpy()
    assert rel_err(got, want, g.interior) < TOL


@pytest.mark.parametrize("grid", GRIDS[:8])
def test_filter_lowpass_shift(capi, orc, port, grid):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    (f,) = seeded_fields(g.shape3, 1)
    allr = (slice(None),) * 3
    for n0 in (0, 1, g.nz // 2):
        assert rel_err(capi.zfilter(g, dev(f), n0).cpu().numpy(), port.zfilter(og, f, n0), allr,
                       floor=1e-3) < TOL
    for zmax, kz in ((2, True), (1, False), (g.nz // 2 - 1, False)):
        assert rel_err(capi.zlowpass(g, dev(f), zmax, kz).cpu().numpy(),
                       port.zlowpass(og, f, zmax, kz), allr) < TOL
    zl = 2 * np.pi
    assert rel_err(capi.zshift(g, dev(f), 0.37, zl).cpu().numpy(), port.zshift(og, f, 0.37, zl), allr) < TOL
    ang = 0.1 + 0.05 * np.arange(g.nx * g.ny, dtype=float).reshape(g.shape2)
    zl2 = np.full(g.shape2, 1.7)
    got = capi.zshift(g, dev(f), dev(ang), dev(zl2)).cpu().numpy()
    assert rel_err(got, port.zshift(og, f, ang, zl2), allr) < TOL


def test_shift_roundtrip_full_size(capi):
    """Size-independent property on a BASELINE-size field: shift by +a then -a is the identity
    up to round-off EXCEPT the Nyquist mode, whose imaginary part c2r drops; so compare after
    removing the Nyquist component with lowPass."""
    g = capi.Grid(64, 68, 256, 2, 2)
    f = torch.rand(g.shape3, dtype=torch.float64, device="cuda")
    f = capi.zlowpass(g, f, g.nz // 2 - 1, True)
    a = capi.zshift(g, f, 0.123, 2 * np.pi)
    b = capi.zshift(g, a, -0.123, 2 * np.pi)
    assert ((b - f).abs().max() / f.abs().max()).item() < 1e-13


def test_dc(capi, orc, port):
    grid = (9, 7, 64, 2, 2)
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    (f,) = seeded_fields(g.shape3, 1)
    got = capi.dc(g, dev(f)).cpu().numpy()
    want = port.dc(og, f)
    assert np.abs(got - want).max() / np.abs(want).max() < TOL


@pytest.mark.parametrize("grid", [(6, 7, 256, 2, 2), (5, 6, 128, 1, 1), (5, 6, 16, 1, 1)])
def test_filter_then_lowpass_one_pass(capi, orc, port, grid):
    """bout_b200_zfilter_lowpass = lowPass(filter(f, n0), zmax, keep_zonal) (elm_pb.cxx:1931-1955)."""
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    (f,) = seeded_fields(g.shape3, 1)
    allr = (slice(None),) * 3
    for n0, zmax, kz in ((1, 16, False), (1, 16, True), (3, 2, True), (0, 4, False), (0, 4, True), (2, g.nz, False)):
        want = port.zlowpass(og, port.zfilter(og, f, n0), zmax, kz)
        got = capi.zfilter_lowpass(g, dev(f), n0, zmax, kz).cpu().numpy()
        assert rel_err(got, want, allr, floor=1e-3) < TOL, (n0, zmax, kz)


@pytest.mark.parametrize("grid", [(6, 9, 256, 2, 2), (5, 8, 128, 1, 2), (4, 7, 512, 1, 1)])
def test_parallel_slices_dual_output_path(capi, grid):
    """ShiftedMetric::calcParallelSlices (shiftedmetric.cxx:226-246, phases :109-123) on the fast path that
    writes yup and ydown from one read of each source row, against a direct numpy restatement; the small-nz
    generic path is pinned by tests/test_gpu_shifted.py against the reference's own output."""
    g = capi.Grid(*grid)
    (f,) = seeded_fields(g.shape3, 1)
    rng = np.random.default_rng(7)
    zs = np.cumsum(0.2 + 0.3 * rng.uniform(size=g.shape2), axis=1)
    zl = 2 * np.pi / 5
    k = np.arange(g.nz // 2 + 1) * 2 * np.pi / zl
    yup, ydown = capi.parallel_slices(g, dev(f), dev(zs), zl, 1)
    yup, ydown = yup[0].cpu().numpy(), ydown[0].cpu().numpy()
    F = np.fft.rfft(f, axis=2)
    for y in range(g.ystart, g.yend + 1):
        for got, off in ((yup, 1), (ydown, -1)):
            ang = (zs[:, y] - zs[:, y + off])[:, None]
            ph = np.exp(-1j * k[None, :] * ang)
            ph[:, 0] = 1.0
            want = np.fft.irfft(F[:, y + off, :] * ph, n=g.nz, axis=1)
            d = np.abs(got[:, y + off, :] - want).max() / np.abs(want).max()
            assert d < TOL, (y, off, d)
