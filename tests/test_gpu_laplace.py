"""GPU parity: Delp2 (FFT and FD) and LaplaceCyclic through the C-ABI vs the oracle.

Tolerances from BASELINE.json north_star: Delp2 1e-12 relative, LaplaceCyclic solutions 1e-10
relative.  Flag combinations follow the reference's own test (tests/integrated/test-laplace/
test_laplace.cxx:19-100: flags 0, 3 (DC+AC grad), INVERT_SET inner/outer, with A, C, D coefficients).
"""
import numpy as np
import pytest
import torch

from util import make_metric, rel_err, seeded_fields

pytestmark = pytest.mark.gpu


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_laplace(capi, g, m, **kw):
    md = {n: dev(getattr(m, n)) for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22")}
    for k in ("A", "C1", "C2", "D"):
        if kw.get(k) is not None:
            kw[k] = dev(kw[k])
    return capi.Laplace(g, md, float(m.dz[0, 0]) * g.nz, **kw)


# nz = 128 .. 1024 take the fast FFT kernels (fft_fast.cuh) and, with nmode >= 127, the shared-memory table
GRIDS = [(20, 6, 32, 2, 1), (9, 5, 7, 2, 1), (12, 3, 16, 1, 1), (68, 4, 64, 2, 1), (20, 4, 128, 2, 1),
         (14, 3, 256, 2, 1), (12, 3, 512, 2, 1), (10, 3, 1024, 2, 1)]


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("trivial", [True, False])
def test_delp2_fft(capi, orc, port, grid, trivial):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz, trivial=trivial)
    (f,) = seeded_fields(g.shape3, 1)
    for all_terms, nonuni in ((True, m.non_uniform), (False, False)):
        want = port.delp2_fft(og, m, f, all_terms=all_terms, nonuniform=nonuni)
        lap = gpu_laplace(capi, g, m, all_terms=all_terms, nonuniform=nonuni)
        got = lap.delp2(dev(f)).cpu().numpy()
        assert rel_err(got, want, g.interior) < 1e-12


# enough (x, y) rows for the register-transform kernel (k_delp2_reg: tiles of 30 / 14 / 30 produced rows per CTA, so
# these row counts leave a partly filled last tile), checked against the marching kernel as well
@pytest.mark.parametrize("grid", [(156, 40, 128, 2, 2), (131, 44, 256, 2, 2), (321, 18, 512, 2, 1)])
def test_delp2_fft_register_transform_path(capi, orc, port, grid, monkeypatch):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz)
    (f,) = seeded_fields(g.shape3, 1)
    for all_terms, nonuni in ((True, m.non_uniform), (False, False)):
        want = port.delp2_fft(og, m, f, all_terms=all_terms, nonuniform=nonuni)
        lap = gpu_laplace(capi, g, m, all_terms=all_terms, nonuniform=nonuni)
        out = torch.full(g.shape3, 7.0, dtype=torch.float64, device="cuda")
        got = lap.delp2(dev(f), out=out).cpu().numpy()
        assert rel_err(got, want, g.interior) < 1e-12
        guard = np.ones(g.shape3, dtype=bool)
        guard[g.interior] = False
        assert np.all(got[guard] == 7.0)          # nothing outside the region is written


@pytest.mark.parametrize("grid", GRIDS)
def test_delp2_fd_matches_composed_reference_operators(capi, orc, port, grid):
    """useFFT=false branch: G1*DDX + G3*DDZ + g11*D2DX2 + g33*D2DZ2 + 2*g13*D2DXDZ
    (coordinates.cxx:1701-1702) composed from the oracle's C2 operators."""
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz)
    (f,) = seeded_fields(g.shape3, 1)
    b = lambda a: a[:, :, None]
    ddx = port.ddx(og, m, f, "C2")
    d2dxdz = port.ddz(og, m, ddx, "C2")
    want = (b(m.G1) * ddx + b(m.G3) * port.ddz(og, m, f, "C2") + b(m.g11) * port.d2dx2(og, m, f, "C2", "C2")
            + b(m.g33) * port.d2dz2(og, m, f, "C2") + b(2 * m.g13) * d2dxdz)
    got = capi.delp2_fd(g, dev(f), dev(m.dx), dev(m.dz), dev(m.g11), dev(m.g33), dev(m.g13), dev(m.G1),
                        dev(m.G3), dev(m.d1_dx)).cpu().numpy()
    assert rel_err(got, want, g.interior) < 1e-12


# (global, inner, outer).  Neumann on BOTH sides with A = 0 is singular for kz = 0 (the reference
# throws "Zero pivot"), so gradient conditions are one-sided here as in test-laplace.
FLAGSETS = [(0, 0, 0), (0, 3, 0), (0, 0, 3), (0, 2, 2), (0, 0, 1), (1, 4, 64), (0, 16, 16), (0, 32 | 3, 32),
            (4, 0, 0), (0, 128, 0), (0, 3 | 16, 32), (0, 1024, 0)]


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("flags", FLAGSETS)
def test_laplace_cyclic_flags(capi, orc, port, grid, flags):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz)
    f, x0 = seeded_fields(g.shape3, 2)
    want = port.laplace_cyclic(og, m, f, x0=x0, flags=flags)
    lap = gpu_laplace(capi, g, m, flags=flags)
    got = lap.solve(dev(f), dev(x0)).cpu().numpy()
    lap.status()
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    assert rel_err(got, want, region) < 1e-10


@pytest.mark.parametrize("grid", GRIDS[:2])
def test_laplace_cyclic_coefficients_and_filter(capi, orc, port, grid):
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz)
    f, x0 = seeded_fields(g.shape3, 2)
    X = (np.arange(g.nx)[:, None] + 0.5) / g.nx * np.ones(g.shape2)
    A = np.ascontiguousarray(-0.1 - 0.05 * X)
    D = np.ascontiguousarray(1 + 0.2 * X)
    Cc = np.ascontiguousarray(1 + 0.3 * np.sin(3 * X))
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    for kw in (dict(A=A), dict(A=A, D=D), dict(C1=Cc, C2=Cc), dict(A=A, C1=Cc, C2=Cc, D=D)):
        for maxmode in (None, max(1, g.nz // 4)):
            want = port.laplace_cyclic(og, m, f, x0=x0, maxmode=maxmode, **kw)
            got = gpu_laplace(capi, g, m, maxmode=maxmode, **kw).solve(dev(f), dev(x0)).cpu().numpy()
            assert rel_err(got, want, region) < 1e-10


@pytest.mark.parametrize("grid", [(20, 1, 32, 2, 0), (36, 3, 16, 2, 1), (260, 1, 256, 2, 0)])
def test_laplace_cyclic_periodic_x(capi, orc, port, grid):
    """hasegawa-wakatani's configuration: periodicX, kz=0 pinned + mean removed
    (invert_laplace.cxx:755-767, cyclic_laplace.cxx:266-287)."""
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz, trivial=True)
    (f,) = seeded_fields(g.shape3, 1)
    want = port.laplace_cyclic(og, m, f, periodic_x=True, nonuniform=False)
    lap = gpu_laplace(capi, g, m, periodic_x=True, nonuniform=False)
    got = lap.solve(dev(f)).cpu().numpy()
    lap.status()
    assert rel_err(got, want, g.interior) < 1e-10


def test_laplace_residual_full_size(capi, orc):
    """Size-independent property at a BASELINE-size grid (256x64x256): applying the forward
    operator Delp2 (FFT form, same coefficients) to the solution returns the right-hand side in
    the interior (the self-consistency check of tests/unit/invert/laplace/test_laplace_cyclic.cxx:36-63)."""
    g = capi.Grid(260, 68, 256, 2, 2)
    m = make_metric(orc, g.shape2, g.nz)
    lap = gpu_laplace(capi, g, m)
    rhs = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
    # drop the Nyquist mode: with g13, G3 != 0 its solution is complex and c2r keeps the real part only
    rhs = capi.zlowpass(g, rhs, g.nz // 2 - 1, True)
    x = lap.solve(rhs)
    lap.status()
    back = lap.delp2(x)
    # rows next to the boundary rows use the Dirichlet-pinned guard values, still consistent
    sl = (slice(g.xstart, g.xend + 1), slice(g.ystart, g.yend + 1), slice(None))
    err = (back[sl] - rhs[sl]).abs().max() / rhs[sl].abs().max()
    assert err.item() < 1e-9


def test_zero_pivot_is_reported(capi, orc):
    g = capi.Grid(12, 3, 8, 2, 1)
    m = make_metric(orc, g.shape2, g.nz, trivial=True)
    m.g11[:] = 0.0  # singular: b = 0 on the kz=0 interior rows
    m.g33[:] = 0.0
    lap = gpu_laplace(capi, g, m, nonuniform=False, all_terms=False, flags=(0, 1, 1))
    lap.solve(torch.ones(g.shape3, dtype=torch.float64, device="cuda"))
    with pytest.raises(capi.BoutB200Error) as e:
        lap.status()
    assert e.value.status == -4


@pytest.mark.parametrize("nxpe", [2, 3, 4, 8])
@pytest.mark.parametrize("flags", [(0, 0, 0), (0, 3, 0), (0, 0, 3), (0, 16, 16), (1, 4, 64)])
def test_laplace_x_decomposed_matches_single_rank(capi, orc, port, nxpe, flags):
    """CyclicReduce across NXPE ranks (cyclic_reduction.hxx:169-440) emulated on one GPU: every
    'rank' reduces its X slab, the interface records are concatenated as the NCCL all-gather would,
    every rank finishes.  The reference's own criterion for its MPI solve is nproc-independence
    (tests/integrated/test-laplace/runtest: nproc 1, 2, 4 against one benchmark)."""
    mxsub, mxg, myg, ny, nz = 6, 2, 1, 3, 16
    gnx = nxpe * mxsub + 2 * mxg
    G, OG = capi.Grid(gnx, ny + 2 * myg, nz, mxg, myg), orc.Grid(gnx, ny + 2 * myg, nz, mxg, myg)
    m = make_metric(orc, G.shape2, nz)
    f, x0 = seeded_fields(G.shape3, 2)
    X = (np.arange(gnx)[:, None] + 0.5) / gnx * np.ones(G.shape2)
    A = np.ascontiguousarray(-0.1 - 0.05 * X)
    D = np.ascontiguousarray(1 + 0.2 * X)
    want = port.laplace_cyclic(OG, m, f, x0=x0, flags=flags, A=A, D=D)
    g = capi.Grid(mxsub + 2 * mxg, ny + 2 * myg, nz, mxg, myg)
    laps, ifaces = [], []
    for p in range(nxpe):
        sl = slice(p * mxsub, p * mxsub + g.nx)
        md = {n: dev(getattr(m, n)[sl]) for n in ("dx", "g11", "g33", "g13", "G1", "G3", "g_11", "g_22")}
        lap = capi.Laplace(g, md, float(m.dz[0, 0]) * nz, flags=flags, A=dev(A[sl]), D=dev(D[sl]),
                           first_x=(p == 0), last_x=(p == nxpe - 1))
        laps.append((lap, sl))
        ifaces.append(lap.reduce(dev(f[sl]), dev(x0[sl])))
    iface_all = torch.cat(ifaces)
    got = np.zeros(G.shape3)
    for p, (lap, sl) in enumerate(laps):
        out = lap.finish(iface_all, nxpe, p, torch.zeros(g.shape3, dtype=torch.float64, device="cuda"))
        lap.status()
        o = out.cpu().numpy()
        xs = 0 if p == 0 else g.xstart
        xe = g.nx if p == nxpe - 1 else g.xend + 1
        got[sl][xs:xe] = o[xs:xe]
    region = (slice(None), slice(G.ystart, G.yend + 1), slice(None))
    assert rel_err(got, want, region) < 1e-10


# nz = 128, 256, 512 with isolated systems take the block-partitioned path (csrc/laplace_block.cuh); row counts
# cover a single block, a trailing separator-only block (n = 16 k + 1), short last blocks and several CTAs per y
BLOCK_GRIDS = [(12, 3, 128, 2, 1), (17, 4, 128, 2, 1), (40, 5, 128, 2, 1), (33, 3, 256, 2, 1), (70, 4, 256, 2, 1),
               (20, 3, 512, 2, 1), (49, 4, 512, 2, 1), (36, 4, 512, 1, 1)]


@pytest.mark.parametrize("grid", BLOCK_GRIDS)
@pytest.mark.parametrize("flags", [(0, 0, 0), (0, 3, 0), (4, 0, 0), (0, 16, 16), (0, 32 | 3, 32), (1, 4, 64)])
@pytest.mark.parametrize("filt", [False, True])
def test_laplace_block_path_matches_three_kernel_path(capi, orc, port, grid, flags, filt, monkeypatch):
    """The block-partitioned path (rows kernels with the spectrum on chip + separator system; default for these
    shapes) against the rfft -> Thomas / PCR -> irfft path (BOUT_B200_LAP_BLOCK=0) and against the oracle.
    Truncated spectra (laplace:filter), INVERT_ZERO_DC and INVERT_SET rows are covered; the second solve on the
    same workspace must not re-factor and must give the same bits."""
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz)
    f, x0 = seeded_fields(g.shape3, 2)
    maxmode = g.nz // 3 if filt else None
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    lap = gpu_laplace(capi, g, m, flags=flags, maxmode=maxmode)
    nblocks = -(-g.nx // (8 if g.nz == 128 else 16))
    n0 = capi.launch_count()
    first = lap.solve(dev(f), dev(x0)).cpu().numpy()
    lap.status()
    # prepare / factor (one cooperative kernel) + [dots + reduced: separators and / or the Nyquist mode] + solve
    assert capi.launch_count() - n0 == (4 if (nblocks > 1 or not filt) else 2)
    again = lap.solve(dev(f), dev(x0)).cpu().numpy()
    lap.status()
    assert np.array_equal(first[region], again[region])
    monkeypatch.setenv("BOUT_B200_LAP_BLOCK", "0")
    plain = lap.solve(dev(f), dev(x0)).cpu().numpy()
    lap.status()
    assert rel_err(first, plain, region) < 1e-11
    want = port.laplace_cyclic(og, m, f, x0=x0, flags=flags, maxmode=maxmode)
    assert rel_err(first, want, region) < 1e-10


def test_laplace_block_path_follows_coefficient_changes(capi, orc, port):
    """The matrix-only tables are kept in the workspace between solves; every solve compares the Field2D inputs
    with the copy taken at factorisation time on the device, so a coefficient array rewritten IN PLACE (setCoefA
    of the host layer), changed settings on the same workspace, or an explicit factor() all give the new operator."""
    grid = (40, 5, 128, 2, 1)
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz)
    f, x0 = seeded_fields(g.shape3, 2)
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    A = np.full(g.shape2, -0.1)
    lap = gpu_laplace(capi, g, m, A=A).factor()
    got = lap.solve(dev(f)).cpu().numpy()
    assert rel_err(got, port.laplace_cyclic(og, m, f, A=A), region) < 1e-10
    lap.keep["A"].fill_(-0.7)                 # same array, new contents, no factor() call
    got = lap.solve(dev(f)).cpu().numpy()
    A2 = np.full(g.shape2, -0.7)
    assert rel_err(got, port.laplace_cyclic(og, m, f, A=A2), region) < 1e-10
    lap.c.inner_boundary_flags = 3            # settings no longer match the held factorisation
    got = lap.solve(dev(f)).cpu().numpy()
    assert rel_err(got, port.laplace_cyclic(og, m, f, A=A2, flags=(0, 3, 0)), region) < 1e-10
    lap.keep["A"][7, 2] = float("nan")        # a NaN input must not read as "unchanged" for ever after
    lap.solve(dev(f))
    lap.keep["A"].fill_(-0.7)
    got = lap.solve(dev(f)).cpu().numpy()
    assert rel_err(got, port.laplace_cyclic(og, m, f, A=A2, flags=(0, 3, 0)), region) < 1e-10


def test_laplace_block_path_y_independent_tables(capi, orc, port):
    """No coefficient array varies along y: one [x][kz] table serves every y (the HW3D configuration)."""
    grid = (70, 6, 256, 2, 2)
    g, og = capi.Grid(*grid), orc.Grid(*grid)
    m = make_metric(orc, g.shape2, g.nz, trivial=True)
    f, x0 = seeded_fields(g.shape3, 2)
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    A = np.full(g.shape2, -0.3)
    for flags in ((0, 0, 0), (0, 3, 0), (0, 16, 16)):
        want = port.laplace_cyclic(og, m, f, x0=x0, flags=flags, A=A, nonuniform=False)
        got = gpu_laplace(capi, g, m, flags=flags, A=A, nonuniform=False).solve(dev(f), dev(x0)).cpu().numpy()
        assert rel_err(got, want, region) < 1e-10


def test_zero_pivot_is_reported_by_the_block_path(capi, orc):
    g = capi.Grid(20, 3, 128, 2, 1)
    m = make_metric(orc, g.shape2, g.nz, trivial=True)
    m.g11[:] = 0.0
    m.g33[:] = 0.0
    lap = gpu_laplace(capi, g, m, nonuniform=False, all_terms=False, flags=(0, 1, 1))
    lap.solve(torch.ones(g.shape3, dtype=torch.float64, device="cuda"))
    with pytest.raises(capi.BoutB200Error) as e:
        lap.status()
    assert e.value.status == -4


@pytest.mark.parametrize("dims", [(516, 132, 512, 2, 2), (516, 68, 256, 2, 2)])
def test_laplace_block_path_full_size(capi, orc, dims, monkeypatch):
    """BASELINE-size grids (the HW3D and elm-pb workloads): Delp2 of the solution returns the right-hand side, and
    the block path agrees with the three-kernel path to round-off."""
    g = capi.Grid(*dims)
    m = make_metric(orc, g.shape2, g.nz)
    lap = gpu_laplace(capi, g, m)
    rhs = torch.rand(g.shape3, dtype=torch.float64, device="cuda") - 0.5
    rhs = capi.zlowpass(g, rhs, g.nz // 2 - 1, True)
    x = lap.solve(rhs)
    lap.status()
    back = lap.delp2(x)
    sl = (slice(g.xstart, g.xend + 1), slice(g.ystart, g.yend + 1), slice(None))
    err = (back[sl] - rhs[sl]).abs().max() / rhs[sl].abs().max()
    assert err.item() < 1e-9
    monkeypatch.setenv("BOUT_B200_LAP_BLOCK", "0")
    y = lap.solve(rhs)
    region = (slice(None), slice(g.ystart, g.yend + 1), slice(None))
    assert ((x[region] - y[region]).abs().max() / y[region].abs().max()).item() < 1e-11
