// LaplaceCyclic, block-partitioned path: the spectrum never visits HBM and every (y, kz) system is split
// along x into blocks so that the grid is (blocks in x) * ny CTAs.
//
// Replaces, for isolated systems (one rank in X, not periodic in X) with nz in {128, 256, 512}:
// LaplaceCyclic::solve (reference src/invert/laplace/impls/cyclic/cyclic_laplace.cxx:187-290),
// FFTTransform::forward / backward (common_transform.cxx:29-115), Laplacian::tridagMatrix
// (invert_laplace.cxx:416-768) and CyclicReduce<dcomplex> (include/bout/cyclic_reduction.hxx:119-637).
//
// Round 1 measured the three-kernel path (rfft -> Thomas -> irfft) at 2.54 GB of DRAM traffic per solve for
// 0.54 GB of algorithmic traffic (the spectrum and the Thomas d' each make a round trip), with the Thomas
// kernel at 257 CTAs.  Here the rows of a system are cut into blocks of MB rows; the first row of every block
// but the first is a SEPARATOR.  With the separator values known the blocks decouple, so:
//
//   rows(DOTS)   CTA = (block, y): loads its rows, rfft in registers (fft_reg.cuh), spectrum of the block in
//                shared memory; thread kz forms  yF = gF . r,  yL = gL . r  over the block's interior rows,
//                where gF / gL are the first / last rows of the inverse of the block's interior matrix
//                (matrix-only vectors, precomputed), and hands out the separator row's spectrum r_s.
//                Output: 3 complex per (y, block, kz).  The right-hand side is READ ONLY.
//   reduced      thread (y, kz): the separators obey a tridiagonal system of P - 1 unknowns,
//                A_j x_{s,j-1} + B_j x_{s,j} + C_j x_{s,j+1} = r_s - a_s yL_{j-1} - c_s yF_j   (factored once).
//   rows(SOLVE)  CTA = (block, y): loads its rows AGAIN, rfft, Thomas forward / backward inside the block
//                in shared memory starting from the known separator values (a separator is an identity row
//                of the modified matrix: d' = x_s, c' = 0), irfft, stores the solution rows.
//
// HBM traffic: rhs twice + result once = 24 B per point and no spectral workspace at all (the round-1 path:
// 75 B per point measured), for one extra forward transform.  The Thomas recurrences carry only 2 fused
// multiply-adds per row on the dependency chain (pivot reciprocals come from the table).
//
// Matrix-only tables ([ysel][row][kz], ysel = y when any coefficient array varies along y, else 0 -- then the
// tables are a few MB and live in L2): inv = 1 / (b_i - a_i c'_{i-1}) of the modified matrix, gF, gL; per
// separator: the factored reduced system.  They are rebuilt only when an input array or a setting changed:
// k_lapb_prepare compares every Field2D input with the copy taken at the last factorisation (bitwise, on the
// device, ~13 MB of traffic) and the factor kernels return at once when nothing changed -- the reference
// re-eliminates the matrix in every solve (cyclic_reduction.hxx:119-167).
//
// All per-thread logic is in B200_HD functions taking the thread's coordinates as arguments, with no barrier
// inside: the kernels (laplace_cyclic.cu) place the barriers, tests/emul/ steps simulated threads through
// the same functions on the host.
#pragma once

#include "laplace_coefs.cuh"
#include "fft_reg.cuh"

#include <cstring>

namespace b200 {

constexpr int LAP_NK = 6;     // table entries per (x,y): k0..k4 of tridagCoef5 and A
constexpr int LAP_IFACE = 6;  // interface coefficients per system: sL, cL, dL, aU, cU, dU
constexpr int LAP_MAXPE = 8;  // ranks in X
constexpr int LAP_NFP = 12;   // Field2D inputs of a solve (fingerprinted by the block path)
constexpr int LAPB_NRED = 10; // per-block / per-separator arrays: RA, RBU, RBD, RC (assembly), the factored system, MID
enum { LAPB_RA = 0, LAPB_RBU, LAPB_RBD, LAPB_RC, LAPB_RINV, LAPB_RAI, LAPB_RCP, LAPB_ASI, LAPB_CSI, LAPB_MID };
enum { LAPW_STATUS = 0, LAPW_YDEP = 1, LAPW_YDEP_GEN = 2, LAPW_DIRTY_GEN = 3 }; // 32-bit words at the workspace start

struct LapParams {
  LapCoefs m;
  LapBndry bd;
  const double* __restrict__ rhs;
  const double* __restrict__ x0;
  double* __restrict__ out;
  double* __restrict__ table; // [ny][nx][LAP_NK] (x contiguous for one y: a system's rows are one run)
  cd* __restrict__ spec;   // [nxs][nys][nmode]  rhs spectrum in, solution spectrum out
  cd* __restrict__ cprime; // [nxs][nys][nmode]
  cd* __restrict__ dprime; // [nxs][nys][nmode]
  cd* __restrict__ sprime; // coupled only: [nxs][nys][nmode]
  cd* __restrict__ inv;    // block path: pivot reciprocals
  cd* __restrict__ bnd;    // block path: a, c of the boundary rows, [inbndry+outbndry][nys][nmode][2]
  cd* __restrict__ iface;  // coupled only: this rank's [LAP_IFACE][nsys]
  const cd* __restrict__ iface_all; // coupled only: [nxpe][LAP_IFACE][nsys]
  int* __restrict__ status;
  int* __restrict__ ydep;  // 0 while no metric / coefficient array was seen to vary along y (k_lap_ycheck)
  const cd* __restrict__ tw;
  int nx, ny, nz;
  int ys, ye;     // y range solved
  int nmode;
  int zero_dc;
  int in_set, out_set;
  int tpr, rows, pow2;
  int nxpe, pe_xind;
  // block path
  double* __restrict__ fpcopy; // [LAP_NFP][nx*ny] inputs as of the last factorisation
  cd* __restrict__ bif;        // [nys][P][2][nmode]  F0 (scaled separator right-hand side), F1 = yL
  cd* __restrict__ bxs;        // [nys][P][nmode]     separator values
  cd* __restrict__ bred;       // [LAPB_NRED][nys][P][nmode]
  double* __restrict__ bnq;    // [nys][n]      Nyquist mode kz = nz/2: Re of the rhs spectrum in, of the solution out
  double kscale;               // 2 pi / zlength: kwave = kz * kscale (common_transform.cxx:73)
  unsigned gen;                // generation of this solve (ydep / dirty words hold the generation that set them)
  int force;                   // settings changed on the host side: factor regardless of the comparison
  int bP, bMB, bNT;            // blocks per system, rows per block, threads per CTA of the rows kernels
  int bPD;                     // rows kernels: CTA b prefetches the rows of CTA b + bPD into L2 (0 = off)
  // images of the solution's edge rows (mesh ystart / yend in mys / mye): the bottom rows also go to the upper guard rows
  // of `out` (wrap_yguards) or of the lower Y neighbour's mapped arena (ypeer), the top rows to the lower guard rows of
  // `out` / of the upper neighbour; offsets in doubles, on = 0 disables a side
  int mys, mye, on_up, on_dn;
  long long off_up, off_dn;
  unsigned long long* peer_flag[2]; // ypeer: raised to `signal` by the last CTA that delivered rows to that neighbour
  unsigned long long signal;
  unsigned int* counters;
  unsigned int expect[2];
  int dbg;                     // timing ablations (BOUT_B200_LAPB_DBG; results are wrong when set): 1 = no per-mode
                               // phase, 2 = no inverse transform, 4 = no forward transform, 8 = no row loads
};

B200_HD cd crecip(cd z) {
  const double d = 1.0 / (z.re * z.re + z.im * z.im);
  return {z.re * d, -z.im * d};
}
B200_HD cd ldcd(const cd* __restrict__ p) {
  const double2 t = B200_LDG(reinterpret_cast<const double2*>(p));
  return {t.x, t.y};
}
B200_HD bool tiny(cd z) { return !((z.re * z.re + z.im * z.im) > 1e-300); }
B200_HD cd cneg(cd z) { return {-z.re, -z.im}; }

// Row ix of the system (jy, kz): interior rows from the table, boundary rows (and the doubly
// periodic kz = 0 pin) through the full tridagMatrix logic.  `tab` is the table block of this
// system's jy, indexed by the row number ix (either staged in shared memory or in global memory).
B200_HD void lapRow(const LapParams& p, const double* tab, int ix, int jy, int kz, double kwave, cd& a, cd& b, cd& c,
                    cd& r) {
  const int ncx = p.bd.xe - p.bd.xs;
  const bool special = p.bd.periodic_x
                           ? (kz == 0 && p.bd.first_x && ix == 0)
                           : ((p.bd.first_x && ix < p.bd.inbndry) || (p.bd.last_x && (ncx - ix) < p.bd.outbndry));
  if (special) {
    tridagRow(p.m, p.bd, p.bd.xs + ix, jy, kz, kwave, a, b, c, r);
    if (!p.bd.periodic_x) { // nothing lies beyond a physical boundary
      if (p.bd.first_x && ix == 0) {
        a = {0.0, 0.0};
      }
      if (p.bd.last_x && ix == ncx) {
        c = {0.0, 0.0};
      }
    }
    return;
  }
  const double2* t = reinterpret_cast<const double2*>(tab + static_cast<long long>(ix) * LAP_NK);
  const double2 t01 = t[0], t23 = t[1], t45 = t[2];
  a = {t01.x - t23.y, -kwave * t23.x};
  b = {-2.0 * t01.x - kwave * kwave * t01.y + t45.y, kwave * t45.x};
  c = {t01.x + t23.y, kwave * t23.x};
}

// Is row ix a boundary row of tridagMatrix, and which slot of the `bnd` array holds its a, c?
B200_HD int lapBndRow(const LapParams& p, int ix, bool& inner, bool& outer) {
  const int ncx = p.bd.xe - p.bd.xs;
  inner = p.bd.first_x && ix < p.bd.inbndry;
  outer = p.bd.last_x && (ncx - ix) < p.bd.outbndry;
  return outer ? p.bd.inbndry + (ix - (ncx - p.bd.outbndry + 1)) : ix;
}

// K-table entry of (jx, jy) (Laplacian::tridagCoefs without the kwave factors, + A)
B200_HD void lapTableEntry(const LapParams& p, int i) {
  double k[5];
  const int jx = i / p.ny, jy = i % p.ny;
  tridagCoef5(p.m, jx, jy, true, true, k);
  double* __restrict__ t = p.table + (static_cast<long long>(jy) * p.nx + jx) * LAP_NK;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    t[j] = k[j];
  }
  t[5] = p.m.A != nullptr ? p.m.A[i] : 0.0;
}

// ---- geometry of the block path -----------------------------------------------------------------------------
// A CTA of NT threads transforms ROWS = NT / T rows at once and then turns into one thread per kz mode.
//   NT >= N: NBLK = NT / N blocks per CTA, thread tid -> (block tid / N, mode tid % N), one chain per thread;
//   NT <  N: one block per CTA, KPT = N / NT modes per thread (independent chains, interleaved).
// Rows per block MB = ROWS / NBLK.  The Nyquist mode kz = N of a block is one more chain of the block's thread 0.
template <int N, int NT>
struct LapbShape {
  using F = RegFft<N>;
  static constexpr int T = F::T, B = F::B, SROW = F::SROW;
  static constexpr int ROWS = NT / T;
  static constexpr int NBLK = NT >= N ? NT / N : 1;
  static constexpr int KPT = NT >= N ? 1 : N / NT;
  static constexpr int MB = ROWS / NBLK;
  static constexpr int NRC = 4; // staged per-row constants: k0 - k3, k0 + k3, k2 (+ padding to 32 bytes)
  static constexpr size_t smemBytes = static_cast<size_t>(ROWS) * SROW * sizeof(cd) + ROWS * NRC * sizeof(double);
};
// rows per block for a row length and CTA size (the two are tied by the shapes above)
__host__ __device__ inline int lapbRowsPerBlock(int nz, int nt) { return nz == 512 ? nt / 16 : (nz == 256 ? 16 : 8); }
// The rows of a block (local indices li0 .. cnt-1 after its separator) split into a top half swept downward and a
// bottom half swept upward.  Boundary rows of tridagMatrix (the first inbndry / last outbndry rows of the whole
// system) have their own coefficient logic: [ra, rb) is the run of regular rows, leading boundary rows always fall
// into the top half and trailing ones into the bottom half, so the unrolled sweeps only ever see regular rows.
struct LapbSplit {
  int ra, rb, mid; // regular rows [ra, rb); top half [li0, mid), bottom half [mid, cnt)
};
B200_HD LapbSplit lapbSplit(const LapBndry& bd, int n, int r0, int li0, int cnt) {
  LapbSplit sp;
  const int inb = bd.first_x ? bd.inbndry : 0, outb = bd.last_x ? bd.outbndry : 0;
  sp.ra = max(li0, min(cnt, inb - r0));
  sp.rb = min(cnt, max(sp.ra, n - outb - r0));
  sp.mid = sp.ra + (sp.rb - sp.ra + 1) / 2;
  return sp;
}

struct LapbCta {
  int iyl, jy; // y row of this CTA (local to the solved range / mesh index)
  int row0;    // first system row of this CTA
  int ysel;    // y index into the matrix-only tables
};
template <int N, int NT>
B200_HD LapbCta lapbCta(const LapParams& p, int bid) {
  // x tile fastest: CTAs in flight together then read DIFFERENT slices of the shared matrix tables (with the y row
  // fastest, ~128 CTAs ask the same few L2 lines for their pivots at the same moment)
  const int n = p.bd.xe - p.bd.xs + 1;
  const int ntile = (n + LapbShape<N, NT>::ROWS - 1) / LapbShape<N, NT>::ROWS;
  LapbCta c;
  c.iyl = bid / ntile;
  c.jy = p.ys + c.iyl;
  c.row0 = (bid % ntile) * LapbShape<N, NT>::ROWS;
  c.ysel = 0; // set by lapbSetYsel once the prepare kernel's verdict may be read
  return c;
}
B200_HD void lapbSetYsel(const LapParams& p, LapbCta& c) {
  c.ysel = (static_cast<unsigned>(p.status[LAPW_YDEP_GEN]) == p.gen) ? c.iyl : 0;
}
B200_HD long long lapbTabIndex(const LapParams& p, int ysel, int row, int kz) {
  const int n = p.bd.xe - p.bd.xs + 1;
  return (static_cast<long long>(ysel) * n + row) * p.nmode + kz;
}
B200_HD long long lapbRedIndex(const LapParams& p, int which, int ysel, int jb, int kz) {
  const int nys = p.ye - p.ys + 1;
  return ((static_cast<long long>(which) * nys + ysel) * p.bP + jb) * p.nmode + kz;
}
B200_HD long long lapbSysIndex(const LapParams& p, int iyl, int jb, int kz) { // bxs
  return (static_cast<long long>(iyl) * p.bP + jb) * p.nmode + kz;
}
B200_HD long long lapbBndIndex(const LapParams& p, int sr, int ysel, int kz) {
  const int nys = p.ye - p.ys + 1;
  return ((static_cast<long long>(sr) * nys + ysel) * p.nmode + kz) * 2;
}
B200_HD long long bitsOf(double v) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(v);
#else
  long long r;
  memcpy(&r, &v, sizeof(r));
  return r;
#endif
}

// ---- prepare: y-dependence of the matrix, change detection ----------------------------------------------------
// element i of nx*ny.  Every load is issued before any is used (one memory round trip, not twelve).
B200_HD void lapbCompareThread(const LapParams& p, int i) {
  const int jx = i / p.ny, jy = i % p.ny;
  const long long i0 = static_cast<long long>(jx) * p.ny + p.ys;
  const bool solved = jy >= p.ys && jy <= p.ye;
  const double* arrs[LAP_NFP] = {p.m.dx, p.m.g11, p.m.g33, p.m.g13, p.m.G1, p.m.G3, p.m.g_11, p.m.g_22,
                                 p.m.A,  p.m.C1,  p.m.C2,  p.m.D};
  const long long n2 = static_cast<long long>(p.nx) * p.ny;
  double v[LAP_NFP], v0[LAP_NFP], old[LAP_NFP];
#pragma unroll
  for (int q = 0; q < LAP_NFP; ++q) {
    const bool have = arrs[q] != nullptr;
    v[q] = have ? B200_LDG(arrs[q] + i) : 0.0;
    v0[q] = have ? B200_LDG(arrs[q] + i0) : 0.0;
    old[q] = have ? p.fpcopy[q * n2 + i] : 0.0;
  }
  int ydiff = 0, changed = 0;
#pragma unroll
  for (int q = 0; q < LAP_NFP; ++q) {
    ydiff |= (solved && !(v[q] == v0[q])) ? 1 : 0;
    const bool ch = bitsOf(old[q]) != bitsOf(v[q]);
    changed |= ch ? 1 : 0;
    if (ch && arrs[q] != nullptr) {
      p.fpcopy[q * n2 + i] = v[q];
    }
  }
  unsigned* w = reinterpret_cast<unsigned*>(p.status);
  if (ydiff) {
    w[LAPW_YDEP_GEN] = p.gen;
  }
  if (changed || (i == 0 && p.force)) {
    w[LAPW_DIRTY_GEN] = p.gen;
    w[LAPW_STATUS] = 0;
  }
}
B200_HD bool lapbDirty(const LapParams& p) { return static_cast<unsigned>(p.status[LAPW_DIRTY_GEN]) == p.gen; }
B200_HD bool lapbYdep(const LapParams& p) { return static_cast<unsigned>(p.status[LAPW_YDEP_GEN]) == p.gen; }

// ---- factor 1: thread (ysel, block, kz): pivots, gF, gL, boundary-row coefficients, reduced-row pieces --------
// The pivot table serves the two-sided elimination of the SOLVE kernel: rows of the top half of a block
// (local index < lapbMid) hold the reciprocal pivots of the downward elimination started below the block's
// separator, rows of the bottom half those of the upward elimination started above the next separator; MID
// closes the 2x2 system where the two meet.  Returns true when a pivot vanished.
B200_HD bool lapbFactor1Thread(const LapParams& p, long long idx) {
  const int n = p.bd.xe - p.bd.xs + 1, nys = p.ye - p.ys + 1, nmode = p.nmode, P = p.bP, MB = p.bMB;
  const int ysel = static_cast<int>(idx / (static_cast<long long>(P) * nmode));
  if (ysel >= (lapbYdep(p) ? nys : 1)) {
    return false;
  }
  const int jb = static_cast<int>((idx / nmode) % P), kz = static_cast<int>(idx % nmode);
  const int jy = p.ys + ysel;
  const double* tab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  const double kwave = kz * p.kscale; // common_transform.cxx:73
  const int r0 = jb * MB, cnt = min(MB, n - r0);
  const int li0 = jb > 0 ? 1 : 0;
  const int lo = r0 + li0, hi = r0 + cnt - 1, mid = r0 + lapbSplit(p.bd, n, r0, li0, cnt).mid;
  cd* __restrict__ INV = p.inv;
  cd* __restrict__ GF = p.cprime;
  cd* __restrict__ GL = p.dprime;
  bool bad = false;
  cd a, b, c, r = {0.0, 0.0};
  if (jb > 0) { // the separator: an identity row of the modified matrix
    const long long o = lapbTabIndex(p, ysel, r0, kz);
    INV[o] = {1.0, 0.0};
    GF[o] = {0.0, 0.0};
    GL[o] = {0.0, 0.0};
  }
  p.bred[lapbRedIndex(p, LAPB_MID, ysel, jb, kz)] = {1.0, 0.0};
  if (hi < lo) { // a trailing block that is a separator only
    p.bred[lapbRedIndex(p, LAPB_RBD, ysel, jb, kz)] = {0.0, 0.0};
    p.bred[lapbRedIndex(p, LAPB_RC, ysel, jb, kz)] = {0.0, 0.0};
    return false;
  }
  // downward elimination inside the block (c' of the row above the block is 0)
  cd cprev = {0.0, 0.0}, cp_above_mid = {0.0, 0.0};
  for (int i = lo; i <= hi; ++i) {
    lapRow(p, tab, i, jy, kz, kwave, a, b, c, r);
    const cd denom = csub(b, cmul(a, cprev));
    bad = bad || tiny(denom);
    const cd inv = crecip(denom);
    cprev = cmul(c, inv);
    if (i == mid - 1) {
      cp_above_mid = cprev;
    }
    INV[lapbTabIndex(p, ysel, i, kz)] = inv;
    bool inner, outer;
    const int sr = lapBndRow(p, i, inner, outer);
    if (inner || outer) {
      cd* q = p.bnd + lapbBndIndex(p, sr, ysel, kz);
      q[0] = a;
      q[1] = c;
    }
  }
  // gL = last row of the inverse: gL_hi = inv_hi, gL_k = -a_{k+1} inv_k gL_{k+1}
  cd g = INV[lapbTabIndex(p, ysel, hi, kz)];
  GL[lapbTabIndex(p, ysel, hi, kz)] = g;
  for (int k = hi - 1; k >= lo; --k) {
    lapRow(p, tab, k + 1, jy, kz, kwave, a, b, c, r);
    g = cneg(cmul(cmul(a, INV[lapbTabIndex(p, ysel, k, kz)]), g));
    GL[lapbTabIndex(p, ysel, k, kz)] = g;
  }
  const cd mLF = GL[lapbTabIndex(p, ysel, lo, kz)], mLL = GL[lapbTabIndex(p, ysel, hi, kz)];
  // upward elimination: invU_k = 1 / (b_k - c_k a_{k+1} invU_{k+1}), kept in GF for the next loop; the
  // bottom-half rows of the pivot table take these
  cd e = {0.0, 0.0};
  for (int k = hi; k >= lo; --k) {
    lapRow(p, tab, k, jy, kz, kwave, a, b, c, r);
    const cd denom = csub(b, cmul(c, e));
    bad = bad || tiny(denom);
    const cd inv = crecip(denom);
    e = cmul(a, inv);
    GF[lapbTabIndex(p, ysel, k, kz)] = inv;
    if (k >= mid) {
      INV[lapbTabIndex(p, ysel, k, kz)] = inv;
    }
    if (k == mid) { // x_mid = (u_mid - a'_mid d'_{mid-1}) / (1 - a'_mid c'_{mid-1})
      const cd den = csub(cd{1.0, 0.0}, cmul(e, cp_above_mid));
      bad = bad || tiny(den);
      p.bred[lapbRedIndex(p, LAPB_MID, ysel, jb, kz)] = crecip(den);
    }
  }
  // gF = first row of the inverse: gF_lo = invU_lo, gF_k = -c_{k-1} invU_k gF_{k-1}
  g = GF[lapbTabIndex(p, ysel, lo, kz)];
  for (int k = lo + 1; k <= hi; ++k) {
    lapRow(p, tab, k - 1, jy, kz, kwave, a, b, c, r);
    g = cneg(cmul(cmul(c, GF[lapbTabIndex(p, ysel, k, kz)]), g));
    GF[lapbTabIndex(p, ysel, k, kz)] = g;
  }
  // pieces of the reduced rows of the separators above and below this block
  const cd mFF = GF[lapbTabIndex(p, ysel, lo, kz)], mFL = GF[lapbTabIndex(p, ysel, hi, kz)];
  lapRow(p, tab, lo, jy, kz, kwave, a, b, c, r);
  const cd a_lo = a;
  lapRow(p, tab, hi, jy, kz, kwave, a, b, c, r);
  const cd c_hi = c;
  if (jb + 1 < P) {
    lapRow(p, tab, r0 + MB, jy, kz, kwave, a, b, c, r);
    p.bred[lapbRedIndex(p, LAPB_RA, ysel, jb + 1, kz)] = cneg(cmul(cmul(a, mLF), a_lo));
    p.bred[lapbRedIndex(p, LAPB_RBU, ysel, jb + 1, kz)] = cmul(cmul(a, mLL), c_hi);
  }
  if (jb > 0) {
    lapRow(p, tab, r0, jy, kz, kwave, a, b, c, r);
    p.bred[lapbRedIndex(p, LAPB_RBD, ysel, jb, kz)] = cmul(cmul(c, mFF), a_lo);
    p.bred[lapbRedIndex(p, LAPB_RC, ysel, jb, kz)] = cneg(cmul(cmul(c, mFL), c_hi));
  }
  return bad;
}

// ---- factor 2: thread (ysel, kz): Thomas factors of the separator system --------------------------------------
// Row j:  A_j x_{j-1} + B_j x_j + C_j x_{j+1} = r_s - a_s yL_{j-1} - c_s yF_j.  Stored scaled by the pivot:
// RINV, CSI = c_s rinv (DOTS forms F0 = r_s rinv - CSI yF), ASI = a_s rinv, RAI = A rinv, RCP = C rinv.
B200_HD bool lapbFactor2Thread(const LapParams& p, int idx) {
  const int nys = p.ye - p.ys + 1, nmode = p.nmode, P = p.bP, MB = p.bMB;
  const int ysel = idx / nmode, kz = idx % nmode;
  if (ysel >= (lapbYdep(p) ? nys : 1)) {
    return false;
  }
  const int jy = p.ys + ysel;
  const double* tab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  const double kwave = kz * p.kscale;
  bool bad = false;
  cd cprev = {0.0, 0.0};
  for (int j = 1; j < P; ++j) {
    cd a, b, c, r = {0.0, 0.0};
    lapRow(p, tab, j * MB, jy, kz, kwave, a, b, c, r);
    const cd RA = p.bred[lapbRedIndex(p, LAPB_RA, ysel, j, kz)];
    const cd Bj = csub(csub(b, p.bred[lapbRedIndex(p, LAPB_RBU, ysel, j, kz)]),
                       p.bred[lapbRedIndex(p, LAPB_RBD, ysel, j, kz)]);
    const cd denom = csub(Bj, cmul(RA, cprev));
    bad = bad || tiny(denom);
    const cd rinv = crecip(denom);
    cprev = cmul(p.bred[lapbRedIndex(p, LAPB_RC, ysel, j, kz)], rinv);
    p.bred[lapbRedIndex(p, LAPB_RINV, ysel, j, kz)] = rinv;
    p.bred[lapbRedIndex(p, LAPB_RAI, ysel, j, kz)] = cmul(RA, rinv);
    p.bred[lapbRedIndex(p, LAPB_RCP, ysel, j, kz)] = cprev;
    p.bred[lapbRedIndex(p, LAPB_ASI, ysel, j, kz)] = cmul(a, rinv);
    p.bred[lapbRedIndex(p, LAPB_CSI, ysel, j, kz)] = cmul(c, rinv);
  }
  return bad;
}

// ---- reduced: thread (iyl, kz): the separator values of one system --------------------------------------------
//   d_j = F0_j - F1_{j-1} - RAI_j d_{j-1},   x_j = d_j - RCP_j x_{j+1}
// with F0_j = (r_s - c_s yF_j) rinv_j and F1_{j-1} = a_s rinv_j yL_{j-1} from the DOTS kernel.
// D: this thread's scratch column (P entries, stride dstride): shared memory on the device.
constexpr int LAPB_PF = 8; // steps whose operands are fetched together
B200_HD void lapbReducedThread(const LapParams& p, int idx, cd* D, int dstride) {
  const int nmode = p.nmode, P = p.bP;
  const int iyl = idx / nmode, kz = idx % nmode;
  const int ysel = lapbYdep(p) ? iyl : 0;
  const cd* __restrict__ IF = p.bif + static_cast<long long>(iyl) * P * 2 * nmode + kz; // [j][2][nmode]
  const cd* __restrict__ rai = p.bred + lapbRedIndex(p, LAPB_RAI, ysel, 0, kz);
  const cd* __restrict__ rcp = p.bred + lapbRedIndex(p, LAPB_RCP, ysel, 0, kz);
  cd dp = {0.0, 0.0};
  for (int j0 = 1; j0 < P; j0 += LAPB_PF) {
    cd f0[LAPB_PF], f1[LAPB_PF], ra[LAPB_PF];
#pragma unroll
    for (int k = 0; k < LAPB_PF; ++k) {
      const int j = min(j0 + k, P - 1);
      f0[k] = ldcd(IF + static_cast<long long>(j) * 2 * nmode);
      f1[k] = ldcd(IF + (static_cast<long long>(j - 1) * 2 + 1) * nmode);
      ra[k] = ldcd(rai + static_cast<long long>(j) * nmode);
    }
#pragma unroll
    for (int k = 0; k < LAPB_PF; ++k) {
      if (j0 + k < P) {
        dp = csub(csub(f0[k], f1[k]), cmul(ra[k], dp));
        D[(j0 + k) * dstride] = dp;
      }
    }
  }
  cd x = {0.0, 0.0};
  cd* __restrict__ xs = p.bxs + lapbSysIndex(p, iyl, 0, kz);
  for (int j0 = P - 1; j0 >= 1; j0 -= LAPB_PF) {
    cd rc[LAPB_PF];
#pragma unroll
    for (int k = 0; k < LAPB_PF; ++k) {
      rc[k] = ldcd(rcp + static_cast<long long>(max(j0 - k, 1)) * nmode);
    }
#pragma unroll
    for (int k = 0; k < LAPB_PF; ++k) {
      const int j = j0 - k;
      if (j >= 1) {
        x = csub(D[j * dstride], cmul(rc[k], x));
        xs[static_cast<long long>(j) * nmode] = x;
      }
    }
  }
}

// ---- rows kernels: per-thread phases --------------------------------------------------------------------------
enum { LAPB_DOTS = 0, LAPB_SOLVE = 1 };

// per-row constants of the CTA's rows -> shared memory (threads tid < ROWS): k0 - k3, k0 + k3, k2, boundary flag
template <int N, int NT>
B200_HD void lapbStageRows(const LapParams& p, const LapbCta& c, int tid, double* rowc) {
  using Sh = LapbShape<N, NT>;
  if (tid >= Sh::ROWS) {
    return;
  }
  const int n = p.bd.xe - p.bd.xs + 1;
  const int ix = c.row0 + tid;
  double* rc = rowc + tid * Sh::NRC;
  if (ix >= n) {
    rc[0] = rc[1] = rc[2] = rc[3] = 0.0;
    return;
  }
  const double2* t = reinterpret_cast<const double2*>(p.table + (static_cast<long long>(c.jy) * p.nx + p.bd.xs + ix) * LAP_NK);
  const double2 t01 = B200_LDG(t), t23 = B200_LDG(t + 1);
  rc[0] = t01.x - t23.y;
  rc[1] = t01.x + t23.y;
  rc[2] = t23.x;
  rc[3] = 0.0;
}

// the row of this thread's group as packed complex numbers: v[n2] = (row[2 (t + T n2)], row[2 (t + T n2) + 1])
template <int N, int NT, int MODE>
B200_HD void lapbLoadRow(const LapParams& p, const LapbCta& c, int tid, cd (&v)[RegFft<N>::B]) {
  using Sh = LapbShape<N, NT>;
  const int grp = tid / Sh::T, t = tid % Sh::T;
  const int n = p.bd.xe - p.bd.xs + 1, ncx = n - 1;
  const int ix = c.row0 + grp;
  bool zero = ix >= n;
  if (!zero) { // rows whose right-hand side tridagMatrix zeroes (invert_laplace.cxx:482-484, 628-630)
    const bool inner = p.bd.first_x && ix < p.bd.inbndry, outer = p.bd.last_x && (ncx - ix) < p.bd.outbndry;
    zero = (inner && !(p.bd.inner_flags & (INV_RHS | INV_SET))) || (outer && !(p.bd.outer_flags & (INV_RHS | INV_SET)));
    if (MODE == LAPB_SOLVE && ix > 0 && ix % Sh::MB == 0) {
      zero = true; // separator: its value comes from the reduced solve
    }
  }
  if (zero) {
#pragma unroll
    for (int n2 = 0; n2 < Sh::B; ++n2) {
      v[n2] = {0.0, 0.0};
    }
    return;
  }
  const int x = p.bd.xs + ix;
  // common_transform.cxx:49-57
  const bool from_x0 = ((x < p.bd.inbndry) && p.in_set) || ((p.nx - x - 1 < p.bd.outbndry) && p.out_set);
  const double2* __restrict__ s2 = reinterpret_cast<const double2*>(
      (from_x0 ? p.x0 : p.rhs) + (static_cast<long long>(x) * p.ny + c.jy) * p.nz);
#pragma unroll
  for (int n2 = 0; n2 < Sh::B; ++n2) {
    const double2 q = B200_LDG(s2 + t + Sh::T * n2);
    v[n2] = {q.x, q.y};
  }
}

template <int N, int NT>
B200_HD void lapbStoreRow(const LapParams& p, const LapbCta& c, int tid, const cd (&v)[RegFft<N>::B]) {
  using Sh = LapbShape<N, NT>;
  const int grp = tid / Sh::T, t = tid % Sh::T;
  const int n = p.bd.xe - p.bd.xs + 1;
  const int ix = c.row0 + grp;
  if (ix >= n) {
    return;
  }
  double2* __restrict__ d2 =
      reinterpret_cast<double2*>(p.out + (static_cast<long long>(p.bd.xs + ix) * p.ny + c.jy) * p.nz);
#pragma unroll
  for (int q = 0; q < Sh::B; ++q) {
    d2[t + Sh::T * q] = make_double2(v[q].re, v[q].im);
  }
  if (p.on_up && c.jy - p.mys < p.ny - 1 - p.mye) { // images of this row (see LapParams)
    double2* __restrict__ u2 = reinterpret_cast<double2*>(reinterpret_cast<double*>(d2) + p.off_up);
#pragma unroll
    for (int q = 0; q < Sh::B; ++q) {
      u2[t + Sh::T * q] = make_double2(v[q].re, v[q].im);
    }
  }
  if (p.on_dn && p.mye - c.jy < p.mys) {
    double2* __restrict__ l2 = reinterpret_cast<double2*>(reinterpret_cast<double*>(d2) + p.off_dn);
#pragma unroll
    for (int q = 0; q < Sh::B; ++q) {
      l2[t + Sh::T * q] = make_double2(v[q].re, v[q].im);
    }
  }
}

// ---- per-mode phase --------------------------------------------------------------------------------------------
// A chain = one kz mode of one block, owned by one thread.  on[ch] = false marks an idle chain (kz beyond the
// kept modes, or a lane that has no second-pass chain): it reads mode 0's operands and stores nothing.
//
// DOTS: yF = gF . r and yL = gL . r over the block's interior rows (gF / gL: first / last row of the inverse of the
// block's interior matrix, so yF, yL are the first / last interior values of the block's solution with zero
// separator values), then F0 = (r_s - c_s yF) rinv for this block's separator row of the reduced system and
// F1 = a_s' rinv' yL for the next one's.  No dependency chain at all.  The first LAPB_DB rows' table entries and
// the scaling factors are fetched before the CTA barrier that ends the transform phase.
constexpr int LAPB_DB = 8; // rows per batch of table fetches
template <int NCH>
struct LapbDots {
  cd gf[NCH][LAPB_DB], gl[NCH][LAPB_DB];
  cd rinv[NCH], csi[NCH], asi[NCH];
};
template <int N, int NT, int NCH>
B200_HD void lapbDotsFetch(const LapParams& p, const LapbCta& c, int blk, const int (&kzs)[NCH], int l0,
                           LapbDots<NCH>& st) {
  using Sh = LapbShape<N, NT>;
  const int n = p.bd.xe - p.bd.xs + 1, nmode = p.nmode;
  const int r0 = c.row0 + blk * Sh::MB;
  const int cnt = max(min(Sh::MB, n - r0), 1);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int kz = kzs[ch] < nmode ? kzs[ch] : 0;
    const long long o = lapbTabIndex(p, c.ysel, min(r0, n - 1), kz);
#pragma unroll
    for (int k = 0; k < LAPB_DB; ++k) {
      const long long ro = static_cast<long long>(min(l0 + k, cnt - 1)) * nmode; // rows past the end weigh nothing
      st.gf[ch][k] = ldcd(p.cprime + o + ro);
      st.gl[ch][k] = ldcd(p.dprime + o + ro);
    }
  }
}
template <int N, int NT, int NCH>
B200_HD void lapbDotsFetchScale(const LapParams& p, const LapbCta& c, int blk, const int (&kzs)[NCH], LapbDots<NCH>& st) {
  using Sh = LapbShape<N, NT>;
  const int n = p.bd.xe - p.bd.xs + 1, nmode = p.nmode;
  const int r0 = c.row0 + blk * Sh::MB, jb = min(r0 / Sh::MB, p.bP - 1);
  const bool next = r0 + Sh::MB < n;
  const cd zero = {0.0, 0.0};
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int kz = kzs[ch] < nmode ? kzs[ch] : 0;
    st.rinv[ch] = jb > 0 ? ldcd(p.bred + lapbRedIndex(p, LAPB_RINV, c.ysel, jb, kz)) : zero;
    st.csi[ch] = jb > 0 ? ldcd(p.bred + lapbRedIndex(p, LAPB_CSI, c.ysel, jb, kz)) : zero;
    st.asi[ch] = next ? ldcd(p.bred + lapbRedIndex(p, LAPB_ASI, c.ysel, jb + 1, kz)) : zero;
  }
}
B200_HD void cfma(cd& acc, cd a, cd b) { // acc += a b, four fused multiply-adds
  acc.re = fma(a.re, b.re, acc.re);
  acc.re = fma(-a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im);
  acc.im = fma(a.im, b.re, acc.im);
}
// st holds batch 0 on entry
template <int N, int NT, int NCH>
B200_HD void lapbDotsRun(const LapParams& p, const LapbCta& c, int blk, const int (&kzs)[NCH], const bool (&on)[NCH],
                         LapbDots<NCH>& st, const cd* S) {
  using Sh = LapbShape<N, NT>;
  const int n = p.bd.xe - p.bd.xs + 1, nmode = p.nmode;
  const int r0 = c.row0 + blk * Sh::MB, jb = r0 / Sh::MB;
  const int cnt = min(Sh::MB, n - r0);
  if (cnt <= 0) {
    return;
  }
  const int li0 = jb > 0 ? 1 : 0;
  const cd* Sb[NCH];
  cd yF[NCH], yL[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    Sb[ch] = S + static_cast<size_t>(blk) * Sh::MB * Sh::SROW + (on[ch] ? kzs[ch] : 0);
    yF[ch] = yL[ch] = {0.0, 0.0};
  }
#pragma unroll
  for (int l0 = 0; l0 < Sh::MB; l0 += LAPB_DB) {
    if (l0 > 0) {
      lapbDotsFetch<N, NT, NCH>(p, c, blk, kzs, l0, st);
    }
#pragma unroll
    for (int k = 0; k < LAPB_DB; ++k) {
      const int li = l0 + k;
      if (li < Sh::MB) {
        const bool act = li >= li0 && li < cnt;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const cd s = act ? Sb[ch][li * Sh::SROW] : cd{0.0, 0.0};
          cfma(yF[ch], st.gf[ch][k], s);
          cfma(yL[ch], st.gl[ch][k], s);
        }
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (on[ch]) {
      cd* __restrict__ f = p.bif + (static_cast<long long>(c.iyl) * p.bP + jb) * 2 * nmode + kzs[ch];
      f[0] = jb > 0 ? csub(cmul(Sb[ch][0], st.rinv[ch]), cmul(st.csi[ch], yF[ch])) : cd{0.0, 0.0};
      f[nmode] = cmul(st.asi[ch], yL[ch]);
    }
  }
}

// SOLVE: two-sided elimination between the known separator values.  Top half: d_i = (r_i - a_i d_{i-1}) inv_i
// downward from d = x_s; bottom half: u_i = (r_i - c_i u_{i+1}) inv_i upward from u = x_{s+1}; where they meet
// x_mid = (u_mid - a'_mid d_{mid-1}) MID; then x_i = d_i - c'_i x_{i+1} upward through the top half and
// x_i = u_i - a'_i x_{i-1} downward through the bottom half.  The two sweeps (and the NCH modes of a thread) are
// independent dependency chains.  S: rhs spectrum in, solution spectrum out.  The operands from global memory are
// fetched by lapbChainLoad before the CTA barrier that ends the transform phase.
template <int MB, int NCH>
struct LapbChain {
  cd w[NCH][MB]; // pivot reciprocals, replaced row by row by c' (top half) / a' (bottom half)
  cd xs[NCH];    // x_s
  cd xn[NCH];    // x_{s+1}
  cd midinv[NCH];
};

template <int N, int NT, int NCH>
B200_HD void lapbChainLoad(const LapParams& p, const LapbCta& c, int blk, const int (&kzs)[NCH],
                           LapbChain<LapbShape<N, NT>::MB, NCH>& st) {
  using Sh = LapbShape<N, NT>;
  constexpr int MB = Sh::MB;
  const int n = p.bd.xe - p.bd.xs + 1, nmode = p.nmode;
  const int r0 = c.row0 + blk * MB, jb = min(r0 / MB, p.bP - 1);
  const int cnt = max(min(MB, n - r0), 1);
  const bool next = r0 + MB < n; // a separator follows this block
  const cd zero = {0.0, 0.0};
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int kz = kzs[ch] < nmode ? kzs[ch] : 0;
    const cd* ip = p.inv + lapbTabIndex(p, c.ysel, min(r0, n - 1), kz);
#pragma unroll
    for (int li = 0; li < MB; ++li) {
      st.w[ch][li] = ldcd(ip + static_cast<long long>(min(li, cnt - 1)) * nmode);
    }
    st.midinv[ch] = ldcd(p.bred + lapbRedIndex(p, LAPB_MID, c.ysel, jb, kz));
    st.xs[ch] = jb > 0 ? ldcd(p.bxs + lapbSysIndex(p, c.iyl, jb, kz)) : zero;
    st.xn[ch] = next ? ldcd(p.bxs + lapbSysIndex(p, c.iyl, jb + 1, kz)) : zero;
  }
}

// The boundary rows of a block's sweeps, outside the unrolled loops (they exist in the first and the last block of
// a system only).  Sc: this mode's column of the block in S.  phase 0: top-half rows [li0, ra) downward (carry =
// d); 1: bottom-half rows [rb, cnt) upward (carry = u; amid = a' of row mid when that is a boundary row);
// 2: bottom-half rows after mid, downward substitution (carry = x); 3: top-half rows [li0, ra) upward (carry = x).
struct LapbEdge { // what lapbEdgeRows needs of the parameter block, by value (a reference to the kernel's parameter
                  // struct would make the compiler copy all of it to local memory in every thread)
  const cd* bnd;
  const cd* inv;
  long long inv0;  // index of (ysel, row r0, kz) in the pivot table
  long long bnd0;  // index of (slot 0, ysel, kz) in bnd
  long long bnd_slot; // index stride between boundary-row slots
  int nmode, r0, li0, cnt, srow;
  int ncx, inbndry, outbndry, first_x, last_x;
};
B200_HD_COLD void lapbEdgeRows(LapbEdge e, LapbSplit sp, bool on, int phase, cd* Sc, cd* carry, cd* amid) {
  const int first = phase == 0 ? e.li0 : (phase == 1 ? e.cnt - 1 : (phase == 2 ? max(sp.rb, sp.mid + 1) : sp.ra - 1));
  const int last = phase == 0 ? sp.ra - 1 : (phase == 1 ? sp.rb : (phase == 2 ? e.cnt - 1 : e.li0));
  const int step = (phase == 0 || phase == 2) ? 1 : -1;
  cd x = *carry;
  for (int li = first; step > 0 ? li <= last : li >= last; li += step) {
    const int ix = e.r0 + li;
    const bool outer = e.last_x && (e.ncx - ix) < e.outbndry;
    const int sr = outer ? e.inbndry + (ix - (e.ncx - e.outbndry + 1)) : ix; // lapBndRow
    const cd* q = e.bnd + e.bnd0 + sr * e.bnd_slot;
    const cd a = ldcd(q), cc = ldcd(q + 1);
    const cd inv = ldcd(e.inv + e.inv0 + static_cast<long long>(li) * e.nmode);
    const cd s = Sc[li * e.srow];
    if (phase == 0) {
      x = csub(cmul(s, inv), cmul(cmul(a, inv), x));
    } else if (phase == 1) {
      x = csub(cmul(s, inv), cmul(cmul(cc, inv), x));
      if (li == sp.mid) {
        *amid = cmul(a, inv);
      }
    } else if (phase == 2) {
      x = csub(s, cmul(cmul(a, inv), x));
    } else {
      x = csub(s, cmul(cmul(cc, inv), x));
    }
    if (on) {
      Sc[li * e.srow] = x;
    }
  }
  *carry = x;
}

// The common case -- a full block of regular rows behind a separator (every block but the first and the last of
// a system) -- with every bound a compile-time constant: straight-line code.  The shared-memory operands of row
// s + 1 are read before row s is computed and stored (the compiler keeps shared loads behind earlier shared stores,
// which would put a full load latency into every step of the recurrences).
template <int N, int NT, int NCH>
B200_HD void lapbChainRunFull(const LapParams& p, int blk, const int (&kzs)[NCH], const bool (&on)[NCH],
                              LapbChain<LapbShape<N, NT>::MB, NCH>& st, cd* S, const double* rowc) {
  using Sh = LapbShape<N, NT>;
  constexpr int MB = Sh::MB, MID = 1 + MB / 2, NTOP = MID - 1, NBOT = MB - MID, SR = Sh::SROW;
  const double* rc0 = rowc + blk * MB * Sh::NRC;
  double kw[NCH];
  cd* Sb[NCH];
  cd dt[NCH], ub[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int kz = on[ch] ? kzs[ch] : 0;
    kw[ch] = kz * p.kscale; // common_transform.cxx:73
    Sb[ch] = S + static_cast<size_t>(blk) * MB * SR + kz;
    dt[ch] = st.xs[ch];
    ub[ch] = st.xn[ch];
  }
  struct RowIn {
    double2 r01;
    double k2;
    cd s[NCH];
  };
  auto fetch = [&](int li) {
    RowIn in;
    in.r01 = *reinterpret_cast<const double2*>(rc0 + li * Sh::NRC);
    in.k2 = rc0[li * Sh::NRC + 2];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      in.s[ch] = Sb[ch][li * SR];
    }
    return in;
  };
  RowIn tn = fetch(1), bn = fetch(MB - 1);
#pragma unroll
  for (int s = 0; s < NTOP; ++s) {
    const RowIn tc = tn, bc = bn;
    if (s + 1 < NTOP) {
      tn = fetch(2 + s);
    }
    if (s + 1 < NBOT) {
      bn = fetch(MB - 2 - s);
    }
    { // top half, downward: d_i = (r_i - a_i d_{i-1}) inv_i
      const int li = 1 + s;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const double im = kw[ch] * tc.k2;
        const cd a = {tc.r01.x, -im}, cc = {tc.r01.y, im};
        const cd inv = st.w[ch][li];
        const cd ri = cmul(tc.s[ch], inv);
        const cd ai = cmul(a, inv);
        st.w[ch][li] = cmul(cc, inv);
        dt[ch] = csub(ri, cmul(ai, dt[ch]));
        if (on[ch]) {
          Sb[ch][li * SR] = dt[ch];
        }
      }
    }
    if (s < NBOT) { // bottom half, upward: u_i = (r_i - c_i u_{i+1}) inv_i
      const int li = MB - 1 - s;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const double im = kw[ch] * bc.k2;
        const cd a = {bc.r01.x, -im}, cc = {bc.r01.y, im};
        const cd inv = st.w[ch][li];
        const cd ri = cmul(bc.s[ch], inv);
        const cd ci = cmul(cc, inv);
        st.w[ch][li] = cmul(a, inv);
        ub[ch] = csub(ri, cmul(ci, ub[ch]));
        if (on[ch] && li != MID) {
          Sb[ch][li * SR] = ub[ch];
        }
      }
    }
  }
  // the second sweeps read back what this thread stored above; again one row ahead of the stores
  cd xt[NCH], xb[NCH], tnx[NCH], bnx[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    tnx[ch] = dt[ch]; // d of row MID - 1 is still in a register
    bnx[ch] = NBOT > 1 ? Sb[ch][(MID + 1) * SR] : cd{0.0, 0.0};
    const cd xm = cmul(csub(ub[ch], cmul(st.w[ch][MID], dt[ch])), st.midinv[ch]);
    xt[ch] = xm;
    xb[ch] = xm;
  }
#pragma unroll
  for (int s = 0; s < NTOP; ++s) {
    cd tcx[NCH], bcx[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      tcx[ch] = tnx[ch];
      bcx[ch] = bnx[ch];
      if (s + 1 < NTOP) {
        tnx[ch] = Sb[ch][(MID - 2 - s) * SR];
      }
      if (s + 1 < NBOT - 1) {
        bnx[ch] = Sb[ch][(MID + 2 + s) * SR];
      }
    }
    { // top half, upward: x_i = d_i - c'_i x_{i+1}
      const int li = MID - 1 - s;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        xt[ch] = csub(tcx[ch], cmul(st.w[ch][li], xt[ch]));
        if (on[ch]) {
          Sb[ch][li * SR] = xt[ch];
        }
      }
    }
    if (s < NBOT - 1) { // bottom half, downward: x_i = u_i - a'_i x_{i-1}
      const int li = MID + 1 + s;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        xb[ch] = csub(bcx[ch], cmul(st.w[ch][li], xb[ch]));
        if (on[ch]) {
          Sb[ch][li * SR] = xb[ch];
        }
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    if (on[ch]) {
      Sb[ch][MID * SR] = xb[ch] = cmul(csub(ub[ch], cmul(st.w[ch][MID], dt[ch])), st.midinv[ch]);
      Sb[ch][0] = st.xs[ch];
    }
  }
}

B200_HD LapbEdge lapbEdgeOf(const LapParams& p, int ysel, int r0, int li0, int cnt, int srow, int kz) {
  LapbEdge e;
  e.bnd = p.bnd;
  e.inv = p.inv;
  e.inv0 = lapbTabIndex(p, ysel, r0, kz);
  e.bnd0 = lapbBndIndex(p, 0, ysel, kz);
  e.bnd_slot = lapbBndIndex(p, 1, ysel, kz) - e.bnd0;
  e.nmode = p.nmode;
  e.r0 = r0;
  e.li0 = li0;
  e.cnt = cnt;
  e.srow = srow;
  e.ncx = p.bd.xe - p.bd.xs;
  e.inbndry = p.bd.inbndry;
  e.outbndry = p.bd.outbndry;
  e.first_x = p.bd.first_x;
  e.last_x = p.bd.last_x;
  return e;
}

template <int N, int NT, int NCH>
B200_HD void lapbChainRun(const LapParams& p, const LapbCta& c, int blk, const int (&kzs)[NCH], const bool (&on)[NCH],
                          LapbChain<LapbShape<N, NT>::MB, NCH>& st, cd* S, const double* rowc) {
  using Sh = LapbShape<N, NT>;
  constexpr int MB = Sh::MB;
  const int n = p.bd.xe - p.bd.xs + 1;
  const int r0 = c.row0 + blk * MB, jb = r0 / MB;
  const int cnt = min(MB, n - r0);
  if (cnt <= 0) {
    return;
  }
  const int li0 = jb > 0 ? 1 : 0;
  const LapbSplit sp = lapbSplit(p.bd, n, r0, li0, cnt);
  const bool edges = sp.ra > li0 || sp.rb < cnt; // boundary rows in this block (uniform over the CTA's block)
  if (jb > 0 && cnt == MB && !edges) {
    lapbChainRunFull<N, NT, NCH>(p, blk, kzs, on, st, S, rowc);
    return;
  }
  int ra = sp.ra, rb = sp.rb, mid = sp.mid;
#ifdef __CUDA_ARCH__
  asm volatile("" : "+r"(ra), "+r"(rb), "+r"(mid)); // fixed registers: the unrolled rows only compare against them
#endif
  const double* rc0 = rowc + blk * MB * Sh::NRC;
  int kz[NCH];
  double kw[NCH];
  cd* Sb[NCH];
  cd dt[NCH], ub[NCH], amid[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    kz[ch] = on[ch] ? kzs[ch] : 0;
    kw[ch] = kz[ch] * p.kscale; // common_transform.cxx:73
#ifdef __CUDA_ARCH__
    asm volatile("" : "+d"(kw[ch])); // keep it in a register (the compiler otherwise re-derives it in every row)
#endif
    Sb[ch] = S + static_cast<size_t>(blk) * MB * Sh::SROW + kz[ch];
    dt[ch] = st.xs[ch];
    ub[ch] = st.xn[ch];
    amid[ch] = {0.0, 0.0};
  }
  if (edges) { // (through copies: the sweeps' own variables must not have their address taken)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      cd t0 = dt[ch], t1 = ub[ch], t2 = amid[ch];
      const LapbEdge e = lapbEdgeOf(p, c.ysel, r0, li0, cnt, Sh::SROW, kz[ch]);
      lapbEdgeRows(e, sp, on[ch], 0, Sb[ch], &t0, &t2);
      lapbEdgeRows(e, sp, on[ch], 1, Sb[ch], &t1, &t2);
      dt[ch] = t0;
      ub[ch] = t1;
      amid[ch] = t2;
    }
  }
#pragma unroll
  for (int s = 0; s < MB; ++s) {
    { // top half, downward: row s
      const int li = s;
      if (li >= ra && li < mid) {
        const double2 r01 = *reinterpret_cast<const double2*>(rc0 + li * Sh::NRC);
        const double k2 = rc0[li * Sh::NRC + 2];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const double im = kw[ch] * k2;
          const cd a = {r01.x, -im}, cc = {r01.y, im};
          const cd inv = st.w[ch][li];
          const cd ri = cmul(Sb[ch][li * Sh::SROW], inv);
          const cd ai = cmul(a, inv);
          st.w[ch][li] = cmul(cc, inv);
          dt[ch] = csub(ri, cmul(ai, dt[ch]));
          if (on[ch]) {
            Sb[ch][li * Sh::SROW] = dt[ch];
          }
        }
      }
    }
    { // bottom half, upward: row MB - 1 - s
      const int li = MB - 1 - s;
      if (li >= mid && li < rb) {
        const double2 r01 = *reinterpret_cast<const double2*>(rc0 + li * Sh::NRC);
        const double k2 = rc0[li * Sh::NRC + 2];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const double im = kw[ch] * k2;
          const cd a = {r01.x, -im}, cc = {r01.y, im};
          const cd inv = st.w[ch][li];
          const cd ri = cmul(Sb[ch][li * Sh::SROW], inv);
          const cd ci = cmul(cc, inv);
          st.w[ch][li] = cmul(a, inv);
          if (li == mid) {
            amid[ch] = st.w[ch][li];
          }
          ub[ch] = csub(ri, cmul(ci, ub[ch]));
          if (on[ch]) {
            Sb[ch][li * Sh::SROW] = ub[ch];
          }
        }
      }
    }
  }
  // the meeting point.  dt = d_{mid-1}; ub = u_mid (or, with an empty bottom half, the next separator's value)
  cd xt[NCH], xb[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    cd xm = ub[ch];
    if (mid < cnt) {
      xm = cmul(csub(ub[ch], cmul(amid[ch], dt[ch])), st.midinv[ch]);
      if (on[ch]) {
        Sb[ch][mid * Sh::SROW] = xm;
      }
    }
    xt[ch] = xm;
    xb[ch] = xm;
  }
#pragma unroll
  for (int s = 0; s < MB; ++s) {
    { // top half, upward: x_i = d_i - c'_i x_{i+1}
      const int li = MB - 1 - s;
      if (li >= ra && li < mid) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          xt[ch] = csub(Sb[ch][li * Sh::SROW], cmul(st.w[ch][li], xt[ch]));
          if (on[ch]) {
            Sb[ch][li * Sh::SROW] = xt[ch];
          }
        }
      }
    }
    { // bottom half, downward: x_i = u_i - a'_i x_{i-1}
      const int li = s;
      if (li > mid && li < rb) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          xb[ch] = csub(Sb[ch][li * Sh::SROW], cmul(st.w[ch][li], xb[ch]));
          if (on[ch]) {
            Sb[ch][li * Sh::SROW] = xb[ch];
          }
        }
      }
    }
  }
  if (edges) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      cd t0 = xb[ch], t1 = xt[ch], t2 = amid[ch];
      const LapbEdge e = lapbEdgeOf(p, c.ysel, r0, li0, cnt, Sh::SROW, kz[ch]);
      lapbEdgeRows(e, sp, on[ch], 2, Sb[ch], &t0, &t2);
      lapbEdgeRows(e, sp, on[ch], 3, Sb[ch], &t1, &t2);
    }
  }
  if (jb > 0) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (on[ch]) {
        Sb[ch][0] = st.xs[ch];
      }
    }
  }
}

// ---- the Nyquist mode kz = nz/2 ----------------------------------------------------------------------------------
// N + 1 modes do not divide over the N per-mode threads of a block, and a second pass for one mode would hold the
// whole CTA at its barrier.  The mode's right-hand side is real and only the real part of its solution is used
// (c2r ignores Im X[nz/2]), so it travels as one double per row: the DOTS kernel exports Re X[N] of every row, the
// (tiny) systems -- one per y -- are solved next to the separator systems in k_lapb_reduced, and the SOLVE kernel
// puts the result into X[N] before its inverse transform.  One warp solves one system with the SAME block
// algorithm and tables (column kz = N of inv, gF, gL and of the separator factors): lane = block for the dot
// products, one lane for the separator sweep, lane = block for the two-sided block solves; everything staged in
// shared memory first.
struct LapbNyq {
  double* r;    // [n]   Re X[N]: right-hand side in, solution out
  cd* D;        // [n]   sweep values
  double* ktab; // [n][LAP_NK] K-table rows of this y
  cd* inv;      // [n]   column N of the pivot table
  cd* gf;       // [n]
  cd* gl;       // [n]
  cd* f0;       // [P]   separator right-hand sides, then separator values
  cd* f1;       // [P]
  cd* rai;      // [P]
  cd* rcp;      // [P]
  cd* bac;      // [LAPB_NQB][2] a, c of the boundary rows
  cd* ends;     // [P][3] a block's two sweep ends and a' of its row mid
};
constexpr int LAPB_NQB = 16; // boundary rows (inbndry + outbndry) whose a, c are staged; more are read from global memory
inline size_t lapbNyqBytes(int n, int P) {
  return static_cast<size_t>(n) * (sizeof(double) * (1 + LAP_NK) + sizeof(cd) * 4)
         + (static_cast<size_t>(P) * 7 + 2 * LAPB_NQB) * sizeof(cd);
}
B200_HD LapbNyq lapbNyqCarve(void* mem, int n, int P) {
  LapbNyq w;
  w.D = reinterpret_cast<cd*>(mem);
  w.inv = w.D + n;
  w.gf = w.inv + n;
  w.gl = w.gf + n;
  w.f0 = w.gl + n;
  w.f1 = w.f0 + P;
  w.rai = w.f1 + P;
  w.rcp = w.rai + P;
  w.bac = w.rcp + P;
  w.ends = w.bac + 2 * LAPB_NQB;
  w.ktab = reinterpret_cast<double*>(w.ends + 3 * P);
  w.r = w.ktab + static_cast<size_t>(n) * LAP_NK;
  return w;
}
B200_HD void lapbNyqStage(const LapParams& p, int iyl, int ysel, int tid, int nt, const LapbNyq& w) {
  const int n = p.bd.xe - p.bd.xs + 1, kz = p.nz / 2, jy = p.ys + iyl;
  const long long o0 = lapbTabIndex(p, ysel, 0, kz);
#pragma unroll 8
  for (int i = tid; i < n; i += nt) {
    const long long o = o0 + static_cast<long long>(i) * p.nmode;
    w.r[i] = p.bnq[static_cast<long long>(iyl) * n + i];
    w.inv[i] = ldcd(p.inv + o);
    w.gf[i] = ldcd(p.cprime + o);
    w.gl[i] = ldcd(p.dprime + o);
  }
  const double* tg = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
#pragma unroll 8
  for (int i = tid; i < n * LAP_NK; i += nt) {
    w.ktab[i] = B200_LDG(tg + i);
  }
  for (int j = tid; j < p.bP; j += nt) {
    w.rai[j] = j > 0 ? ldcd(p.bred + lapbRedIndex(p, LAPB_RAI, ysel, j, kz)) : cd{0.0, 0.0};
    w.rcp[j] = j > 0 ? ldcd(p.bred + lapbRedIndex(p, LAPB_RCP, ysel, j, kz)) : cd{0.0, 0.0};
  }
  const int nb = (p.bd.first_x ? p.bd.inbndry : 0) + (p.bd.last_x ? p.bd.outbndry : 0);
  for (int q = tid; q < 2 * min(nb, LAPB_NQB); q += nt) {
    w.bac[q] = ldcd(p.bnd + lapbBndIndex(p, q / 2, ysel, kz) + q % 2);
  }
}
// block jb: F0 = (r_s - c_s yF) rinv, F1 = a_s' rinv' yL
B200_HD void lapbNyqDots(const LapParams& p, int ysel, int jb, const LapbNyq& w) {
  const int n = p.bd.xe - p.bd.xs + 1, kz = p.nz / 2, MB = p.bMB;
  const int r0 = jb * MB, cnt = min(MB, n - r0), li0 = jb > 0 ? 1 : 0;
  const int jc = max(jb, 1), jn = min(jb + 1, p.bP - 1); // (clamped: the loads below are issued before the loop)
  const cd rinv = ldcd(p.bred + lapbRedIndex(p, LAPB_RINV, ysel, jc, kz));
  const cd csi = ldcd(p.bred + lapbRedIndex(p, LAPB_CSI, ysel, jc, kz));
  const cd asi = ldcd(p.bred + lapbRedIndex(p, LAPB_ASI, ysel, max(jn, 1), kz));
  cd yF = {0.0, 0.0}, yL = {0.0, 0.0};
  for (int li = li0; li < cnt; ++li) {
    yF = cadd(yF, cscale(w.gf[r0 + li], w.r[r0 + li]));
    yL = cadd(yL, cscale(w.gl[r0 + li], w.r[r0 + li]));
  }
  cd f0 = {0.0, 0.0}, f1 = {0.0, 0.0};
  if (jb > 0) {
    f0 = csub(cscale(rinv, w.r[r0]), cmul(csi, yF));
  }
  if (jb + 1 < p.bP) {
    f1 = cmul(asi, yL);
  }
  w.f0[jb] = f0;
  w.f1[jb] = f1;
}
// the separator sweep (one thread): f0[j] becomes the value of separator j
B200_HD void lapbNyqSeparators(const LapParams& p, const LapbNyq& w) {
  const int P = p.bP;
  cd d = {0.0, 0.0};
  for (int j = 1; j < P; ++j) {
    d = csub(csub(w.f0[j], w.f1[j - 1]), cmul(w.rai[j], d));
    w.f0[j] = d;
  }
  cd x = {0.0, 0.0};
  for (int j = P - 1; j >= 1; --j) {
    x = csub(w.f0[j], cmul(w.rcp[j], x));
    w.f0[j] = x;
  }
}
// a, c of row i for the Nyquist mode: boundary rows from the bnd array (filled by the factor kernel), others from the
// staged K-table
B200_HD void lapbNyqAC(const LapParams& p, int ysel, const LapbNyq& w, int i, double kwave, cd& a, cd& c) {
  bool inner, outer;
  const int sr = lapBndRow(p, i, inner, outer);
  if (inner || outer) {
    if (sr < LAPB_NQB) {
      a = w.bac[2 * sr];
      c = w.bac[2 * sr + 1];
    } else {
      const cd* q = p.bnd + lapbBndIndex(p, sr, ysel, p.nz / 2);
      a = ldcd(q);
      c = ldcd(q + 1);
    }
    return;
  }
  const double* t = w.ktab + static_cast<size_t>(i) * LAP_NK;
  a = {t[0] - t[3], -kwave * t[2]};
  c = {t[0] + t[3], kwave * t[2]};
}
// block jb between its separator values: two-sided sweeps with the shared pivot table, one thread per half
// (half 0: top rows downward, half 1: bottom rows upward).  Sweep: stores the sweep values, leaves the half's end
// value (and a' of row mid) in `ends`; after a barrier, Subst: both halves form x_mid and substitute outward,
// writing Re x into r.
B200_HD void lapbNyqSweepHalf(const LapParams& p, int ysel, int jb, int half, const LapbNyq& w, cd* ends) {
  const int n = p.bd.xe - p.bd.xs + 1, kz = p.nz / 2, MB = p.bMB;
  const int r0 = jb * MB, cnt = min(MB, n - r0), li0 = jb > 0 ? 1 : 0;
  const double kwave = kz * p.kscale;
  const LapbSplit sp = lapbSplit(p.bd, n, r0, li0, cnt);
  cd a, c;
  if (half == 0) {
    cd dt = jb > 0 ? w.f0[jb] : cd{0.0, 0.0};
    for (int li = li0; li < sp.mid; ++li) {
      const int i = r0 + li;
      lapbNyqAC(p, ysel, w, i, kwave, a, c);
      dt = cmul(csub(cd{w.r[i], 0.0}, cmul(a, dt)), w.inv[i]);
      w.D[i] = dt;
    }
    ends[jb * 3] = dt;
  } else {
    cd ub = jb + 1 < p.bP ? w.f0[jb + 1] : cd{0.0, 0.0}, amid = {0.0, 0.0};
    for (int li = cnt - 1; li >= sp.mid; --li) {
      const int i = r0 + li;
      lapbNyqAC(p, ysel, w, i, kwave, a, c);
      ub = cmul(csub(cd{w.r[i], 0.0}, cmul(c, ub)), w.inv[i]);
      w.D[i] = ub;
      if (li == sp.mid) {
        amid = cmul(a, w.inv[i]);
      }
    }
    ends[jb * 3 + 1] = ub;
    ends[jb * 3 + 2] = amid;
  }
}
B200_HD void lapbNyqSubstHalf(const LapParams& p, int ysel, int jb, int half, const LapbNyq& w, const cd* ends) {
  const int n = p.bd.xe - p.bd.xs + 1, kz = p.nz / 2, MB = p.bMB;
  const int r0 = jb * MB, cnt = min(MB, n - r0), li0 = jb > 0 ? 1 : 0;
  const double kwave = kz * p.kscale;
  const LapbSplit sp = lapbSplit(p.bd, n, r0, li0, cnt);
  const cd dt = ends[jb * 3], ub = ends[jb * 3 + 1], amid = ends[jb * 3 + 2];
  cd xm = ub;
  if (sp.mid < cnt) {
    xm = cmul(csub(ub, cmul(amid, dt)), ldcd(p.bred + lapbRedIndex(p, LAPB_MID, ysel, jb, kz)));
  }
  cd a, c, x = xm;
  if (half == 0) {
    for (int li = sp.mid - 1; li >= li0; --li) { // top half, upward
      const int i = r0 + li;
      lapbNyqAC(p, ysel, w, i, kwave, a, c);
      x = csub(w.D[i], cmul(cmul(c, w.inv[i]), x));
      w.r[i] = x.re;
    }
    if (jb > 0) {
      w.r[r0] = w.f0[jb].re;
    }
  } else {
    if (sp.mid < cnt) {
      w.r[r0 + sp.mid] = xm.re;
    }
    for (int li = sp.mid + 1; li < cnt; ++li) { // bottom half, downward
      const int i = r0 + li;
      lapbNyqAC(p, ysel, w, i, kwave, a, c);
      x = csub(w.D[i], cmul(cmul(a, w.inv[i]), x));
      w.r[i] = x.re;
    }
  }
}

// the modes of thread tid in the per-mode phase, and the block they belong to
template <int N, int NT>
B200_HD int lapbModes(int tid, int (&kzs)[LapbShape<N, NT>::KPT]) {
  using Sh = LapbShape<N, NT>;
#pragma unroll
  for (int ch = 0; ch < Sh::KPT; ++ch) {
    kzs[ch] = (NT >= N ? tid % N : tid) + ch * NT;
  }
  return NT >= N ? tid / N : 0;
}

// ---- host side: is this configuration served by the block path, and how large are its arrays? ---------------
inline bool lapbShapeOk(int nz) { return nz == 128 || nz == 256 || nz == 512; }
inline int lapbBlocks(int n, int mb) { return (n + mb - 1) / mb; }
// complex elements of the per-solve / per-separator arrays: bif (2) + bxs + bred (LAPB_NRED), for the smallest
// block size (8 rows) any CTA size uses
inline size_t lapbSmallElems(int nx, int ny, int nmode) {
  return static_cast<size_t>(2 + 1 + LAPB_NRED) * ny * lapbBlocks(nx, 8) * nmode;
}
// threads per CTA of the rows kernels for nz = 512 (256: blocks of 16 rows; 128: blocks of 8 rows, two modes per
// thread).  Other row lengths always use 256.
inline int& lapbThreads512() {
  static int nt = 256;
  return nt;
}

// ---- host side: workspace layout and the parameter block (no CUDA calls: shared with tests/emul/) -------------
inline bool lapCoupled(const bout_b200_laplace* L) { return L->periodic_x || !L->first_x || !L->last_x; }
inline size_t lapRoundUp(size_t n, size_t a) { return (n + a - 1) / a * a; }

inline void lapBndryWidths(const bout_b200_grid* g, const bout_b200_laplace* L, int* inb, int* outb) {
  *inb = *outb = g->xstart;
  if ((L->global_flags & INV_BOTH_BNDRY_ONE) || g->xstart < 2) {
    *inb = *outb = 1;
  }
  if (L->inner_boundary_flags & INV_BNDRY_ONE) {
    *inb = 1;
  }
  if (L->outer_boundary_flags & INV_BNDRY_ONE) {
    *outb = 1;
  }
}

// Byte offsets of the workspace regions.  The block path overlays its matrix-only tables on the regions the
// three-kernel paths use for c', d' (gF, gL) and keeps its small per-solve arrays where they keep the spectrum.
struct LapLayout {
  size_t table, fpcopy, spec, cprime, dprime, inv, bnd, nyq, sprime, iface, total;
  size_t nspec; // complex elements of one spectral array
};
inline LapLayout lapLayout(const bout_b200_grid* g, const bout_b200_laplace* L) {
  LapLayout l{};
  const size_t nmode = static_cast<size_t>(L->maxmode) + 1;
  l.nspec = static_cast<size_t>(g->nx) * g->ny * nmode;
  const size_t n2 = static_cast<size_t>(g->nx) * g->ny;
  size_t spec_elems = l.nspec;
  if (lapbShapeOk(g->nz)) {
    const size_t small = lapbSmallElems(g->nx, g->ny, static_cast<int>(nmode));
    spec_elems = small > spec_elems ? small : spec_elems;
  }
  size_t o = 256;
  l.table = o;
  o += lapRoundUp(n2 * LAP_NK * sizeof(double), 256);
  l.fpcopy = o;
  o += lapRoundUp(n2 * LAP_NFP * sizeof(double), 256);
  l.spec = o;
  o += lapRoundUp(spec_elems * sizeof(cd), 256);
  l.cprime = o;
  o += lapRoundUp(l.nspec * sizeof(cd), 256);
  l.dprime = o;
  o += lapRoundUp(l.nspec * sizeof(cd), 256);
  l.inv = o;
  o += lapRoundUp(l.nspec * sizeof(cd), 256);
  l.bnd = o;
  o += lapRoundUp(static_cast<size_t>(4) * (g->xstart > 1 ? g->xstart : 1) * g->ny * nmode * sizeof(cd), 256);
  l.nyq = o; // block path, Nyquist mode: one double per (x, y)
  o += lapRoundUp(n2 * sizeof(double), 256);
  if (lapCoupled(L)) {
    l.sprime = o;
    o += lapRoundUp(l.nspec * sizeof(cd), 256);
    l.iface = o;
    o += lapRoundUp(static_cast<size_t>(LAP_IFACE) * g->ny * nmode * sizeof(cd), 256);
  }
  l.total = o;
  return l;
}

// Everything of LapParams that follows from the grid, the settings and the workspace address.
inline void lapFillParams(const bout_b200_grid* g, const bout_b200_laplace* L, void* workspace, LapParams* pp) {
  LapParams& p = *pp;
  p = LapParams{};
  p.m.dx = L->dx;
  p.m.g11 = L->g11;
  p.m.g33 = L->g33;
  p.m.g13 = L->g13;
  p.m.G1 = L->G1;
  p.m.G3 = L->G3;
  p.m.g_11 = L->g_11;
  p.m.g_22 = L->g_22;
  p.m.A = L->A;
  p.m.C1 = L->C1;
  p.m.C2 = L->C2;
  p.m.D = L->D;
  p.m.zlength = L->zlength;
  p.m.nx = g->nx;
  p.m.ny = g->ny;
  p.m.all_terms = L->all_terms;
  p.m.nonuniform = L->nonuniform;
  p.bd.inner_flags = L->inner_boundary_flags;
  p.bd.outer_flags = L->outer_boundary_flags;
  p.bd.global_flags = L->global_flags;
  lapBndryWidths(g, L, &p.bd.inbndry, &p.bd.outbndry);
  p.bd.periodic_x = L->periodic_x;
  p.bd.first_x = L->first_x;
  p.bd.last_x = L->last_x;
  // cyclic_laplace.cxx:88-101
  p.bd.xs = (L->first_x && !L->periodic_x) ? 0 : g->xstart;
  p.bd.xe = (L->last_x && !L->periodic_x) ? g->nx - 1 : g->xend;
  p.bd.nx_local = g->nx;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  // cyclic_laplace.cxx:216-232 (include_yguards only matters where a y boundary exists)
  p.ys = L->include_yguards_lower ? 0 : g->ystart;
  p.ye = L->include_yguards_upper ? g->ny - 1 : g->yend;
  p.mys = g->ystart;
  p.mye = g->yend;
  if (!L->include_yguards_lower && !L->include_yguards_upper) {
    const long long image = static_cast<long long>(g->yend - g->ystart + 1) * g->nz;
    if (L->wrap_yguards) {
      p.on_up = p.on_dn = 1;
      p.off_up = image;
      p.off_dn = -image;
    } else if (L->ypeer != nullptr) {
      const bout_b200_ypeer* yp = L->ypeer;
      if (yp->peer_flag[0] != nullptr) {
        p.on_up = 1;
        p.off_up = image + yp->delta[0];
      }
      if (yp->peer_flag[1] != nullptr) {
        p.on_dn = 1;
        p.off_dn = -image + yp->delta[1];
      }
      p.peer_flag[0] = yp->peer_flag[0];
      p.peer_flag[1] = yp->peer_flag[1];
      p.signal = yp->signal;
      p.counters = yp->counters;
    }
  }
  p.nmode = L->maxmode + 1;
  p.zero_dc = (L->global_flags & INV_ZERO_DC) ? 1 : 0;
  p.in_set = (L->first_x && (L->inner_boundary_flags & INV_SET)) ? 1 : 0;
  p.out_set = (L->last_x && (L->outer_boundary_flags & INV_SET)) ? 1 : 0;
  p.nxpe = 1;
  p.pe_xind = 0;
  p.kscale = 2.0 * M_PI / L->zlength;

  const LapLayout l = lapLayout(g, L);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  p.status = reinterpret_cast<int*>(ws);
  p.ydep = p.status + LAPW_YDEP;
  p.table = reinterpret_cast<double*>(ws + l.table);
  p.fpcopy = reinterpret_cast<double*>(ws + l.fpcopy);
  p.spec = reinterpret_cast<cd*>(ws + l.spec);
  p.cprime = reinterpret_cast<cd*>(ws + l.cprime);
  p.dprime = reinterpret_cast<cd*>(ws + l.dprime);
  p.inv = reinterpret_cast<cd*>(ws + l.inv);
  p.bnd = reinterpret_cast<cd*>(ws + l.bnd);
  if (lapCoupled(L)) {
    p.sprime = reinterpret_cast<cd*>(ws + l.sprime);
    p.iface = reinterpret_cast<cd*>(ws + l.iface);
  }
  if (lapbShapeOk(g->nz)) {
    const int n = p.bd.xe - p.bd.xs + 1;
    p.bNT = g->nz == 512 ? lapbThreads512() : 256;
    p.bMB = lapbRowsPerBlock(g->nz, p.bNT);
    p.bP = lapbBlocks(n, p.bMB);
    const size_t per = static_cast<size_t>(p.ye - p.ys + 1) * p.bP * p.nmode;
    p.bif = p.spec;
    p.bxs = p.bif + 2 * per;
    p.bred = p.bxs + per;
    p.bnq = reinterpret_cast<double*>(ws + l.nyq);
  }
}

} // namespace b200
