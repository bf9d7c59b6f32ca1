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
constexpr int LAPB_NRED = 9;  // reduced-system arrays: RA, RBU, RBD, RC (assembly), RINV, RAI, RCP, RAS, RCS
enum { LAPB_RA = 0, LAPB_RBU, LAPB_RBD, LAPB_RC, LAPB_RINV, LAPB_RAI, LAPB_RCP, LAPB_RAS, LAPB_RCS };
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
  cd* __restrict__ bif;        // [nys][P][3][nmode]  yF, yL, r_s
  cd* __restrict__ bxs;        // [nys][P][nmode]     separator values
  cd* __restrict__ bds;        // [nys][P][nmode]     d' of the reduced forward sweep
  cd* __restrict__ bred;       // [LAPB_NRED][nys][P][nmode]
  unsigned gen;                // generation of this solve (ydep / dirty words hold the generation that set them)
  int force;                   // settings changed on the host side: factor regardless of the comparison
  int bP, bMB;                 // blocks per system, rows per block
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
template <int N>
struct LapbShape {
  using F = RegFft<N>;
  static constexpr int NT = 256;          // threads per CTA
  static constexpr int T = F::T, B = F::B, SROW = F::SROW;
  static constexpr int ROWS = NT / T;     // rows per CTA (16, 32, 32)
  static constexpr int NBLK = NT / N;     // blocks per CTA (1, 2, 4): thread tid -> block tid / N, mode tid % N
  static constexpr int MB = ROWS / NBLK;  // rows per block (16, 16, 8)
  static_assert(MB == B, "rows per block = points per thread");
  static constexpr size_t smemBytes = static_cast<size_t>(ROWS) * SROW * sizeof(cd) + ROWS * LAP_NK * sizeof(double);
};
__host__ __device__ inline int lapbRowsPerBlock(int nz) { return nz == 128 ? 8 : 16; }

struct LapbCta {
  int iyl, jy; // y row of this CTA (local to the solved range / mesh index)
  int row0;    // first system row of this CTA
  int ysel;    // y index into the matrix-only tables
};
template <int N>
B200_HD LapbCta lapbCta(const LapParams& p, int bid) {
  const int nys = p.ye - p.ys + 1;
  LapbCta c;
  c.iyl = bid % nys;
  c.jy = p.ys + c.iyl;
  c.row0 = (bid / nys) * LapbShape<N>::ROWS;
  c.ysel = (static_cast<unsigned>(p.status[LAPW_YDEP_GEN]) == p.gen) ? c.iyl : 0;
  return c;
}
B200_HD long long lapbTabIndex(const LapParams& p, int ysel, int row, int kz) {
  const int n = p.bd.xe - p.bd.xs + 1;
  return (static_cast<long long>(ysel) * n + row) * p.nmode + kz;
}
B200_HD long long lapbRedIndex(const LapParams& p, int which, int ysel, int jb, int kz) {
  const int nys = p.ye - p.ys + 1;
  return ((static_cast<long long>(which) * nys + ysel) * p.bP + jb) * p.nmode + kz;
}
B200_HD long long lapbSysIndex(const LapParams& p, int iyl, int jb, int kz) { // bxs, bds
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

// ---- prepare: K-table, y-dependence of the matrix, change detection -------------------------------------------
// thread i of nx*ny
B200_HD void lapbPrepareThread(const LapParams& p, int i) {
  lapTableEntry(p, i);
  const int jx = i / p.ny, jy = i % p.ny;
  const long long i0 = static_cast<long long>(jx) * p.ny + p.ys;
  const bool solved = jy >= p.ys && jy <= p.ye;
  const double* arrs[LAP_NFP] = {p.m.dx, p.m.g11, p.m.g33, p.m.g13, p.m.G1, p.m.G3, p.m.g_11, p.m.g_22,
                                 p.m.A,  p.m.C1,  p.m.C2,  p.m.D};
  const long long n2 = static_cast<long long>(p.nx) * p.ny;
  bool ydiff = false, changed = false;
#pragma unroll
  for (int q = 0; q < LAP_NFP; ++q) {
    if (arrs[q] != nullptr) {
      const double v = B200_LDG(arrs[q] + i);
      ydiff = ydiff || (solved && !(v == B200_LDG(arrs[q] + i0)));
      double* cp = p.fpcopy + q * n2 + i;
      changed = changed || bitsOf(*cp) != bitsOf(v);
      *cp = v;
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
// Returns true when a pivot vanished.
B200_HD bool lapbFactor1Thread(const LapParams& p, long long idx) {
  const int n = p.bd.xe - p.bd.xs + 1, nys = p.ye - p.ys + 1, nmode = p.nmode, P = p.bP, MB = p.bMB;
  const int ysel = static_cast<int>(idx / (static_cast<long long>(P) * nmode));
  if (ysel >= (lapbYdep(p) ? nys : 1)) {
    return false;
  }
  const int jb = static_cast<int>((idx / nmode) % P), kz = static_cast<int>(idx % nmode);
  const int jy = p.ys + ysel;
  const double* tab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength; // common_transform.cxx:73
  const int r0 = jb * MB, cnt = min(MB, n - r0);
  const int lo = r0 + (jb > 0 ? 1 : 0), hi = r0 + cnt - 1;
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
  if (hi < lo) { // a trailing block that is a separator only
    p.bred[lapbRedIndex(p, LAPB_RBD, ysel, jb, kz)] = {0.0, 0.0};
    p.bred[lapbRedIndex(p, LAPB_RC, ysel, jb, kz)] = {0.0, 0.0};
    return false;
  }
  // downward elimination inside the block (c' of the row above the block is 0)
  cd cprev = {0.0, 0.0};
  for (int i = lo; i <= hi; ++i) {
    lapRow(p, tab, i, jy, kz, kwave, a, b, c, r);
    const cd denom = csub(b, cmul(a, cprev));
    bad = bad || tiny(denom);
    const cd inv = crecip(denom);
    cprev = cmul(c, inv);
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
  // upward elimination: invU_k = 1 / (b_k - c_k a_{k+1} invU_{k+1}), kept in GF for the next loop
  cd e = {0.0, 0.0};
  for (int k = hi; k >= lo; --k) {
    lapRow(p, tab, k, jy, kz, kwave, a, b, c, r);
    const cd denom = csub(b, cmul(c, e));
    bad = bad || tiny(denom);
    const cd inv = crecip(denom);
    e = cmul(a, inv);
    GF[lapbTabIndex(p, ysel, k, kz)] = inv;
  }
  // gF = first row of the inverse: gF_lo = invU_lo, gF_k = -c_{k-1} invU_k gF_{k-1}
  g = GF[lapbTabIndex(p, ysel, lo, kz)];
  for (int k = lo + 1; k <= hi; ++k) {
    lapRow(p, tab, k - 1, jy, kz, kwave, a, b, c, r);
    g = cneg(cmul(cmul(c, GF[lapbTabIndex(p, ysel, k, kz)]), g));
    GF[lapbTabIndex(p, ysel, k, kz)] = g;
  }
  // pieces of the reduced rows of the separators above and below this block
  const cd mLF = GL[lapbTabIndex(p, ysel, lo, kz)], mLL = GL[lapbTabIndex(p, ysel, hi, kz)];
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
B200_HD bool lapbFactor2Thread(const LapParams& p, int idx) {
  const int nys = p.ye - p.ys + 1, nmode = p.nmode, P = p.bP, MB = p.bMB;
  const int ysel = idx / nmode, kz = idx % nmode;
  if (ysel >= (lapbYdep(p) ? nys : 1)) {
    return false;
  }
  const int jy = p.ys + ysel;
  const double* tab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength;
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
    p.bred[lapbRedIndex(p, LAPB_RAS, ysel, j, kz)] = a;
    p.bred[lapbRedIndex(p, LAPB_RCS, ysel, j, kz)] = c;
  }
  return bad;
}

// ---- reduced: thread (iyl, kz): the separator values of one system --------------------------------------------
B200_HD void lapbReducedThread(const LapParams& p, int idx) {
  const int nmode = p.nmode, P = p.bP;
  const int iyl = idx / nmode, kz = idx % nmode;
  const int ysel = lapbYdep(p) ? iyl : 0;
  const cd* __restrict__ IF = p.bif;
  const cd* __restrict__ red = p.bred;
  cd* DS = p.bds; // written and read back by this thread: plain loads and stores
  cd dp = {0.0, 0.0};
#pragma unroll 4
  for (int j = 1; j < P; ++j) {
    const long long f = (static_cast<long long>(iyl) * P + j) * 3 * nmode + kz;
    const cd yLprev = ldcd(IF + f - 2 * nmode); // [iyl][j-1][1]
    const cd yF = ldcd(IF + f), rs = ldcd(IF + f + 2 * nmode);
    const cd as = ldcd(red + lapbRedIndex(p, LAPB_RAS, ysel, j, kz));
    const cd cs = ldcd(red + lapbRedIndex(p, LAPB_RCS, ysel, j, kz));
    const cd R = csub(csub(rs, cmul(as, yLprev)), cmul(cs, yF));
    const cd d = csub(cmul(R, ldcd(red + lapbRedIndex(p, LAPB_RINV, ysel, j, kz))),
                      cmul(ldcd(red + lapbRedIndex(p, LAPB_RAI, ysel, j, kz)), dp));
    DS[lapbSysIndex(p, iyl, j, kz)] = d;
    dp = d;
  }
  cd x = {0.0, 0.0};
#pragma unroll 4
  for (int j = P - 1; j >= 1; --j) {
    x = csub(DS[lapbSysIndex(p, iyl, j, kz)], cmul(ldcd(red + lapbRedIndex(p, LAPB_RCP, ysel, j, kz)), x));
    p.bxs[lapbSysIndex(p, iyl, j, kz)] = x;
  }
}

// ---- rows kernels: per-thread phases --------------------------------------------------------------------------
enum { LAPB_DOTS = 0, LAPB_SOLVE = 1 };

// this thread's part of the CTA's K-table rows -> shared memory
template <int N>
B200_HD void lapbStageTable(const LapParams& p, const LapbCta& c, int tid, double* ktab) {
  using Sh = LapbShape<N>;
  const int n = p.bd.xe - p.bd.xs + 1;
  for (int i = tid; i < Sh::ROWS * LAP_NK; i += Sh::NT) {
    const int r = i / LAP_NK;
    ktab[i] = (c.row0 + r < n)
                  ? B200_LDG(p.table + (static_cast<long long>(c.jy) * p.nx + p.bd.xs + c.row0 + r) * LAP_NK + i % LAP_NK)
                  : 0.0;
  }
}

// the row of this thread's group as packed complex numbers: v[n2] = (row[2 (t + T n2)], row[2 (t + T n2) + 1])
template <int N, int MODE>
B200_HD void lapbLoadRow(const LapParams& p, const LapbCta& c, int tid, cd (&v)[RegFft<N>::B]) {
  using Sh = LapbShape<N>;
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

template <int N>
B200_HD void lapbStoreRow(const LapParams& p, const LapbCta& c, int tid, const cd (&v)[RegFft<N>::B]) {
  using Sh = LapbShape<N>;
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
}

// a, c of local row r of the CTA for mode kz: boundary rows from the bnd array, interior rows from the staged table
template <int N>
B200_HD void lapbRowAC(const LapParams& p, const LapbCta& c, const double* ktab, int r, int kz, double kwave, cd& a,
                       cd& cc) {
  const int ix = c.row0 + r;
  bool inner, outer;
  const int sr = lapBndRow(p, ix, inner, outer);
  if (inner || outer) {
    const cd* q = p.bnd + lapbBndIndex(p, sr, c.ysel, kz);
    a = ldcd(q);
    cc = ldcd(q + 1);
    return;
  }
  const double* t = ktab + r * LAP_NK;
  a = {t[0] - t[3], -kwave * t[2]};
  cc = {t[0] + t[3], kwave * t[2]};
}

// DOTS, thread (blk, kz): yF, yL over the block's interior rows and the separator row's spectrum
template <int N>
B200_HD void lapbDotsThread(const LapParams& p, const LapbCta& c, int blk, int kz, const cd* S) {
  using Sh = LapbShape<N>;
  const int n = p.bd.xe - p.bd.xs + 1;
  const int r0 = c.row0 + blk * Sh::MB, jb = r0 / Sh::MB;
  const int cnt = min(Sh::MB, n - r0);
  if (cnt <= 0) {
    return;
  }
  const int li0 = jb > 0 ? 1 : 0;
  const cd* __restrict__ GF = p.cprime;
  const cd* __restrict__ GL = p.dprime;
  const cd* Sb = S + static_cast<size_t>(blk) * Sh::MB * Sh::SROW + kz;
  cd yF = {0.0, 0.0}, yL = {0.0, 0.0};
#pragma unroll
  for (int li = 0; li < Sh::MB; ++li) {
    if (li >= li0 && li < cnt) {
      const long long o = lapbTabIndex(p, c.ysel, r0 + li, kz);
      const cd s = Sb[li * Sh::SROW];
      yF = cadd(yF, cmul(ldcd(GF + o), s));
      yL = cadd(yL, cmul(ldcd(GL + o), s));
    }
  }
  cd* __restrict__ f = p.bif + (static_cast<long long>(c.iyl) * p.bP + jb) * 3 * p.nmode + kz;
  f[0] = yF;
  f[p.nmode] = yL;
  f[2 * p.nmode] = jb > 0 ? Sb[0] : cd{0.0, 0.0};
}

// SOLVE, thread (blk, kz): Thomas inside the block between the known separator values; S: rhs spectrum in,
// solution spectrum out.
template <int N>
B200_HD void lapbSolveThread(const LapParams& p, const LapbCta& c, int blk, int kz, cd* S, const double* ktab) {
  using Sh = LapbShape<N>;
  const int n = p.bd.xe - p.bd.xs + 1;
  const int r0 = c.row0 + blk * Sh::MB, jb = r0 / Sh::MB;
  const int cnt = min(Sh::MB, n - r0);
  if (cnt <= 0) {
    return;
  }
  const int li0 = jb > 0 ? 1 : 0;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength; // common_transform.cxx:73
  cd* Sb = S + static_cast<size_t>(blk) * Sh::MB * Sh::SROW + kz;
  cd inv[Sh::MB];
#pragma unroll
  for (int li = 0; li < Sh::MB; ++li) {
    inv[li] = li < cnt ? ldcd(p.inv + lapbTabIndex(p, c.ysel, r0 + li, kz)) : cd{1.0, 0.0};
  }
  const cd xs = jb > 0 ? ldcd(p.bxs + lapbSysIndex(p, c.iyl, jb, kz)) : cd{0.0, 0.0};
  cd x = (r0 + cnt < n) ? ldcd(p.bxs + lapbSysIndex(p, c.iyl, jb + 1, kz)) : cd{0.0, 0.0};
  cd dprev = xs;
#pragma unroll
  for (int li = 0; li < Sh::MB; ++li) {
    if (li >= li0 && li < cnt) {
      cd a, cc;
      lapbRowAC<N>(p, c, ktab, blk * Sh::MB + li, kz, kwave, a, cc);
      const cd ri = cmul(Sb[li * Sh::SROW], inv[li]);
      const cd ai = cmul(a, inv[li]);
      dprev = csub(ri, cmul(ai, dprev));
      Sb[li * Sh::SROW] = dprev;
    }
  }
#pragma unroll
  for (int li = Sh::MB - 1; li >= 0; --li) {
    if (li >= li0 && li < cnt) {
      cd a, cc;
      lapbRowAC<N>(p, c, ktab, blk * Sh::MB + li, kz, kwave, a, cc);
      x = csub(Sb[li * Sh::SROW], cmul(cmul(cc, inv[li]), x));
      Sb[li * Sh::SROW] = x;
    }
  }
  if (jb > 0) {
    Sb[0] = xs;
  }
}

// ---- host side: is this configuration served by the block path, and how large are its arrays? ---------------
inline bool lapbShapeOk(int nz) { return nz == 128 || nz == 256 || nz == 512; }
inline int lapbBlocks(int n, int nz) { return (n + lapbRowsPerBlock(nz) - 1) / lapbRowsPerBlock(nz); }
// complex elements of the per-solve / reduced arrays: bif (3) + bxs + bds + bred (LAPB_NRED)
inline size_t lapbSmallElems(int nx, int ny, int nz, int nmode) {
  return static_cast<size_t>(3 + 2 + LAPB_NRED) * ny * lapbBlocks(nx, nz) * nmode;
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
  size_t table, fpcopy, spec, cprime, dprime, inv, bnd, sprime, iface, total;
  size_t nspec; // complex elements of one spectral array
};
inline LapLayout lapLayout(const bout_b200_grid* g, const bout_b200_laplace* L) {
  LapLayout l{};
  const size_t nmode = static_cast<size_t>(L->maxmode) + 1;
  l.nspec = static_cast<size_t>(g->nx) * g->ny * nmode;
  const size_t n2 = static_cast<size_t>(g->nx) * g->ny;
  size_t spec_elems = l.nspec;
  if (lapbShapeOk(g->nz)) {
    const size_t small = lapbSmallElems(g->nx, g->ny, g->nz, static_cast<int>(nmode));
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
  p.nmode = L->maxmode + 1;
  p.zero_dc = (L->global_flags & INV_ZERO_DC) ? 1 : 0;
  p.in_set = (L->first_x && (L->inner_boundary_flags & INV_SET)) ? 1 : 0;
  p.out_set = (L->last_x && (L->outer_boundary_flags & INV_SET)) ? 1 : 0;
  p.nxpe = 1;
  p.pe_xind = 0;

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
    p.bMB = lapbRowsPerBlock(g->nz);
    p.bP = lapbBlocks(n, g->nz);
    const size_t per = static_cast<size_t>(p.ye - p.ys + 1) * p.bP * p.nmode;
    p.bif = p.spec;
    p.bxs = p.bif + 3 * per;
    p.bds = p.bxs + per;
    p.bred = p.bds + per;
  }
}

} // namespace b200
