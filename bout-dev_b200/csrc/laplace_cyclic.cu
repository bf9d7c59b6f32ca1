// LaplaceCyclic (subsystem 3): batched z-FFT in shared memory -> batched complex tridiagonal solve
// per (y, kz) with coefficients regenerated on chip -> inverse z-FFT.
//
// Replaces LaplaceCyclic::solve(Field3D, Field3D)
// (reference src/invert/laplace/impls/cyclic/cyclic_laplace.cxx:187-290), FFTTransform::forward /
// backward (src/invert/laplace/common_transform.cxx:29-115), Laplacian::tridagMatrix
// (src/invert/laplace/invert_laplace.cxx:416-768) and CyclicReduce
// (include/bout/cyclic_reduction.hxx:119-637), including its X-decomposed form.
//
// Data layout: the spectral work arrays are [x - xs][y - ys][kz] complex (kz fastest), i.e. the
// natural output order of the row FFTs.  The reference transposes to [y*nmode+kz][x] so that each
// system is contiguous for its serial CPU solver; on the GPU one thread owns one (y,kz) system and
// walks x, so at every x the 32 lanes of a warp read 32 consecutive kz -- fully coalesced with no
// transpose at all.
//
// Tridiagonal algorithm.  A thread's walk along x is a serial dependency chain, so the kernels are
// written to keep HBM latency off that chain: the kz-independent parts of a, b, c come from a
// per-(x,y) table built once per solve (k_lap_table), the right-hand sides are read from a buffer
// that is never written by the same kernel and are prefetched PF rows ahead into registers.
//   * isolated systems (one rank in X, not periodic): Thomas; forward sweep stores c', d'; back
//     substitution writes x.
//   * coupled systems (X-decomposed over NXPE ranks and/or periodic in X): each rank eliminates
//     downward AND upward in one pass (two independent chains per thread) keeping, per row,
//         x_i + c'_i x_{i+1} + s_i x_L = d'_i          (x_L = last unknown of the rank below)
//     and produces two interface equations in (x_L, x_first, x_last, x_R); the 6 complex
//     coefficients per system and rank are exchanged (one all-gather -- the reference's
//     CyclicReduce needs a gather and a scatter, cyclic_reduction.hxx:215-440), every rank solves
//     the 2*NXPE-unknown interface system redundantly (dense, partial pivoting) and back-substitutes
//     x_i = d'_i - c'_i x_{i+1} - s_i x_L.  Periodic X on one rank is the NXPE = 1 case (x_L = x_last,
//     x_R = x_first).  Both are direct solvers for the reference's matrix and agree with it to
//     round-off (parity tolerance for this path: 1e-10 relative).
#include "laplace_coefs.cuh"
#include "fft_fast.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace b200 {

constexpr int LAP_NK = 6;     // table entries per (x,y): k0..k4 of tridagCoef5 and A
constexpr int LAP_IFACE = 6;  // interface coefficients per system: sL, cL, dL, aU, cU, dU
constexpr int LAP_MAXPE = 8;  // ranks in X

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
  cd* __restrict__ inv;    // fused sweeps: 1 / (b_i - a_i c'_{i-1}), [nxs][nys][nmode]
  cd* __restrict__ bnd;    // fused sweeps: a, c of the boundary rows, [inbndry+outbndry][nys][nmode][2]
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
};

__device__ __forceinline__ RowFFT makeRow(const LapParams& p, unsigned char* smem_raw, int& rloc) {
  const int nbuf = fftBufElems(p.nz);
  rloc = threadIdx.x / p.tpr;
  RowFFT r;
  r.nz = p.nz;
  r.N = p.nz / 2;
  r.pow2 = p.pow2 != 0;
  r.tpr = p.tpr;
  r.lane = threadIdx.x % p.tpr;
  r.A = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * 2 * nbuf;
  r.B = r.A + nbuf;
  r.tw = p.tw;
  r.wsync = p.tpr == 32;
  return r;
}

__device__ __forceinline__ void loadRealRow(const RowFFT& r, const double* __restrict__ src) {
  double* a = reinterpret_cast<double*>(r.A);
  if (r.nz % 2 == 0 && aligned16(src)) {
    for (int j = r.lane; j < r.nz / 2; j += r.tpr) {
      reinterpret_cast<double2*>(a)[j] = __ldg(reinterpret_cast<const double2*>(src) + j);
    }
  } else {
    for (int j = r.lane; j < r.nz; j += r.tpr) {
      a[j] = __ldg(src + j);
    }
  }
}

// K1: rfft of every (x,y) row of rhs (or x0 in boundary rows when INVERT_SET) -> spec
__global__ void __launch_bounds__(128) k_lap_forward(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int rloc;
  const RowFFT r = makeRow(p, smem_raw, rloc);
  const int nys = p.ye - p.ys + 1;
  const int nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  const long long rowid = static_cast<long long>(blockIdx.x) * p.rows + rloc;
  const bool active = rowid < nrows;
  const long long rid = active ? rowid : nrows - 1;
  const int ixl = static_cast<int>(rid / nys), iyl = static_cast<int>(rid % nys);
  const int x = p.bd.xs + ixl, y = p.ys + iyl;
  // common_transform.cxx:49-57
  const bool from_x0 = ((x < p.bd.inbndry) && p.in_set) || ((p.nx - x - 1 < p.bd.outbndry) && p.out_set);
  const double* __restrict__ src = (from_x0 ? p.x0 : p.rhs) + (static_cast<long long>(x) * p.ny + y) * p.nz;
  loadRealRow(r, src);
  rowSync(r);
  const cd* X = rfftForward(r);
  if (active) {
    cd* __restrict__ dst = p.spec + rid * p.nmode;
    for (int k = r.lane; k < p.nmode; k += r.tpr) {
      dst[k] = X[k];
    }
  }
}

// K1 / K3, fast path (nz in {128, 256, 512, 1024}): persistent CTAs, register twiddles, swizzled buffers
template <int N>
__global__ void __launch_bounds__(128) k_lap_forward_fast(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TPR = FftShape<N>::TPR, ROWS = 128 / TPR;
  const int rloc = threadIdx.x / TPR, lane = threadIdx.x % TPR;
  cd* A = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * 2 * (N + 8);
  cd* B = A + (N + 8);
  FastTw<N> tw;
  fastTwInit<N>(tw, p.tw, lane);
  const int nys = p.ye - p.ys + 1;
  const int nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  for (long long base = static_cast<long long>(blockIdx.x) * ROWS; base < nrows;
       base += static_cast<long long>(gridDim.x) * ROWS) {
    const long long rowid = base + rloc;
    const bool active = rowid < nrows;
    const long long rid = active ? rowid : nrows - 1;
    const int ixl = static_cast<int>(rid / nys), iyl = static_cast<int>(rid % nys);
    const int x = p.bd.xs + ixl, y = p.ys + iyl;
    // common_transform.cxx:49-57
    const bool from_x0 = ((x < p.bd.inbndry) && p.in_set) || ((p.nx - x - 1 < p.bd.outbndry) && p.out_set);
    const double* __restrict__ src = (from_x0 ? p.x0 : p.rhs) + (static_cast<long long>(x) * p.ny + y) * p.nz;
    rfftForwardFast<N, SyncGroup>(src, p.spec + rid * p.nmode, p.nmode, active, A, B, tw, lane,
                                  SyncGroup{1 + rloc, TPR});
  }
}

template <int N>
__global__ void __launch_bounds__(128) k_lap_backward_fast(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TPR = FftShape<N>::TPR, ROWS = 128 / TPR;
  const int rloc = threadIdx.x / TPR, lane = threadIdx.x % TPR;
  cd* A = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * 2 * (N + 8);
  cd* B = A + (N + 8);
  FastTw<N> tw;
  fastTwInit<N>(tw, p.tw, lane);
  const int nys = p.ye - p.ys + 1;
  const int nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  for (long long base = static_cast<long long>(blockIdx.x) * ROWS; base < nrows;
       base += static_cast<long long>(gridDim.x) * ROWS) {
    const long long rowid = base + rloc;
    const bool active = rowid < nrows;
    const long long rid = active ? rowid : nrows - 1;
    const int ixl = static_cast<int>(rid / nys), iyl = static_cast<int>(rid % nys);
    const int x = p.bd.xs + ixl, y = p.ys + iyl;
    rfftBackwardFast<N, SyncGroup>(p.spec + rid * p.nmode, p.out + (static_cast<long long>(x) * p.ny + y) * p.nz,
                                   p.nmode, p.zero_dc != 0, active, A, B, tw, lane, SyncGroup{1 + rloc, TPR});
  }
}

__device__ __forceinline__ cd crecip(cd z) {
  const double d = 1.0 / (z.re * z.re + z.im * z.im);
  return {z.re * d, -z.im * d};
}
__device__ __forceinline__ cd ldcd(const cd* __restrict__ p) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  return {t.x, t.y};
}
__device__ __forceinline__ bool tiny(cd z) { return !((z.re * z.re + z.im * z.im) > 1e-300); }

// K0: kz-independent coefficient table (Laplacian::tridagCoefs without the kwave factors, + A)
__global__ void __launch_bounds__(256) k_lap_table(const LapParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.nx * p.ny) {
    return;
  }
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

// Row ix of the system (jy, kz): interior rows from the table, boundary rows (and the doubly
// periodic kz = 0 pin) through the full tridagMatrix logic.  `tab` is the table block of this
// system's jy, indexed by the row number ix (either staged in shared memory or in global memory).
__device__ __forceinline__ void lapRow(const LapParams& p, const double* tab, int ix, int jy, int kz,
                                       double kwave, cd& a, cd& b, cd& c, cd& r) {
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

// Stages the table rows of the (at most two) jy values a CTA's systems belong to into shared memory
// and returns this thread's block; falls back to global memory when the CTA spans more than two
// jy (tiny nmode) -- `smem_ok` is decided by the host for the whole launch.
__device__ __forceinline__ const double* stageTable(const LapParams& p, double* stab, int n, int t, int nsys,
                                                    bool smem_ok) {
  const int tc = min(t, nsys - 1);
  const int jy = p.ys + tc / p.nmode;
  const double* gtab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  if (!smem_ok) {
    return gtab;
  }
  const int t0 = blockIdx.x * blockDim.x;
  const int jy0 = p.ys + t0 / p.nmode;
  const int jy1 = p.ys + min(t0 + static_cast<int>(blockDim.x) - 1, nsys - 1) / p.nmode;
  const int per = n * LAP_NK; // doubles per jy block
  for (int q = 0; q <= jy1 - jy0; ++q) {
    const double2* src = reinterpret_cast<const double2*>(p.table + (static_cast<long long>(jy0 + q) * p.nx + p.bd.xs) * LAP_NK);
    double2* dst = reinterpret_cast<double2*>(stab + static_cast<size_t>(q) * per);
    for (int i = threadIdx.x; i < per / 2; i += blockDim.x) {
      dst[i] = __ldg(src + i);
    }
  }
  __syncthreads();
  return stab + static_cast<size_t>(jy - jy0) * per;
}

#ifndef B200_LAP_PF
#define B200_LAP_PF 8
#endif
constexpr int LAP_PF = B200_LAP_PF; // rows prefetched into registers ahead of the dependency chain

// K2 (isolated systems): one thread per (y, kz) system, Thomas along x.
// Do the matrix coefficients depend on y?  When no metric / coefficient array does, the Thomas factors c' are the
// same for every y and the sweeps keep ONE [x][kz] copy of them (2 MB at 516 x 257: L2-resident) instead of
// streaming an [x][y][kz] array through HBM twice -- a third of the kernel's DRAM traffic.  Checked on the
// device for every solve (the arrays are the caller's and may change between solves); bitwise comparison.
__global__ void __launch_bounds__(256) k_lap_ycheck(const LapParams p) {
  const int nys = p.ye - p.ys + 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nx * nys) {
    return;
  }
  const int x = t / nys, y = p.ys + t % nys;
  const long long i = static_cast<long long>(x) * p.ny + y, i0 = static_cast<long long>(x) * p.ny + p.ys;
  const double* arrs[12] = {p.m.dx, p.m.g11, p.m.g33, p.m.g13, p.m.G1, p.m.G3, p.m.g_11, p.m.g_22,
                            p.m.A,  p.m.C1,  p.m.C2,  p.m.D};
  bool differs = false;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    if (arrs[q] != nullptr) {
      differs = differs || !(__ldg(arrs[q] + i) == __ldg(arrs[q] + i0));
    }
  }
  if (differs) {
    *p.ydep = 1;
  }
}

__global__ void __launch_bounds__(128) k_lap_thomas(const LapParams p, const bool smem_ok) {
  extern __shared__ __align__(16) double stab[];
  const int nys = p.ye - p.ys + 1;
  const int nsys = nys * p.nmode;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = p.bd.xe - p.bd.xs + 1;
  const double* tab = stageTable(p, stab, n, t, nsys, smem_ok);
  if (t >= nsys) {
    return;
  }
  const int iyl = t / p.nmode, kz = t % p.nmode;
  const int jy = p.ys + iyl;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const cd* __restrict__ rin = p.spec + t;
  // y-independent matrix: every y of a kz writes (identical values) to / reads from one [x][kz] copy of c'
  const bool yind = *p.ydep == 0;
  const long long xsc = yind ? p.nmode : xs;
  cd* __restrict__ cp = p.cprime + (yind ? kz : t);
  cd* __restrict__ dp = p.dprime + t;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength; // common_transform.cxx:73
  bool bad = false;

  cd nxt[LAP_PF];
#pragma unroll
  for (int k = 0; k < LAP_PF; ++k) {
    nxt[k] = ldcd(rin + static_cast<long long>(min(k, n - 1)) * xs);
  }
  cd cprev = {0, 0}, dprev = {0, 0};
  for (int j0 = 0; j0 < n; j0 += LAP_PF) {
    cd cur[LAP_PF];
#pragma unroll
    for (int k = 0; k < LAP_PF; ++k) {
      cur[k] = nxt[k];
    }
#pragma unroll
    for (int k = 0; k < LAP_PF; ++k) {
      nxt[k] = ldcd(rin + static_cast<long long>(min(j0 + LAP_PF + k, n - 1)) * xs);
    }
#pragma unroll
    for (int k = 0; k < LAP_PF; ++k) {
      const int ix = j0 + k;
      if (ix < n) {
        cd a, b, c, r = cur[k];
        lapRow(p, tab, ix, jy, kz, kwave, a, b, c, r);
        const cd denom = csub(b, cmul(a, cprev));
        bad = bad || tiny(denom);
        const cd inv = crecip(denom);
        cprev = cmul(c, inv);
        dprev = cmul(csub(r, cmul(a, dprev)), inv);
        cp[ix * xsc] = cprev;
        dp[ix * xs] = dprev;
      }
    }
  }
  // back substitution, reading c', d' (never written below) PF rows ahead
  cd* __restrict__ xo = p.spec + t;
  const cd* __restrict__ cpr = p.cprime + (yind ? kz : t);
  const cd* __restrict__ dpr = p.dprime + t;
  cd x = dprev;
  xo[static_cast<long long>(n - 1) * xs] = x;
  cd nc[LAP_PF], nd[LAP_PF];
#pragma unroll
  for (int k = 0; k < LAP_PF; ++k) {
    const long long o = static_cast<long long>(max(n - 2 - k, 0));
    nc[k] = ldcd(cpr + o * xsc);
    nd[k] = ldcd(dpr + o * xs);
  }
  for (int j0 = n - 2; j0 >= 0; j0 -= LAP_PF) {
    cd cc[LAP_PF], dd[LAP_PF];
#pragma unroll
    for (int k = 0; k < LAP_PF; ++k) {
      cc[k] = nc[k];
      dd[k] = nd[k];
    }
#pragma unroll
    for (int k = 0; k < LAP_PF; ++k) {
      const long long o = static_cast<long long>(max(j0 - LAP_PF - k, 0));
      nc[k] = ldcd(cpr + o * xsc);
      nd[k] = ldcd(dpr + o * xs);
    }
#pragma unroll
    for (int k = 0; k < LAP_PF; ++k) {
      const int ix = j0 - k;
      if (ix >= 0) {
        x = csub(dd[k], cmul(cc[k], x));
        xo[ix * xs] = x;
      }
    }
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

// ---- few long systems (2-D models: ny = 1, e.g. blob2d 1024 x 1024 has 513 systems of 1028 rows) ----------------
// One thread per system leaves the GPU empty and is bound by the 1000-step dependency chain.  Here ONE CTA
// solves ONE system by parallel cyclic reduction in shared memory: ceil(log2 n) steps, in each of which every
// row i eliminates its couplings to rows i -+ s (s = 1, 2, 4, ...), after which x_i = d_i / b_i.  Same rows
// (lapRow) as the Thomas kernel; a different but equally stable elimination order for these diagonally dominant
// systems, so solutions agree with the reference's CyclicReduce to round-off (tests: 1e-10).  Periodic-X systems
// (one rank, n a power of two) wrap the row indices; after the last step (s = n/2) a row's a and c refer to
// x_i itself.
__global__ void __launch_bounds__(256) k_lap_pcr(const LapParams p) {
  extern __shared__ __align__(16) unsigned char pcr_raw[];
  const int n = p.bd.xe - p.bd.xs + 1;
  const bool periodic = p.bd.periodic_x != 0; // single rank, n a power of two (checked by the host)
  cd* buf[2][4]; // [ping-pong][a, b, c, d]
  for (int h = 0; h < 2; ++h) {
    for (int q = 0; q < 4; ++q) {
      buf[h][q] = reinterpret_cast<cd*>(pcr_raw) + static_cast<size_t>(h * 4 + q) * n;
    }
  }
  const int nys = p.ye - p.ys + 1;
  const int t = blockIdx.x; // system index
  const int iyl = t / p.nmode, kz = t % p.nmode;
  const int jy = p.ys + iyl;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength;
  const double* tab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  cd* __restrict__ sp = p.spec + t;
  for (int ix = threadIdx.x; ix < n; ix += blockDim.x) {
    cd a, b, c, r = ldcd(sp + ix * xs);
    lapRow(p, tab, ix, jy, kz, kwave, a, b, c, r);
    if (!periodic && ix == 0) {
      a = {0.0, 0.0};
    }
    if (!periodic && ix == n - 1) {
      c = {0.0, 0.0};
    }
    buf[0][0][ix] = a;
    buf[0][1][ix] = b;
    buf[0][2][ix] = c;
    buf[0][3][ix] = r;
  }
  __syncthreads();
  bool bad = false;
  int cur = 0;
  for (int sft = 1; sft < n; sft <<= 1) {
    cd* const* in = buf[cur];
    cd* const* out = buf[cur ^ 1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const cd a = in[0][i], c = in[2][i];
      cd b = in[1][i], d = in[3][i];
      cd an = {0.0, 0.0}, cn = {0.0, 0.0};
      int im = i - sft, ip = i + sft;
      if (periodic) { // n is a power of two: rows wrap around
        im &= n - 1;
        ip &= n - 1;
      }
      if (im >= 0) {
        const cd bm = in[1][im];
        bad = bad || tiny(bm);
        const cd k1 = cmul(a, crecip(bm));
        an = cmul(in[0][im], k1);
        an = {-an.re, -an.im};
        b = csub(b, cmul(in[2][im], k1));
        d = csub(d, cmul(in[3][im], k1));
      }
      if (ip < n) {
        const cd bp = in[1][ip];
        bad = bad || tiny(bp);
        const cd k2 = cmul(c, crecip(bp));
        cn = cmul(in[2][ip], k2);
        cn = {-cn.re, -cn.im};
        b = csub(b, cmul(in[0][ip], k2));
        d = csub(d, cmul(in[3][ip], k2));
      }
      out[0][i] = an;
      out[1][i] = b;
      out[2][i] = cn;
      out[3][i] = d;
    }
    __syncthreads();
    cur ^= 1;
  }
  // every row is now decoupled: its remaining a and c multiply x_i itself (periodic, s = n) or are zero
  __shared__ double psum[256];
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const cd b = cadd(cadd(buf[cur][0][i], buf[cur][1][i]), buf[cur][2][i]);
    bad = bad || tiny(b);
    const cd x = cmul(buf[cur][3][i], crecip(b));
    buf[cur][3][i] = x;
    part += x.re;
  }
  double avg = 0.0;
  if (periodic && kz == 0) { // subtract the x-average of the kz = 0 mode (cyclic_laplace.cxx:266-287); uniform branch
    psum[threadIdx.x] = part;
    __syncthreads();
    double sum = 0.0;
    for (int q = 0; q < static_cast<int>(blockDim.x); ++q) {
      sum += psum[q];
    }
    avg = sum / static_cast<double>(n);
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) { // rows this thread wrote itself
    cd x = buf[cur][3][i];
    x.re -= avg;
    sp[i * xs] = x;
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

// ---- isolated systems, fused sweeps: the spectrum never visits HBM -----------------------------------
// The Thomas recurrence c'_i = c_i / (b_i - a_i c'_{i-1}) depends on the matrix only, not on the
// right-hand side: k_lap_factor stores inv_i = 1 / (b_i - a_i c'_{i-1}) per (x, y, kz) once
// (bout_b200_laplace_cyclic_factor; every later solve with unchanged coefficients reuses it), which
// takes the division and a 16-deep dependency chain out of every solve.  Two sweep kernels remain.
// One CTA owns one y index (all kz systems of that y share the same input rows) and marches along x
// in batches of eight rows:
//   forward  : rows (x..x+7, y, :) --rfft in shared memory--> 8 x nmode spectrum values on chip;
//              thread kz advances d'_i = (r_i - a_i d'_{i-1}) inv_i and stores d' (the only write);
//   backward : thread kz back-substitutes x_i = d'_i - (c_i inv_i) x_{i+1} into the on-chip spectrum
//              --irfft--> output rows.
// The transform of one batch and the Thomas steps of the neighbouring batch are interleaved stage
// by stage in the same threads (cfftFast2's hook), so the recurrence runs in the shadow of the
// shared-memory latency of the butterflies; row data, inv and d' are register-prefetched one batch
// ahead.  HBM traffic per point: 8 (rhs) + 8 (inv) + 8 (d' out) and 8 (inv) + 8 (d' in) + 8 (out)
// = 48 B (32 B when the coefficients do not depend on y: inv is then served by L2), instead of the
// 3-kernel path's 80 B.  Every row's arithmetic is the same code as the 3-kernel path: results are
// equal to it to round-off (FMA contraction may differ).  Opt-in, see lapFusedEligible.
// CTA = N threads (N = nz/2): four row groups of N/4 threads, each transforming two rows of the
// batch; Thomas thread kz = tid; thread 0 also carries the Nyquist system kz = N, whose inv / d'
// values are staged through shared memory by the last eight threads.
constexpr int LAPF_ROWS = 8;

template <int N>
struct LapFusedShape {
  static constexpr int NT = N;
  static constexpr int SROW = N + 2; // spectrum row pitch (N + 1 modes)
  static constexpr int NSLOT = fftSwaps<N>();                          // hook calls per batch
  static constexpr int RPS = (LAPF_ROWS + NSLOT - 1) / NSLOT;          // Thomas rows per hook call
  static constexpr size_t fftElems = static_cast<size_t>(LAPF_ROWS) * 2 * (N + 8);
  static constexpr size_t specElems = static_cast<size_t>(2) * LAPF_ROWS * SROW; // double buffered
  static constexpr size_t nyqElems = static_cast<size_t>(2) * 2 * LAPF_ROWS;     // [2][inv, d'][row]
  static size_t smemBytes(int n) {
    return (fftElems + specElems + nyqElems) * sizeof(cd) + static_cast<size_t>(n) * LAP_NK * sizeof(double);
  }
};

// Boundary rows are rare (inbndry + outbndry of n) but their coefficient logic is long: k_lap_factor
// stores their a and c per (boundary row, y, kz) in `bnd`, the sweeps read them back, and the
// unrolled hot loops only carry the interior-row formula.  `r` is zeroed where tridagMatrix does.
__device__ __forceinline__ int lapBndRow(const LapParams& p, int ix, bool& inner, bool& outer) {
  const int ncx = p.bd.xe - p.bd.xs;
  inner = p.bd.first_x && ix < p.bd.inbndry;
  outer = p.bd.last_x && (ncx - ix) < p.bd.outbndry;
  return outer ? p.bd.inbndry + (ix - (ncx - p.bd.outbndry + 1)) : ix;
}
__device__ __forceinline__ void lapRowFused(const LapParams& p, const double* tab, int ix, int iyl, int kz,
                                            double kwave, cd& a, cd& c, cd& r) {
  bool inner, outer;
  const int sr = lapBndRow(p, ix, inner, outer);
  if (inner || outer) {
    const int nys = p.ye - p.ys + 1;
    const cd* q = p.bnd + ((static_cast<long long>(sr) * nys + iyl) * p.nmode + kz) * 2;
    a = ldcd(q);
    c = ldcd(q + 1);
    const int fl = outer ? p.bd.outer_flags : p.bd.inner_flags;
    if (!(fl & (INV_RHS | INV_SET))) {
      r = {0.0, 0.0};
    }
    return;
  }
  const double2* t = reinterpret_cast<const double2*>(tab + static_cast<long long>(ix) * LAP_NK);
  const double2 t01 = t[0], t23 = t[1];
  a = {t01.x - t23.y, -kwave * t23.x};
  c = {t01.x + t23.y, kwave * t23.x};
}

// inv_i for every row of every (y, kz) system (the matrix-only half of the Thomas forward sweep)
__global__ void __launch_bounds__(128) k_lap_factor(const LapParams p, const bool smem_ok) {
  extern __shared__ __align__(16) double stab[];
  const int nys = p.ye - p.ys + 1;
  const int nsys = nys * p.nmode;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = p.bd.xe - p.bd.xs + 1;
  const double* tab = stageTable(p, stab, n, t, nsys, smem_ok);
  if (t >= nsys) {
    return;
  }
  const int iyl = t / p.nmode, kz = t % p.nmode;
  const int jy = p.ys + iyl;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength;
  cd* __restrict__ ip = p.inv + t;
  cd cprev = {0, 0};
  bool bad = false;
  for (int ix = 0; ix < n; ++ix) {
    cd a, b, c, r = {0.0, 0.0};
    lapRow(p, tab, ix, jy, kz, kwave, a, b, c, r);
    const cd denom = csub(b, cmul(a, cprev));
    bad = bad || tiny(denom);
    const cd inv = crecip(denom);
    cprev = cmul(c, inv);
    ip[ix * xs] = inv;
    bool inner, outer;
    const int sr = lapBndRow(p, ix, inner, outer);
    if (inner || outer) {
      cd* q = p.bnd + ((static_cast<long long>(sr) * nys + iyl) * p.nmode + kz) * 2;
      q[0] = a;
      q[1] = c;
    }
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

template <int N>
struct LapFusedCtx {
  cd *a0, *b0, *a1, *b1; // transform buffers of this thread group's two rows
  cd* S;                 // [2][LAPF_ROWS][SROW]
  cd* nyq;               // [2][2][LAPF_ROWS]
  double* stab;
  int grp, lane;
};
template <int N>
__device__ __forceinline__ LapFusedCtx<N> lapFusedCtx(unsigned char* smem_raw) {
  using Sh = LapFusedShape<N>;
  LapFusedCtx<N> c;
  c.grp = threadIdx.x / (N / 4);
  c.lane = threadIdx.x % (N / 4);
  cd* bufs = reinterpret_cast<cd*>(smem_raw);
  c.a0 = bufs + static_cast<size_t>(c.grp) * 2 * (N + 8);
  c.b0 = c.a0 + (N + 8);
  c.a1 = bufs + static_cast<size_t>(c.grp + 4) * 2 * (N + 8);
  c.b1 = c.a1 + (N + 8);
  c.S = bufs + Sh::fftElems;
  c.nyq = c.S + Sh::specElems;
  c.stab = reinterpret_cast<double*>(c.nyq + Sh::nyqElems);
  return c;
}

template <int N>
__global__ void __launch_bounds__(N, 1) k_lap_fused_forward(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Sh = LapFusedShape<N>;
  constexpr int TPR = N / 4, SROW = Sh::SROW, R = LAPF_ROWS, RPS = Sh::RPS;
  const int tid = threadIdx.x;
  const LapFusedCtx<N> cx = lapFusedCtx<N>(smem_raw);
  const int lane = cx.lane;
  const int nys = p.ye - p.ys + 1;
  const int n = p.bd.xe - p.bd.xs + 1;
  const int nb = (n + R - 1) / R;
  const int iyl = blockIdx.x, jy = p.ys + iyl;
  { // this y's coefficient table rows
    const double2* src = reinterpret_cast<const double2*>(p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK);
    double2* dst = reinterpret_cast<double2*>(cx.stab);
    for (int i = tid; i < n * (LAP_NK / 2); i += Sh::NT) {
      dst[i] = __ldg(src + i);
    }
  }
  FastTw<N> tw;
  fastTwInit<N>(tw, p.tw, lane);

  const bool sys = tid < p.nmode;
  const bool nyqt = tid == 0 && p.nmode > N;          // second chain of thread 0: kz = N
  const bool nyql = tid >= N - R && p.nmode > N;      // loader of the Nyquist system's inv values
  const double kwave = tid * 2.0 * M_PI / p.m.zlength; // common_transform.cxx:73
  const double kwaveN = N * 2.0 * M_PI / p.m.zlength;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const long long t0 = static_cast<long long>(iyl) * p.nmode;
  const cd* __restrict__ ivp = p.inv + t0 + tid;
  cd* __restrict__ dp = p.dprime + t0 + tid;

  double2 pre[2][4]; // the two rows of batch b this thread group transforms
  cd pinv[R];        // inv of batch b for this thread's system (consumed in iteration b + 1)
  cd pnq = {0.0, 0.0};
  auto prefetch = [&](int b) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ix = b * R + cx.grp + 4 * h;
      if (b < nb && ix < n) {
        const int x = p.bd.xs + ix;
        // common_transform.cxx:49-57
        const bool from_x0 = ((x < p.bd.inbndry) && p.in_set) || ((p.nx - x - 1 < p.bd.outbndry) && p.out_set);
        const double2* __restrict__ s2 = reinterpret_cast<const double2*>(
            (from_x0 ? p.x0 : p.rhs) + (static_cast<long long>(x) * p.ny + jy) * p.nz);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          pre[h][m] = __ldg(s2 + lane + m * TPR);
        }
      } else {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          pre[h][m] = make_double2(0.0, 0.0);
        }
      }
    }
  };
  auto prefetchInv = [&](int b) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int ix = b * R + r;
      pinv[r] = (sys && b < nb && ix < n) ? ldcd(ivp + ix * xs) : cd{0.0, 0.0};
    }
    if (nyql) {
      const int ix = b * R + (tid - (N - R));
      pnq = (b < nb && ix < n) ? ldcd(p.inv + t0 + N + ix * xs) : cd{0.0, 0.0};
    }
  };
  prefetch(0);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    pinv[r] = {0.0, 0.0};
  }
  cd dprev = {0, 0}, dprevN = {0, 0};

  for (int b = 0; b <= nb; ++b) { // iteration b: transform batch b, Thomas steps of batch b - 1
    cd cinv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cinv[r] = pinv[r];
    }
    if (nyql) {
      cx.nyq[(b & 1) * 2 * R + (tid - (N - R))] = pnq; // batch b - 1, read after the barrier below
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      cx.a0[fsw(lane + m * TPR)] = {pre[0][m].x, pre[0][m].y};
      cx.a1[fsw(lane + m * TPR)] = {pre[1][m].x, pre[1][m].y};
    }
    prefetch(b + 1);
    prefetchInv(b);
    __syncthreads(); // rows staged (first time round: the table too); spectrum of batch b - 1 visible
    const cd* Sp = cx.S + ((b - 1) & 1) * R * SROW;
    const cd* nq = cx.nyq + (b & 1) * 2 * R;
    cfftFast2<N, false>(cx.a0, cx.b0, cx.a1, cx.b1, tw, lane, b < nb, [&](int slot) {
      if (b > 0) {
#pragma unroll
        for (int q = 0; q < RPS; ++q) {
          const int r = slot * RPS + q;
          const int ix = (b - 1) * R + r;
          if (r < R && ix < n) {
            if (sys) {
              cd a, c, rr = Sp[r * SROW + tid];
              lapRowFused(p, cx.stab, ix, iyl, tid, kwave, a, c, rr);
              dprev = cmul(csub(rr, cmul(a, dprev)), cinv[r]);
              dp[ix * xs] = dprev;
            }
            if (nyqt) {
              cd a, c, rr = Sp[r * SROW + N];
              lapRowFused(p, cx.stab, ix, iyl, N, kwaveN, a, c, rr);
              dprevN = cmul(csub(rr, cmul(a, dprevN)), nq[r]);
              p.dprime[t0 + N + ix * xs] = dprevN;
            }
          }
        }
      }
    });
    if (b < nb) {
      constexpr bool inB = (fftSwaps<N>() % 2) != 0;
      cd* Sc = cx.S + (b & 1) * R * SROW;
      rfftSplitTo<N>(inB ? cx.b0 : cx.a0, Sc + cx.grp * SROW, p.nmode, tw, lane);
      rfftSplitTo<N>(inB ? cx.b1 : cx.a1, Sc + (cx.grp + 4) * SROW, p.nmode, tw, lane);
      if (!inB) {
        __syncthreads(); // the split read buffer A, which the next batch's rows are staged into
      }
    }
  }
}

template <int N>
__global__ void __launch_bounds__(N, 1) k_lap_fused_backward(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Sh = LapFusedShape<N>;
  constexpr int TPR = N / 4, SROW = Sh::SROW, R = LAPF_ROWS, RPS = Sh::RPS;
  const int tid = threadIdx.x;
  const LapFusedCtx<N> cx = lapFusedCtx<N>(smem_raw);
  const int lane = cx.lane;
  const int nys = p.ye - p.ys + 1;
  const int n = p.bd.xe - p.bd.xs + 1;
  const int nb = (n + R - 1) / R;
  const int iyl = blockIdx.x, jy = p.ys + iyl;
  {
    const double2* src = reinterpret_cast<const double2*>(p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK);
    double2* dst = reinterpret_cast<double2*>(cx.stab);
    for (int i = tid; i < n * (LAP_NK / 2); i += Sh::NT) {
      dst[i] = __ldg(src + i);
    }
  }
  FastTw<N> tw;
  fastTwInit<N>(tw, p.tw, lane);

  const bool sys = tid < p.nmode;
  const bool nyqt = tid == 0 && p.nmode > N;
  const bool nyql = tid >= N - 2 * R && p.nmode > N; // loaders: R inv values, then R d' values
  const double kwave = tid * 2.0 * M_PI / p.m.zlength;
  const double kwaveN = N * 2.0 * M_PI / p.m.zlength;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const long long t0 = static_cast<long long>(iyl) * p.nmode;
  const cd* __restrict__ ivp = p.inv + t0 + tid;
  const cd* __restrict__ dpr = p.dprime + t0 + tid;

  cd pinv[R], pd[R];
  cd pnq = {0.0, 0.0};
  auto prefetch = [&](int b) { // batch b of this thread's system
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int ix = b * R + r;
      const bool ok = sys && b >= 0 && ix < n;
      pinv[r] = ok ? ldcd(ivp + ix * xs) : cd{0.0, 0.0};
      pd[r] = ok ? ldcd(dpr + ix * xs) : cd{0.0, 0.0};
    }
    if (nyql) {
      const int q = tid - (N - 2 * R);
      const int ix = b * R + (q % R);
      const cd* src = (q < R ? p.inv : p.dprime) + t0 + N;
      pnq = (b >= 0 && ix < n) ? ldcd(src + ix * xs) : cd{0.0, 0.0};
    }
  };
  prefetch(nb - 1);
  cd x = {0.0, 0.0}, xN = {0.0, 0.0}; // x_n = 0: the last row of an isolated system has c = 0

  for (int i = 0; i <= nb; ++i) { // iteration i: Thomas steps of batch nb-1-i, inverse transform of batch nb-i
    const int bt = nb - 1 - i, bf = nb - i;
    cd cinv[R], cdp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cinv[r] = pinv[r];
      cdp[r] = pd[r];
    }
    if (nyql) {
      cx.nyq[(i & 1) * 2 * R + (tid - (N - 2 * R))] = pnq;
    }
    prefetch(bt - 1);
    if (i > 0) { // solution spectrum of batch bf (complete since the last barrier of iteration i - 1)
      const cd* Sp = cx.S + ((i - 1) & 1) * R * SROW;
      rfftMergeFrom<N>(Sp + cx.grp * SROW, cx.a0, p.nmode, p.zero_dc != 0, tw, lane);
      rfftMergeFrom<N>(Sp + (cx.grp + 4) * SROW, cx.a1, p.nmode, p.zero_dc != 0, tw, lane);
    }
    __syncthreads();
    cd* Sc = cx.S + (i & 1) * R * SROW;
    const cd* nq = cx.nyq + (i & 1) * 2 * R;
    cfftFast2<N, true>(cx.a0, cx.b0, cx.a1, cx.b1, tw, lane, i > 0, [&](int slot) {
      if (bt >= 0) {
#pragma unroll
        for (int q = 0; q < RPS; ++q) {
          const int r = R - 1 - (slot * RPS + q);
          const int ix = bt * R + r;
          if (r >= 0 && ix < n) {
            if (sys) {
              cd a, c, rr = {0.0, 0.0};
              lapRowFused(p, cx.stab, ix, iyl, tid, kwave, a, c, rr);
              x = csub(cdp[r], cmul(cmul(c, cinv[r]), x));
              Sc[r * SROW + tid] = x;
            }
            if (nyqt) {
              cd a, c, rr = {0.0, 0.0};
              lapRowFused(p, cx.stab, ix, iyl, N, kwaveN, a, c, rr);
              xN = csub(nq[R + r], cmul(cmul(c, nq[r]), xN));
              Sc[r * SROW + N] = xN;
            }
          }
        }
      }
    });
    if (i > 0) {
      constexpr bool inB = (fftSwaps<N>() % 2) != 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ix = bf * R + cx.grp + 4 * h;
        if (ix < n) {
          const cd* z = h == 0 ? (inB ? cx.b0 : cx.a0) : (inB ? cx.b1 : cx.a1);
          double2* __restrict__ d2 = reinterpret_cast<double2*>(
              p.out + (static_cast<long long>(p.bd.xs + ix) * p.ny + jy) * p.nz);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int j = lane + m * TPR;
            const cd v = z[fsw(j)];
            d2[j] = make_double2(v.re, v.im);
          }
        }
      }
    }
    // no barrier: the next merge writes buffer A at exactly the positions this thread read above
    // (or the result sits in buffer B), and S[(i+1)&1] was last read by the merge of iteration i
  }
}

constexpr int LAP_PFC = 4; // coupled kernels carry two or three rings: keep them in registers

// K2a (coupled systems): downward and upward elimination in one pass -> interface equations
//   last row :  x_last  + sL x_L + cL x_R = dL      (downward sweep; also stores c', d', s per row)
//   first row:  x_first + aU x_L + cU x_R = dU      (upward sweep, nothing stored)
__global__ void __launch_bounds__(128) k_lap_reduce(const LapParams p, const bool smem_ok) {
  extern __shared__ __align__(16) double stab[];
  const int nys = p.ye - p.ys + 1;
  const int nsys = nys * p.nmode;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = p.bd.xe - p.bd.xs + 1;
  const double* tab = stageTable(p, stab, n, t, nsys, smem_ok);
  if (t >= nsys) {
    return;
  }
  const int iyl = t / p.nmode, kz = t % p.nmode;
  const int jy = p.ys + iyl;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const cd* __restrict__ rin = p.spec + t;
  cd* __restrict__ cp = p.cprime + t;
  cd* __restrict__ dp = p.dprime + t;
  cd* __restrict__ sp = p.sprime + t;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength;
  bool bad = false;

  cd nd[LAP_PFC], nu[LAP_PFC];
#pragma unroll
  for (int k = 0; k < LAP_PFC; ++k) {
    nd[k] = ldcd(rin + static_cast<long long>(min(k, n - 1)) * xs);
    nu[k] = ldcd(rin + static_cast<long long>(max(n - 1 - k, 0)) * xs);
  }
  cd cprev = {0, 0}, dprev = {0, 0}, sprev = {-1.0, 0};   // s_{-1} = -1 makes s_0 = a_0 / b_0
  cd aup = {0, 0}, rup = {0, 0}, tup = {-1.0, 0};         // t_n = -1 makes t_{n-1} = c_{n-1} / b_{n-1}
  for (int j0 = 0; j0 < n; j0 += LAP_PFC) {
    cd cd_[LAP_PFC], cu_[LAP_PFC];
#pragma unroll
    for (int k = 0; k < LAP_PFC; ++k) {
      cd_[k] = nd[k];
      cu_[k] = nu[k];
    }
#pragma unroll
    for (int k = 0; k < LAP_PFC; ++k) {
      const int j = j0 + LAP_PFC + k;
      nd[k] = ldcd(rin + static_cast<long long>(min(j, n - 1)) * xs);
      nu[k] = ldcd(rin + static_cast<long long>(max(n - 1 - j, 0)) * xs);
    }
#pragma unroll
    for (int k = 0; k < LAP_PFC; ++k) {
      const int j = j0 + k;
      if (j < n) {
        { // downward, row j
          cd a, b, c, r = cd_[k];
          lapRow(p, tab, j, jy, kz, kwave, a, b, c, r);
          const cd denom = csub(b, cmul(a, cprev));
          bad = bad || tiny(denom);
          const cd inv = crecip(denom);
          cprev = cmul(c, inv);
          dprev = cmul(csub(r, cmul(a, dprev)), inv);
          const cd as = cmul(a, sprev);
          sprev = cmul(cd{-as.re, -as.im}, inv);
          cp[j * xs] = cprev;
          dp[j * xs] = dprev;
          sp[j * xs] = sprev;
        }
        { // upward, row n-1-j
          const int i = n - 1 - j;
          cd a, b, c, r = cu_[k];
          lapRow(p, tab, i, jy, kz, kwave, a, b, c, r);
          const cd denom = csub(b, cmul(c, aup));
          bad = bad || tiny(denom);
          const cd inv = crecip(denom);
          aup = cmul(a, inv);
          rup = cmul(csub(r, cmul(c, rup)), inv);
          const cd ct = cmul(c, tup);
          tup = cmul(cd{-ct.re, -ct.im}, inv);
        }
      }
    }
  }
  cd* __restrict__ f = p.iface + t;
  f[0 * nsys] = sprev;
  f[1 * nsys] = cprev;
  f[2 * nsys] = dprev;
  f[3 * nsys] = aup;
  f[4 * nsys] = tup;
  f[5 * nsys] = rup;
  if (bad) {
    atomicExch(p.status, 1);
  }
}

// K2b (coupled systems): interface system over the NXPE ranks (dense, partial pivoting), then back
// substitution of this rank's rows.
__global__ void __launch_bounds__(128) k_lap_finish(const LapParams p) {
  const int nys = p.ye - p.ys + 1;
  const int nsys = nys * p.nmode;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsys) {
    return;
  }
  const int kz = t % p.nmode;
  const int n = p.bd.xe - p.bd.xs + 1;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const int P = p.nxpe, N2 = 2 * P;
  const bool per = p.bd.periodic_x != 0;
  bool bad = false;
  // unknowns: z[2q] = first unknown of rank q, z[2q+1] = last unknown of rank q
  cd M[2 * LAP_MAXPE][2 * LAP_MAXPE + 1];
  for (int i = 0; i < N2; ++i) {
    for (int j = 0; j <= N2; ++j) {
      M[i][j] = {0.0, 0.0};
    }
  }
  for (int q = 0; q < P; ++q) {
    const cd* __restrict__ f = p.iface_all + (static_cast<long long>(q) * LAP_IFACE) * nsys + t;
    const cd sL = ldcd(f), cL = ldcd(f + nsys), dL = ldcd(f + 2 * nsys);
    const cd aU = ldcd(f + 3 * nsys), cU = ldcd(f + 4 * nsys), dU = ldcd(f + 5 * nsys);
    const int prev = q > 0 ? q - 1 : (per ? P - 1 : -1);
    const int next = q < P - 1 ? q + 1 : (per ? 0 : -1);
    M[2 * q][2 * q].re += 1.0;
    M[2 * q + 1][2 * q + 1].re += 1.0;
    if (prev >= 0) {
      M[2 * q][2 * prev + 1] = cadd(M[2 * q][2 * prev + 1], aU);
      M[2 * q + 1][2 * prev + 1] = cadd(M[2 * q + 1][2 * prev + 1], sL);
    }
    if (next >= 0) {
      M[2 * q][2 * next] = cadd(M[2 * q][2 * next], cU);
      M[2 * q + 1][2 * next] = cadd(M[2 * q + 1][2 * next], cL);
    }
    M[2 * q][N2] = dU;
    M[2 * q + 1][N2] = dL;
  }
  for (int col = 0; col < N2; ++col) { // Gaussian elimination, partial pivoting
    int piv = col;
    double best = M[col][col].re * M[col][col].re + M[col][col].im * M[col][col].im;
    for (int r = col + 1; r < N2; ++r) {
      const double v = M[r][col].re * M[r][col].re + M[r][col].im * M[r][col].im;
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    bad = bad || !(best > 1e-300);
    if (piv != col) {
      for (int j = col; j <= N2; ++j) {
        const cd tmp = M[col][j];
        M[col][j] = M[piv][j];
        M[piv][j] = tmp;
      }
    }
    const cd inv = crecip(M[col][col]);
    for (int j = col; j <= N2; ++j) {
      M[col][j] = cmul(M[col][j], inv);
    }
    for (int r = 0; r < N2; ++r) {
      if (r != col) {
        const cd fct = M[r][col];
        if (fct.re != 0.0 || fct.im != 0.0) {
          for (int j = col; j <= N2; ++j) {
            M[r][j] = csub(M[r][j], cmul(fct, M[col][j]));
          }
        }
      }
    }
  }
  const int q = p.pe_xind;
  const int prev = q > 0 ? q - 1 : (per ? P - 1 : -1);
  const cd xL = prev >= 0 ? M[2 * prev + 1][N2] : cd{0.0, 0.0};
  cd x = M[2 * q + 1][N2]; // this rank's last unknown

  cd* __restrict__ xo = p.spec + t;
  const cd* __restrict__ cpr = p.cprime + t;
  const cd* __restrict__ dpr = p.dprime + t;
  const cd* __restrict__ spr = p.sprime + t;
  xo[static_cast<long long>(n - 1) * xs] = x;
  double sum = x.re;
  cd nc[LAP_PFC], ndd[LAP_PFC], ns[LAP_PFC];
#pragma unroll
  for (int k = 0; k < LAP_PFC; ++k) {
    const long long o = static_cast<long long>(max(n - 2 - k, 0)) * xs;
    nc[k] = ldcd(cpr + o);
    ndd[k] = ldcd(dpr + o);
    ns[k] = ldcd(spr + o);
  }
  for (int j0 = n - 2; j0 >= 0; j0 -= LAP_PFC) {
    cd cc[LAP_PFC], dd[LAP_PFC], ss[LAP_PFC];
#pragma unroll
    for (int k = 0; k < LAP_PFC; ++k) {
      cc[k] = nc[k];
      dd[k] = ndd[k];
      ss[k] = ns[k];
    }
#pragma unroll
    for (int k = 0; k < LAP_PFC; ++k) {
      const long long o = static_cast<long long>(max(j0 - LAP_PFC - k, 0)) * xs;
      nc[k] = ldcd(cpr + o);
      ndd[k] = ldcd(dpr + o);
      ns[k] = ldcd(spr + o);
    }
#pragma unroll
    for (int k = 0; k < LAP_PFC; ++k) {
      const int ix = j0 - k;
      if (ix >= 0) {
        x = csub(csub(dd[k], cmul(cc[k], x)), cmul(ss[k], xL));
        xo[ix * xs] = x;
        sum += x.re;
      }
    }
  }
  if (per && kz == 0 && P == 1) { // subtract the x-average of the kz=0 mode (cyclic_laplace.cxx:266-287)
    const double avg = sum / static_cast<double>(n);
    for (int ix = 0; ix < n; ++ix) {
      xo[ix * xs].re -= avg;
    }
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

// K3: spec -> irfft -> out rows (common_transform.cxx:82-115)
__global__ void __launch_bounds__(128) k_lap_backward(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int rloc;
  const RowFFT r = makeRow(p, smem_raw, rloc);
  const int nys = p.ye - p.ys + 1;
  const int nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  const long long rowid = static_cast<long long>(blockIdx.x) * p.rows + rloc;
  const bool active = rowid < nrows;
  const long long rid = active ? rowid : nrows - 1;
  const int ixl = static_cast<int>(rid / nys), iyl = static_cast<int>(rid % nys);
  const int x = p.bd.xs + ixl, y = p.ys + iyl;
  const cd* __restrict__ src = p.spec + rid * p.nmode;
  for (int k = r.lane; k <= p.nz / 2; k += r.tpr) {
    cd v = {0.0, 0.0};
    if (k < p.nmode && !(k == 0 && p.zero_dc)) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(src) + k);
      v = {t.x, t.y};
    }
    r.A[k] = v;
  }
  rowSync(r);
  const double* res = rfftBackward(r, r.A);
  if (active) {
    double* __restrict__ dst = p.out + (static_cast<long long>(x) * p.ny + y) * p.nz;
    if (p.nz % 2 == 0 && aligned16(dst)) {
      for (int j = r.lane; j < p.nz / 2; j += r.tpr) {
        reinterpret_cast<double2*>(dst)[j] = reinterpret_cast<const double2*>(res)[j];
      }
    } else {
      for (int j = r.lane; j < p.nz; j += r.tpr) {
        dst[j] = res[j];
      }
    }
  }
}

} // namespace b200

using namespace b200;

static void lapBndryWidths(const bout_b200_grid* g, const bout_b200_laplace* L, int* inb, int* outb) {
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

static bool lapCoupled(const bout_b200_laplace* L) {
  return L->periodic_x || !L->first_x || !L->last_x;
}

static size_t roundUp(size_t n, size_t a) { return (n + a - 1) / a * a; }

extern "C" size_t bout_b200_laplace_cyclic_workspace(const bout_b200_grid* g, const bout_b200_laplace* L) {
  if (g == nullptr || L == nullptr) {
    return 0;
  }
  const size_t nspec = static_cast<size_t>(g->nx) * g->ny * (L->maxmode + 1) * sizeof(cd);
  const size_t table = roundUp(static_cast<size_t>(g->nx) * g->ny * LAP_NK * sizeof(double), 256);
  const size_t nbnd = static_cast<size_t>(4) * g->xstart * g->ny * (L->maxmode + 1) * sizeof(cd); // boundary rows' a, c
  size_t n = 256 + table + 4 * nspec + nbnd; // spec, c', d', inv
  if (lapCoupled(L)) {
    n += nspec + static_cast<size_t>(LAP_IFACE) * g->ny * (L->maxmode + 1) * sizeof(cd);
  }
  return n;
}

extern "C" size_t bout_b200_laplace_cyclic_iface_doubles(const bout_b200_grid* g, const bout_b200_laplace* L) {
  if (g == nullptr || L == nullptr) {
    return 0;
  }
  const int ys = L->include_yguards_lower ? 0 : g->ystart;
  const int ye = L->include_yguards_upper ? g->ny - 1 : g->yend;
  return static_cast<size_t>(2 * LAP_IFACE) * (ye - ys + 1) * (L->maxmode + 1);
}

namespace {
struct LapPlan {
  LapParams p;
  RowGeom geo;
  unsigned rblocks;
  int nsys;
  size_t tab_smem; // bytes of shared memory for the staged coefficient table (0 = read it from global memory)
};

int lapSetup(const bout_b200_grid* g, const bout_b200_laplace* L, void* workspace, size_t workspace_bytes,
             cudaStream_t s, LapPlan* plan) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(L && workspace, BOUT_B200_ERR_ARG, "laplace: null argument");
  B200_REQUIRE(L->dx && L->g11 && L->g33 && L->g13, BOUT_B200_ERR_ARG, "laplace: metric missing");
  B200_REQUIRE(!L->all_terms || (L->G1 && L->G3), BOUT_B200_ERR_ARG, "laplace: all_terms needs G1, G3");
  B200_REQUIRE(L->maxmode >= 0 && L->maxmode <= g->nz / 2, BOUT_B200_ERR_ARG, "laplace: bad maxmode");
  B200_REQUIRE(workspace_bytes >= bout_b200_laplace_cyclic_workspace(g, L), BOUT_B200_ERR_ARG,
               "laplace: workspace too small");
  B200_REQUIRE(aligned16(workspace), BOUT_B200_ERR_ARG, "laplace: workspace must be 16-byte aligned");
  const int gradflags = INV_DC_GRAD | INV_AC_GRAD;
  if (((L->inner_boundary_flags | L->outer_boundary_flags) & gradflags)
      && ((L->inner_boundary_flags | L->outer_boundary_flags) & (INV_SET | INV_RHS))) {
    B200_REQUIRE(L->g_11 != nullptr, BOUT_B200_ERR_ARG, "laplace: gradient-with-value boundary needs g_11");
  }
  if ((L->inner_boundary_flags | L->outer_boundary_flags) & (INV_DC_GRADPAR | INV_DC_GRADPARINV)) {
    B200_REQUIRE(L->g_22 != nullptr, BOUT_B200_ERR_ARG, "laplace: GRADPAR boundary needs g_22");
  }
  LapParams& p = plan->p;
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

  const size_t nspec = static_cast<size_t>(g->nx) * g->ny * p.nmode;
  const size_t table = roundUp(static_cast<size_t>(g->nx) * g->ny * LAP_NK * sizeof(double), 256);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  p.status = reinterpret_cast<int*>(ws);
  p.ydep = p.status + 1;
  p.table = reinterpret_cast<double*>(ws + 256);
  p.spec = reinterpret_cast<cd*>(ws + 256 + table);
  p.cprime = p.spec + nspec;
  p.dprime = p.cprime + nspec;
  p.inv = p.dprime + nspec;
  p.bnd = p.inv + nspec;
  if (lapCoupled(L)) {
    p.sprime = p.bnd + static_cast<size_t>(4) * g->xstart * g->ny * p.nmode;
    p.iface = p.sprime + nspec;
  }
  int st = 0;
  p.tw = twiddleTable(g->nz, s, &st);
  if (st != 0) {
    return st;
  }
  plan->geo = rowGeometry(g->nz, 128); // (one warp per row measured slower: 315 vs 280 us at nz = 512)
  p.tpr = plan->geo.tpr;
  p.rows = plan->geo.rows;
  p.pow2 = plan->geo.pow2 ? 1 : 0;
  B200_REQUIRE(plan->geo.smemBytes <= 200 * 1024, BOUT_B200_ERR_UNSUPPORTED, "laplace: nz=%d too large", g->nz);
  if (plan->geo.smemBytes > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_forward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(plan->geo.smemBytes)));
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_backward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(plan->geo.smemBytes)));
  }
  const int nys = p.ye - p.ys + 1, nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  plan->rblocks = static_cast<unsigned>((nrows + plan->geo.rows - 1) / plan->geo.rows);
  plan->nsys = nys * p.nmode;
  // a CTA of 128 consecutive systems spans at most two jy when nmode >= 127
  plan->tab_smem = 0;
  const size_t tab_bytes = static_cast<size_t>(2) * nxs * LAP_NK * sizeof(double);
  if (p.nmode >= 127 && tab_bytes <= 160 * 1024) {
    plan->tab_smem = tab_bytes;
    if (tab_bytes > 48 * 1024) {
      B200_CUDA_OK(cudaFuncSetAttribute(k_lap_thomas, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(tab_bytes)));
      B200_CUDA_OK(cudaFuncSetAttribute(k_lap_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(tab_bytes)));
    }
  }
  return 0;
}

// FFT launches: fast path for nz in {128, 256, 512, 1024} with 16-byte aligned rows
template <bool FORWARD>
int launchFft(const LapPlan* plan, cudaStream_t s) {
  const LapParams& p = plan->p;
  const int N = p.nz / 2;
  const bool aligned = FORWARD ? (aligned16(p.rhs) && (p.x0 == nullptr || aligned16(p.x0))) : aligned16(p.out);
  static const bool no_fast = getenv("BOUT_B200_NO_FAST_FFT") != nullptr;
  if (!no_fast && p.pow2 && aligned && (N == 64 || N == 128 || N == 256 || N == 512)) {
    const int rows = 128 / (N / 4);
    const long long nrows = static_cast<long long>(p.bd.xe - p.bd.xs + 1) * (p.ye - p.ys + 1);
    const long long want = (nrows + rows - 1) / rows;
    const unsigned blocks = static_cast<unsigned>(want < 148LL * 12 ? want : 148LL * 12);
    const size_t smem = static_cast<size_t>(rows) * 2 * (N + 8) * sizeof(cd);
#define B200_FFT_CASE(NN)                                                                              \
  case NN:                                                                                             \
    if (FORWARD) {                                                                                     \
      k_lap_forward_fast<NN><<<blocks, 128, smem, s>>>(p);                                             \
    } else {                                                                                           \
      k_lap_backward_fast<NN><<<blocks, 128, smem, s>>>(p);                                            \
    }                                                                                                  \
    break;
    switch (N) {
      B200_FFT_CASE(64)
      B200_FFT_CASE(128)
      B200_FFT_CASE(256)
      B200_FFT_CASE(512)
    }
#undef B200_FFT_CASE
  } else if (FORWARD) {
    k_lap_forward<<<plan->rblocks, 128, plan->geo.smemBytes, s>>>(p);
  } else {
    k_lap_backward<<<plan->rblocks, 128, plan->geo.smemBytes, s>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}

// Fused sweeps (k_lap_fused_forward / _backward): one CTA per solved y row, so the path needs enough
// y rows to occupy the SMs; below the threshold the row-parallel 3-kernel path is faster.
// BOUT_B200_LAP_FUSED=0 / 1 forces the choice (tests run both).
template <int N>
int lapFusedFits(int n) {
  return LapFusedShape<N>::smemBytes(n) <= 220 * 1024;
}
bool lapFusedEligible(const LapPlan* plan, const bout_b200_laplace* L) {
  const LapParams& p = plan->p;
  const int N = p.nz / 2, n = p.bd.xe - p.bd.xs + 1, nys = p.ye - p.ys + 1;
  if (lapCoupled(L) || !p.pow2 || !(N == 64 || N == 128 || N == 256 || N == 512)
      || n <= p.bd.inbndry + p.bd.outbndry) {
    return false;
  }
  const bool fits = N == 64 ? lapFusedFits<64>(n) : N == 128 ? lapFusedFits<128>(n)
                  : N == 256 ? lapFusedFits<256>(n) : lapFusedFits<512>(n);
  if (!fits) {
    return false;
  }
  // Measured on B200 at 516x132x512 (tools/time_laplace.py): 3-kernel path 592 us; fused sweeps 650 us
  // on a held factorisation, 785 us one-shot.  One CTA of 8 warps per SM leaves the butterflies and
  // the recurrence latency-bound (issue slots 25 % busy, FP64 pipe 17 %), which costs more than the
  // 40 % of HBM traffic the sweeps save -- so the path is opt-in (BOUT_B200_LAP_FUSED=1) until it
  // runs >= 16 warps per SM.
  (void)nys;
  const char* env = getenv("BOUT_B200_LAP_FUSED");
  return env != nullptr && env[0] == '1';
}
bool lapUseFused(const LapPlan* plan, const bout_b200_laplace* L) {
  const LapParams& p = plan->p;
  if (!aligned16(p.rhs) || (p.x0 != nullptr && !aligned16(p.x0)) || !aligned16(p.out)) {
    return false;
  }
  return lapFusedEligible(plan, L);
}

// Factorisations held in caller workspaces (bout_b200_laplace_cyclic_factor): workspace -> the grid
// and settings it was computed for.  Pointer members are compared by value: the caller vouches for
// the contents of the metric / coefficient arrays being unchanged (it re-factors when it changes them).
struct FactorKey {
  bout_b200_grid g;
  bout_b200_laplace L;
  size_t bytes;
};
std::mutex g_factor_mutex;
std::unordered_map<const void*, FactorKey> g_factors;

bool factorMatches(const void* ws, const bout_b200_grid* g, const bout_b200_laplace* L, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_factor_mutex);
  auto it = g_factors.find(ws);
  if (it == g_factors.end()) {
    return false;
  }
  FactorKey k{};
  k.g = *g;
  k.L = *L;
  k.bytes = bytes;
  return it->second.bytes == bytes && memcmp(&it->second.g, &k.g, sizeof(k.g)) == 0
         && it->second.L.dx == L->dx && it->second.L.g11 == L->g11 && it->second.L.g33 == L->g33
         && it->second.L.g13 == L->g13 && it->second.L.G1 == L->G1 && it->second.L.G3 == L->G3
         && it->second.L.g_11 == L->g_11 && it->second.L.g_22 == L->g_22 && it->second.L.A == L->A
         && it->second.L.C1 == L->C1 && it->second.L.C2 == L->C2 && it->second.L.D == L->D
         && it->second.L.zlength == L->zlength && it->second.L.all_terms == L->all_terms
         && it->second.L.nonuniform == L->nonuniform && it->second.L.global_flags == L->global_flags
         && it->second.L.inner_boundary_flags == L->inner_boundary_flags
         && it->second.L.outer_boundary_flags == L->outer_boundary_flags && it->second.L.maxmode == L->maxmode
         && it->second.L.periodic_x == L->periodic_x && it->second.L.first_x == L->first_x
         && it->second.L.last_x == L->last_x
         && it->second.L.include_yguards_lower == L->include_yguards_lower
         && it->second.L.include_yguards_upper == L->include_yguards_upper;
}

// status reset + coefficient table + inv (the matrix-only half of the solve)
int lapFactor(LapPlan* plan, cudaStream_t s) {
  LapParams& p = plan->p;
  B200_CUDA_OK(cudaMemsetAsync(p.status, 0, sizeof(int), s));
  const int n2 = p.nx * p.ny;
  k_lap_table<<<(n2 + 255) / 256, 256, 0, s>>>(p);
  B200_LAUNCH_OK();
  if (plan->tab_smem > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_factor, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(plan->tab_smem)));
  }
  k_lap_factor<<<static_cast<unsigned>((plan->nsys + 127) / 128), 128, plan->tab_smem, s>>>(p, plan->tab_smem != 0);
  B200_LAUNCH_OK();
  return 0;
}

template <int N>
int launchFusedT(const LapPlan* plan, cudaStream_t s) {
  const LapParams& p = plan->p;
  const int n = p.bd.xe - p.bd.xs + 1, nys = p.ye - p.ys + 1;
  const size_t smem = LapFusedShape<N>::smemBytes(n);
  static size_t attr = 0; // largest size the two kernels have been configured for
  if (smem > 48 * 1024 && smem > attr) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_fused_forward<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_fused_backward<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    attr = smem;
  }
  k_lap_fused_forward<N><<<nys, N, smem, s>>>(p);
  B200_LAUNCH_OK();
  k_lap_fused_backward<N><<<nys, N, smem, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}
int launchFused(const LapPlan* plan, cudaStream_t s) {
  switch (plan->p.nz / 2) {
  case 64: return launchFusedT<64>(plan, s);
  case 128: return launchFusedT<128>(plan, s);
  case 256: return launchFusedT<256>(plan, s);
  default: return launchFusedT<512>(plan, s);
  }
}

// status reset + coefficient table (+ forward FFT of the 3-kernel paths)
int lapForward(const bout_b200_laplace* L, const double* rhs, const double* x0, LapPlan* plan, cudaStream_t s,
               bool fft = true) {
  LapParams& p = plan->p;
  B200_REQUIRE(rhs != nullptr, BOUT_B200_ERR_ARG, "laplace: null rhs");
  B200_REQUIRE(!((L->inner_boundary_flags | L->outer_boundary_flags) & INV_SET) || x0 != nullptr,
               BOUT_B200_ERR_ARG, "laplace: INVERT_SET needs x0");
  p.rhs = rhs;
  p.x0 = x0;
  B200_CUDA_OK(cudaMemsetAsync(p.status, 0, 2 * sizeof(int), s)); // status and ydep
  const int n2 = p.nx * p.ny;
  k_lap_table<<<(n2 + 255) / 256, 256, 0, s>>>(p);
  B200_LAUNCH_OK();
  const int nys = p.ye - p.ys + 1;
  if (nys > 1) {
    k_lap_ycheck<<<(p.nx * nys + 255) / 256, 256, 0, s>>>(p);
    B200_LAUNCH_OK();
  }
  return fft ? launchFft<true>(plan, s) : 0;
}
} // namespace

extern "C" int bout_b200_laplace_cyclic_solve(const bout_b200_grid* g, const bout_b200_laplace* L,
                                              const double* rhs, const double* x0, double* out,
                                              void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = asStream(stream);
  LapPlan plan;
  int rc = lapSetup(g, L, workspace, workspace_bytes, s, &plan);
  if (rc != 0) {
    return rc;
  }
  B200_REQUIRE(out != nullptr, BOUT_B200_ERR_ARG, "laplace: null output");
  B200_REQUIRE(L->first_x && L->last_x, BOUT_B200_ERR_UNSUPPORTED,
               "laplace: X-decomposed solves go through bout_b200_laplace_cyclic_reduce/_finish");
  LapParams& p = plan.p;
  p.rhs = rhs;
  p.x0 = x0;
  p.out = out;
  if (lapUseFused(&plan, L)) {
    B200_REQUIRE(rhs != nullptr, BOUT_B200_ERR_ARG, "laplace: null rhs");
    B200_REQUIRE(!((L->inner_boundary_flags | L->outer_boundary_flags) & INV_SET) || x0 != nullptr,
                 BOUT_B200_ERR_ARG, "laplace: INVERT_SET needs x0");
    if (!factorMatches(workspace, g, L, workspace_bytes)) { // one-shot: factor as part of this solve
      rc = lapFactor(&plan, s);
      if (rc != 0) {
        return rc;
      }
    }
    return launchFused(&plan, s);
  }
  {
    std::lock_guard<std::mutex> lock(g_factor_mutex);
    g_factors.erase(workspace); // the 3-kernel paths overwrite the table and status
  }
  rc = lapForward(L, rhs, x0, &plan, s);
  if (rc != 0) {
    return rc;
  }
  const unsigned sblocks = static_cast<unsigned>((plan.nsys + 127) / 128);
  const int nrows_sys = p.bd.xe - p.bd.xs + 1;
  const size_t pcr_smem = static_cast<size_t>(8) * nrows_sys * sizeof(cd);
  static const bool no_pcr = getenv("BOUT_B200_LAP_NO_PCR") != nullptr;
  const bool pcr_shape = !lapCoupled(L) || (L->periodic_x && (nrows_sys & (nrows_sys - 1)) == 0);
  if (pcr_shape && !no_pcr && plan.nsys <= 2048 && nrows_sys >= 8 && pcr_smem <= 200 * 1024) {
    static size_t pcr_attr = 0;
    if (pcr_smem > 48 * 1024 && pcr_smem > pcr_attr) {
      B200_CUDA_OK(cudaFuncSetAttribute(k_lap_pcr, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(pcr_smem)));
      pcr_attr = pcr_smem;
    }
    k_lap_pcr<<<static_cast<unsigned>(plan.nsys), 256, pcr_smem, s>>>(p);
    B200_LAUNCH_OK();
  } else if (!lapCoupled(L)) {
    k_lap_thomas<<<sblocks, 128, plan.tab_smem, s>>>(p, plan.tab_smem != 0);
    B200_LAUNCH_OK();
  } else { // periodic in X on a single rank: the NXPE = 1 case of the coupled algorithm
    k_lap_reduce<<<sblocks, 128, plan.tab_smem, s>>>(p, plan.tab_smem != 0);
    B200_LAUNCH_OK();
    p.iface_all = p.iface;
    k_lap_finish<<<sblocks, 128, 0, s>>>(p);
    B200_LAUNCH_OK();
  }
  return launchFft<false>(&plan, s);
}

extern "C" int bout_b200_laplace_cyclic_factor(const bout_b200_grid* g, const bout_b200_laplace* L,
                                               void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = asStream(stream);
  LapPlan plan;
  int rc = lapSetup(g, L, workspace, workspace_bytes, s, &plan);
  if (rc != 0) {
    return rc;
  }
  {
    std::lock_guard<std::mutex> lock(g_factor_mutex);
    g_factors.erase(workspace);
  }
  if (!lapFusedEligible(&plan, L)) {
    return 0; // this configuration regenerates its coefficients inside every solve
  }
  rc = lapFactor(&plan, s);
  if (rc != 0) {
    return rc;
  }
  FactorKey k{};
  k.g = *g;
  k.L = *L;
  k.bytes = workspace_bytes;
  std::lock_guard<std::mutex> lock(g_factor_mutex);
  g_factors[workspace] = k;
  return 0;
}

extern "C" int bout_b200_laplace_cyclic_invalidate(void* workspace) {
  std::lock_guard<std::mutex> lock(g_factor_mutex);
  g_factors.erase(workspace);
  return 0;
}

extern "C" int bout_b200_laplace_cyclic_reduce(const bout_b200_grid* g, const bout_b200_laplace* L,
                                               const double* rhs, const double* x0, void* workspace,
                                               size_t workspace_bytes, double* iface_out, void* stream) {
  cudaStream_t s = asStream(stream);
  LapPlan plan;
  int rc = lapSetup(g, L, workspace, workspace_bytes, s, &plan);
  if (rc != 0) {
    return rc;
  }
  B200_REQUIRE(iface_out != nullptr && aligned16(iface_out), BOUT_B200_ERR_ARG, "laplace reduce: bad iface_out");
  B200_REQUIRE(lapCoupled(L), BOUT_B200_ERR_ARG, "laplace reduce: system is not coupled (use _solve)");
  rc = lapForward(L, rhs, x0, &plan, s);
  if (rc != 0) {
    return rc;
  }
  plan.p.iface = reinterpret_cast<cd*>(iface_out);
  k_lap_reduce<<<static_cast<unsigned>((plan.nsys + 127) / 128), 128, plan.tab_smem, s>>>(plan.p, plan.tab_smem != 0);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_laplace_cyclic_finish(const bout_b200_grid* g, const bout_b200_laplace* L,
                                               const double* iface_all, int nxpe, int pe_xind, double* out,
                                               void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = asStream(stream);
  LapPlan plan;
  int rc = lapSetup(g, L, workspace, workspace_bytes, s, &plan);
  if (rc != 0) {
    return rc;
  }
  B200_REQUIRE(out != nullptr && iface_all != nullptr && aligned16(iface_all), BOUT_B200_ERR_ARG,
               "laplace finish: null argument");
  B200_REQUIRE(nxpe >= 1 && nxpe <= LAP_MAXPE && pe_xind >= 0 && pe_xind < nxpe, BOUT_B200_ERR_ARG,
               "laplace finish: NXPE must be 1..%d", LAP_MAXPE);
  B200_REQUIRE(lapCoupled(L), BOUT_B200_ERR_ARG, "laplace finish: system is not coupled");
  B200_REQUIRE(!(L->periodic_x && nxpe > 1), BOUT_B200_ERR_UNSUPPORTED,
               "laplace: periodic X over several ranks is not supported yet (kz=0 mean needs a reduction)");
  LapParams& p = plan.p;
  p.out = out;
  p.iface_all = reinterpret_cast<const cd*>(iface_all);
  p.nxpe = nxpe;
  p.pe_xind = pe_xind;
  k_lap_finish<<<static_cast<unsigned>((plan.nsys + 127) / 128), 128, 0, s>>>(p);
  B200_LAUNCH_OK();
  return launchFft<false>(&plan, s);
}

// Synchronises the stream and reports a zero pivot seen by the last solve that used `workspace`
// (reference: BoutException "Zero pivot in CyclicReduce::reduce", cyclic_reduction.hxx:553).
extern "C" int bout_b200_laplace_cyclic_status(void* workspace, void* stream) {
  B200_REQUIRE(workspace != nullptr, BOUT_B200_ERR_ARG, "laplace status: null workspace");
  int flag = 0;
  B200_CUDA_OK(cudaMemcpyAsync(&flag, workspace, sizeof(int), cudaMemcpyDeviceToHost, asStream(stream)));
  B200_CUDA_OK(cudaStreamSynchronize(asStream(stream)));
  if (flag != 0) {
    return fail(BOUT_B200_ERR_PIVOT, "Zero pivot in the cyclic tridiagonal solve");
  }
  return 0;
}
