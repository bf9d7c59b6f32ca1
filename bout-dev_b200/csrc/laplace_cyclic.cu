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
#include "laplace_block.cuh"
#include "fft_fast.cuh"

#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace cg = cooperative_groups;

namespace b200 {

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

// K0: kz-independent coefficient table (Laplacian::tridagCoefs without the kwave factors, + A)
__global__ void __launch_bounds__(256) k_lap_table(const LapParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.nx * p.ny) {
    lapTableEntry(p, i);
  }
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
constexpr int LAP_PCR_NT = 512; // the 8 n complex numbers of a system allow one CTA per SM anyway: make it a wide one
// 1 / z through the Newton-refined hardware reciprocal (common.cuh): each row inverts two pivots in each of the
// log2 n steps
__device__ __forceinline__ cd crecipFast(cd z) {
  const double d = rcpPositive(z.re * z.re + z.im * z.im);
  return {z.re * d, -z.im * d};
}
__global__ void __launch_bounds__(LAP_PCR_NT) k_lap_pcr(const LapParams p) {
  extern __shared__ __align__(16) unsigned char pcr_raw[];
  const int n = p.bd.xe - p.bd.xs + 1;
  const bool periodic = p.bd.periodic_x != 0; // single rank, n a power of two (checked by the host)
  // Rows are kept NORMALISED, (a, c, d) / b, so that a step reads nine and writes three complex numbers per row
  // (twelve and four with b carried along) and inverts one pivot instead of two: the load/store pipe bounds this
  // kernel.  [ping-pong][a, c, d][n], addressed by arithmetic on one base pointer (an array of pointers indexed by the
  // ping-pong bit would live in local memory: six more loads per row and step through the same pipe).
  cd* const base = reinterpret_cast<cd*>(pcr_raw);
  auto arr = [&](int h, int q) { return base + static_cast<size_t>(h * 3 + q) * n; };
  const int nys = p.ye - p.ys + 1;
  const int t = blockIdx.x; // system index
  const int iyl = t / p.nmode, kz = t % p.nmode;
  const int jy = p.ys + iyl;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength;
  const double* tab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  cd* __restrict__ sp = p.spec + t;
  bool bad = false;
  for (int ix = threadIdx.x; ix < n; ix += blockDim.x) {
    cd a, b, c, r = ldcd(sp + ix * xs);
    lapRow(p, tab, ix, jy, kz, kwave, a, b, c, r);
    if (!periodic && ix == 0) {
      a = {0.0, 0.0};
    }
    if (!periodic && ix == n - 1) {
      c = {0.0, 0.0};
    }
    bad = bad || tiny(b);
    const cd ib = crecip(b);
    arr(0, 0)[ix] = cmul(a, ib);
    arr(0, 1)[ix] = cmul(c, ib);
    arr(0, 2)[ix] = cmul(r, ib);
  }
  __syncthreads();
  int cur = 0;
  for (int sft = 1; sft < n; sft <<= 1) {
    const cd* const ina = arr(cur, 0);
    const cd* const inc = arr(cur, 1);
    const cd* const ind = arr(cur, 2);
    cd* const outa = arr(cur ^ 1, 0);
    cd* const outc = arr(cur ^ 1, 1);
    cd* const outd = arr(cur ^ 1, 2);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      // x_i + a_i x_{i-s} + c_i x_{i+s} = d_i; rows i -+ s eliminate those two couplings.  A coupling that would
      // leave the system is exactly zero (a_0 = c_{n-1} = 0 and a step multiplies a_i by a_{i-s}), so the
      // neighbour index may simply be clamped: whatever it reads is multiplied by that zero.
      int im = i - sft, ip = i + sft;
      if (periodic) { // n is a power of two: rows wrap around
        im &= n - 1;
        ip &= n - 1;
      } else {
        im = max(im, 0);
        ip = min(ip, n - 1);
      }
      const cd a = ina[i], c = inc[i], d = ind[i];
      const cd am = ina[im], cm = inc[im], dm = ind[im];
      const cd ap = ina[ip], cp = inc[ip], dp = ind[ip];
      const cd bb = csub(csub(cd{1.0, 0.0}, cmul(a, cm)), cmul(c, ap));
      bad = bad || tiny(bb);
      const cd ib = tiny(bb) ? crecip(bb) : crecipFast(bb);
      const cd an = cmul(cmul(a, am), ib);
      const cd cn = cmul(cmul(c, cp), ib);
      const cd dn = cmul(csub(csub(d, cmul(a, dm)), cmul(c, dp)), ib);
      outa[i] = {-an.re, -an.im};
      outc[i] = {-cn.re, -cn.im};
      outd[i] = dn;
    }
    __syncthreads();
    cur ^= 1;
  }
  // every row is now decoupled: its remaining a and c multiply x_i itself (periodic, s = n) or are zero
  __shared__ double psum[LAP_PCR_NT];
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const cd b = cadd(cadd(arr(cur, 0)[i], cd{1.0, 0.0}), arr(cur, 1)[i]);
    bad = bad || tiny(b);
    const cd x = cmul(arr(cur, 2)[i], crecip(b));
    arr(cur, 2)[i] = x;
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
    cd x = arr(cur, 2)[i];
    x.re -= avg;
    sp[i * xs] = x;
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

// The same elimination with every thread's own rows (i = tid + k * 256, k < R) carried in REGISTERS from step to step:
// a step then reads only the six neighbour values of a row from shared memory and stores its three new ones (six loads
// + three stores where k_lap_pcr has nine + three through the pipe that bounds it).  Systems of up to 256 R rows.
constexpr int LAP_PCRR_NT = 256;
template <int R>
__global__ void __launch_bounds__(LAP_PCRR_NT, 2) k_lap_pcr_reg(const LapParams p) {
  extern __shared__ __align__(16) unsigned char pcr_raw[];
  const int n = p.bd.xe - p.bd.xs + 1;
  const bool periodic = p.bd.periodic_x != 0;
  cd* const base = reinterpret_cast<cd*>(pcr_raw);
  auto arr = [&](int h, int q) { return base + static_cast<size_t>(h * 3 + q) * n; };
  const int nys = p.ye - p.ys + 1;
  const int t = blockIdx.x;
  const int iyl = t / p.nmode, kz = t % p.nmode;
  const int jy = p.ys + iyl;
  const long long xs = static_cast<long long>(nys) * p.nmode;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength;
  const double* tab = p.table + (static_cast<long long>(jy) * p.nx + p.bd.xs) * LAP_NK;
  cd* __restrict__ sp = p.spec + t;
  const int tid = threadIdx.x;
  bool bad = false;
  cd ra[R], rc[R], rd[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int ix = tid + k * LAP_PCRR_NT;
    ra[k] = rc[k] = rd[k] = cd{0.0, 0.0};
    if (ix < n) {
      cd a, b, c, r = ldcd(sp + ix * xs);
      lapRow(p, tab, ix, jy, kz, kwave, a, b, c, r);
      if (!periodic && ix == 0) {
        a = {0.0, 0.0};
      }
      if (!periodic && ix == n - 1) {
        c = {0.0, 0.0};
      }
      bad = bad || tiny(b);
      const cd ib = crecip(b);
      ra[k] = cmul(a, ib);
      rc[k] = cmul(c, ib);
      rd[k] = cmul(r, ib);
      arr(0, 0)[ix] = ra[k];
      arr(0, 1)[ix] = rc[k];
      arr(0, 2)[ix] = rd[k];
    }
  }
  __syncthreads();
  int cur = 0;
  for (int sft = 1; sft < n; sft <<= 1) {
    const cd* const ina = arr(cur, 0);
    const cd* const inc = arr(cur, 1);
    const cd* const ind = arr(cur, 2);
    cd* const outa = arr(cur ^ 1, 0);
    cd* const outc = arr(cur ^ 1, 1);
    cd* const outd = arr(cur ^ 1, 2);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int i = tid + k * LAP_PCRR_NT;
      if (i < n) {
        int im = i - sft, ip = i + sft;
        if (periodic) {
          im &= n - 1;
          ip &= n - 1;
        } else { // a coupling that would leave the system is exactly zero: the clamped neighbour is multiplied by it
          im = max(im, 0);
          ip = min(ip, n - 1);
        }
        const cd a = ra[k], c = rc[k], d = rd[k];
        const cd am = ina[im], cm = inc[im], dm = ind[im];
        const cd ap = ina[ip], cp = inc[ip], dp = ind[ip];
        const cd bb = csub(csub(cd{1.0, 0.0}, cmul(a, cm)), cmul(c, ap));
        bad = bad || tiny(bb);
        const cd ib = tiny(bb) ? crecip(bb) : crecipFast(bb);
        const cd an = cmul(cmul(a, am), ib);
        const cd cn = cmul(cmul(c, cp), ib);
        ra[k] = {-an.re, -an.im};
        rc[k] = {-cn.re, -cn.im};
        rd[k] = cmul(csub(csub(d, cmul(a, dm)), cmul(c, dp)), ib);
        outa[i] = ra[k];
        outc[i] = rc[k];
        outd[i] = rd[k];
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  // every row is decoupled: its remaining a and c multiply x_i itself (periodic, s = n) or are zero
  __shared__ double psum[LAP_PCRR_NT];
  double part = 0.0;
  cd x[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = tid + k * LAP_PCRR_NT;
    x[k] = {0.0, 0.0};
    if (i < n) {
      const cd b = cadd(cadd(ra[k], cd{1.0, 0.0}), rc[k]);
      bad = bad || tiny(b);
      x[k] = cmul(rd[k], crecip(b));
      part += x[k].re;
    }
  }
  double avg = 0.0;
  if (periodic && kz == 0) { // subtract the x-average of the kz = 0 mode (cyclic_laplace.cxx:266-287); uniform branch
    psum[tid] = part;
    __syncthreads();
    double sum = 0.0;
    for (int q = 0; q < LAP_PCRR_NT; ++q) {
      sum += psum[q];
    }
    avg = sum / static_cast<double>(n);
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int i = tid + k * LAP_PCRR_NT;
    if (i < n) {
      x[k].re -= avg;
      sp[i * xs] = x[k];
    }
  }
  if (bad) {
    atomicExch(p.status, 1);
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


// ---- block path (laplace_block.cuh) ---------------------------------------------------------------------------
// One cooperative kernel per solve decides whether the matrix-only tables are still valid and rebuilds them when
// not: compare the Field2D inputs with the copy of the last factorisation -> grid barrier -> (only when something
// changed) K-table -> barrier -> block factors -> barrier -> separator-system factors.
__global__ void __launch_bounds__(256) k_lapb_prepare(const LapParams p) {
  // the DOTS kernel may start now: it reads only the right-hand side until it waits for this grid to finish
  asm volatile("griddepcontrol.launch_dependents;");
  cg::grid_group grid = cg::this_grid();
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  const int n2 = p.nx * p.ny;
  for (long long i = tid; i < n2; i += nth) {
    lapbCompareThread(p, static_cast<int>(i));
  }
  grid.sync();
  if (!lapbDirty(p)) {
    return;
  }
  for (long long i = tid; i < n2; i += nth) {
    lapTableEntry(p, static_cast<int>(i));
  }
  grid.sync();
  bool bad = false;
  const long long t1 = static_cast<long long>(p.ye - p.ys + 1) * p.bP * p.nmode;
  for (long long idx = tid; idx < t1; idx += nth) {
    bad = lapbFactor1Thread(p, idx) || bad;
  }
  grid.sync();
  if (p.bP > 1) {
    const long long t2 = static_cast<long long>(p.ye - p.ys + 1) * p.nmode;
    for (long long idx = tid; idx < t2; idx += nth) {
      bad = lapbFactor2Thread(p, static_cast<int>(idx)) || bad;
    }
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

constexpr int LAPB_RED_NT = 64;   // separator systems per CTA
constexpr int LAPB_RED_CTA = 256; // threads per CTA (the Nyquist CTAs use them all)
// Separator systems: one thread per (y, kz < N), LAPB_RED_NT per CTA (the other threads leave at once).  The CTAs
// after those solve the Nyquist systems, one per y, with the block algorithm (laplace_block.cuh): staging by all
// threads, dot products by one thread per block, the separator sweep by one thread, the block solves by one thread
// per half block.
__global__ void __launch_bounds__(LAPB_RED_CTA) k_lapb_reduced(const LapParams p, const int nred_ctas) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the SOLVE kernel may start now: its row loads and forward transforms do not depend on this grid
  asm volatile("griddepcontrol.launch_dependents;");
  const int nys = p.ye - p.ys + 1, nk = min(p.nmode, p.nz / 2);
  const int tid = threadIdx.x;
  if (static_cast<int>(blockIdx.x) < nred_ctas) {
    cd* D = reinterpret_cast<cd*>(smem_raw); // [P][LAPB_RED_NT]: a thread's column is conflict-free
    const int idx = blockIdx.x * LAPB_RED_NT + tid;
    if (tid < LAPB_RED_NT && idx < nys * nk) {
      lapbReducedThread(p, (idx / nk) * p.nmode + idx % nk, D + tid, LAPB_RED_NT);
    }
    return;
  }
  if (p.dbg & 256) {
    return;
  }
  const int iyl = blockIdx.x - nred_ctas, n = p.bd.xe - p.bd.xs + 1;
  const int ysel = lapbYdep(p) ? iyl : 0;
  const LapbNyq w = lapbNyqCarve(smem_raw, n, p.bP);
  lapbNyqStage(p, iyl, ysel, tid, LAPB_RED_CTA, w);
  __syncthreads();
  for (int jb = tid; jb < p.bP; jb += LAPB_RED_CTA) {
    lapbNyqDots(p, ysel, jb, w);
  }
  __syncthreads();
  if (tid == 0) {
    lapbNyqSeparators(p, w);
  }
  __syncthreads();
  for (int q = tid; q < 2 * p.bP; q += LAPB_RED_CTA) {
    lapbNyqSweepHalf(p, ysel, q >> 1, q & 1, w, w.ends);
  }
  __syncthreads();
  for (int q = tid; q < 2 * p.bP; q += LAPB_RED_CTA) {
    lapbNyqSubstHalf(p, ysel, q >> 1, q & 1, w, w.ends);
  }
  __syncthreads();
  double* __restrict__ rg = p.bnq + static_cast<long long>(iyl) * n;
  for (int i = tid; i < n; i += LAPB_RED_CTA) {
    rg[i] = w.r[i];
  }
}

// rows kernels: CTA = (NBLK consecutive blocks of one y row); the phases of laplace_block.cuh / fft_reg.cuh
// with their barriers.  A row is owned by T lanes of one warp, so the transform phases only need __syncwarp.
// While a CTA works, the rows of the CTA that will take its place (block index + bPD) are pulled into L2 by
// the bulk-prefetch unit: DRAM latency then overlaps the transforms without costing registers or shared memory.
template <int N, int NT, int MODE>
__global__ void __launch_bounds__(NT, 512 / NT) k_lapb_rows(const LapParams p) {
  using Sh = LapbShape<N, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* S = reinterpret_cast<cd*>(smem_raw);
  double* rowc = reinterpret_cast<double*>(S + Sh::ROWS * Sh::SROW);
  const int tid = threadIdx.x, t = tid % Sh::T;
  LapbCta c = lapbCta<N, NT>(p, blockIdx.x);
  if (p.bPD > 0 && tid < Sh::ROWS && blockIdx.x + p.bPD < gridDim.x) {
    const LapbCta cn = lapbCta<N, NT>(p, blockIdx.x + p.bPD);
    const int ix = cn.row0 + tid;
    if (ix <= p.bd.xe - p.bd.xs) {
      const double* src = p.rhs + (static_cast<long long>(p.bd.xs + ix) * p.ny + cn.jy) * p.nz;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(p.nz * 8) : "memory");
    }
  }
  cd* row = S + (tid / Sh::T) * Sh::SROW;
  cd v[Sh::B];
  if (!(p.dbg & 8)) {
    lapbLoadRow<N, NT, MODE>(p, c, tid, v);
  } else {
#pragma unroll
    for (int q = 0; q < Sh::B; ++q) {
      v[q] = {1.0 + tid, 2.0 + q};
    }
  }
  if (!(p.dbg & 4)) {
    regfftStage1<N, false>(v, p.tw, t);
    regfftXchgWrite<N>(v, row, t);
    __syncwarp();
    regfftStage2<N, false>(v, row, t);
    __syncwarp();
    regfftSplitWriteZ<N>(v, row, t);
    __syncwarp();
    {
      cd pc[Sh::B / 2 + 1];
      regfftSplitReadPartners<N>(pc, row, t);
      __syncwarp();
      regfftSplitWriteX<N>(v, pc, row, p.tw, t);
    }
  } else {
    regfftSplitWriteZ<N>(v, row, t);
  }
  // Everything above read the right-hand side only.  From here on the results of the previous kernel in the stream
  // are needed (DOTS: the matrix tables of k_lapb_prepare; SOLVE: the separator and Nyquist values of
  // k_lapb_reduced): this grid was launched with programmatic stream serialisation and waits for them here.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  lapbSetYsel(p, c);
  if (MODE == LAPB_SOLVE) {
    lapbStageRows<N, NT>(p, c, tid, rowc);
  }
  const bool nyquist = p.nmode > N;
  if (nyquist && t == 0 && c.row0 + tid / Sh::T <= p.bd.xe - p.bd.xs) { // this thread wrote X[N] of its row itself
    double* q = p.bnq + static_cast<long long>(c.iyl) * (p.bd.xe - p.bd.xs + 1) + c.row0 + tid / Sh::T;
    if (MODE == LAPB_DOTS) {
      *q = row[N].re;
    } else {
      row[N] = {*q, 0.0};
    }
  }
  if (!(p.dbg & 1)) { // per-mode phase; its operands are fetched on this side of the barrier
    int kzs[Sh::KPT];
    bool on[Sh::KPT];
    const int blk = lapbModes<N, NT>(tid, kzs);
#pragma unroll
    for (int ch = 0; ch < Sh::KPT; ++ch) {
      on[ch] = kzs[ch] < p.nmode;
    }
    if (MODE == LAPB_DOTS) {
      LapbDots<Sh::KPT> st;
      lapbDotsFetch<N, NT, Sh::KPT>(p, c, blk, kzs, 0, st);
      lapbDotsFetchScale<N, NT, Sh::KPT>(p, c, blk, kzs, st);
      __syncthreads();
      lapbDotsRun<N, NT, Sh::KPT>(p, c, blk, kzs, on, st, S);
    } else {
      LapbChain<Sh::MB, Sh::KPT> st;
      if (!(p.dbg & 64)) {
        lapbChainLoad<N, NT, Sh::KPT>(p, c, blk, kzs, st);
      } else {
#pragma unroll
        for (int ch = 0; ch < Sh::KPT; ++ch) {
#pragma unroll
          for (int li = 0; li < Sh::MB; ++li) {
            st.w[ch][li] = {0.5 + tid, 0.25};
          }
          st.xs[ch] = st.xn[ch] = st.midinv[ch] = {0.1, 0.2};
        }
      }
      __syncthreads();
      lapbChainRun<N, NT, Sh::KPT>(p, c, blk, kzs, on, st, S, rowc);
    }
  } else {
    __syncthreads();
  }
  if (MODE == LAPB_DOTS) {
    return;
  }
  __syncthreads();
  if (!(p.dbg & 2)) {
    regfftMergeRead<N>(v, row, p.nmode, p.zero_dc != 0, p.tw, t);
    __syncwarp();
    regfftStage1<N, true>(v, p.tw, t);
    regfftXchgWrite<N>(v, row, t);
    __syncwarp();
    regfftStage2<N, true>(v, row, t);
  } else {
#pragma unroll
    for (int q = 0; q < Sh::B; ++q) {
      v[q] = row[t + Sh::T * q];
    }
  }
  lapbStoreRow<N, NT>(p, c, tid, v);
  // rows delivered to a Y neighbour on another GPU (bout_b200_ypeer): fenced system-wide, then the last CTA of that
  // side raises the neighbour's flag
  const bool to_lower = p.peer_flag[0] != nullptr && p.on_up && c.jy - p.mys < p.ny - 1 - p.mye;
  const bool to_upper = p.peer_flag[1] != nullptr && p.on_dn && p.mye - c.jy < p.mys;
  if (to_lower || to_upper) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        if (d == 0 ? to_lower : to_upper) {
          const unsigned int prev = atomicAdd(p.counters + d, 1u);
          if (prev == p.expect[d] - 1) {
            p.counters[d] = 0;
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_flag[d]), "l"(p.signal) : "memory");
          }
        }
      }
    }
  }
}

} // namespace b200

using namespace b200;

extern "C" size_t bout_b200_laplace_cyclic_workspace(const bout_b200_grid* g, const bout_b200_laplace* L) {
  if (g == nullptr || L == nullptr) {
    return 0;
  }
  return lapLayout(g, L).total;
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
  if (const char* env = getenv("BOUT_B200_LAPB_NT")) { // rows-kernel CTA size at nz = 512: 256 (default) or 128
    lapbThreads512() = atoi(env) == 128 ? 128 : 256;
  }
  lapFillParams(g, L, workspace, &p);
  if (const char* env = getenv("BOUT_B200_LAPB_DBG")) {
    p.dbg = atoi(env);
  }
  p.bPD = 2 * 148; // CTAs in flight: 2 per SM
  if (const char* env = getenv("BOUT_B200_LAPB_PD")) {
    p.bPD = atoi(env);
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

// Factorisations held in caller workspaces (block path): workspace -> the grid and settings its tables were
// computed for.  Pointer members are compared by value; the CONTENTS of the arrays they point to are compared
// on the device by k_lapb_prepare in every solve, so a caller that rewrites a coefficient array in place gets
// a new factorisation without having to say so.
struct FactorKey {
  bout_b200_grid g;
  bout_b200_laplace L;
  size_t bytes;
};
std::mutex g_factor_mutex;
std::unordered_map<const void*, FactorKey> g_factors;
std::atomic<unsigned> g_lapb_gen{0};

bool factorMatches(const void* ws, const bout_b200_grid* g, const bout_b200_laplace* L, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_factor_mutex);
  auto it = g_factors.find(ws);
  if (it == g_factors.end()) {
    return false;
  }
  const bout_b200_laplace& K = it->second.L;
  return it->second.bytes == bytes && memcmp(&it->second.g, g, sizeof(*g)) == 0 && K.dx == L->dx && K.g11 == L->g11
         && K.g33 == L->g33 && K.g13 == L->g13 && K.G1 == L->G1 && K.G3 == L->G3 && K.g_11 == L->g_11
         && K.g_22 == L->g_22 && K.A == L->A && K.C1 == L->C1 && K.C2 == L->C2 && K.D == L->D
         && K.zlength == L->zlength && K.all_terms == L->all_terms && K.nonuniform == L->nonuniform
         && K.global_flags == L->global_flags && K.inner_boundary_flags == L->inner_boundary_flags
         && K.outer_boundary_flags == L->outer_boundary_flags && K.maxmode == L->maxmode
         && K.periodic_x == L->periodic_x && K.first_x == L->first_x && K.last_x == L->last_x
         && K.include_yguards_lower == L->include_yguards_lower
         && K.include_yguards_upper == L->include_yguards_upper;
}
void factorForget(const void* ws) {
  std::lock_guard<std::mutex> lock(g_factor_mutex);
  g_factors.erase(ws);
}

// Is this solve served by the block path?  Isolated systems, nz in {128, 256, 512}, 16-byte aligned rows.
// BOUT_B200_LAP_BLOCK=0 keeps the three-kernel paths (tests compare the two).
bool lapUseBlock(const LapPlan* plan, const bout_b200_laplace* L, bool with_fields) {
  const LapParams& p = plan->p;
  if (lapCoupled(L) || !lapbShapeOk(p.nz)) {
    return false;
  }
  // k_lapb_reduced keeps P sweep values per thread, and one whole Nyquist system with its tables, on chip
  if (static_cast<size_t>(p.bP) * 64 * sizeof(cd) > 160 * 1024 || lapbNyqBytes(p.bd.xe - p.bd.xs + 1, p.bP) > 160 * 1024) {
    return false;
  }
  if (with_fields && (!aligned16(p.rhs) || (p.x0 != nullptr && !aligned16(p.x0)) || !aligned16(p.out))) {
    return false;
  }
  const char* env = getenv("BOUT_B200_LAP_BLOCK");
  return env == nullptr || env[0] != '0';
}

template <int N, int NT, int MODE>
int launchRows(const LapParams& p_in, cudaStream_t s) {
  using Sh = LapbShape<N, NT>;
  LapParams p = p_in;
  if (MODE == LAPB_SOLVE && (p.peer_flag[0] != nullptr || p.peer_flag[1] != nullptr)) {
    // CTAs that deliver rows to the lower / upper Y neighbour: every x tile of the bottom / top rows
    const unsigned ntile = static_cast<unsigned>((p.bd.xe - p.bd.xs + 1 + Sh::ROWS - 1) / Sh::ROWS);
    int rows_lo = 0, rows_hi = 0;
    for (int y = p.ys; y <= p.ye; ++y) {
      rows_lo += (y - p.mys < p.ny - 1 - p.mye) ? 1 : 0;
      rows_hi += (p.mye - y < p.mys) ? 1 : 0;
    }
    p.expect[0] = static_cast<unsigned>(rows_lo) * ntile;
    p.expect[1] = static_cast<unsigned>(rows_hi) * ntile;
  }
  static bool configured = false;
  if (!configured) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_lapb_rows<N, NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(Sh::smemBytes)));
    configured = true;
  }
  const int n = p.bd.xe - p.bd.xs + 1, nys = p.ye - p.ys + 1;
  const unsigned blocks = static_cast<unsigned>(nys) * ((n + Sh::ROWS - 1) / Sh::ROWS);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = Sh::smemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; // may start while the previous kernel still runs
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, k_lapb_rows<N, NT, MODE>, p));
  countLaunch();
  return 0;
}
template <int MODE>
int launchRowsN(const LapParams& p, cudaStream_t s) {
  switch (p.nz) {
  case 128: return launchRows<64, 256, MODE>(p, s);
  case 256: return launchRows<128, 256, MODE>(p, s);
  default: return p.bNT == 128 ? launchRows<256, 128, MODE>(p, s) : launchRows<256, 256, MODE>(p, s);
  }
}

// prepare + (conditional) factorisation; leaves p.gen / p.force set for the solve kernels
int lapBlockFactor(const bout_b200_grid* g, const bout_b200_laplace* L, void* workspace, size_t workspace_bytes,
                   LapPlan* plan, bool force, cudaStream_t s) {
  LapParams& p = plan->p;
  unsigned gen = 0;
  while (gen == 0) {
    gen = ++g_lapb_gen;
  }
  p.gen = gen;
  p.force = (force || !factorMatches(workspace, g, L, workspace_bytes)) ? 1 : 0;
  {
    FactorKey k{};
    k.g = *g;
    k.L = *L;
    k.bytes = workspace_bytes;
    std::lock_guard<std::mutex> lock(g_factor_mutex);
    g_factors[workspace] = k;
  }
  static int coop_blocks = 0; // co-resident CTAs of the cooperative kernel on this device
  if (coop_blocks == 0) {
    int dev = 0, sms = 0, per = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    B200_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B200_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lapb_prepare, 256, 0));
    B200_REQUIRE(per >= 1, BOUT_B200_ERR_CUDA, "laplace: the prepare kernel does not fit on an SM");
    coop_blocks = sms * (per < 2 ? per : 2);
  }
  void* args[] = {&p};
  B200_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_lapb_prepare), dim3(coop_blocks), dim3(256), args, 0, s));
  countLaunch();
  return 0;
}

int lapBlockSolve(const LapPlan* plan, cudaStream_t s) {
  const LapParams& p = plan->p;
  int rc = 0;
  if (p.bP > 1 || p.nmode > p.nz / 2) { // separators and / or the Nyquist mode to be solved between the two passes
    rc = launchRowsN<LAPB_DOTS>(p, s);
    if (rc != 0) {
      return rc;
    }
    const int nys = p.ye - p.ys + 1, nk = p.nmode < p.nz / 2 ? p.nmode : p.nz / 2, n = p.bd.xe - p.bd.xs + 1;
    const int nred = (nys * nk + LAPB_RED_NT - 1) / LAPB_RED_NT;
    const bool nyquist = p.nmode > p.nz / 2;
    size_t dsm = static_cast<size_t>(p.bP) * LAPB_RED_NT * sizeof(cd);
    const size_t nsm = lapbNyqBytes(n, p.bP);
    if (nyquist && nsm > dsm) {
      dsm = nsm;
    }
    B200_REQUIRE(dsm <= 200 * 1024, BOUT_B200_ERR_UNSUPPORTED, "laplace: nx = %d too large for the block path", p.nx);
    static size_t dsm_attr = 48 * 1024;
    if (dsm > dsm_attr) {
      B200_CUDA_OK(cudaFuncSetAttribute(k_lapb_reduced, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dsm)));
      dsm_attr = dsm;
    }
    k_lapb_reduced<<<nred + (nyquist ? nys : 0), LAPB_RED_CTA, dsm, s>>>(p, nred);
    B200_LAUNCH_OK();
  }
  return launchRowsN<LAPB_SOLVE>(p, s);
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
  if (lapUseBlock(&plan, L, true)) {
    B200_REQUIRE(rhs != nullptr, BOUT_B200_ERR_ARG, "laplace: null rhs");
    B200_REQUIRE(!((L->inner_boundary_flags | L->outer_boundary_flags) & INV_SET) || x0 != nullptr,
                 BOUT_B200_ERR_ARG, "laplace: INVERT_SET needs x0");
    rc = lapBlockFactor(g, L, workspace, workspace_bytes, &plan, false, s);
    if (rc != 0) {
      return rc;
    }
    return lapBlockSolve(&plan, s);
  }
  factorForget(workspace); // the three-kernel paths overwrite the regions the block path keeps its tables in
  rc = lapForward(L, rhs, x0, &plan, s);
  if (rc != 0) {
    return rc;
  }
  const unsigned sblocks = static_cast<unsigned>((plan.nsys + 127) / 128);
  const int nrows_sys = p.bd.xe - p.bd.xs + 1;
  const size_t pcr_smem = static_cast<size_t>(6) * nrows_sys * sizeof(cd); // [2][a, c, d][n]
  static const bool no_pcr = getenv("BOUT_B200_LAP_NO_PCR") != nullptr;
  const bool pcr_shape = !lapCoupled(L) || (L->periodic_x && (nrows_sys & (nrows_sys - 1)) == 0);
  if (pcr_shape && !no_pcr && plan.nsys <= 2048 && nrows_sys >= 8 && pcr_smem <= 200 * 1024) {
    static size_t pcr_attr = 0;
    if (pcr_smem > 48 * 1024 && pcr_smem > pcr_attr) {
      B200_CUDA_OK(cudaFuncSetAttribute(k_lap_pcr, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(pcr_smem)));
      pcr_attr = pcr_smem;
    }
    static const bool pcr_noreg = getenv("BOUT_B200_PCR_REG") != nullptr && getenv("BOUT_B200_PCR_REG")[0] == '0';
    const int rpt = (nrows_sys + LAP_PCRR_NT - 1) / LAP_PCRR_NT; // rows per thread of the register variant
    if (!pcr_noreg && rpt <= 6 && 2 * (pcr_smem + 3072) <= 220 * 1024) {
      auto launch = [&](auto kern) {
        static size_t attr = 0; // (one static per instantiation)
        if (pcr_smem > 48 * 1024 && pcr_smem > attr) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pcr_smem));
          attr = pcr_smem;
        }
        kern<<<static_cast<unsigned>(plan.nsys), LAP_PCRR_NT, pcr_smem, s>>>(p);
      };
      switch (rpt) {
      case 1: launch(k_lap_pcr_reg<1>); break;
      case 2: launch(k_lap_pcr_reg<2>); break;
      case 3: launch(k_lap_pcr_reg<3>); break;
      case 4: launch(k_lap_pcr_reg<4>); break;
      case 5: launch(k_lap_pcr_reg<5>); break;
      default: launch(k_lap_pcr_reg<6>); break;
      }
      B200_LAUNCH_OK();
    } else {
    // two 256-thread CTAs per SM while two systems fit into shared memory (measured at 1028 rows: 93 us per solve
    // against 107 us with one 512-thread CTA), one wide CTA beyond that; BOUT_B200_PCR_NT overrides
    static const int pcr_env = getenv("BOUT_B200_PCR_NT") ? atoi(getenv("BOUT_B200_PCR_NT")) : 0;
    int pcr_nt = 2 * (pcr_smem + 1024) <= 220 * 1024 ? 256 : LAP_PCR_NT;
    if (pcr_env == 256 || pcr_env == LAP_PCR_NT) {
      pcr_nt = pcr_env;
    }
    k_lap_pcr<<<static_cast<unsigned>(plan.nsys), pcr_nt, pcr_smem, s>>>(p);
    B200_LAUNCH_OK();
    }
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
  B200_REQUIRE(L->ypeer == nullptr, BOUT_B200_ERR_UNSUPPORTED,
               "laplace: rows are delivered to Y neighbours by the block path only (isolated systems, nz = 128 / 256 / 512)");
  rc = launchFft<false>(&plan, s);
  if (rc == 0 && L->wrap_yguards && p.on_up) { // the block path stores the wrapped rows itself; here they are copied afterwards
    double* f[1] = {out};
    rc = bout_b200_halo_self(g, 1, f, 0, 1, stream);
  }
  return rc;
}

// Builds the matrix-only tables of the block path now (they are otherwise built by the first solve and
// rebuilt by any solve that finds an input array changed).  A no-op for configurations on the other paths.
extern "C" int bout_b200_laplace_cyclic_factor(const bout_b200_grid* g, const bout_b200_laplace* L,
                                               void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = asStream(stream);
  LapPlan plan;
  int rc = lapSetup(g, L, workspace, workspace_bytes, s, &plan);
  if (rc != 0) {
    return rc;
  }
  if (!lapUseBlock(&plan, L, false)) {
    factorForget(workspace);
    return 0;
  }
  return lapBlockFactor(g, L, workspace, workspace_bytes, &plan, true, s);
}

extern "C" int bout_b200_laplace_cyclic_invalidate(void* workspace) {
  factorForget(workspace);
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

// =====================================================================================================================
// LaplaceXZ (cyclic) and InvertPar (cyclic) on the same machinery -- row rfft (k_lap_forward), one thread
// per (line, kz) tridiagonal system with the coefficients formed on the fly, row irfft (k_lap_backward).
// =====================================================================================================================
namespace b200 {
struct XzParams {
  const double *A, *B, *J, *g11, *g33, *dx;
  double kscale;
  int inner, outer;
  int ny, xstart, xend, ys, nys, nmode;
  cd* spec;          // [row = x - (xstart - 1)][iyl][kz]: right-hand side in, solution out
  cd* cp;            // c' of the sweep, same layout
  const cd* x0spec;  // INVERT_SET on the inner boundary: spectra of x0 at x = xstart - 1, xstart ([2][iyl][kz])
  int* status;
};

// laplacexz-cyclic.cxx:63-143 (coefficients), :146-250 (right-hand side), CyclicReduce -> Thomas here
__global__ void __launch_bounds__(128) k_xz_thomas(const XzParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nys * p.nmode) {
    return;
  }
  const int iyl = t / p.nmode, kz = t % p.nmode, y = p.ys + iyl;
  const double kwave = kz * p.kscale;
  const long long rs = static_cast<long long>(p.nys) * p.nmode; // row stride
  const int n = p.xend - p.xstart + 3;
  cd* __restrict__ S = p.spec + t;
  cd* __restrict__ CP = p.cp + t;
  const bool grad_in = (kz == 0 && (p.inner & INV_DC_GRAD)) || (kz != 0 && (p.inner & INV_AC_GRAD));
  const bool grad_out = (kz == 0 && (p.outer & INV_DC_GRAD)) || (kz != 0 && (p.outer & INV_AC_GRAD));
  auto at = [&](const double* f, int x) { return __ldg(f + static_cast<long long>(x) * p.ny + y); };
  // row 0: x = xstart - 1
  double b0 = grad_in ? 1.0 : 0.5, c0 = grad_in ? -1.0 : 0.5;
  cd r0 = {0.0, 0.0};
  if (p.inner & INV_SET) {
    const cd u = p.x0spec[t], v = p.x0spec[rs + t];
    r0 = {b0 * u.re + c0 * v.re, b0 * u.im + c0 * v.im};
  } else if (p.inner & INV_RHS) {
    const cd u = S[0], v = S[rs];
    r0 = {b0 * u.re + c0 * v.re, b0 * u.im + c0 * v.im};
  }
  bool bad = b0 == 0.0;
  double cprev = c0 / b0;
  cd dprev = {r0.re / b0, r0.im / b0};
  CP[0] = {cprev, 0.0};
  S[0] = dprev;
  for (int i = 1; i < n; ++i) {
    double a, b, c;
    cd r;
    if (i < n - 1) {
      const int x = p.xstart - 1 + i;
      const double Jc = at(p.J, x), dxc = at(p.dx, x), Ac = at(p.A, x), gc = at(p.g11, x);
      const double vp = (0.5 * (Ac + at(p.A, x + 1))) * (0.5 * (Jc + at(p.J, x + 1))) * (0.5 * (gc + at(p.g11, x + 1)))
                        / (Jc * (0.5 * (dxc + at(p.dx, x + 1))) * dxc);
      const double vm = (0.5 * (Ac + at(p.A, x - 1))) * (0.5 * (Jc + at(p.J, x - 1))) * (0.5 * (gc + at(p.g11, x - 1)))
                        / (Jc * (0.5 * (dxc + at(p.dx, x - 1))) * dxc);
      a = vm;
      c = vp;
      b = at(p.B, x) - vp - vm - Ac * kwave * kwave * at(p.g33, x);
      r = S[i * rs];
    } else { // x = xend + 1; its right-hand side is always zeroed (laplacexz-cyclic.cxx:237-239)
      a = grad_out ? -1.0 : 0.5;
      b = grad_out ? 1.0 : 0.5;
      c = 0.0;
      r = {0.0, 0.0};
    }
    const double den = b - a * cprev;
    bad = bad || !(fabs(den) > 1e-300);
    const double inv = 1.0 / den;
    cprev = c * inv;
    dprev = {(r.re - a * dprev.re) * inv, (r.im - a * dprev.im) * inv};
    CP[i * rs] = {cprev, 0.0};
    S[i * rs] = dprev;
  }
  cd x = dprev;
  for (int i = n - 2; i >= 0; --i) {
    const double cpi = CP[i * rs].re;
    const cd d = S[i * rs];
    x = {d.re - cpi * x.re, d.im - cpi * x.im};
    S[i * rs] = x;
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

struct ParParams {
  const double *A, *B, *C, *D, *E, *sg, *g_22, *dy, *d1_dy, *twist;
  double kscale;
  int closed;
  int ny, xs, nxs, ystart, yend, nmode;
  int ys, nys;   // y range transformed (closed: the interior rows; open: every row)
  cd *spec, *cp, *zs;
  int* status;
};

// parderiv/impls/cyclic/cyclic.cxx:60-245.  Closed field lines: periodic system with the twist-shift phases on its
// two corner elements, solved by Thomas + the Sherman-Morrison correction; open: boundary rows in the y guard cells.
__global__ void __launch_bounds__(128) k_par_thomas(const ParParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nxs * p.nmode) {
    return;
  }
  const int ixl = t / p.nmode, kz = t % p.nmode, x = p.xs + ixl;
  const double kwave = kz * p.kscale;
  const long long rs = p.nmode;
  const long long base = static_cast<long long>(ixl) * p.nys * p.nmode + kz;
  cd* __restrict__ S = p.spec + base;
  cd* __restrict__ CP = p.cp + base;
  cd* __restrict__ Z = p.zs + base;
  const int n = p.nys;
  auto rowOf = [&](int j, cd& a, cd& b, cd& c) { // row j <-> y = ys + j
    const int y = p.ys + j;
    if (!p.closed && y < p.ystart) {
      a = {0.0, 0.0}, b = {1.0, 0.0}, c = {-1.0, 0.0};
      return false;
    }
    if (!p.closed && y > p.yend) {
      a = {-1.0, 0.0}, b = {1.0, 0.0}, c = {0.0, 0.0};
      return false;
    }
    const long long i2 = static_cast<long long>(x) * p.ny + y;
    const double Bv = p.B ? __ldg(p.B + i2) : 0.0, dy = __ldg(p.dy + i2);
    const double acoef = p.A ? __ldg(p.A + i2) : 1.0;
    double bcoef = Bv / __ldg(p.g_22 + i2);
    double ccoef = p.C ? __ldg(p.C + i2) : 0.0;
    const double dcoef = p.D ? __ldg(p.D + i2) : 0.0;
    double ecoef = (p.E ? __ldg(p.E + i2) : 0.0) + __ldg(p.sg + i2) * Bv;
    if (p.d1_dy != nullptr) {
      ecoef += bcoef * __ldg(p.d1_dy + i2);
    }
    bcoef /= dy * dy;
    ccoef /= dy;
    ecoef /= dy;
    a = {bcoef - 0.5 * ecoef, -0.5 * kwave * ccoef};
    b = {acoef - 2. * bcoef - kwave * kwave * dcoef, 0.0};
    c = {bcoef + 0.5 * ecoef, 0.5 * kwave * ccoef};
    return true;
  };
  bool bad = false;
  cd a, b, c;
  if (!p.closed) {
    cd cprev = {0.0, 0.0}, dprev = {0.0, 0.0};
    for (int j = 0; j < n; ++j) {
      const bool interior = rowOf(j, a, b, c);
      const cd r = interior ? S[j * rs] : cd{0.0, 0.0};
      const cd den = j == 0 ? b : csub(b, cmul(a, cprev));
      bad = bad || tiny(den);
      const cd inv = crecip(den);
      cprev = cmul(c, inv);
      dprev = cmul(j == 0 ? r : csub(r, cmul(a, dprev)), inv);
      CP[j * rs] = cprev;
      S[j * rs] = dprev;
    }
    cd xv = dprev;
    for (int j = n - 2; j >= 0; --j) {
      xv = csub(S[j * rs], cmul(CP[j * rs], xv));
      S[j * rs] = xv;
    }
  } else {
    const double ts = p.twist ? __ldg(p.twist + x) : 0.0;
    double sn, cs;
    sincos(kwave * ts, &sn, &cs);
    // corner elements: row 0 couples to x_{n-1} through a_0 exp(-i k ts), row n-1 to x_0 through c_{n-1} exp(+i k ts)
    rowOf(0, a, b, c);
    const cd beta = cmul(a, cd{cs, -sn});
    const cd gamma = cneg(b);
    cd an, bn, cn;
    rowOf(n - 1, an, bn, cn);
    const cd alpha = cmul(cn, cd{cs, sn});
    bad = bad || tiny(gamma);
    const cd ginv = crecip(gamma);
    // T' = T - u v^T with u = (gamma, 0, .., 0, alpha), v = (1, 0, .., 0, beta / gamma)
    cd den = csub(b, gamma);
    bad = bad || tiny(den);
    cd inv = crecip(den);
    cd cprev = cmul(c, inv), dprev = cmul(S[0], inv), zprev = cmul(gamma, inv);
    CP[0] = cprev;
    S[0] = dprev;
    Z[0] = zprev;
    for (int j = 1; j < n; ++j) {
      rowOf(j, a, b, c);
      cd u = {0.0, 0.0};
      if (j == n - 1) {
        b = csub(b, cmul(cmul(alpha, beta), ginv));
        u = alpha;
      }
      den = csub(b, cmul(a, cprev));
      bad = bad || tiny(den);
      inv = crecip(den);
      cprev = cmul(c, inv);
      dprev = cmul(csub(S[j * rs], cmul(a, dprev)), inv);
      zprev = cmul(csub(u, cmul(a, zprev)), inv);
      CP[j * rs] = cprev;
      S[j * rs] = dprev;
      Z[j * rs] = zprev;
    }
    cd yv = dprev, zv = zprev;
    const cd ylast = yv, zlast = zv;
    for (int j = n - 2; j >= 0; --j) {
      const cd cpj = CP[j * rs];
      yv = csub(S[j * rs], cmul(cpj, yv));
      zv = csub(Z[j * rs], cmul(cpj, zv));
      S[j * rs] = yv;
      Z[j * rs] = zv;
    }
    const cd bg = cmul(beta, ginv);
    const cd num = cadd(yv, cmul(bg, ylast));
    const cd dn = cadd(cadd(cd{1.0, 0.0}, zv), cmul(bg, zlast));
    bad = bad || tiny(dn);
    const cd fact = cmul(num, crecip(dn));
    for (int j = 0; j < n; ++j) {
      S[j * rs] = csub(S[j * rs], cmul(fact, Z[j * rs]));
    }
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}
} // namespace b200

namespace {
// LapParams for the generic row transforms over [xs, xe] x [ys, ye], every mode kept
int rowTransformPlan(const bout_b200_grid* g, int xs, int xe, int ys, int ye, cd* spec, cudaStream_t s, LapParams* pp,
                     RowGeom* geo, unsigned* blocks) {
  LapParams& p = *pp;
  p = LapParams{};
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.bd.xs = xs;
  p.bd.xe = xe;
  p.ys = ys;
  p.ye = ye;
  p.nmode = g->nz / 2 + 1;
  p.spec = spec;
  int st = 0;
  p.tw = twiddleTable(g->nz, s, &st);
  if (st != 0) {
    return st;
  }
  *geo = rowGeometry(g->nz, 128);
  p.tpr = geo->tpr;
  p.rows = geo->rows;
  p.pow2 = geo->pow2 ? 1 : 0;
  B200_REQUIRE(geo->smemBytes <= 200 * 1024, BOUT_B200_ERR_UNSUPPORTED, "row transform: nz=%d too large", g->nz);
  if (geo->smemBytes > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_forward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(geo->smemBytes)));
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_backward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(geo->smemBytes)));
  }
  const long long nrows = static_cast<long long>(xe - xs + 1) * (ye - ys + 1);
  *blocks = static_cast<unsigned>((nrows + geo->rows - 1) / geo->rows);
  return 0;
}
size_t specBytes(const bout_b200_grid* g, int nrows_x, int nrows_y) {
  return lapRoundUp(static_cast<size_t>(nrows_x) * nrows_y * (g->nz / 2 + 1) * sizeof(cd), 256);
}
} // namespace

extern "C" size_t bout_b200_laplacexz_workspace(const bout_b200_grid* g) {
  if (g == nullptr) {
    return 0;
  }
  const int nrx = g->xend - g->xstart + 3, nys = g->yend - g->ystart + 1;
  return 256 + 2 * specBytes(g, nrx, nys) + specBytes(g, 2, nys);
}

extern "C" int bout_b200_laplacexz_solve(const bout_b200_grid* g, const bout_b200_laplacexz* xz, const double* rhs,
                                         const double* x0, double* out, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  cudaStream_t s = asStream(stream);
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(xz && rhs && out && workspace, BOUT_B200_ERR_ARG, "laplacexz: null argument");
  B200_REQUIRE(xz->A && xz->B && xz->J && xz->g11 && xz->g33 && xz->dx, BOUT_B200_ERR_ARG, "laplacexz: coefficient or metric missing");
  B200_REQUIRE(g->xstart >= 1 && g->xend + 1 < g->nx, BOUT_B200_ERR_GUARDS, "laplacexz: needs one guard cell in x");
  B200_REQUIRE(!(xz->inner_boundary_flags & INV_SET) || x0 != nullptr, BOUT_B200_ERR_ARG, "laplacexz: INVERT_SET needs x0");
  B200_REQUIRE(workspace_bytes >= bout_b200_laplacexz_workspace(g), BOUT_B200_ERR_ARG, "laplacexz: workspace too small");
  B200_REQUIRE(aligned16(workspace), BOUT_B200_ERR_ARG, "laplacexz: workspace must be 16-byte aligned");
  const int nrx = g->xend - g->xstart + 3, nys = g->yend - g->ystart + 1;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  XzParams q{};
  q.status = reinterpret_cast<int*>(ws);
  q.spec = reinterpret_cast<cd*>(ws + 256);
  q.cp = reinterpret_cast<cd*>(ws + 256 + specBytes(g, nrx, nys));
  cd* x0spec = reinterpret_cast<cd*>(ws + 256 + 2 * specBytes(g, nrx, nys));
  q.x0spec = x0spec;
  B200_CUDA_OK(cudaMemsetAsync(q.status, 0, sizeof(int), s));
  LapParams p;
  RowGeom geo;
  unsigned blocks = 0;
  int rc = rowTransformPlan(g, g->xstart - 1, g->xend + 1, g->ystart, g->yend, q.spec, s, &p, &geo, &blocks);
  if (rc != 0) {
    return rc;
  }
  p.rhs = rhs;
  p.out = out;
  k_lap_forward<<<blocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  if (xz->inner_boundary_flags & INV_SET) {
    LapParams p0;
    RowGeom geo0;
    unsigned blocks0 = 0;
    rc = rowTransformPlan(g, g->xstart - 1, g->xstart, g->ystart, g->yend, x0spec, s, &p0, &geo0, &blocks0);
    if (rc != 0) {
      return rc;
    }
    p0.rhs = x0;
    k_lap_forward<<<blocks0, 128, geo0.smemBytes, s>>>(p0);
    B200_LAUNCH_OK();
  }
  q.A = xz->A, q.B = xz->B, q.J = xz->J, q.g11 = xz->g11, q.g33 = xz->g33, q.dx = xz->dx;
  q.kscale = 2.0 * M_PI / xz->zlength;
  q.inner = xz->inner_boundary_flags;
  q.outer = xz->outer_boundary_flags;
  q.ny = g->ny, q.xstart = g->xstart, q.xend = g->xend, q.ys = g->ystart, q.nys = nys, q.nmode = p.nmode;
  const int nsys = nys * p.nmode;
  k_xz_thomas<<<(nsys + 127) / 128, 128, 0, s>>>(q);
  B200_LAUNCH_OK();
  k_lap_backward<<<blocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" size_t bout_b200_invert_par_workspace(const bout_b200_grid* g) {
  if (g == nullptr) {
    return 0;
  }
  return 256 + 3 * specBytes(g, g->xend - g->xstart + 1, g->ny);
}

extern "C" int bout_b200_invert_par_solve(const bout_b200_grid* g, const bout_b200_invert_par* ip, const double* f,
                                          double* out, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = asStream(stream);
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(ip && f && out && workspace, BOUT_B200_ERR_ARG, "invert_par: null argument");
  B200_REQUIRE(ip->g_22 && ip->dy && ip->sg, BOUT_B200_ERR_ARG, "invert_par: metric missing");
  B200_REQUIRE(workspace_bytes >= bout_b200_invert_par_workspace(g), BOUT_B200_ERR_ARG, "invert_par: workspace too small");
  B200_REQUIRE(aligned16(workspace), BOUT_B200_ERR_ARG, "invert_par: workspace must be 16-byte aligned");
  const int nxs = g->xend - g->xstart + 1;
  const int ys = ip->closed ? g->ystart : 0, ye = ip->closed ? g->yend : g->ny - 1;
  const int nys = ye - ys + 1;
  B200_REQUIRE(nys >= 3, BOUT_B200_ERR_ARG, "invert_par: needs at least 3 rows in y");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const size_t sb = specBytes(g, nxs, g->ny);
  ParParams q{};
  q.status = reinterpret_cast<int*>(ws);
  q.spec = reinterpret_cast<cd*>(ws + 256);
  q.cp = reinterpret_cast<cd*>(ws + 256 + sb);
  q.zs = reinterpret_cast<cd*>(ws + 256 + 2 * sb);
  B200_CUDA_OK(cudaMemsetAsync(q.status, 0, sizeof(int), s));
  LapParams p;
  RowGeom geo;
  unsigned blocks = 0;
  int rc = rowTransformPlan(g, g->xstart, g->xend, ys, ye, q.spec, s, &p, &geo, &blocks);
  if (rc != 0) {
    return rc;
  }
  p.rhs = f;
  p.out = out;
  k_lap_forward<<<blocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  q.A = ip->A, q.B = ip->B, q.C = ip->C, q.D = ip->D, q.E = ip->E, q.sg = ip->sg;
  q.g_22 = ip->g_22, q.dy = ip->dy, q.d1_dy = ip->d1_dy, q.twist = ip->twist_shift;
  q.kscale = 2.0 * M_PI / ip->zlength;
  q.closed = ip->closed ? 1 : 0;
  q.ny = g->ny, q.xs = g->xstart, q.nxs = nxs, q.ystart = g->ystart, q.yend = g->yend, q.nmode = p.nmode;
  q.ys = ys, q.nys = nys;
  const int nsys = nxs * p.nmode;
  k_par_thomas<<<(nsys + 127) / 128, 128, 0, s>>>(q);
  B200_LAUNCH_OK();
  k_lap_backward<<<blocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

// =====================================================================================================================
// FieldPerp (one x-z plane at a y index): slice copies and LaplaceCyclic::solve(const FieldPerp&, const FieldPerp&)
// (reference include/bout/fieldperp.hxx:54-330, src/field/fieldperp.cxx sliceXZ, cyclic_laplace.cxx:120-185)
// =====================================================================================================================
extern "C" int bout_b200_copy_xz(const bout_b200_grid* g, int jy, double* field3d, double* perp, int to_perp,
                                 void* stream) {
  cudaStream_t s = asStream(stream);
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(field3d && perp, BOUT_B200_ERR_ARG, "copy_xz: null argument");
  B200_REQUIRE(jy >= 0 && jy < g->ny, BOUT_B200_ERR_ARG, "copy_xz: y index %d outside 0..%d", jy, g->ny - 1);
  const size_t row = static_cast<size_t>(g->nz) * sizeof(double), pitch = row * g->ny;
  double* plane = field3d + static_cast<size_t>(jy) * g->nz;
  if (to_perp) {
    B200_CUDA_OK(cudaMemcpy2DAsync(perp, row, plane, pitch, row, g->nx, cudaMemcpyDeviceToDevice, s));
  } else {
    B200_CUDA_OK(cudaMemcpy2DAsync(plane, pitch, perp, row, row, g->nx, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

namespace {
constexpr int kPerpFields = 12;
bout_b200_grid perpGrid(const bout_b200_grid* g) { return bout_b200_grid{g->nx, 1, g->nz, g->xstart, g->xend, 0, 0}; }
size_t perpHead(const bout_b200_grid* g) { return lapRoundUp(static_cast<size_t>(kPerpFields) * g->nx * sizeof(double), 256); }
} // namespace

extern "C" size_t bout_b200_laplace_cyclic_workspace_perp(const bout_b200_grid* g, const bout_b200_laplace* L) {
  if (g == nullptr || L == nullptr) {
    return 0;
  }
  const bout_b200_grid g1 = perpGrid(g);
  return perpHead(g) + bout_b200_laplace_cyclic_workspace(&g1, L);
}

extern "C" int bout_b200_laplace_cyclic_solve_perp(const bout_b200_grid* g, const bout_b200_laplace* L, int jy,
                                                   const double* rhs, const double* x0, double* out, void* workspace,
                                                   size_t workspace_bytes, void* stream) {
  cudaStream_t s = asStream(stream);
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(L && rhs && out && workspace, BOUT_B200_ERR_ARG, "laplace (FieldPerp): null argument");
  B200_REQUIRE(jy >= 0 && jy < g->ny, BOUT_B200_ERR_ARG, "laplace (FieldPerp): y index %d outside 0..%d", jy, g->ny - 1);
  B200_REQUIRE(workspace_bytes >= bout_b200_laplace_cyclic_workspace_perp(g, L), BOUT_B200_ERR_ARG,
               "laplace (FieldPerp): workspace too small");
  // the y = jy row of every Field2D input, as [nx][1] arrays at the head of the workspace
  bout_b200_laplace L1 = *L;
  const double** src[kPerpFields] = {&L1.dx, &L1.g11, &L1.g33, &L1.g13, &L1.G1, &L1.G3,
                                     &L1.g_11, &L1.g_22, &L1.A, &L1.C1, &L1.C2, &L1.D};
  double* head = static_cast<double*>(workspace);
  for (int q = 0; q < kPerpFields; ++q) {
    if (*src[q] == nullptr) {
      continue;
    }
    double* dst = head + static_cast<size_t>(q) * g->nx;
    B200_CUDA_OK(cudaMemcpy2DAsync(dst, sizeof(double), *src[q] + jy, static_cast<size_t>(g->ny) * sizeof(double),
                                   sizeof(double), g->nx, cudaMemcpyDeviceToDevice, s));
    *src[q] = dst;
  }
  L1.include_yguards_lower = L1.include_yguards_upper = 0;
  const bout_b200_grid g1 = perpGrid(g);
  unsigned char* tail = static_cast<unsigned char*>(workspace) + perpHead(g);
  return bout_b200_laplace_cyclic_solve(&g1, &L1, rhs, x0, out, tail, workspace_bytes - perpHead(g), stream);
}
