// z-FFT row operations: rfft -> per-mode multiplier -> irfft in ONE kernel (one read + one write
// of the field; the spectrum never leaves shared memory).
//
// Replaces: FFTDerivativeType / FFT2ndDerivativeType (src/mesh/index_derivs.cxx:418-541),
// filter / lowPass / shiftZ (src/field/field3d.cxx:620-741), and the phase-shift core of
// ShiftedMetric::shiftZ (src/mesh/parallel/shiftedmetric.cxx:200-224).
#include "fft.cuh"
#include "fft_fast.cuh"
#include "fft_reg.cuh"

#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace b200 {

const cd* twiddleTable(int nz, cudaStream_t stream, int* status) {
  static std::mutex mtx;
  static std::map<std::pair<int, int>, cd*> cache;
  std::lock_guard<std::mutex> lock(mtx);
  int dev = 0;
  *status = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    *status = fail(BOUT_B200_ERR_CUDA, "cudaGetDevice failed (no CUDA device?)");
    return nullptr;
  }
  auto it = cache.find({dev, nz});
  if (it != cache.end()) {
    return it->second;
  }
  std::vector<cd> host(nz);
  for (int k = 0; k < nz; ++k) {
    const long double a = -2.0L * 3.141592653589793238462643383279502884L * k / nz;
    host[k] = {static_cast<double>(cosl(a)), static_cast<double>(sinl(a))};
  }
  cd* devp = nullptr;
  if (cudaMalloc(&devp, sizeof(cd) * nz) != cudaSuccess
      || cudaMemcpyAsync(devp, host.data(), sizeof(cd) * nz, cudaMemcpyHostToDevice, stream) != cudaSuccess
      || cudaStreamSynchronize(stream) != cudaSuccess) {
    *status = fail(BOUT_B200_ERR_CUDA, "twiddle table upload failed: %s",
                   cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  cache[{dev, nz}] = devp;
  return devp;
}

enum ZMode { Z_DDZ1 = 0, Z_DDZ2 = 1, Z_FILTER = 2, Z_LOWPASS = 3, Z_SHIFT = 4, Z_SLICE = 5, Z_ZEROLAP = 6,
             Z_FILTER_LOWPASS = 7, Z_SLICE2 = 8 };

struct ZfftParams {
  const double* __restrict__ f;
  const double* __restrict__ scale;  // Field2D output multiplier or NULL
  const double* __restrict__ angle;  // Field2D shift angle (Z_SHIFT) or NULL -> angle0
  const double* __restrict__ zlength; // Field2D zlength (Z_SHIFT) or NULL -> zlength0
  double* __restrict__ out;
  double* __restrict__ out2; // Z_SLICE2: ydown (out = yup)
  const cd* __restrict__ tw;
  int nx, ny, nz, xs, xe, ys, ye;
  int mode, kmax, n0, keep_zonal;
  int yoff; // Z_SLICE: this row is the y + yoff neighbour of the row whose frame it is shifted into
  // Z_ZEROLAP: guard rows are continued from the last interior row xin (outward direction bdir)
  const double* __restrict__ m_dx;
  const double* __restrict__ m_g11;
  const double* __restrict__ m_g33;
  int xin, bdir;
  // Z_ZEROLAP on both X sides in one launch: zl_win >= 0 = number of inner guard rows; row groups 0 .. zl_win-1 are the
  // inner guard cells (continued from zl_xstart inwards), the others the outer ones (from zl_xend outwards)
  int zl_win, zl_xstart, zl_xend;
  double angle0, zlength0;
  int tpr, rows, pow2;
};

__global__ void __launch_bounds__(128) k_zfft_rows(const ZfftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nbuf = fftBufElems(p.nz);
  const int rloc = threadIdx.x / p.tpr;
  RowFFT r;
  r.nz = p.nz;
  r.N = p.nz / 2;
  r.pow2 = p.pow2 != 0;
  r.tpr = p.tpr;
  r.lane = threadIdx.x % p.tpr;
  r.A = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * 2 * nbuf;
  r.B = r.A + nbuf;
  r.tw = p.tw;

  const int nyb = p.ye - p.ys + 1;
  const long long nrows = static_cast<long long>(p.xe - p.xs + 1) * nyb;
  const long long rowid = static_cast<long long>(blockIdx.x) * p.rows + rloc;
  const bool active = rowid < nrows;
  const long long rid = active ? rowid : nrows - 1; // idle groups redo the last row, never store
  int x = p.xs + static_cast<int>(rid / nyb);
  int xin = p.xin, bdir = p.bdir;
  if (p.mode == Z_ZEROLAP && p.zl_win >= 0) {
    const int r = static_cast<int>(rid / nyb);
    if (r < p.zl_win) {
      x = r;
      xin = p.zl_xstart;
      bdir = -1;
    } else {
      x = p.zl_xend + 1 + (r - p.zl_win);
      xin = p.zl_xend;
      bdir = 1;
    }
  }
  const int y = p.ys + static_cast<int>(rid % nyb);
  const long long i2 = static_cast<long long>(x) * p.ny + y;
  const long long isrc2 = p.mode == Z_ZEROLAP ? static_cast<long long>(xin) * p.ny + y : i2;
  const double* __restrict__ src = p.f + isrc2 * p.nz;

  // load the real row into A
  double* a = reinterpret_cast<double*>(r.A);
  if (p.nz % 2 == 0 && aligned16(src)) {
    for (int j = r.lane; j < p.nz / 2; j += r.tpr) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(src) + j);
      reinterpret_cast<double2*>(a)[j] = t;
    }
  } else {
    for (int j = r.lane; j < p.nz; j += r.tpr) {
      a[j] = __ldg(src + j);
    }
  }
  __syncthreads();
  cd* X = rfftForward(r);
  __shared__ double zsum[128];
  if (p.mode == Z_ZEROLAP) { // z-sum of the second interior row, all threads of the row group (uniform branch)
    const double* __restrict__ in2 = p.f + (static_cast<long long>(xin - bdir) * p.ny + y) * p.nz;
    double part = 0.0;
    for (int j = r.lane; j < p.nz; j += r.tpr) {
      part += __ldg(in2 + j);
    }
    zsum[threadIdx.x] = part;
    __syncthreads();
  }

  const int nk = p.nz / 2;
  const double kfac = 2.0 * M_PI / p.nz;
  for (int k = r.lane; k <= nk; k += r.tpr) {
    cd v = X[k];
    switch (p.mode) {
    case Z_DDZ1: {
      if (k <= p.kmax) {
        const double kw = k * kfac;
        v = {-v.im * kw, v.re * kw}; // * (0, kwave)
      } else {
        v = {0.0, 0.0};
      }
      break;
    }
    case Z_DDZ2: {
      const double kw = k * kfac;
      const double m = -kw * kw;
      v = {v.re * m, v.im * m};
      break;
    }
    case Z_FILTER:
      if (k != p.n0) {
        v = {0.0, 0.0};
      }
      break;
    case Z_LOWPASS:
      if (k > p.kmax || (k == 0 && !p.keep_zonal)) {
        v = {0.0, 0.0};
      }
      break;
    case Z_FILTER_LOWPASS:
      if (k != p.n0 || k > p.kmax || (k == 0 && !p.keep_zonal)) {
        v = {0.0, 0.0};
      }
      break;
    case Z_SHIFT: {
      if (k >= 1) {
        const double zl = p.zlength ? __ldg(p.zlength + i2) : p.zlength0;
        const double ang = p.angle ? __ldg(p.angle + i2) : p.angle0;
        const double kwave = k * 2.0 * M_PI / zl;
        double s, c;
        sincos(kwave * ang, &s, &c);
        v = cmul(v, cd{c, -s});
      }
      break;
    }
    case Z_SLICE: { // ShiftedMetric::cachePhases, shiftedmetric.cxx:109-123: zShift(y) - zShift(y + offset)
      if (k >= 1) {
        const double ang = __ldg(p.angle + i2 - p.yoff) - __ldg(p.angle + i2);
        const double kwave = k * 2.0 * M_PI / p.zlength0;
        double s, c;
        sincos(kwave * ang, &s, &c);
        v = cmul(v, cd{c, -s});
      }
      break;
    }
    case Z_ZEROLAP: { // BoundaryZeroLaplace::apply(Field3D), boundary_standard.cxx:2630-2686
      const int steps = (x - xin) * bdir;
      if (k == 0) { // straight line through the last two interior z-means
        double m2 = 0.0;
        for (int j = 0; j < p.tpr; ++j) {
          m2 += zsum[rloc * p.tpr + j];
        }
        m2 /= p.nz;
        const double grad = v.re - m2;
        for (int s = 0; s < steps; ++s) {
          v.re += grad;
        }
      } else { // decaying solution of g11 d2f/dx2 - g33 kz^2 f = 0, applied guard cell by guard cell
        const double kwave = k * 2.0 * M_PI / p.zlength0;
        for (int s = 1; s <= steps; ++s) {
          const long long ig = static_cast<long long>(xin + s * bdir) * p.ny + y;
          const double coef = -1.0 * sqrt(__ldg(p.m_g33 + ig) / __ldg(p.m_g11 + ig)) * __ldg(p.m_dx + ig);
          const double e = exp(coef * kwave);
          v.re *= e;
          v.im *= e;
        }
      }
      break;
    }
    default:
      break;
    }
    X[k] = v;
  }
  __syncthreads();
  const double* res = rfftBackward(r, X);
  if (active) {
    const double sc = p.scale ? __ldg(p.scale + i2) : 1.0;
    double* __restrict__ dst = p.out + i2 * p.nz;
    if (p.nz % 2 == 0 && aligned16(dst)) {
      for (int j = r.lane; j < p.nz / 2; j += r.tpr) {
        double2 t = reinterpret_cast<const double2*>(res)[j];
        t.x *= sc;
        t.y *= sc;
        reinterpret_cast<double2*>(dst)[j] = t;
      }
    } else {
      for (int j = r.lane; j < p.nz; j += r.tpr) {
        dst[j] = res[j] * sc;
      }
    }
  }
}


// ---- fast path (nz in {128, 256, 512, 1024}, 16-byte aligned rows): fft_fast.cuh transforms, persistent
// CTAs of 256 threads, the spectrum of a row held in shared memory between the forward and the one or two
// backward transforms.  Z_SLICE2 produces both parallel slices (y+1 -> yup, y-1 -> ydown) of a source row
// from ONE read and ONE forward transform of that row.
template <int N>
__device__ __forceinline__ void zmodSpectrum(const ZfftParams& p, const cd* S, cd* D, int lane, int mode, double ang,
                                             double zl) {
  constexpr int nz = 2 * N;
  const double kfac = 2.0 * M_PI / nz;
  for (int k = lane; k <= N; k += N / 4) {
    cd v = S[k];
    switch (mode) {
    case Z_DDZ1:
      if (k <= p.kmax) {
        const double kw = k * kfac;
        v = {-v.im * kw, v.re * kw};
      } else {
        v = {0.0, 0.0};
      }
      break;
    case Z_DDZ2: {
      const double kw = k * kfac;
      const double m = -kw * kw;
      v = {v.re * m, v.im * m};
      break;
    }
    case Z_FILTER:
      if (k != p.n0) {
        v = {0.0, 0.0};
      }
      break;
    case Z_LOWPASS:
      if (k > p.kmax || (k == 0 && !p.keep_zonal)) {
        v = {0.0, 0.0};
      }
      break;
    case Z_FILTER_LOWPASS: // filter(f, n0) then lowPass(., kmax, keep_zonal)
      if (k != p.n0 || k > p.kmax || (k == 0 && !p.keep_zonal)) {
        v = {0.0, 0.0};
      }
      break;
    default: // Z_SHIFT, Z_SLICE: multiply by exp(-i kwave ang)
      if (k >= 1) {
        const double kwave = k * 2.0 * M_PI / zl;
        double sn, cs;
        sincos(kwave * ang, &sn, &cs);
        v = cmul(v, cd{cs, -sn});
      }
      break;
    }
    D[k] = v;
  }
}

template <int N>
__global__ void __launch_bounds__(256) k_zrows_fast(const ZfftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TPR = FftShape<N>::TPR, ROWS = 256 / TPR, STRIDE = 4 * (N + 8);
  const int rloc = threadIdx.x / TPR, lane = threadIdx.x % TPR;
  cd* A = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * STRIDE;
  cd* B = A + (N + 8);
  cd* S = B + (N + 8);   // spectrum of the row (N + 1 modes)
  cd* D = S + (N + 8);   // modified spectrum
  FastTw<N> tw;
  fastTwInit<N>(tw, p.tw, lane);
  const SyncGroup sync{1 + rloc, TPR}; // rows of the CTA run decoupled
  const int nyb = p.ye - p.ys + 1;
  const long long nrows = static_cast<long long>(p.xe - p.xs + 1) * nyb;
  for (long long base = static_cast<long long>(blockIdx.x) * ROWS; base < nrows;
       base += static_cast<long long>(gridDim.x) * ROWS) {
    const long long rowid = base + rloc;
    const bool active = rowid < nrows;
    const long long rid = active ? rowid : nrows - 1;
    const int x = p.xs + static_cast<int>(rid / nyb);
    const int y = p.ys + static_cast<int>(rid % nyb);
    const long long i2 = static_cast<long long>(x) * p.ny + y;
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(p.f + i2 * p.nz);
    cd v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double2 q = __ldg(s2 + lane + m * (N / 4));
      v[m] = {q.x, q.y};
    }
    const cd* Z = cfftFastFromRegs<N, false, SyncGroup>(v, A, B, tw, lane, sync);
    rfftSplitTo<N>(Z, S, N + 1, tw, lane);
    sync();
    const int nout = p.mode == Z_SLICE2 ? 2 : 1;
    for (int o = 0; o < nout; ++o) {
      double ang = 0.0, zl = p.zlength0;
      int mode = p.mode;
      bool store = active;
      double* __restrict__ dstf = p.out;
      if (p.mode == Z_SHIFT) {
        zl = p.zlength ? __ldg(p.zlength + i2) : p.zlength0;
        ang = p.angle ? __ldg(p.angle + i2) : p.angle0;
      } else if (p.mode == Z_SLICE) {
        ang = __ldg(p.angle + i2 - p.yoff) - __ldg(p.angle + i2);
      } else if (p.mode == Z_SLICE2) {
        // o = 0: this row is the y+1 neighbour of row y-1 (yup); o = 1: the y-1 neighbour of row y+1 (ydown)
        const int yoff = o == 0 ? 1 : -1;
        const int yt = y - yoff;                       // the row whose frame we shift into
        const bool ok = yt >= p.n0 && yt <= p.kmax;    // n0..kmax = ystart..yend of the target rows
        store = active && ok;
        const int ytc = ok ? yt : y;
        ang = __ldg(p.angle + static_cast<long long>(x) * p.ny + ytc) - __ldg(p.angle + i2);
        mode = Z_SLICE;
        dstf = o == 0 ? p.out : p.out2;
      }
      zmodSpectrum<N>(p, S, D, lane, mode, ang, zl);
      sync();
      cd w[4], r[4];
      rfftMergeRegs<N>([&](int k) -> cd { return D[k]; }, w, tw, lane);
      cfftFastRegsToRegs<N, true, SyncGroup>(w, r, A, B, tw, lane, sync);
      if (store) {
        const double sc = p.scale ? __ldg(p.scale + i2) : 1.0;
        double2* __restrict__ d2 = reinterpret_cast<double2*>(dstf + i2 * p.nz);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          d2[lane + m * (N / 4)] = make_double2(r[m].re * sc, r[m].im * sc);
        }
      }
    }
  }
}

// ---- register path (nz in {128, 256, 512}): fft_reg.cuh transforms ---------------------------------------------
// A row is owned by T lanes of ONE warp holding 16 (8) points each in registers; a transform passes through shared
// memory once; there is no CTA-wide barrier anywhere, so the warps of a CTA run decoupled.  The spectrum of the row
// sits in the row's buffer X between the forward and the backward transform; TWO = Z_SLICE2 keeps it there for the
// second output and modifies into a second buffer.  The phase factors exp(-i k theta) of the shift modes are built
// by repeated multiplication from two sincos per thread (round 1 called sincos for every mode of every row).
template <int N, bool TWO>
__global__ void __launch_bounds__(128, 4) k_zrows_reg(const ZfftParams p) {
  using F = RegFft<N>;
  constexpr int T = F::T, B = F::B, SROW = F::SROW, ROWS = 128 / T, NBUF = TWO ? 2 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, grp = tid / T, t = tid % T;
  cd* X = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(grp) * NBUF * SROW;
  cd* D = TWO ? X + SROW : X; // modified spectrum and exchange area of the backward transform
  const int nyb = p.ye - p.ys + 1;
  const long long nrows = static_cast<long long>(p.xe - p.xs + 1) * nyb;
  constexpr double kfac = 2.0 * M_PI / (2 * N);
  for (long long base = static_cast<long long>(blockIdx.x) * ROWS; base < nrows;
       base += static_cast<long long>(gridDim.x) * ROWS) {
    const long long rowid = base + grp;
    const bool active = rowid < nrows;
    const long long rid = active ? rowid : nrows - 1; // idle groups redo the last row, never store
    const int x = p.xs + static_cast<int>(rid / nyb);
    const int y = p.ys + static_cast<int>(rid % nyb);
    const long long i2 = static_cast<long long>(x) * p.ny + y;
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(p.f + i2 * p.nz);
    cd v[B];
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) {
      const double2 q = __ldg(s2 + t + T * n2);
      v[n2] = {q.x, q.y};
    }
    regfftStage1<N, false>(v, p.tw, t);
    regfftXchgWrite<N>(v, X, t);
    __syncwarp();
    regfftStage2<N, false>(v, X, t);
    __syncwarp();
    regfftSplitWriteZ<N>(v, X, t);
    __syncwarp();
    bool fused_mod = false; // the per-mode multiplier applied while the spectrum is written (single-output modes)
    {
      cd pc[B / 2 + 1];
      regfftSplitReadPartners<N>(pc, X, t);
      __syncwarp();
      if (!TWO && p.mode != Z_SHIFT && p.mode != Z_SLICE) {
        fused_mod = true;
        const int mode = p.mode, kmax = p.kmax, n0 = p.n0, keep_zonal = p.keep_zonal;
        regfftSplitWriteXMod<N>(v, pc, X, p.tw, t, [&](int k, int, bool, cd w) -> cd {
          const double kw = k * kfac;
          switch (mode) {
          case Z_DDZ1: return k <= kmax ? cd{-w.im * kw, w.re * kw} : cd{0.0, 0.0};
          case Z_DDZ2: return cscale(w, -kw * kw);
          case Z_FILTER: return k != n0 ? cd{0.0, 0.0} : w;
          case Z_LOWPASS: return (k > kmax || (k == 0 && !keep_zonal)) ? cd{0.0, 0.0} : w;
          default: return (k != n0 || k > kmax || (k == 0 && !keep_zonal)) ? cd{0.0, 0.0} : w; // Z_FILTER_LOWPASS
          }
        });
      } else {
        regfftSplitWriteX<N>(v, pc, X, p.tw, t);
      }
    }
    __syncwarp();
#pragma unroll
    for (int o = 0; o < NBUF; ++o) {
      double ang = 0.0, zl = p.zlength0;
      int mode = p.mode;
      bool store = active;
      double* __restrict__ dstf = p.out;
      if (p.mode == Z_SHIFT) {
        zl = p.zlength ? __ldg(p.zlength + i2) : p.zlength0;
        ang = p.angle ? __ldg(p.angle + i2) : p.angle0;
      } else if (p.mode == Z_SLICE) {
        ang = __ldg(p.angle + i2 - p.yoff) - __ldg(p.angle + i2);
      } else if (TWO) {
        // o = 0: this row is the y+1 neighbour of row y-1 (yup); o = 1: the y-1 neighbour of row y+1 (ydown)
        const int yoff = o == 0 ? 1 : -1;
        const int yt = y - yoff;                       // the row whose frame we shift into
        const bool ok = yt >= p.n0 && yt <= p.kmax;    // n0..kmax = ystart..yend of the target rows
        store = active && ok;
        const int ytc = ok ? yt : y;
        ang = __ldg(p.angle + static_cast<long long>(x) * p.ny + ytc) - __ldg(p.angle + i2);
        mode = Z_SLICE;
        dstf = o == 0 ? p.out : p.out2;
      }
      // this thread's bins k = t + T q (and the Nyquist bin for t = 0): X -> D
      if (fused_mod) {
        // already applied by regfftSplitWriteXMod (D is X)
      } else if (mode == Z_SHIFT || mode == Z_SLICE) { // * exp(-i kwave ang), kwave = 2 pi k / zl
        const double th = 2.0 * M_PI / zl * ang;
        double sn, cs;
        sincos(th * t, &sn, &cs);
        cd ph = {cs, -sn};            // exp(-i t th)
        sincos(th * T, &sn, &cs);
        const cd rot = {cs, -sn};     // exp(-i T th)
#pragma unroll
        for (int q = 0; q < B; ++q) {
          const int k = t + T * q;
          D[k] = cmul(X[k], ph);     // (k = 0: ph = 1 exactly)
          ph = cmul(ph, rot);
        }
        if (t == 0) {
          D[N] = cmul(X[N], ph);     // ph = exp(-i N th) after B steps
        }
      } else {
#pragma unroll
        for (int q = 0; q <= B; ++q) {
          const int k = q < B ? t + T * q : N;
          if (q < B || t == 0) {
            cd w = X[k];
            const double kw = k * kfac;
            switch (mode) {
            case Z_DDZ1:
              w = k <= p.kmax ? cd{-w.im * kw, w.re * kw} : cd{0.0, 0.0};
              break;
            case Z_DDZ2:
              w = cscale(w, -kw * kw);
              break;
            case Z_FILTER:
              if (k != p.n0) {
                w = {0.0, 0.0};
              }
              break;
            case Z_LOWPASS:
              if (k > p.kmax || (k == 0 && !p.keep_zonal)) {
                w = {0.0, 0.0};
              }
              break;
            default: // Z_FILTER_LOWPASS: filter(f, n0) then lowPass(., kmax, keep_zonal)
              if (k != p.n0 || k > p.kmax || (k == 0 && !p.keep_zonal)) {
                w = {0.0, 0.0};
              }
              break;
            }
            D[k] = w;
          }
        }
      }
      if (!fused_mod) {
        __syncwarp();
      }
      regfftMergeRead<N>(v, D, N + 1, false, p.tw, t);
      __syncwarp();
      regfftStage1<N, true>(v, p.tw, t);
      regfftXchgWrite<N>(v, D, t);
      __syncwarp();
      regfftStage2<N, true>(v, D, t);
      if (store) {
        const double sc = p.scale ? __ldg(p.scale + i2) : 1.0;
        double2* __restrict__ d2 = reinterpret_cast<double2*>(dstf + i2 * p.nz);
#pragma unroll
        for (int q = 0; q < B; ++q) {
          d2[t + T * q] = make_double2(v[q].re * sc, v[q].im * sc);
        }
      }
      __syncwarp(); // the next output (or row) reuses the buffers
    }
  }
}

template <int N, bool TWO>
int launchZrowsRegT(const ZfftParams& p, long long nrows, cudaStream_t s) {
  using F = RegFft<N>;
  constexpr int ROWS = 128 / F::T;
  const size_t smem = static_cast<size_t>(ROWS) * (TWO ? 2 : 1) * F::SROW * sizeof(cd);
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_zrows_reg<N, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    attr_set = true;
  }
  const long long want = (nrows + ROWS - 1) / ROWS;
  const unsigned blocks = static_cast<unsigned>(want < 148LL * 12 ? want : 148LL * 12);
  k_zrows_reg<N, TWO><<<blocks, 128, smem, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}
template <int N>
int launchZrowsReg(const ZfftParams& p, long long nrows, cudaStream_t s) {
  return p.mode == Z_SLICE2 ? launchZrowsRegT<N, true>(p, nrows, s) : launchZrowsRegT<N, false>(p, nrows, s);
}

template <int N>
int launchZrowsFast(const ZfftParams& p, long long nrows, cudaStream_t s) {
  constexpr int ROWS = 256 / FftShape<N>::TPR;
  const size_t smem = static_cast<size_t>(ROWS) * 4 * (N + 8) * sizeof(cd);
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_zrows_fast<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    attr_set = true;
  }
  const long long want = (nrows + ROWS - 1) / ROWS;
  const unsigned blocks = static_cast<unsigned>(want < 148LL * 4 ? want : 148LL * 4);
  k_zrows_fast<N><<<blocks, 256, smem, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

int launchZfft(const bout_b200_grid* g, int region, ZfftParams p, void* stream, const Box* box = nullptr) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(p.f && p.out, BOUT_B200_ERR_ARG, "zfft: null field");
  Box b;
  if (box != nullptr) {
    b = *box;
  } else if (regionBox(g, region, &b) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  cudaStream_t s = asStream(stream);
  int st = 0;
  p.tw = twiddleTable(g->nz, s, &st);
  if (st != 0) {
    return st;
  }
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = b.xs;
  p.xe = b.xe;
  p.ys = b.ys;
  p.ye = b.ye;
  const long long nrows_all = static_cast<long long>(b.xe - b.xs + 1) * (b.ye - b.ys + 1);
  const int N = g->nz / 2;
  const bool fast_ok = p.mode != Z_ZEROLAP && (g->nz == 128 || g->nz == 256 || g->nz == 512 || g->nz == 1024)
                       && aligned16(p.f) && aligned16(p.out) && (p.out2 == nullptr || aligned16(p.out2));
  static const bool no_reg = getenv("BOUT_B200_ZROWS_REG") != nullptr && getenv("BOUT_B200_ZROWS_REG")[0] == '0';
  if (fast_ok && !no_reg && N <= 256) { // nz = 128, 256, 512: register-resident transforms
    switch (N) {
    case 64: return launchZrowsReg<64>(p, nrows_all, s);
    case 128: return launchZrowsReg<128>(p, nrows_all, s);
    default: return launchZrowsReg<256>(p, nrows_all, s);
    }
  }
  if (fast_ok) {
    switch (N) {
    case 64: return launchZrowsFast<64>(p, nrows_all, s);
    case 128: return launchZrowsFast<128>(p, nrows_all, s);
    case 256: return launchZrowsFast<256>(p, nrows_all, s);
    default: return launchZrowsFast<512>(p, nrows_all, s);
    }
  }
  B200_REQUIRE(p.mode != Z_SLICE2, BOUT_B200_ERR_UNSUPPORTED, "zfft: dual-slice mode needs the fast path");
  const RowGeom geo = rowGeometry(g->nz, 128);
  p.tpr = geo.tpr;
  p.rows = geo.rows;
  p.pow2 = geo.pow2 ? 1 : 0;
  B200_REQUIRE(geo.smemBytes <= 200 * 1024, BOUT_B200_ERR_UNSUPPORTED, "zfft: nz=%d too large", g->nz);
  if (geo.smemBytes > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_zfft_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(geo.smemBytes)));
  }
  const long long nrows = static_cast<long long>(b.xe - b.xs + 1) * (b.ye - b.ys + 1);
  const unsigned blocks = static_cast<unsigned>((nrows + geo.rows - 1) / geo.rows);
  k_zfft_rows<<<blocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

// ---- one Fourier mode of every (x,y) row: M = sum_j f_j exp(-2 pi i n0 j / nz) (unnormalised) ---------------
// Where a term is linear in a field with z-independent coefficients and the result is filtered to the single
// mode n0 afterwards (elm-pb: the y-derivative terms of U, P, Jpar), only this one coefficient of each row is
// needed instead of the field's full parallel slices.
struct RowModeParams {
  const double* f[4];
  double* out[4];
  const cd* __restrict__ tw;
  long long nrows;
  int nf, nz, n0;
};
__global__ void __launch_bounds__(256) k_row_mode(const RowModeParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (w >= p.nrows * p.nf) {
    return;
  }
  const int fi = static_cast<int>(w / p.nrows);
  const long long row = w % p.nrows;
  // (selected without indexing the parameter arrays by a run-time value: that copies the struct to local memory)
  const double* __restrict__ fsel = fi == 0 ? p.f[0] : (fi == 1 ? p.f[1] : (fi == 2 ? p.f[2] : p.f[3]));
  double* __restrict__ osel = fi == 0 ? p.out[0] : (fi == 1 ? p.out[1] : (fi == 2 ? p.out[2] : p.out[3]));
  const double* __restrict__ src = fsel + row * p.nz;
  double re = 0.0, im = 0.0;
  if (p.nz % 64 == 0 && aligned16(src)) {
    // 16-byte loads; three table reads per lane (its two starting factors and the common rotation by n0 * 64
    // points), the factors then advance by that rotation: nz / 64 <= 16 steps, a few ulp -- the table reads per
    // element were four times the L1 wavefronts of the data itself
    const int i0 = static_cast<int>((static_cast<long long>(p.n0) * 2 * lane) % p.nz);
    const int i1 = (i0 + p.n0) % p.nz;
    const cd ws = p.tw[static_cast<int>((static_cast<long long>(p.n0) * 64) % p.nz)];
    cd t0 = p.tw[i0], t1 = p.tw[i1];
    double re1 = 0.0, im1 = 0.0;
#pragma unroll 4
    for (int j = 2 * lane; j < p.nz; j += 64) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(src + j));
      re += v.x * t0.re;
      im += v.x * t0.im;
      re1 += v.y * t1.re;
      im1 += v.y * t1.im;
      const cd n0 = {t0.re * ws.re - t0.im * ws.im, t0.re * ws.im + t0.im * ws.re};
      const cd n1 = {t1.re * ws.re - t1.im * ws.im, t1.re * ws.im + t1.im * ws.re};
      t0 = n0;
      t1 = n1;
    }
    re += re1;
    im += im1;
  } else {
    for (int j = lane; j < p.nz; j += 32) {
      const cd t = p.tw[(static_cast<long long>(p.n0) * j) % p.nz];
      const double v = __ldg(src + j);
      re += v * t.re;
      im += v * t.im;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if (lane == 0) {
    osel[2 * row] = re;
    osel[2 * row + 1] = im;
  }
}

// ---- DC: z-mean of every (x,y) row (src/field/field3d.cxx:785-803) --------------------------
__global__ void __launch_bounds__(256) k_dc(const double* __restrict__ f, double* __restrict__ out,
                                            long long nrows, int nz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (row >= nrows) {
    return;
  }
  const double* __restrict__ src = f + row * nz;
  // sequential-order sum per lane then tree: differs from the reference's left-to-right sum only
  // in rounding (<= nz * eps relative)
  double acc = 0.0;
  for (int j = lane; j < nz; j += 32) {
    acc += __ldg(src + j);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if (lane == 0) {
    out[row] = acc / nz;
  }
}

} // namespace b200

using namespace b200;

extern "C" int bout_b200_ddz_fft(const bout_b200_grid* g, int order, double fft_filter, int region,
                                 const double* f, const double* scale, double* out, void* stream) {
  B200_REQUIRE(order == 1 || order == 2, BOUT_B200_ERR_ARG, "ddz_fft: order must be 1 or 2");
  B200_REQUIRE(g != nullptr, BOUT_B200_ERR_ARG, "ddz_fft: null grid");
  if (g->nz == 1) { // single point: zero field (index_derivs_interface.hxx:152-156)
    B200_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * g->nx * g->ny * g->nz, asStream(stream)));
    return 0;
  }
  ZfftParams p{};
  p.f = f;
  p.scale = scale;
  p.out = out;
  const int ncz = g->nz;
  if (order == 1) { // index_derivs.cxx:435-444
    int kfilter = static_cast<int>(fft_filter * ncz / 2);
    kfilter = kfilter < 0 ? 0 : (kfilter > ncz / 2 ? ncz / 2 : kfilter);
    p.kmax = ncz / 2 - kfilter;
    p.mode = Z_DDZ1;
  } else {
    p.kmax = ncz / 2;
    p.mode = Z_DDZ2;
  }
  return launchZfft(g, region, p, stream);
}

extern "C" int bout_b200_zfilter(const bout_b200_grid* g, int region, int n0, const double* f,
                                 double* out, void* stream) {
  ZfftParams p{};
  p.f = f;
  p.out = out;
  p.mode = Z_FILTER;
  p.n0 = n0;
  return launchZfft(g, region, p, stream);
}

extern "C" int bout_b200_zlowpass(const bout_b200_grid* g, int region, int zmax, int keep_zonal,
                                  const double* f, double* out, void* stream) {
  ZfftParams p{};
  p.f = f;
  p.out = out;
  p.mode = Z_LOWPASS;
  p.kmax = zmax;
  p.keep_zonal = keep_zonal;
  return launchZfft(g, region, p, stream);
}

/* filter(f, n0) followed by lowPass(., zmax, keep_zonal) in one pass (elm_pb.cxx:1931-1955). */
extern "C" int bout_b200_zfilter_lowpass(const bout_b200_grid* g, int region, int n0, int zmax, int keep_zonal,
                                         const double* f, double* out, void* stream) {
  ZfftParams p{};
  p.f = f;
  p.out = out;
  p.mode = Z_FILTER_LOWPASS;
  p.n0 = n0;
  p.kmax = zmax;
  p.keep_zonal = keep_zonal;
  return launchZfft(g, region, p, stream);
}

extern "C" int bout_b200_zshift(const bout_b200_grid* g, int region, const double* f,
                                const double* angle2d, double angle0, const double* zlength2d,
                                double zlength0, double* out, void* stream) {
  ZfftParams p{};
  p.f = f;
  p.out = out;
  p.mode = Z_SHIFT;
  p.angle = angle2d;
  p.angle0 = angle0;
  p.zlength = zlength2d;
  p.zlength0 = zlength0;
  return launchZfft(g, region, p, stream);
}

// ShiftedMetric::calcParallelSlices (src/mesh/parallel/shiftedmetric.cxx:226-246): slice +-(i+1) holds,
// at row y + offset, the field's own row y + offset shifted in z into the frame of row y.
extern "C" int bout_b200_parallel_slices(const bout_b200_grid* g, const double* f, const double* zshift2d,
                                         double zlength, int nslices, double* const* yup,
                                         double* const* ydown, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && zshift2d && yup && ydown, BOUT_B200_ERR_ARG, "parallel_slices: null argument");
  B200_REQUIRE(nslices >= 1 && nslices <= g->ystart && nslices <= g->ny - 1 - g->yend, BOUT_B200_ERR_GUARDS,
               "parallel_slices: %d slices need as many y guard cells", nslices);
  const bool fast_dual = nslices == 1 && (g->nz == 128 || g->nz == 256 || g->nz == 512 || g->nz == 1024)
                         && aligned16(f) && yup[0] && ydown[0] && aligned16(yup[0]) && aligned16(ydown[0])
                         && yup[0] != f && ydown[0] != f && g->ystart >= 1;
  if (fast_dual) { // one read + one forward transform per source row, both slices written
    ZfftParams p{};
    p.f = f;
    p.out = yup[0];
    p.out2 = ydown[0];
    p.mode = Z_SLICE2;
    p.angle = zshift2d;
    p.zlength0 = zlength;
    p.n0 = g->ystart; // target rows ystart..yend
    p.kmax = g->yend;
    const Box b = {0, g->nx - 1, g->ystart - 1, g->yend + 1};
    return launchZfft(g, BOUT_B200_RGN_ALL, p, stream, &b);
  }
  for (int i = 0; i < nslices; ++i) {
    for (int sign = 1; sign >= -1; sign -= 2) {
      double* out = sign > 0 ? yup[i] : ydown[i];
      B200_REQUIRE(out != nullptr && out != f, BOUT_B200_ERR_ARG, "parallel_slices: bad slice pointer");
      const int off = sign * (i + 1);
      if (g->nz == 1) { // shifting does not change the values (shiftedmetric.cxx:159-163)
        B200_CUDA_OK(cudaMemcpyAsync(out, f, sizeof(double) * g->nx * g->ny, cudaMemcpyDeviceToDevice,
                                     asStream(stream)));
        continue;
      }
      ZfftParams p{};
      p.f = f;
      p.out = out;
      p.mode = Z_SLICE;
      p.angle = zshift2d;
      p.zlength0 = zlength;
      p.yoff = off;
      const Box b = {0, g->nx - 1, g->ystart + off, g->yend + off}; // RGN_NOY rows, displaced
      const int rc = launchZfft(g, BOUT_B200_RGN_ALL, p, stream, &b);
      if (rc != 0) {
        return rc;
      }
    }
  }
  return 0;
}

// BoundaryZeroLaplace::apply(Field3D) (src/mesh/boundary_standard.cxx:2630-2686) on the x-inner (side 0)
// or x-outer (side 1) guard cells of rows y = lo..hi (default ystart..yend), in place.
extern "C" int bout_b200_boundary_zerolaplace(const bout_b200_grid* g, int side, const double* dx2d,
                                              const double* g11, const double* g33, double zlength, int lo,
                                              int hi, double* f, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && dx2d && g11 && g33 && (side == 0 || side == 1 || side == 2), BOUT_B200_ERR_ARG,
               "boundary_zerolaplace: bad argument (X boundaries only; 2 = both in one launch)");
  if (side == 2 && (g->xstart == 0 || g->nx - 1 - g->xend == 0)) { // only one side has guard cells
    side = g->xstart > 0 ? 0 : 1;
  }
  const int width = side == 0 ? g->xstart : (side == 1 ? g->nx - 1 - g->xend : g->xstart + g->nx - 1 - g->xend);
  if (width <= 0) {
    return 0;
  }
  B200_REQUIRE(g->xend - g->xstart >= 1, BOUT_B200_ERR_GUARDS, "boundary_zerolaplace: needs 2 interior points");
  if (lo < 0 || hi < 0) {
    lo = g->ystart;
    hi = g->yend;
  }
  ZfftParams p{};
  p.f = f;
  p.out = f;
  p.mode = Z_ZEROLAP;
  p.m_dx = dx2d;
  p.m_g11 = g11;
  p.m_g33 = g33;
  p.zlength0 = zlength;
  p.xin = side == 0 ? g->xstart : g->xend;
  p.bdir = side == 0 ? -1 : 1;
  p.zl_win = -1;
  if (side == 2) { // row groups 0 .. width-1 cover the inner guard cells, then the outer ones
    p.zl_win = g->xstart;
    p.zl_xstart = g->xstart;
    p.zl_xend = g->xend;
  }
  const Box b = side == 0 ? Box{0, g->xstart - 1, lo, hi}
                          : (side == 1 ? Box{g->xend + 1, g->nx - 1, lo, hi} : Box{0, width - 1, lo, hi});
  return launchZfft(g, BOUT_B200_RGN_ALL, p, stream, &b);
}

extern "C" int bout_b200_row_mode(const bout_b200_grid* g, int nfields, const double* const* fields, int n0,
                                  double* const* out_complex2d, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(fields && out_complex2d && nfields >= 1 && nfields <= 4 && n0 >= 0 && n0 <= g->nz / 2,
               BOUT_B200_ERR_ARG, "row_mode: bad argument");
  RowModeParams p{};
  for (int i = 0; i < nfields; ++i) {
    B200_REQUIRE(fields[i] && out_complex2d[i], BOUT_B200_ERR_ARG, "row_mode: null field");
    p.f[i] = fields[i];
    p.out[i] = out_complex2d[i];
  }
  cudaStream_t s = asStream(stream);
  int st = 0;
  p.tw = twiddleTable(g->nz, s, &st);
  if (st != 0) {
    return st;
  }
  p.nrows = static_cast<long long>(g->nx) * g->ny;
  p.nf = nfields;
  p.nz = g->nz;
  p.n0 = n0;
  const long long warps = p.nrows * nfields;
  k_row_mode<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_dc(const bout_b200_grid* g, const double* f, double* out2d, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && out2d, BOUT_B200_ERR_ARG, "dc: null argument");
  const long long nrows = static_cast<long long>(g->nx) * g->ny;
  k_dc<<<static_cast<unsigned>((nrows + 7) / 8), 256, 0, asStream(stream)>>>(f, out2d, nrows, g->nz);
  B200_LAUNCH_OK();
  return 0;
}
