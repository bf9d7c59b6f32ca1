// Fused Delp2 kernels (subsystem 2b): one kernel per RHS term, no temporaries.
//
//  * bout_b200_delp2_fft : Coordinates::Delp2(Field3D, useFFT=true)
//    (reference src/mesh/coordinates.cxx:1662-1699): rfft each x row, a*F[x-1] + b*F[x] + c*F[x+1]
//    with Laplacian::tridagCoefs of the default instance, irfft.  A thread group marches along x
//    keeping the spectra of rows x-1, x, x+1 in shared memory, so each row is transformed once.
//  * bout_b200_delp2_fd  : the useFFT=false branch (coordinates.cxx:1701-1702) for C2 methods,
//    G1 DDX + G3 DDZ + g11 D2DX2 + g33 D2DZ2 + 2 g13 D2DXDZ as one 9-point X-Z stencil pass
//    (same arithmetic as include/bout/single_index_ops.hxx:222-246).
#include <cstdlib>

#include "bracket.cuh"
#include "fft_fast.cuh"
#include "fft_reg.cuh"
#include "laplace_coefs.cuh"

namespace b200 {

struct Delp2FftParams {
  LapCoefs m;
  const double* __restrict__ f;
  double* __restrict__ out;
  const cd* __restrict__ tw;
  int nx, ny, nz, xs, xe, ys, ye;
  int chunk, nchunks;
  int tpr, rows, pow2;
};

__global__ void __launch_bounds__(128) k_delp2_fft(const Delp2FftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nbuf = fftBufElems(p.nz);
  const int rloc = threadIdx.x / p.tpr;
  RowFFT r;
  r.nz = p.nz;
  r.N = p.nz / 2;
  r.pow2 = p.pow2 != 0;
  r.tpr = p.tpr;
  r.lane = threadIdx.x % p.tpr;
  cd* base = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * 5 * nbuf;
  r.A = base;
  r.B = base + nbuf;
  r.tw = p.tw;
  cd* S[3] = {base + 2 * nbuf, base + 3 * nbuf, base + 4 * nbuf};

  const int nyb = p.ye - p.ys + 1;
  const long long ntasks = static_cast<long long>(nyb) * p.nchunks;
  const long long task0 = static_cast<long long>(blockIdx.x) * p.rows + rloc;
  const bool task_ok = task0 < ntasks;
  const long long task = task_ok ? task0 : ntasks - 1;
  const int y = p.ys + static_cast<int>(task / p.nchunks);
  const int x0 = p.xs + static_cast<int>(task % p.nchunks) * p.chunk;
  const int x1 = min(x0 + p.chunk - 1, p.xe);
  const int nk = p.nz / 2 + 1;
  const long long plane = static_cast<long long>(p.ny) * p.nz;
  const double* __restrict__ col = p.f + static_cast<long long>(y) * p.nz;

  auto forwardTo = [&](int x, cd* dst) {
    // the CTA is synchronised on entry; on exit dst holds the spectrum and the CTA is synchronised
    double* a = reinterpret_cast<double*>(r.A);
    const double* __restrict__ src = col + x * plane;
    if (p.nz % 2 == 0 && aligned16(src)) {
      for (int j = r.lane; j < p.nz / 2; j += r.tpr) {
        reinterpret_cast<double2*>(a)[j] = __ldg(reinterpret_cast<const double2*>(src) + j);
      }
    } else {
      for (int j = r.lane; j < p.nz; j += r.tpr) {
        a[j] = __ldg(src + j);
      }
    }
    __syncthreads();
    const cd* X = rfftForward(r);
    for (int k = r.lane; k < nk; k += r.tpr) {
      dst[k] = X[k];
    }
    __syncthreads();
  };

  forwardTo(x0 - 1, S[(x0 - 1) % 3]);
  forwardTo(x0, S[x0 % 3]);
  for (int it = 0; it < p.chunk; ++it) {
    const int x = min(x0 + it, x1); // short chunks repeat their last row (never stored twice)
    const bool store = task_ok && (x0 + it <= x1);
    forwardTo(x + 1, S[(x + 1) % 3]);
    const cd* Sm = S[(x + 2) % 3]; // (x - 1) mod 3
    const cd* S0 = S[x % 3];
    const cd* Sp = S[(x + 1) % 3];
    for (int k = r.lane; k < nk; k += r.tpr) {
      cd a, b, c;
      const double kwave = k * 2.0 * M_PI / p.m.zlength; // invert_laplace.cxx:263
      tridagCoefs(p.m, x, y, kwave, false, false, a, b, c);
      r.A[k] = cadd(cadd(cmul(a, Sm[k]), cmul(b, S0[k])), cmul(c, Sp[k]));
    }
    __syncthreads();
    const double* res = rfftBackward(r, r.A);
    if (store) {
      double* __restrict__ dst = p.out + x * plane + static_cast<long long>(y) * p.nz;
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
    __syncthreads();
  }
}


// Fast path (nz in {128, 256, 512, 1024}, 16-byte aligned rows): same x-march, with the fft_fast.cuh
// transforms (register twiddles, swizzled buffers), 256-thread CTAs and per-row barriers, so the row groups of
// a CTA run decoupled.  Shared memory per row group: the two FFT buffers, the three spectra, the combination.
template <int N>
__global__ void __launch_bounds__(256) k_delp2_fast(const Delp2FftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TPR = FftShape<N>::TPR, BUF = N + 8;
  const int ROWS = blockDim.x / TPR; // row groups (tasks) per CTA: 256 / TPR, fewer when there are few tasks
  const int rloc = threadIdx.x / TPR, lane = threadIdx.x % TPR;
  cd* base = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * 6 * BUF;
  cd* A = base;
  cd* B = base + BUF;
  cd* S[3] = {base + 2 * BUF, base + 3 * BUF, base + 4 * BUF};
  cd* D = base + 5 * BUF;
  FastTw<N> tw;
  fastTwInit<N>(tw, p.tw, lane);
  const SyncGroup sync{1 + rloc, TPR};

  const int nyb = p.ye - p.ys + 1;
  const long long ntasks = static_cast<long long>(nyb) * p.nchunks;
  const long long task0 = static_cast<long long>(blockIdx.x) * ROWS + rloc;
  const bool task_ok = task0 < ntasks;
  const long long task = task_ok ? task0 : ntasks - 1;
  const int y = p.ys + static_cast<int>(task / p.nchunks);
  const int x0 = p.xs + static_cast<int>(task % p.nchunks) * p.chunk;
  const int x1 = min(x0 + p.chunk - 1, p.xe);
  const long long plane = static_cast<long long>(p.ny) * p.nz;
  const double* __restrict__ col = p.f + static_cast<long long>(y) * p.nz;

  auto forwardTo = [&](int x, cd* dst) { // the row group is synchronised on entry and on exit
    const double2* __restrict__ s2 = reinterpret_cast<const double2*>(col + x * plane);
    cd v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double2 q = __ldg(s2 + lane + m * (N / 4));
      v[m] = {q.x, q.y};
    }
    const cd* Z = cfftFastFromRegs<N, false, SyncGroup>(v, A, B, tw, lane, sync);
    rfftSplitTo<N>(Z, dst, N + 1, tw, lane);
    sync();
  };

  forwardTo(x0 - 1, S[(x0 - 1) % 3]);
  forwardTo(x0, S[x0 % 3]);
  for (int it = 0; it < p.chunk; ++it) {
    const int x = min(x0 + it, x1); // short chunks repeat their last row (never stored twice)
    const bool store = task_ok && (x0 + it <= x1);
    forwardTo(x + 1, S[(x + 1) % 3]);
    const cd* Sm = S[(x + 2) % 3]; // (x - 1) mod 3
    const cd* S0 = S[x % 3];
    const cd* Sp = S[(x + 1) % 3];
    for (int k = lane; k <= N; k += TPR) {
      cd a, b, c;
      const double kwave = k * 2.0 * M_PI / p.m.zlength; // invert_laplace.cxx:263
      tridagCoefs(p.m, x, y, kwave, false, false, a, b, c);
      D[k] = cadd(cadd(cmul(a, Sm[k]), cmul(b, S0[k])), cmul(c, Sp[k]));
    }
    sync();
    cd w[4], r[4];
    rfftMergeRegs<N>([&](int k) -> cd { return D[k]; }, w, tw, lane);
    cfftFastRegsToRegs<N, true, SyncGroup>(w, r, A, B, tw, lane, sync);
    if (store) {
      double2* __restrict__ d2 = reinterpret_cast<double2*>(p.out + x * plane + static_cast<long long>(y) * p.nz);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        d2[lane + m * (N / 4)] = make_double2(r[m].re, r[m].im);
      }
    }
  }
}

template <int N>
int launchDelp2Fast(Delp2FftParams p, cudaStream_t s) {
  constexpr int TPR = FftShape<N>::TPR, ROWS_MAX = 256 / TPR;
  const int len = p.xe - p.xs + 1, nyb = p.ye - p.ys + 1;
  // (y, x-chunk) tasks: enough of them to occupy the SMs; each chunk costs two extra forward transforms, so chunks
  // shrink only for small problems and not below 4 rows (measured on blob2d's 1028 x 1 x 1024, the one shape this
  // kernel still serves: 272 / 258 / 247 / 293 us per RHS with chunks of 1 / 2 / 4 / 8)
  int chunk = 32;
  while (chunk > 4 && static_cast<long long>(nyb) * ((len + chunk - 1) / chunk) < 148LL * 4 * ROWS_MAX) {
    chunk /= 2;
  }
  static const int chunk_env = getenv("BOUT_B200_DELP2_CHUNK") ? atoi(getenv("BOUT_B200_DELP2_CHUNK")) : 0;
  if (chunk_env > 0) {
    chunk = chunk_env;
  }
  p.chunk = chunk;
  p.nchunks = (len + chunk - 1) / chunk;
  const long long ntasks = static_cast<long long>(nyb) * p.nchunks;
  int rows = ROWS_MAX; // tasks per CTA: spread few tasks over many SMs
  while (rows > 1 && (ntasks + rows - 1) / rows < 148LL * 2) {
    rows /= 2;
  }
  const int threads = rows * TPR < 32 ? 32 : rows * TPR;
  const int rows_eff = threads / TPR;
  const size_t smem = static_cast<size_t>(rows_eff) * 6 * (N + 8) * sizeof(cd);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_delp2_fast<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    attr = smem;
  }
  k_delp2_fast<N><<<static_cast<unsigned>((ntasks + rows_eff - 1) / rows_eff), threads, smem, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

// Register-transform path (nz in {128, 256, 512}): no march.  A CTA takes ROWS = NT / T consecutive x rows of one y
// (the first and the last are the x - 1 / x + 1 neighbours of the ROWS - 2 rows it produces), every row group of T
// threads transforms its row in registers (fft_reg.cuh: one shared-memory exchange per transform) and leaves the
// spectrum in the row's buffer; after ONE CTA barrier each group forms the packed input of its inverse transform
// straight from the three neighbouring spectra, a X[x-1] + b X[x] + c X[x+1] for bins k and N - k, and transforms
// back.  2 / ROWS of the forward transforms are redundant (6 % at nz = 256); shared memory holds only the spectra.
template <int N, int NT>
__global__ void __launch_bounds__(NT) k_delp2_reg(const Delp2FftParams p) {
  using F = RegFft<N>;
  constexpr int T = F::T, B = F::B, SROW = F::SROW, ROWS = NT / T, OUTR = ROWS - 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* S = reinterpret_cast<cd*>(smem_raw);
  const int tid = threadIdx.x, grp = tid / T, t = tid % T;
  const int ntile = p.nchunks;
  const int y = p.ys + blockIdx.x / ntile;
  const int x = p.xs + (blockIdx.x % ntile) * OUTR - 1 + grp; // this group's row
  const bool have = x <= p.xe + 1;
  cd* row = S + grp * SROW;
  const long long plane = static_cast<long long>(p.ny) * p.nz;
  cd v[B];
  if (have) {
    const double2* __restrict__ s2 =
        reinterpret_cast<const double2*>(p.f + x * plane + static_cast<long long>(y) * p.nz);
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) {
      const double2 q = B200_LDG(s2 + t + T * n2);
      v[n2] = {q.x, q.y};
    }
  } else {
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) {
      v[n2] = {0.0, 0.0};
    }
  }
  regfftStage1<N, false>(v, p.tw, t);
  regfftXchgWrite<N>(v, row, t);
  __syncwarp();
  regfftStage2<N, false>(v, row, t);
  __syncwarp();
  regfftSplitWriteZ<N>(v, row, t);
  __syncwarp();
  {
    cd pc[B / 2 + 1];
    regfftSplitReadPartners<N>(pc, row, t);
    __syncwarp();
    regfftSplitWriteX<N>(v, pc, row, p.tw, t);
  }
  __syncthreads();
  const bool produce = grp >= 1 && grp <= OUTR && x <= p.xe;
  if (produce) {
    double k5[5];
    tridagCoef5(p.m, x, y, false, false, k5);
    const double kscale = 2.0 * M_PI / p.m.zlength; // invert_laplace.cxx:263
    const cd *Xm = row - SROW, *X0 = row, *Xp = row + SROW;
    auto bin = [&](int k) -> cd {
      cd a, b, c;
      coefsFromK(k5, k * kscale, a, b, c);
      return cadd(cadd(cmul(a, Xm[k]), cmul(b, X0[k])), cmul(c, Xp[k]));
    };
    const cd ws = ldtw(p.tw, t);
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) { // regfftMergeRead with the combined spectrum as its input
      const int k = t + T * n2;
      cd xk = bin(k);
      cd xc = cconj(bin(N - k));
      if (k == 0) { // imaginary parts of DC and Nyquist are ignored (FFTW c2r)
        xk.im = 0.0;
        xc.im = 0.0;
      }
      const cd wk = cconj(regfftSplitTw<N>(ws, n2));
      const cd d = cmul(wk, csub(xk, xc));
      v[n2] = cadd(cadd(xk, xc), cd{-d.im, d.re});
    }
  }
  __syncthreads(); // every group has read its neighbours' spectra: the buffers become exchange areas again
  if (!produce) {
    return;
  }
  regfftStage1<N, true>(v, p.tw, t);
  regfftXchgWrite<N>(v, row, t);
  __syncwarp();
  regfftStage2<N, true>(v, row, t);
  double2* __restrict__ d2 = reinterpret_cast<double2*>(p.out + x * plane + static_cast<long long>(y) * p.nz);
#pragma unroll
  for (int q = 0; q < B; ++q) {
    d2[t + T * q] = make_double2(v[q].re, v[q].im);
  }
}

template <int N, int NT>
int launchDelp2Reg(Delp2FftParams p, cudaStream_t s) {
  using F = RegFft<N>;
  constexpr int ROWS = NT / F::T, OUTR = ROWS - 2;
  const int len = p.xe - p.xs + 1, nyb = p.ye - p.ys + 1;
  p.nchunks = (len + OUTR - 1) / OUTR;
  const size_t smem = static_cast<size_t>(ROWS) * F::SROW * sizeof(cd);
  static bool configured = false;
  if (!configured) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_delp2_reg<N, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    configured = true;
  }
  k_delp2_reg<N, NT><<<static_cast<unsigned>(nyb) * p.nchunks, NT, smem, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

struct Delp2FdParams {
  const double* __restrict__ f;
  const double* __restrict__ dx;
  const double* __restrict__ dz;
  const double* __restrict__ g11;
  const double* __restrict__ g33;
  const double* __restrict__ g13;
  const double* __restrict__ G1;
  const double* __restrict__ G3;
  const double* __restrict__ d1_dx; // NULL when the mesh is uniform (coords->non_uniform false)
  double* __restrict__ out;
  int nx, ny, nz, xs, xe, ys, ye, chunk;
};

// Per-point arithmetic, shared with the fused RHS kernels.  Term order follows
// coordinates.cxx:1701-1702 with each operator as derivs.cxx defines it for C2.
__device__ __forceinline__ double delp2FdPoint(const Row4& fm, const Row4& f0, const Row4& fp, int i,
                                               double rdx, double rdz, double g11, double g33,
                                               double g13, double G1, double G3, double d1_dx,
                                               bool nonuni) {
  const int zm = i, z = i + 1, zp = i + 2;
  const double ddx = (0.5 * (fp.v[z] - fm.v[z])) * rdx;
  const double ddz = (0.5 * (f0.v[zp] - f0.v[zm])) * rdz;
  double d2dx2 = (fp.v[z] + fm.v[z] - 2. * f0.v[z]) * (rdx * rdx);
  if (nonuni) {
    d2dx2 += (d1_dx * (0.5 * (fp.v[z] - fm.v[z]))) * rdx;
  }
  const double d2dz2 = (f0.v[zp] + f0.v[zm] - 2. * f0.v[z]) * (rdz * rdz);
  // D2DXDZ = DDZ(DDX(f))
  const double ddx_zp = (0.5 * (fp.v[zp] - fm.v[zp])) * rdx;
  const double ddx_zm = (0.5 * (fp.v[zm] - fm.v[zm])) * rdx;
  const double d2dxdz = (0.5 * (ddx_zp - ddx_zm)) * rdz;
  return G1 * ddx + G3 * ddz + g11 * d2dx2 + g33 * d2dz2 + 2 * g13 * d2dxdz;
}

template <int VEC>
__global__ void __launch_bounds__(256) k_delp2_fd(const Delp2FdParams p) {
  const int nzv = p.nz / VEC;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(p.ye - p.ys + 1) * nzv) {
    return;
  }
  const int z = static_cast<int>(t % nzv) * VEC;
  const int y = p.ys + static_cast<int>(t / nzv);
  const int x0 = p.xs + blockIdx.y * p.chunk;
  const int x1 = min(x0 + p.chunk - 1, p.xe);
  const long long plane = static_cast<long long>(p.ny) * p.nz;
  const long long rowoff = static_cast<long long>(y) * p.nz;
  Row4 Fm, F0, Fp;
  F0 = loadRow4<VEC>(p.f + (x0 - 1) * plane + rowoff, z, p.nz);
  Fp = loadRow4<VEC>(p.f + x0 * plane + rowoff, z, p.nz);
#pragma unroll 2
  for (int x = x0; x <= x1; ++x) {
    Fm = F0;
    F0 = Fp;
    Fp = loadRow4<VEC>(p.f + (x + 1) * plane + rowoff, z, p.nz);
    const long long i2 = static_cast<long long>(x) * p.ny + y;
    const double rdx = 1.0 / __ldg(p.dx + i2), rdz = 1.0 / __ldg(p.dz + i2);
    const double g11 = __ldg(p.g11 + i2), g33 = __ldg(p.g33 + i2), g13 = __ldg(p.g13 + i2);
    const double G1 = __ldg(p.G1 + i2), G3 = __ldg(p.G3 + i2);
    const double d1 = p.d1_dx ? __ldg(p.d1_dx + i2) : 0.0;
    Vz<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      r.v[i] = delp2FdPoint(Fm, F0, Fp, i, rdx, rdz, g11, g33, g13, G1, G3, d1, p.d1_dx != nullptr);
    }
    r.store(p.out + x * plane + rowoff + z);
  }
}

} // namespace b200

using namespace b200;

extern "C" int bout_b200_delp2_fft(const bout_b200_grid* g, const bout_b200_laplace* L,
                                   const double* f, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(L && f && out && f != out, BOUT_B200_ERR_ARG, "delp2_fft: bad argument");
  B200_REQUIRE(L->dx && L->g11 && L->g33 && L->g13, BOUT_B200_ERR_ARG, "delp2_fft: metric missing");
  B200_REQUIRE(!L->all_terms || (L->G1 && L->G3), BOUT_B200_ERR_ARG, "delp2_fft: all_terms needs G1, G3");
  B200_REQUIRE(g->xstart >= 1 && g->xend <= g->nx - 2, BOUT_B200_ERR_GUARDS,
               "delp2_fft: needs one x guard cell each side"); // ASSERT2(xstart > 0), coordinates.cxx:1657
  cudaStream_t s = asStream(stream);
  Delp2FftParams p{};
  p.m.dx = L->dx;
  p.m.g11 = L->g11;
  p.m.g33 = L->g33;
  p.m.g13 = L->g13;
  p.m.G1 = L->G1;
  p.m.G3 = L->G3;
  p.m.zlength = L->zlength;
  p.m.nx = g->nx;
  p.m.ny = g->ny;
  p.m.all_terms = L->all_terms;
  p.m.nonuniform = L->nonuniform;
  p.f = f;
  p.out = out;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = g->xstart;
  p.xe = g->xend;
  p.ys = g->ystart;
  p.ye = g->yend;
  int st = 0;
  p.tw = twiddleTable(g->nz, s, &st);
  if (st != 0) {
    return st;
  }
  const bool fast_ok = aligned16(f) && aligned16(out) && getenv("BOUT_B200_NO_FAST_FFT") == nullptr;
  static const bool no_reg = getenv("BOUT_B200_DELP2_REG") != nullptr && getenv("BOUT_B200_DELP2_REG")[0] == '0';
  // the register-transform kernel for every problem size it supports: it has 2 / ROWS redundant transforms where a
  // marching task has 2 per chunk, and on small 2-D grids (hasegawa-wakatani 256 x 256: four Delp2 per RHS) its one
  // round of transforms is a third of the latency of a march (166 -> 148 us per RHS through the host layer)
  if (fast_ok && !no_reg && (g->nz == 128 || g->nz == 256 || g->nz == 512)) {
    switch (g->nz / 2) {
    case 64: return launchDelp2Reg<64, 256>(p, s);
    case 128: return launchDelp2Reg<128, 256>(p, s);
    default: return launchDelp2Reg<256, 256>(p, s);
    }
  }
  if ((g->nz == 128 || g->nz == 256 || g->nz == 512 || g->nz == 1024) && fast_ok) {
    switch (g->nz / 2) {
    case 64: return launchDelp2Fast<64>(p, s);
    case 128: return launchDelp2Fast<128>(p, s);
    case 256: return launchDelp2Fast<256>(p, s);
    default: return launchDelp2Fast<512>(p, s);
    }
  }
  RowGeom geo = rowGeometry(g->nz, 128);
  geo.smemBytes = static_cast<size_t>(geo.rows) * 5 * fftBufElems(g->nz) * sizeof(cd);
  p.tpr = geo.tpr;
  p.rows = geo.rows;
  p.pow2 = geo.pow2 ? 1 : 0;
  B200_REQUIRE(geo.smemBytes <= 200 * 1024, BOUT_B200_ERR_UNSUPPORTED, "delp2_fft: nz=%d too large", g->nz);
  if (geo.smemBytes > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_delp2_fft, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(geo.smemBytes)));
  }
  const int len = p.xe - p.xs + 1, nyb = p.ye - p.ys + 1;
  // enough (y, chunk) tasks to fill 148 SMs a few times over, chunks no shorter than 8 rows
  int chunk = 32;
  while (chunk > 8 && static_cast<long long>(nyb) * ((len + chunk - 1) / chunk) < 148LL * 8 * geo.rows) {
    chunk /= 2;
  }
  p.chunk = chunk;
  p.nchunks = (len + chunk - 1) / chunk;
  const long long ntasks = static_cast<long long>(nyb) * p.nchunks;
  const unsigned blocks = static_cast<unsigned>((ntasks + geo.rows - 1) / geo.rows);
  k_delp2_fft<<<blocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_delp2_fd(const bout_b200_grid* g, const double* f, const double* dx,
                                  const double* dz, const double* g11, const double* g33,
                                  const double* g13, const double* G1, const double* G3,
                                  const double* d1_dx, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && dx && dz && g11 && g33 && g13 && G1 && G3 && out && f != out, BOUT_B200_ERR_ARG,
               "delp2_fd: bad argument");
  B200_REQUIRE(g->xstart >= 1 && g->xend <= g->nx - 2, BOUT_B200_ERR_GUARDS,
               "delp2_fd: needs one x guard cell each side");
  Delp2FdParams p{f, dx, dz, g11, g33, g13, G1, G3, d1_dx, out, g->nx, g->ny, g->nz,
                  g->xstart, g->xend, g->ystart, g->yend, 16};
  const bool vec2 = (g->nz % 2 == 0) && aligned16(f) && aligned16(out);
  const int nzv = vec2 ? g->nz / 2 : g->nz;
  const long long north = static_cast<long long>(p.ye - p.ys + 1) * nzv;
  const int len = p.xe - p.xs + 1;
  int chunk = 8;
  for (int c : {64, 32, 16}) {
    if (north * ((len + c - 1) / c) >= 148LL * 2048) {
      chunk = c;
      break;
    }
  }
  p.chunk = chunk;
  dim3 grid(static_cast<unsigned>((north + 255) / 256), static_cast<unsigned>((len + chunk - 1) / chunk));
  if (vec2) {
    k_delp2_fd<2><<<grid, 256, 0, asStream(stream)>>>(p);
  } else {
    k_delp2_fd<1><<<grid, 256, 0, asStream(stream)>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}
