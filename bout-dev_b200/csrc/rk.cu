// Explicit integrator stage updates (subsystem 4), applied field-by-field on device-resident SoA
// fields over the evolving region (RGN_NOBNDRY): the reference's load_vars / save_derivs
// gather-scatter through an interleaved state vector (src/solver/solver.cxx:1039-1309) is not
// needed when fields stay in place.
// Reference arithmetic: RK4Solver::take_step (src/solver/impls/rk4/rk4.cxx:172-210),
// RK3SSP::take_step (src/solver/impls/rk3-ssp/rk3-ssp.cxx:100-129).
#include "common.cuh"

namespace b200 {

enum RkMode { RK_AXPY = 0, RK4_FINAL = 1, RK3_STAGE2 = 2, RK3_STAGE3 = 3 };

struct RkParams {
  const double* a; // start
  const double* k1;
  const double* k2;
  const double* k3;
  const double* k4;
  double* out;
  double dt, c;
  int mode;
  int ny, nz, xs, xe, ys, ye;
};

template <int VEC>
__global__ void __launch_bounds__(256) k_rk(const RkParams p) {
  const int nzv = p.nz / VEC;
  const int nyb = p.ye - p.ys + 1;
  const long long total = static_cast<long long>(p.xe - p.xs + 1) * nyb * nzv;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    int zv, yb_, xb_;
    splitIndex(t, nzv, nyb, zv, yb_, xb_);
    const int y = p.ys + yb_, x = p.xs + xb_;
    const long long i = (static_cast<long long>(x) * p.ny + y) * p.nz + static_cast<long long>(zv) * VEC;
    const Vz<VEC> s = Vz<VEC>::load(p.a + i);
    Vz<VEC> r;
    if (p.mode == RK_AXPY) { // k5 = start + c*dt*k   (rk4.cxx:179-200; rk3-ssp.cxx:107-109 with c = 1)
      const Vz<VEC> k = Vz<VEC>::load(p.k1 + i);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        r.v[j] = s.v[j] + p.c * p.dt * k.v[j];
      }
    } else if (p.mode == RK4_FINAL) { // rk4.cxx:206-209
      const Vz<VEC> a = Vz<VEC>::load(p.k1 + i), b = Vz<VEC>::load(p.k2 + i);
      const Vz<VEC> c = Vz<VEC>::load(p.k3 + i), d = Vz<VEC>::load(p.k4 + i);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        r.v[j] = s.v[j] + (1. / 6.) * p.dt * (a.v[j] + 2. * b.v[j] + 2. * c.v[j] + d.v[j]);
      }
    } else if (p.mode == RK3_STAGE2) { // u2 = 0.75*start + 0.25*u1 + 0.25*dt*L
      const Vz<VEC> u1 = Vz<VEC>::load(p.k1 + i), L = Vz<VEC>::load(p.k2 + i);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        r.v[j] = 0.75 * s.v[j] + 0.25 * u1.v[j] + 0.25 * p.dt * L.v[j];
      }
    } else { // result = (1/3)*start + (2/3)*(u2 + dt*L)
      const Vz<VEC> u2 = Vz<VEC>::load(p.k1 + i), L = Vz<VEC>::load(p.k2 + i);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        r.v[j] = (1. / 3) * s.v[j] + (2. / 3.) * (u2.v[j] + p.dt * L.v[j]);
      }
    }
    r.store(p.out + i);
  }
}

} // namespace b200

using namespace b200;

extern "C" int bout_b200_rk_stage(const bout_b200_grid* g, int region, int mode, double dt, double c,
                                  const double* start, const double* k1, const double* k2,
                                  const double* k3, const double* k4, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(start && k1 && out && mode >= 0 && mode <= 3, BOUT_B200_ERR_ARG, "rk_stage: bad argument");
  B200_REQUIRE(mode == RK_AXPY || k2 != nullptr, BOUT_B200_ERR_ARG, "rk_stage: missing stage array");
  B200_REQUIRE(mode != RK4_FINAL || (k3 && k4), BOUT_B200_ERR_ARG, "rk_stage: RK4 final needs k1..k4");
  Box b;
  if (regionBox(g, region, &b) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  RkParams p{start, k1, k2, k3, k4, out, dt, c, mode, g->ny, g->nz, b.xs, b.xe, b.ys, b.ye};
  const bool vec2 = (g->nz % 2 == 0) && aligned16(start) && aligned16(k1) && aligned16(out)
                    && (!k2 || aligned16(k2)) && (!k3 || aligned16(k3)) && (!k4 || aligned16(k4));
  const long long total = static_cast<long long>(b.xe - b.xs + 1) * (b.ye - b.ys + 1) * (vec2 ? g->nz / 2 : g->nz);
  const long long want = (total + 255) / 256;
  const unsigned blocks = static_cast<unsigned>(want < 148LL * 64 ? (want < 1 ? 1 : want) : 148LL * 64);
  if (vec2) {
    k_rk<2><<<blocks, 256, 0, asStream(stream)>>>(p);
  } else {
    k_rk<1><<<blocks, 256, 0, asStream(stream)>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}

// ---- error norms of the adaptive integrators ----------------------------------------------------------
// MODE 0: sum over the evolving region of |a - b| / (|a| + |b| + atol): RK4Solver::run (rk4.cxx:103-106),
//         RKScheme::getErr (src/solver/impls/rkgeneric/rkscheme.cxx:150-156), SplitRK::run (split-rk.cxx:128-133).
// MODE 1: max over the evolving region of |a - b|: get_error of the Adams-Bashforth solver
//         (src/solver/impls/adams_bashforth/adams_bashforth.cxx:229-253).
// The reference reduces with an OpenMP `reduction(+)` (order not fixed); here the order is fixed: every CTA owns a
// grid-stride subsequence and writes one partial, one CTA then combines the partials in index order.
namespace b200 {

constexpr int kErrThreads = 256;
constexpr int kErrMaxBlocks = 148 * 8;

template <int MODE>
__device__ __forceinline__ double errCombine(double x, double y) {
  return MODE == 0 ? x + y : fmax(x, y);
}

template <int MODE>
__device__ __forceinline__ double blockReduce(double v, double* sh) { // all kErrThreads threads call; result in thread 0
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v = errCombine<MODE>(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh[warp] = v;
  }
  __syncthreads();
  double r = 0.0; // both norms are >= 0, so 0 is the identity of either reduction
  if (threadIdx.x == 0) {
    for (int w = 0; w < kErrThreads / 32; ++w) {
      r = errCombine<MODE>(r, sh[w]);
    }
  }
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(kErrThreads) k_err_partial(const double* __restrict__ a, const double* __restrict__ b,
                                                             double atol, double* __restrict__ partial, int ny, int nz,
                                                             int xs, int xe, int ys, int ye) {
  __shared__ double sh[kErrThreads / 32];
  const int nyb = ye - ys + 1;
  const long long total = static_cast<long long>(xe - xs + 1) * nyb * nz;
  double acc = 0.0;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    int z, yb_, xb_;
    splitIndex(t, nz, nyb, z, yb_, xb_);
    const int y = ys + yb_, x = xs + xb_;
    const long long i = (static_cast<long long>(x) * ny + y) * nz + z;
    const double va = __ldg(a + i), vb = __ldg(b + i);
    if (MODE == 0) {
      acc += fabs(va - vb) / (fabs(va) + fabs(vb) + atol);
    } else {
      acc = fmax(acc, fabs(va - vb));
    }
  }
  const double s = blockReduce<MODE>(acc, sh);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kErrThreads) k_err_final(const double* __restrict__ partial, int n, int accumulate,
                                                           double* __restrict__ result) {
  __shared__ double sh[kErrThreads / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kErrThreads) {
    acc = errCombine<MODE>(acc, partial[i]);
  }
  const double s = blockReduce<MODE>(acc, sh);
  if (threadIdx.x == 0) {
    result[0] = accumulate != 0 ? errCombine<MODE>(result[0], s) : s;
  }
}

template <int MODE>
int errNorm(const bout_b200_grid* g, int region, const double* a, const double* b, double atol, double* scratch,
            int accumulate, double* result, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(a && b && scratch && result, BOUT_B200_ERR_ARG, "error norm: null argument");
  Box bx;
  if (regionBox(g, region, &bx) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  const long long total = static_cast<long long>(bx.xe - bx.xs + 1) * (bx.ye - bx.ys + 1) * g->nz;
  const long long want = (total + kErrThreads - 1) / kErrThreads;
  const int blocks = static_cast<int>(want < kErrMaxBlocks ? (want < 1 ? 1 : want) : kErrMaxBlocks);
  k_err_partial<MODE><<<blocks, kErrThreads, 0, asStream(stream)>>>(a, b, atol, scratch, g->ny, g->nz, bx.xs, bx.xe,
                                                                    bx.ys, bx.ye);
  B200_LAUNCH_OK();
  k_err_final<MODE><<<1, kErrThreads, 0, asStream(stream)>>>(scratch, blocks, accumulate, result);
  B200_LAUNCH_OK();
  return 0;
}

} // namespace b200

extern "C" size_t bout_b200_err_sum_scratch_doubles(void) { return static_cast<size_t>(kErrMaxBlocks); }

extern "C" int bout_b200_err_sum(const bout_b200_grid* g, int region, const double* a, const double* b, double atol,
                                 double* scratch, int accumulate, double* result, void* stream) {
  return errNorm<0>(g, region, a, b, atol, scratch, accumulate, result, stream);
}

extern "C" int bout_b200_err_max(const bout_b200_grid* g, int region, const double* a, const double* b,
                                 double* scratch, int accumulate, double* result, void* stream) {
  return errNorm<1>(g, region, a, b, 0.0, scratch, accumulate, result, stream);
}
