// Shared device/host helpers for libbout_b200 (sm_100a, fp64).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/bout_b200.h"

namespace b200 {

// ---- error plumbing: C-ABI functions return a status, the message is kept per host thread ----
std::string& lastError();
int fail(int status, const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void countLaunch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define B200_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t err__ = (expr);                                                             \
    if (err__ != cudaSuccess) {                                                             \
      return b200::fail(BOUT_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                        cudaGetErrorString(err__), __FILE__, __LINE__);                     \
    }                                                                                       \
  } while (0)

#define B200_LAUNCH_OK()                                                                    \
  do {                                                                                      \
    cudaError_t err__ = cudaGetLastError();                                                 \
    if (err__ != cudaSuccess) {                                                             \
      return b200::fail(BOUT_B200_ERR_CUDA, "kernel launch failed: %s (%s:%d)",             \
                        cudaGetErrorString(err__), __FILE__, __LINE__);                     \
    }                                                                                       \
    b200::countLaunch();                                                                    \
  } while (0)

#define B200_REQUIRE(cond, status, ...)                                                     \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      return b200::fail(status, __VA_ARGS__);                                               \
    }                                                                                       \
  } while (0)

inline cudaStream_t asStream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- grid / region -------------------------------------------------------------------------
struct Box { // inclusive index box over (x, y); z is always the full periodic row
  int xs, xe, ys, ye;
};

inline int regionBox(const bout_b200_grid* g, int region, Box* b) {
  switch (region) {
  case BOUT_B200_RGN_NOBNDRY: *b = {g->xstart, g->xend, g->ystart, g->yend}; return 0;
  case BOUT_B200_RGN_ALL: *b = {0, g->nx - 1, 0, g->ny - 1}; return 0;
  case BOUT_B200_RGN_NOX: *b = {g->xstart, g->xend, 0, g->ny - 1}; return 0;
  case BOUT_B200_RGN_NOY: *b = {0, g->nx - 1, g->ystart, g->yend}; return 0;
  default: return fail(BOUT_B200_ERR_ARG, "unknown region %d", region);
  }
}

inline int checkGrid(const bout_b200_grid* g) {
  if (g == nullptr || g->nx < 1 || g->ny < 1 || g->nz < 1 || g->xstart < 0 || g->ystart < 0
      || g->xend >= g->nx || g->yend >= g->ny || g->xend < g->xstart || g->yend < g->ystart) {
    return fail(BOUT_B200_ERR_ARG, "invalid grid descriptor");
  }
  return 0;
}

// Functions shared by the kernels and the host-side emulation of tests/emul/ (which steps simulated threads
// through the same code): callable from both sides, read-only loads through the non-coherent path on the device.
#define B200_HD __host__ __device__ __forceinline__
#define B200_HD_COLD __host__ __device__ __noinline__ // rare paths kept out of the unrolled hot loops
#ifdef __CUDA_ARCH__
#define B200_LDG(p) __ldg(p)
#else
#define B200_LDG(p) (*(p))
#endif

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<size_t>(p) & 15u) == 0; }

// ---- z-vectorised access: VEC = 2 uses 16-byte double2 transactions (nz even, base 16B aligned)
template <int VEC>
struct Vz;
template <>
struct Vz<1> {
  double v[1];
  __device__ __forceinline__ static Vz load(const double* __restrict__ p) {
    Vz r;
    r.v[0] = __ldg(p);
    return r;
  }
  __device__ __forceinline__ void store(double* __restrict__ p) const { p[0] = v[0]; }
};
template <>
struct Vz<2> {
  double v[2];
  __device__ __forceinline__ static Vz load(const double* __restrict__ p) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    Vz r;
    r.v[0] = t.x;
    r.v[1] = t.y;
    return r;
  }
  __device__ __forceinline__ void store(double* __restrict__ p) const {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  }
};

__device__ __forceinline__ int wrapz(int z, int nz) {
  // z in [-nz, 2nz)
  return z < 0 ? z + nz : (z >= nz ? z - nz : z);
}

// Flat work-item index t -> (z-vector zv < nzv, row yb < nyb, plane xb): 32-bit unsigned arithmetic whenever t fits
// (a 64-bit division by a run-time value is ~100 instructions on sm_100: four of them per element made the
// element-wise kernels integer-bound), a mask / shift when nzv is a power of two (nz almost always is).
__device__ __forceinline__ void splitIndex(long long t, int nzv, int nyb, int& zv, int& yb, int& xb) {
  if (t <= 0x7fffffffLL) {
    const unsigned u = static_cast<unsigned>(t), n = static_cast<unsigned>(nzv), m = static_cast<unsigned>(nyb);
    unsigned xy;
    if ((n & (n - 1)) == 0) {
      xy = u >> (__ffs(static_cast<int>(n)) - 1);
      zv = static_cast<int>(u & (n - 1));
    } else {
      xy = u / n;
      zv = static_cast<int>(u - xy * n);
    }
    const unsigned xq = xy / m;
    yb = static_cast<int>(xy - xq * m);
    xb = static_cast<int>(xq);
  } else {
    zv = static_cast<int>(t % nzv);
    const long long xy = t / nzv;
    yb = static_cast<int>(xy % nyb);
    xb = static_cast<int>(xy / nyb);
  }
}

// Cycles a receiver waits: BOUT_B200_HALO_TIMEOUT_S seconds (default 60: first-call IPC mapping, module loads, a
// neighbour busy writing output are all far below it; a dead neighbour is not) at the 1.965 GHz B200 clock.
inline long long haloTimeoutCycles() {
  static const long long cycles = [] {
    const char* env = getenv("BOUT_B200_HALO_TIMEOUT_S");
    const double secs = env != nullptr && atof(env) > 0.0 ? atof(env) : 60.0;
    return static_cast<long long>(secs * 1.965e9);
  }();
  return cycles;
}

// Reciprocal of a normal, non-zero double (the callers' arguments are bounded away from 0 by construction): hardware
// seed (rcp.approx.ftz.f64, the upper word only) + two Newton steps = 1 MUFU + 4 DFMA, <= 2 ulp, instead of the IEEE
// division sequence with its special-case path (~25 issue slots).
__device__ __forceinline__ double rcpPositive(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

} // namespace b200
