// Fused physics-model RHS kernels: "one kernel per RHS term set" with the explicit integrator's
// stage update folded into the same pass (subsystems 2 + 4 of the hot path).
//
// The reference ships, next to each operator-at-a-time model, a hand-fused "outer loop" variant
// that evaluates the whole right-hand side in one loop with single-index operators
// (examples/hasegawa-wakatani-3d/hw.cxx:58-74, examples/blob2d-outerloop, examples/elm-pb-outerloop;
// operators in include/bout/single_index_ops.hxx).  These kernels are the B200 statement of those
// loops:
//   * every thread owns a (y, z-pair) column and marches along x with the x-1, x, x+1 planes of
//     each evolving field held in registers (z-1..z+2 of each), so n, vort and phi are read from
//     L2/HBM once per chunk and feed both brackets, both Delp2 terms and DDZ;
//   * y-neighbour values for the parallel term are read directly (L2 hits: the neighbouring
//     CTAs stream the same rows);
//   * all metric combinations that depend on (x,y) only are produced once by a small prologue
//     kernel instead of per point (fp64 divisions are ~12-25 issue slots each on sm_100);
//   * optional epilogue: the RK stage update state' = start + a*ddt, acc' = acc + b*ddt written in
//     the same pass, so no separate axpy kernels run (reference RK4: rk4.cxx:172-210).
#include "bracket.cuh"

#include <algorithm>
#include <cstdlib>

namespace b200 {

// ------------------------------------------------------------------------------------------------
// Per-(x,y) coefficients shared by the fused kernels
// ------------------------------------------------------------------------------------------------
enum { C_SF = 0,   // 1 / (12 dx dz)                         bracket spacing factor
       C_RDZ2,     // 1 / (2 dz)                             DDZ
       C_DPX,      // (G1 + d1_dx g11) / (2 dx)              Delp2: DDX coefficient
       C_DPZ,      // G3 / (2 dz)
       C_DPXX,     // g11 / dx^2
       C_DPZZ,     // g33 / dz^2
       C_DPXZ,     // 2 g13 / (4 dz dx)
       C_UP,       // 2 (J + Jp) / ((dy + dyp)(g_22 + g_22p)) / (dy J)   Div_par_Grad_par upper flux
       C_DN,       // same with the lower neighbour
       C_NCOEF };

struct CoefIn {
  const double* dx;
  const double* dy;
  const double* dz;
  const double* J;
  const double* g_22;
  const double* g11;
  const double* g33;
  const double* g13;
  const double* G1;
  const double* G3;
  const double* d1_dx; // may be NULL (uniform mesh)
  int nx, ny;
};

__global__ void __launch_bounds__(256) k_fused_coefs(const CoefIn m, double* __restrict__ coef, int with_par) {
  const int n2 = m.nx * m.ny;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) {
    return;
  }
  const int y = i % m.ny;
  const double dx = m.dx[i], dz = m.dz[i];
  coef[C_SF * n2 + i] = 1.0 / (12 * dx * dz);
  coef[C_RDZ2 * n2 + i] = 1.0 / (2.0 * dz);
  const double g11 = m.g11[i];
  const double d1 = m.d1_dx ? m.d1_dx[i] : 0.0;
  coef[C_DPX * n2 + i] = (m.G1[i] + d1 * g11) / (2.0 * dx);
  coef[C_DPZ * n2 + i] = m.G3[i] / (2.0 * dz);
  coef[C_DPXX * n2 + i] = g11 / (dx * dx);
  coef[C_DPZZ * n2 + i] = m.g33[i] / (dz * dz);
  coef[C_DPXZ * n2 + i] = 2 * m.g13[i] / (4. * dz * dx);
  if (with_par && y >= 1 && y <= m.ny - 2) {
    const double dy = m.dy[i], J = m.J[i], g22 = m.g_22[i];
    const double up = 2. * (J + m.J[i + 1]) / ((dy + m.dy[i + 1]) * (g22 + m.g_22[i + 1]));
    const double dn = 2. * (J + m.J[i - 1]) / ((dy + m.dy[i - 1]) * (g22 + m.g_22[i - 1]));
    coef[C_UP * n2 + i] = up / (dy * J);
    coef[C_DN * n2 + i] = dn / (dy * J);
  } else {
    coef[C_UP * n2 + i] = 0.0;
    coef[C_DN * n2 + i] = 0.0;
  }
}

// Delp2 of single_index_ops.hxx:222-246 from the three x-planes, with precombined coefficients
__device__ __forceinline__ double delp2Point(const Row4& fm, const Row4& f0, const Row4& fp, int i,
                                             double cx, double cz, double cxx, double czz, double cxz) {
  const int zm = i, z = i + 1, zp = i + 2;
  return cx * (fp.v[z] - fm.v[z]) + cz * (f0.v[zp] - f0.v[zm])
         + cxx * (fp.v[z] - 2.0 * f0.v[z] + fm.v[z]) + czz * (f0.v[zp] - 2.0 * f0.v[z] + f0.v[zm])
         + cxz * ((fp.v[zp] - fm.v[zp]) - (fp.v[zm] - fm.v[zm]));
}

struct RkEpilogue {
  // For each evolving field e (0..1): state_out = start + a*ddt ; acc_out = acc_in + b*ddt.
  // Any pointer may be NULL to skip that output.  In-place acc (acc_out == acc_in) is allowed.
  const double* start[2];
  const double* acc_in[2];
  double* state_out[2];
  double* acc_out[2];
  double a, b;
  int enabled;
};

template <int VEC>
__device__ __forceinline__ void finishPoint(const RkEpilogue& ep, int e, long long idx, const Vz<VEC>& ddt,
                                            double* __restrict__ ddt_out) {
  if (ddt_out != nullptr) {
    ddt.store(ddt_out + idx);
  }
  if (ep.enabled) {
    const Vz<VEC> s = Vz<VEC>::load(ep.start[e] + idx);
    if (ep.state_out[e] != nullptr) {
      Vz<VEC> r;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        r.v[i] = s.v[i] + ep.a * ddt.v[i];
      }
      r.store(ep.state_out[e] + idx);
    }
    if (ep.acc_out[e] != nullptr) {
      Vz<VEC> base = s;
      if (ep.acc_in[e] != nullptr) {
        base = Vz<VEC>::load(ep.acc_in[e] + idx);
      }
      Vz<VEC> r;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        r.v[i] = base.v[i] + ep.b * ddt.v[i];
      }
      r.store(ep.acc_out[e] + idx);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hasegawa-wakatani-3d (examples/hasegawa-wakatani-3d/hw.cxx:50-77)
//   div_current = alpha * Div_par_Grad_par(phi - n)
//   ddt(n)    = -[phi,n]    - div_current - kappa*DDZ(phi) + Dn   *Delp2(n)
//   ddt(vort) = -[phi,vort] - div_current                  + Dvort*Delp2(vort)
// ------------------------------------------------------------------------------------------------
// Guard cells of the next stage state written by the kernel that produces it (bout_b200_guard_fill): the constant-value
// Dirichlet condition Solver::pre_rhs would apply on the x sides (boundary_standard.cxx:527-581: first guard
// 2 val - f_inside, the others val; rows ystart..yend) and the y guard rows a single-rank periodic
// Mesh::communicate would copy (x guard cells of those rows included).
struct GuardFill {
  double val;
  int xin, xout;
  int ys, ye; // the mesh's ystart / yend (the launch may cover fewer rows)
  // Images of the edge rows.  The bottom rows (y - ys < upper guard rows) go to the upper guard rows of: this field
  // (single rank, periodic y) or the same field in the LOWER neighbour's mapped arena; the top rows to the lower guard
  // rows of this field / of the UPPER neighbour.  off = row shift + arena offset, in doubles; on = 0 disables a side.
  long long off_up, off_dn;
  int on_up, on_dn;
  // Y neighbours on other GPUs (bout_b200_ypeer): flags raised once all rows of a side are delivered, flags awaited by
  // the CTAs of the edge rows before they touch anything
  unsigned long long* peer_flag[2];
  const unsigned long long* wait_flag[2][2];
  unsigned long long wait_value[2][2];
  unsigned long long signal;
  unsigned int* counters;
  unsigned int expect[2]; // CTAs that deliver to the lower / upper neighbour in this launch
  int* status;
  long long timeout;
};
__device__ __forceinline__ bool imageUp(const GuardFill& gf, int ny, int y) { return gf.on_up && y - gf.ys < ny - 1 - gf.ye; }
__device__ __forceinline__ bool imageDn(const GuardFill& gf, int y) { return gf.on_dn && gf.ye - y < gf.ys; }
// the images of the point (x, y, z..z+1) (and the point itself when `self`)
__device__ __forceinline__ void storeWrapY(const GuardFill& gf, double* __restrict__ out, int ny, int nz, int x, int y,
                                           int z, double2 v, bool self) {
  double* q = out + (static_cast<long long>(x) * ny + y) * nz + z;
  if (self) {
    *reinterpret_cast<double2*>(q) = v;
  }
  if (imageUp(gf, ny, y)) {
    *reinterpret_cast<double2*>(q + gf.off_up) = v;
  }
  if (imageDn(gf, y)) {
    *reinterpret_cast<double2*>(q + gf.off_dn) = v;
  }
}
// bounded spin on a flag another GPU raises; a time-out is fatal (cf. k_xchg_y2 in halo.cu)
__device__ __noinline__ void waitFlag(const unsigned long long* flag, unsigned long long value, int* status, long long timeout) {
  const long long t0 = clock64();
  unsigned long long v;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= value) {
      return;
    }
    __nanosleep(64);
  } while (clock64() - t0 < timeout);
  *status = 1;
  __threadfence_system();
  __trap();
}
// one thread per CTA, after the CTA's remote stores are fenced: the last CTA of a side raises the neighbour's flag
__device__ __forceinline__ void signalSide(const GuardFill& gf, int d) {
  const unsigned int prev = atomicAdd(gf.counters + d, 1u);
  if (prev == gf.expect[d] - 1) {
    gf.counters[d] = 0;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(gf.peer_flag[d]), "l"(gf.signal) : "memory");
  }
}

struct Hw3dParams {
  const double* __restrict__ n;
  const double* __restrict__ vort;
  const double* __restrict__ phi;
  const double* __restrict__ coef; // [C_NCOEF][nx*ny]
  double* __restrict__ ddt_n;
  double* __restrict__ ddt_vort;
  RkEpilogue ep;
  double alpha, kappa, Dn, Dvort;
  int nx, ny, nz, xs, xe, ys, ye, chunk;
  int with_par; // ny > 1
  GuardFill gf;
};

template <int VEC>
__global__ void __launch_bounds__(128) k_hw3d_rhs(const Hw3dParams p) {
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
  const int n2 = p.nx * p.ny;

  Row4 Nm, N0, Np, Vm, V0, Vp, Pm, P0, Pp;
  N0 = loadRow4<VEC>(p.n + (x0 - 1) * plane + rowoff, z, p.nz);
  Np = loadRow4<VEC>(p.n + x0 * plane + rowoff, z, p.nz);
  V0 = loadRow4<VEC>(p.vort + (x0 - 1) * plane + rowoff, z, p.nz);
  Vp = loadRow4<VEC>(p.vort + x0 * plane + rowoff, z, p.nz);
  P0 = loadRow4<VEC>(p.phi + (x0 - 1) * plane + rowoff, z, p.nz);
  Pp = loadRow4<VEC>(p.phi + x0 * plane + rowoff, z, p.nz);
  for (int x = x0; x <= x1; ++x) {
    Nm = N0; N0 = Np;
    Vm = V0; V0 = Vp;
    Pm = P0; P0 = Pp;
    const long long nxt = (x + 1) * plane + rowoff;
    Np = loadRow4<VEC>(p.n + nxt, z, p.nz);
    Vp = loadRow4<VEC>(p.vort + nxt, z, p.nz);
    Pp = loadRow4<VEC>(p.phi + nxt, z, p.nz);
    const long long idx = x * plane + rowoff + z;
    // y neighbours of (phi - n) at this x
    Vz<VEC> pup, pdn, nup, ndn;
    if (p.with_par) {
      pup = Vz<VEC>::load(p.phi + idx + p.nz);
      pdn = Vz<VEC>::load(p.phi + idx - p.nz);
      nup = Vz<VEC>::load(p.n + idx + p.nz);
      ndn = Vz<VEC>::load(p.n + idx - p.nz);
    }
    const int i2 = x * p.ny + y;
    const double sf = __ldg(p.coef + C_SF * n2 + i2), rdz2 = __ldg(p.coef + C_RDZ2 * n2 + i2);
    const double cx = __ldg(p.coef + C_DPX * n2 + i2), cz = __ldg(p.coef + C_DPZ * n2 + i2);
    const double cxx = __ldg(p.coef + C_DPXX * n2 + i2), czz = __ldg(p.coef + C_DPZZ * n2 + i2);
    const double cxz = __ldg(p.coef + C_DPXZ * n2 + i2);
    const double cup = __ldg(p.coef + C_UP * n2 + i2), cdn = __ldg(p.coef + C_DN * n2 + i2);
    Vz<VEC> dn_, dv_;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      double div_current = 0.0;
      if (p.with_par) {
        const double c = P0.v[i + 1] - N0.v[i + 1];
        const double up = pup.v[i] - nup.v[i];
        const double dn = pdn.v[i] - ndn.v[i];
        div_current = p.alpha * (cup * (up - c) - cdn * (c - dn));
      }
      const double br_n = arakawaCore(Pm, P0, Pp, Nm, N0, Np, i) * sf;
      const double br_v = arakawaCore(Pm, P0, Pp, Vm, V0, Vp, i) * sf;
      const double ddz_phi = (P0.v[i + 2] - P0.v[i]) * rdz2;
      const double d2n = delp2Point(Nm, N0, Np, i, cx, cz, cxx, czz, cxz);
      const double d2v = delp2Point(Vm, V0, Vp, i, cx, cz, cxx, czz, cxz);
      dn_.v[i] = -br_n - div_current - p.kappa * ddz_phi + p.Dn * d2n;
      dv_.v[i] = -br_v - div_current + p.Dvort * d2v;
    }
    finishPoint<VEC>(p.ep, 0, idx, dn_, p.ddt_n);
    finishPoint<VEC>(p.ep, 1, idx, dv_, p.ddt_vort);
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-pipelined variant of k_hw3d_rhs (the default on sm_100a).
//
// A CTA owns one y row and a chunk of x planes; thread t owns the z-pair (2t, 2t+1).  The three
// stencil fields' rows (n, vort, phi at this y) of plane x+1, x+2, ... are brought into a ring of
// NSLOT shared-memory slots by the TMA unit (cp.async.bulk, one 1-D bulk copy per field row, completion
// on an mbarrier per slot), issued by one elected thread up to NSLOT-3 planes ahead of the plane
// being computed.  Memory-level parallelism therefore comes from the copy engine's queue instead
// of from registers or resident warps: the register-marching kernel above holds 9 x 4 neighbour
// values per thread across iterations (158 registers, 12 warps/SM, 44 % of the HBM roofline).
// Stencil neighbours are read from shared memory; the operands that are used exactly once per
// point (y neighbours of n and phi, the RK epilogue's start / accumulator values) are ordinary
// coalesced 16-byte global loads.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smemAddr(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemAddr(dst)),
               "l"(src), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "B200_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra B200_DONE;\n"
               "bra B200_WAIT;\n"
               "B200_DONE:\n"
               "}" ::"r"(smemAddr(bar)),
               "r"(parity)
               : "memory");
}

__device__ __forceinline__ Row4 ldsRow4(const double* row, int z, int nz) {
  Row4 r;
  const double2 c = *reinterpret_cast<const double2*>(row + z);
  r.v[0] = row[z == 0 ? nz - 1 : z - 1];
  r.v[1] = c.x;
  r.v[2] = c.y;
  r.v[3] = row[z + 2 == nz ? 0 : z + 2];
  return r;
}

// x0..x1: the planes this CTA marched over at row y; thread = z pair.  Dirichlet x guard cells and their y-wrapped copies.
__device__ __forceinline__ void guardPass(const GuardFill& gf, double* f0, double* f1, int nx, int ny, int nz, int xs, int xe,
                                       int x0, int x1, int y, int z) {
  const bool xin = gf.xin && x0 == xs, xout = gf.xout && x1 == xe;
  if (!xin && !xout) {
    return;
  }
  for (int e = 0; e < 2; ++e) {
    double* f = e == 0 ? f0 : f1;
    if (f == nullptr) {
      continue;
    }
    if (xin) {
      const double2 v = *reinterpret_cast<const double2*>(f + (static_cast<long long>(xs) * ny + y) * nz + z);
      storeWrapY(gf, f, ny, nz, xs - 1, y, z, make_double2(2 * gf.val - v.x, 2 * gf.val - v.y), true);
      for (int gx = xs - 2; gx >= 0; --gx) {
        storeWrapY(gf, f, ny, nz, gx, y, z, make_double2(gf.val, gf.val), true);
      }
    }
    if (xout) {
      const double2 v = *reinterpret_cast<const double2*>(f + (static_cast<long long>(xe) * ny + y) * nz + z);
      storeWrapY(gf, f, ny, nz, xe + 1, y, z, make_double2(2 * gf.val - v.x, 2 * gf.val - v.y), true);
      for (int gx = xe + 2; gx < nx; ++gx) {
        storeWrapY(gf, f, ny, nz, gx, y, z, make_double2(gf.val, gf.val), true);
      }
    }
  }
}

constexpr int HW3D_NS = 4; // stencil ring: planes x-1, x, x+1 resident + one in flight
constexpr int HW3D_NA = 2; // once-used operand ring: plane x resident + plane x+1 in flight
constexpr int HW3D_AUX = 8; // rows per aux slot: phi(y+1), phi(y-1), n(y+1), n(y-1), start n, start vort, acc n, acc vort

// GMODE: 0 = interior cells only; 1 = + the Dirichlet x guard cells of the epilogue's output, written after the march by
// the CTAs of the first / last x chunk (rows whose y images nobody needs); 2 = + the y images of the rows (periodic
// wrap or delivery to the Y neighbours with its waits and flags) -- for the few edge rows, launched on their own so
// that the march of all the other rows keeps the register budget of the plain kernel.
template <int THREADS, int GMODE>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 ? 2 : 1)) k_hw3d_rhs_tma(const Hw3dParams p) {
  constexpr bool GUARDS = GMODE == 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nz = p.nz;
  double* tile = reinterpret_cast<double*>(smem_raw);               // [NS][3 fields][nz]
  double* aux = tile + static_cast<size_t>(HW3D_NS) * 3 * nz;        // [NA][AUX rows][nz]
  __shared__ __align__(8) unsigned long long full[HW3D_NS];
  __shared__ __align__(8) unsigned long long afull[HW3D_NA];
  const int tid = threadIdx.x;
  const int z = 2 * tid;
  const int y = p.ys + blockIdx.x;
  const int x0 = p.xs + blockIdx.y * p.chunk;
  const int x1 = min(x0 + p.chunk - 1, p.xe);
  const int nplanes = x1 - x0 + 3; // stencil planes x0-1 .. x1+1
  const int ncentre = x1 - x0 + 1;
  const long long plane = static_cast<long long>(p.ny) * nz;
  const long long rowoff = static_cast<long long>(y) * nz;
  const unsigned rowbytes = static_cast<unsigned>(nz) * 8u;
  const int n2 = p.nx * p.ny;
  const bool ep = p.ep.enabled != 0;
  const bool has_acc0 = ep && p.ep.acc_in[0] != nullptr && p.ep.acc_out[0] != nullptr;
  const bool has_acc1 = ep && p.ep.acc_in[1] != nullptr && p.ep.acc_out[1] != nullptr;
  const unsigned aux_rows = (p.with_par ? 4u : 0u) + (ep ? 2u : 0u) + (has_acc0 ? 1u : 0u) + (has_acc1 ? 1u : 0u);
  // GUARDS: does this CTA's row have an image (a y guard row of this field, or of the field in a neighbour's arena)?
  const bool has_up = GUARDS && imageUp(p.gf, p.ny, y), has_dn = GUARDS && imageDn(p.gf, y);

  auto issue = [&](int j) { // stencil plane j (x = x0 - 1 + j) into slot j % NS
    const int slot = j % HW3D_NS;
    const long long off = static_cast<long long>(x0 - 1 + j) * plane + rowoff;
    unsigned long long* bar = &full[slot];
    mbarExpectTx(bar, 3u * rowbytes);
    double* dst = tile + static_cast<size_t>(slot) * 3 * nz;
    bulkLoad(dst, p.n + off, rowbytes, bar);
    bulkLoad(dst + nz, p.vort + off, rowbytes, bar);
    bulkLoad(dst + 2 * nz, p.phi + off, rowbytes, bar);
  };
  auto issueAux = [&](int ja) { // once-used rows of centre plane x = x0 + ja into slot ja % NA
    if (aux_rows == 0) {
      return;
    }
    const int slot = ja % HW3D_NA;
    const long long off = static_cast<long long>(x0 + ja) * plane + rowoff;
    unsigned long long* bar = &afull[slot];
    mbarExpectTx(bar, aux_rows * rowbytes);
    double* dst = aux + static_cast<size_t>(slot) * HW3D_AUX * nz;
    if (p.with_par) {
      bulkLoad(dst, p.phi + off + nz, rowbytes, bar);
      bulkLoad(dst + nz, p.phi + off - nz, rowbytes, bar);
      bulkLoad(dst + 2 * nz, p.n + off + nz, rowbytes, bar);
      bulkLoad(dst + 3 * nz, p.n + off - nz, rowbytes, bar);
    }
    if (ep) {
      bulkLoad(dst + 4 * nz, p.ep.start[0] + off, rowbytes, bar);
      bulkLoad(dst + 5 * nz, p.ep.start[1] + off, rowbytes, bar);
    }
    if (has_acc0) {
      bulkLoad(dst + 6 * nz, p.ep.acc_in[0] + off, rowbytes, bar);
    }
    if (has_acc1) {
      bulkLoad(dst + 7 * nz, p.ep.acc_in[1] + off, rowbytes, bar);
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < HW3D_NS; ++s) {
      mbarInit(&full[s], 1);
    }
#pragma unroll
    for (int s = 0; s < HW3D_NA; ++s) {
      mbarInit(&afull[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (GUARDS && tid == 0) { // rows next to a neighbour on another GPU: its rows must have arrived, and it must be done
                            // with the guard rows this CTA is about to overwrite (bout_b200_ypeer)
    const int width = max(p.gf.ys, p.ny - 1 - p.gf.ye);
    const bool lo = y - p.gf.ys < width, hi = p.gf.ye - y < width;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (lo && p.gf.wait_flag[0][k] != nullptr) {
        waitFlag(p.gf.wait_flag[0][k], p.gf.wait_value[0][k], p.gf.status, p.gf.timeout);
      }
      if (hi && p.gf.wait_flag[1][k] != nullptr) {
        waitFlag(p.gf.wait_flag[1][k], p.gf.wait_value[1][k], p.gf.status, p.gf.timeout);
      }
    }
    // rows another GPU wrote (generic proxy, acquired above) are read by the bulk-copy unit
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (tid == 0) {
    for (int j = 0; j < min(HW3D_NS, nplanes); ++j) {
      issue(j);
    }
    for (int ja = 0; ja < min(HW3D_NA, ncentre); ++ja) {
      issueAux(ja);
    }
  }
  mbarWait(&full[0], 0);
  mbarWait(&full[1], 0);
  for (int x = x0; x <= x1; ++x) {
    const int j = x - x0 + 1; // centre stencil plane index
    const int ja = x - x0;    // aux plane index
    mbarWait(&full[(j + 1) % HW3D_NS], static_cast<unsigned>(((j + 1) / HW3D_NS) & 1));
    if (aux_rows != 0) {
      mbarWait(&afull[ja % HW3D_NA], static_cast<unsigned>((ja / HW3D_NA) & 1));
    }
    {
      const long long idx = x * plane + rowoff + z;
      const int i2 = x * p.ny + y;
      const double sf = __ldg(p.coef + C_SF * n2 + i2), rdz2 = __ldg(p.coef + C_RDZ2 * n2 + i2);
      const double cx = __ldg(p.coef + C_DPX * n2 + i2), cz = __ldg(p.coef + C_DPZ * n2 + i2);
      const double cxx = __ldg(p.coef + C_DPXX * n2 + i2), czz = __ldg(p.coef + C_DPZZ * n2 + i2);
      const double cxz = __ldg(p.coef + C_DPXZ * n2 + i2);
      const double cup = __ldg(p.coef + C_UP * n2 + i2), cdn = __ldg(p.coef + C_DN * n2 + i2);
      const double* sm = tile + static_cast<size_t>((j - 1) % HW3D_NS) * 3 * nz;
      const double* s0 = tile + static_cast<size_t>(j % HW3D_NS) * 3 * nz;
      const double* sp = tile + static_cast<size_t>((j + 1) % HW3D_NS) * 3 * nz;
      const double* ax = aux + static_cast<size_t>(ja % HW3D_NA) * HW3D_AUX * nz;
      const Row4 Pm = ldsRow4(sm + 2 * nz, z, nz), P0 = ldsRow4(s0 + 2 * nz, z, nz), Pp = ldsRow4(sp + 2 * nz, z, nz);
      double dn_[2], dv_[2];
      {
        const Row4 Nm = ldsRow4(sm, z, nz), N0 = ldsRow4(s0, z, nz), Np = ldsRow4(sp, z, nz);
        double2 pup = {0, 0}, pdn = {0, 0}, nup = {0, 0}, ndn = {0, 0};
        if (p.with_par) {
          pup = *reinterpret_cast<const double2*>(ax + z);
          pdn = *reinterpret_cast<const double2*>(ax + nz + z);
          nup = *reinterpret_cast<const double2*>(ax + 2 * nz + z);
          ndn = *reinterpret_cast<const double2*>(ax + 3 * nz + z);
        }
        const double upv[2] = {pup.x - nup.x, pup.y - nup.y};
        const double dnv[2] = {pdn.x - ndn.x, pdn.y - ndn.y};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          double div_current = 0.0;
          if (p.with_par) {
            const double c = P0.v[i + 1] - N0.v[i + 1];
            div_current = p.alpha * (cup * (upv[i] - c) - cdn * (c - dnv[i]));
          }
          const double br_n = arakawaCore(Pm, P0, Pp, Nm, N0, Np, i) * sf;
          const double ddz_phi = (P0.v[i + 2] - P0.v[i]) * rdz2;
          const double d2n = delp2Point(Nm, N0, Np, i, cx, cz, cxx, czz, cxz);
          dn_[i] = -br_n - div_current - p.kappa * ddz_phi + p.Dn * d2n;
          dv_[i] = -div_current;
        }
      }
      {
        const Row4 Vm = ldsRow4(sm + nz, z, nz), V0 = ldsRow4(s0 + nz, z, nz), Vp = ldsRow4(sp + nz, z, nz);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double br_v = arakawaCore(Pm, P0, Pp, Vm, V0, Vp, i) * sf;
          const double d2v = delp2Point(Vm, V0, Vp, i, cx, cz, cxx, czz, cxz);
          dv_[i] = -br_v + dv_[i] + p.Dvort * d2v;
        }
      }
      // outputs: ddt and/or the RK stage update (reads start / acc from the aux slot)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double* d = e == 0 ? dn_ : dv_;
        double* ddt_out = e == 0 ? p.ddt_n : p.ddt_vort;
        if (ddt_out != nullptr) {
          *reinterpret_cast<double2*>(ddt_out + idx) = make_double2(d[0], d[1]);
        }
        if (ep) {
          const double2 st = *reinterpret_cast<const double2*>(ax + (4 + e) * nz + z);
          // GUARDS: the field the next right-hand side reads (the stage state; after the last stage the combination)
          // also goes to the y guard rows this row is the periodic image of
          if (p.ep.state_out[e] != nullptr) {
            const double2 r = make_double2(st.x + p.ep.a * d[0], st.y + p.ep.a * d[1]);
            *reinterpret_cast<double2*>(p.ep.state_out[e] + idx) = r;
            if (GUARDS && has_up) {
              *reinterpret_cast<double2*>(p.ep.state_out[e] + idx + p.gf.off_up) = r;
            }
            if (GUARDS && has_dn) {
              *reinterpret_cast<double2*>(p.ep.state_out[e] + idx + p.gf.off_dn) = r;
            }
          }
          if (p.ep.acc_out[e] != nullptr) {
            double2 base = st;
            if (p.ep.acc_in[e] != nullptr) {
              base = *reinterpret_cast<const double2*>(ax + (6 + e) * nz + z);
            }
            const double2 r = make_double2(base.x + p.ep.b * d[0], base.y + p.ep.b * d[1]);
            *reinterpret_cast<double2*>(p.ep.acc_out[e] + idx) = r;
            if (GUARDS && p.ep.state_out[e] == nullptr) {
              if (has_up) {
                *reinterpret_cast<double2*>(p.ep.acc_out[e] + idx + p.gf.off_up) = r;
              }
              if (has_dn) {
                *reinterpret_cast<double2*>(p.ep.acc_out[e] + idx + p.gf.off_dn) = r;
              }
            }
          }
        }
      }
    }
    __syncthreads(); // every thread is done with stencil plane j-1 and aux plane ja: refill their slots
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (j - 1 + HW3D_NS < nplanes) {
        issue(j - 1 + HW3D_NS);
      }
      if (ja + HW3D_NA < ncentre) {
        issueAux(ja + HW3D_NA);
      }
    }
  }
  // x guard cells of the field the next right-hand side reads, from the edge-plane values this thread has just written
  if (GMODE >= 1 && ep) {
    guardPass(p.gf, p.ep.state_out[0] != nullptr ? p.ep.state_out[0] : p.ep.acc_out[0],
              p.ep.state_out[1] != nullptr ? p.ep.state_out[1] : p.ep.acc_out[1], p.nx, p.ny, nz, p.xs, p.xe, x0, x1, y, z);
  }
  if (GUARDS && ep) {
    const bool to_lower = has_up && p.gf.peer_flag[0] != nullptr, to_upper = has_dn && p.gf.peer_flag[1] != nullptr;
    if (to_lower || to_upper) { // every thread's remote stores are visible system-wide before the flag can rise
      __threadfence_system();
      __syncthreads();
      if (tid == 0) {
        if (to_lower) {
          signalSide(p.gf, 0);
        }
        if (to_upper) {
          signalSide(p.gf, 1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hasegawa-wakatani convective part (examples/hasegawa-wakatani/hw.cxx:94-118), Arakawa bracket:
//   ddt(n)    = -[phi,n]    + alpha*(nzphi - nzn) - kappa*DDZ(phi)
//   ddt(vort) = -[phi,vort] + alpha*(nzphi - nzn)
// with nzf = f - zonal[x,y] when `modified` (zonal = averageY(DC(f)), a Field2D) else f.
// ------------------------------------------------------------------------------------------------
struct HwConvParams {
  const double* __restrict__ n;
  const double* __restrict__ vort;
  const double* __restrict__ phi;
  const double* __restrict__ zonal_n;   // Field2D or NULL
  const double* __restrict__ zonal_phi; // Field2D or NULL
  const double* __restrict__ dx;
  const double* __restrict__ dz;
  double* __restrict__ ddt_n;
  double* __restrict__ ddt_vort;
  RkEpilogue ep;
  double alpha, kappa;
  int nx, ny, nz, xs, xe, ys, ye, chunk;
};

template <int VEC>
__global__ void __launch_bounds__(128) k_hw_convective(const HwConvParams p) {
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
  Row4 Nm, N0, Np, Vm, V0, Vp, Pm, P0, Pp;
  N0 = loadRow4<VEC>(p.n + (x0 - 1) * plane + rowoff, z, p.nz);
  Np = loadRow4<VEC>(p.n + x0 * plane + rowoff, z, p.nz);
  V0 = loadRow4<VEC>(p.vort + (x0 - 1) * plane + rowoff, z, p.nz);
  Vp = loadRow4<VEC>(p.vort + x0 * plane + rowoff, z, p.nz);
  P0 = loadRow4<VEC>(p.phi + (x0 - 1) * plane + rowoff, z, p.nz);
  Pp = loadRow4<VEC>(p.phi + x0 * plane + rowoff, z, p.nz);
  for (int x = x0; x <= x1; ++x) {
    Nm = N0; N0 = Np;
    Vm = V0; V0 = Vp;
    Pm = P0; P0 = Pp;
    const long long nxt = (x + 1) * plane + rowoff;
    Np = loadRow4<VEC>(p.n + nxt, z, p.nz);
    Vp = loadRow4<VEC>(p.vort + nxt, z, p.nz);
    Pp = loadRow4<VEC>(p.phi + nxt, z, p.nz);
    const long long idx = x * plane + rowoff + z;
    const long long i2 = static_cast<long long>(x) * p.ny + y;
    const double dxv = __ldg(p.dx + i2), dzv = __ldg(p.dz + i2);
    const double sf = 1.0 / (12 * dzv * dxv); // difops.cxx:950
    const double rdz = 1.0 / dzv;             // DDZ = index * (1/dz), derivs.cxx:82-86
    const double zn = p.zonal_n ? __ldg(p.zonal_n + i2) : 0.0;
    const double zp = p.zonal_phi ? __ldg(p.zonal_phi + i2) : 0.0;
    Vz<VEC> dn_, dv_;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const double br_n = arakawaCore(Pm, P0, Pp, Nm, N0, Np, i) * sf;
      const double br_v = arakawaCore(Pm, P0, Pp, Vm, V0, Vp, i) * sf;
      const double coupling = p.alpha * ((P0.v[i + 1] - zp) - (N0.v[i + 1] - zn));
      const double ddz_phi = (0.5 * (P0.v[i + 2] - P0.v[i])) * rdz;
      dn_.v[i] = (-1.0 * br_n + coupling) - p.kappa * ddz_phi;
      dv_.v[i] = -1.0 * br_v + coupling;
    }
    finishPoint<VEC>(p.ep, 0, idx, dn_, p.ddt_n);
    finishPoint<VEC>(p.ep, 1, idx, dv_, p.ddt_vort);
  }
}

inline int pickChunk(long long north, int len) {
  int chunk = 8;
  for (int c : {64, 32, 16}) {
    if (north * ((len + c - 1) / c) >= 148LL * 1024) {
      chunk = c;
      break;
    }
  }
  return chunk;
}

inline RkEpilogue makeEpilogue(const bout_b200_rk_epilogue* e) {
  RkEpilogue ep{};
  if (e != nullptr) {
    for (int i = 0; i < 2; ++i) {
      ep.start[i] = e->start[i];
      ep.acc_in[i] = e->acc_in[i];
      ep.state_out[i] = e->state_out[i];
      ep.acc_out[i] = e->acc_out[i];
    }
    ep.a = e->a;
    ep.b = e->b;
    ep.enabled = 1;
  }
  return ep;
}

} // namespace b200

using namespace b200;

extern "C" size_t bout_b200_fused_coef_doubles(const bout_b200_grid* g) {
  return g ? static_cast<size_t>(C_NCOEF) * g->nx * g->ny : 0;
}

static int launchFusedCoefs(const bout_b200_grid* g, const bout_b200_metric* m, double* coef_workspace,
                            cudaStream_t s) {
  B200_REQUIRE(m->dx && m->dy && m->dz && m->J && m->g_22 && m->g11 && m->g33 && m->g13 && m->G1 && m->G3,
               BOUT_B200_ERR_ARG, "fused coefficients: metric missing");
  const int with_par = g->ny > 1 ? 1 : 0;
  CoefIn ci{m->dx, m->dy, m->dz, m->J, m->g_22, m->g11, m->g33, m->g13, m->G1, m->G3, m->d1_dx, g->nx, g->ny};
  const int n2 = g->nx * g->ny;
  k_fused_coefs<<<(n2 + 255) / 256, 256, 0, s>>>(ci, coef_workspace, with_par);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_fused_coefs(const bout_b200_grid* g, const bout_b200_metric* m,
                                     double* coef_workspace, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(m && coef_workspace, BOUT_B200_ERR_ARG, "fused_coefs: null argument");
  return launchFusedCoefs(g, m, coef_workspace, asStream(stream));
}

extern "C" int bout_b200_hw3d_rhs(const bout_b200_grid* g, const bout_b200_metric* m, double alpha,
                                  double kappa, double Dn, double Dvort, const double* n,
                                  const double* vort, const double* phi, double* ddt_n,
                                  double* ddt_vort, const bout_b200_rk_epilogue* epilogue,
                                  double* coef_workspace, void* stream) {
  return bout_b200_hw3d_rhs_guards(g, m, alpha, kappa, Dn, Dvort, n, vort, phi, ddt_n, ddt_vort, epilogue, nullptr,
                                   coef_workspace, stream);
}

extern "C" int bout_b200_hw3d_rhs_guards(const bout_b200_grid* g, const bout_b200_metric* m, double alpha,
                                         double kappa, double Dn, double Dvort, const double* n,
                                         const double* vort, const double* phi, double* ddt_n,
                                         double* ddt_vort, const bout_b200_rk_epilogue* epilogue,
                                         const bout_b200_guard_fill* guards, double* coef_workspace, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  const bout_b200_ypeer* yp = guards != nullptr ? guards->ypeer : nullptr;
  const bool fill = guards != nullptr && (guards->x_inner || guards->x_outer || guards->wrap_y || yp != nullptr);
  if (yp != nullptr) {
    B200_REQUIRE(!guards->wrap_y, BOUT_B200_ERR_ARG, "hw3d_rhs: wrap_y and ypeer exclude each other");
    B200_REQUIRE(yp->counters != nullptr && yp->status != nullptr, BOUT_B200_ERR_ARG, "hw3d_rhs: ypeer without counters / status");
  }
  if (fill) {
    const bool writes = guards->x_inner || guards->x_outer || guards->wrap_y
                        || (yp != nullptr && (yp->peer_flag[0] != nullptr || yp->peer_flag[1] != nullptr));
    B200_REQUIRE(!writes || epilogue != nullptr, BOUT_B200_ERR_ARG, "hw3d_rhs: guard cells are filled on the epilogue's output");
    B200_REQUIRE(guards->ystart >= 0 && guards->ystart <= g->ystart && guards->yend >= g->yend && guards->yend < g->ny,
                 BOUT_B200_ERR_ARG, "hw3d_rhs: the rows of this launch lie outside the guard fill's ystart..yend");
    B200_REQUIRE(!guards->wrap_y || (guards->yend - guards->ystart + 1 >= guards->ystart
                                     && guards->yend - guards->ystart + 1 >= g->ny - 1 - guards->yend),
                 BOUT_B200_ERR_GUARDS, "hw3d_rhs: fewer interior rows than y guard rows");
  }
  B200_REQUIRE(n && vort && phi && coef_workspace, BOUT_B200_ERR_ARG, "hw3d_rhs: null argument");
  B200_REQUIRE((ddt_n && ddt_vort) || epilogue, BOUT_B200_ERR_ARG, "hw3d_rhs: no output requested");
  B200_REQUIRE(g->xstart >= 1 && g->xend <= g->nx - 2, BOUT_B200_ERR_GUARDS, "hw3d_rhs: needs an x guard cell");
  const int with_par = g->ny > 1 ? 1 : 0;
  B200_REQUIRE(!with_par || (g->ystart >= 1 && g->yend <= g->ny - 2), BOUT_B200_ERR_GUARDS,
               "hw3d_rhs: needs a y guard cell");
  cudaStream_t s = asStream(stream);
  if (m != nullptr) { // metric given: (re)build the per-(x,y) coefficients; NULL = already built
    const int rc = launchFusedCoefs(g, m, coef_workspace, s);
    if (rc != 0) {
      return rc;
    }
  }
  Hw3dParams p{};
  p.n = n;
  p.vort = vort;
  p.phi = phi;
  p.coef = coef_workspace;
  p.ddt_n = ddt_n;
  p.ddt_vort = ddt_vort;
  p.ep = makeEpilogue(epilogue);
  p.alpha = alpha;
  p.kappa = kappa;
  p.Dn = Dn;
  p.Dvort = Dvort;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = g->xstart;
  p.xe = g->xend;
  p.ys = g->ystart;
  p.ye = g->yend;
  p.with_par = with_par;
  bool vec2 = (g->nz % 2 == 0) && aligned16(n) && aligned16(vort) && aligned16(phi)
              && (!ddt_n || aligned16(ddt_n)) && (!ddt_vort || aligned16(ddt_vort));
  if (epilogue) {
    for (int i = 0; i < 2; ++i) {
      vec2 = vec2 && aligned16(epilogue->start[i]) && aligned16(epilogue->acc_in[i])
             && aligned16(epilogue->state_out[i]) && aligned16(epilogue->acc_out[i]);
    }
  }
  const int nzv = vec2 ? g->nz / 2 : g->nz;
  const long long north = static_cast<long long>(p.ye - p.ys + 1) * nzv;
  const int len = p.xe - p.xs + 1;
  static const bool no_tma = getenv("BOUT_B200_NO_TMA") != nullptr;
  const int half = g->nz / 2;
  const bool tma = vec2 && !no_tma && (half == 64 || half == 128 || half == 256 || half == 512);
  if (fill && tma) {
    GuardFill& gf = p.gf;
    gf.val = guards->value;
    gf.xin = guards->x_inner;
    gf.xout = guards->x_outer;
    gf.ys = guards->ystart;
    gf.ye = guards->yend;
    const long long image = static_cast<long long>(gf.ye - gf.ys + 1) * g->nz; // interior rows of the mesh, in doubles
    if (guards->wrap_y) {
      gf.on_up = gf.on_dn = 1;
      gf.off_up = image;
      gf.off_dn = -image;
    }
    if (yp != nullptr) {
      if (yp->peer_flag[0] != nullptr) {
        gf.on_up = 1;
        gf.off_up = image + yp->delta[0];
      }
      if (yp->peer_flag[1] != nullptr) {
        gf.on_dn = 1;
        gf.off_dn = -image + yp->delta[1];
      }
      for (int d = 0; d < 2; ++d) {
        gf.peer_flag[d] = yp->peer_flag[d];
        for (int k = 0; k < 2; ++k) {
          gf.wait_flag[d][k] = yp->wait_flag[d][k];
          gf.wait_value[d][k] = yp->wait_value[d][k];
        }
      }
      gf.signal = yp->signal;
      gf.counters = yp->counters;
      gf.status = yp->status;
      gf.timeout = haloTimeoutCycles();
    }
  }
  B200_REQUIRE(yp == nullptr || tma, BOUT_B200_ERR_UNSUPPORTED,
               "hw3d_rhs: rows are delivered to Y neighbours by the bulk-copy kernel only (nz = 128 .. 1024, aligned fields)");
  if (fill && !tma) { // the register-marching kernel writes interior cells only: the guard cells by their own kernels
    const int rc = bout_b200_hw3d_rhs_guards(g, nullptr, alpha, kappa, Dn, Dvort, n, vort, phi, ddt_n, ddt_vort, epilogue,
                                             nullptr, coef_workspace, stream);
    if (rc != 0) {
      return rc;
    }
    bout_b200_grid gm = *g;
    gm.ystart = guards->ystart;
    gm.yend = guards->yend;
    double* outs[2];
    int nout = 0;
    for (int e = 0; e < 2; ++e) {
      double* f = epilogue->state_out[e] != nullptr ? epilogue->state_out[e] : epilogue->acc_out[e];
      if (f == nullptr) {
        continue;
      }
      outs[nout++] = f;
      for (int side = 0; side < 2; ++side) {
        if (side == 0 ? guards->x_inner : guards->x_outer) {
          const int rb = bout_b200_boundary(&gm, 0, side, guards->value, nullptr, g->ystart, g->yend, f, stream);
          if (rb != 0) {
            return rb;
          }
        }
      }
    }
    // (a launch over a subset of the rows re-copies rows other launches own: harmless, the values are final
    // only once every row's launch is done, which is also when the caller may read them)
    return (guards->wrap_y && nout > 0) ? bout_b200_halo_self(&gm, nout, outs, 0, 1, stream) : 0;
  }
  if (tma) {
    // TMA-pipelined kernel: one CTA per (y, x-chunk).  All CTAs run for about the same time (chunk planes + the
    // pipeline fill), so the chunk length is chosen for whole waves of resident CTAs: 512 planes x 128 rows in
    // chunks of 64 are 1024 CTAs = 3.46 waves of 2 x 148 (measured 0.507 ms), in chunks of 57 they are 1152 = 3.9
    // waves (0.474 ms).
    const int nyint = p.ye - p.ys + 1;
    const long long slots = 148LL * (half <= 256 ? 2 : 1);
    int chunk = 64;
    long long best = -1;
    for (int c = 8; c <= 128; ++c) {
      const long long ncta = static_cast<long long>(nyint) * ((len + c - 1) / c);
      const long long cost = ((ncta + slots - 1) / slots) * (c + 4);
      if (best < 0 || cost <= best) {
        best = cost;
        chunk = c;
      }
    }
    static const int chunk_env = getenv("BOUT_B200_HW3D_CHUNK") ? atoi(getenv("BOUT_B200_HW3D_CHUNK")) : 0;
    if (chunk_env > 0) {
      chunk = chunk_env;
    }
    p.chunk = chunk;
    const size_t smem = static_cast<size_t>(HW3D_NS * 3 + HW3D_NA * HW3D_AUX) * g->nz * sizeof(double);
    dim3 grid(static_cast<unsigned>(nyint), static_cast<unsigned>((len + chunk - 1) / chunk));
    // which variant: do any of this launch's rows have a y image or wait for a neighbour?
    int gmode = 0;
    if (fill) {
      const int width = std::max(p.gf.ys, g->ny - 1 - p.gf.ye);
      const bool edge_rows = (p.gf.on_up || p.gf.on_dn || yp != nullptr)
                             && (p.ys - p.gf.ys < width || p.gf.ye - p.ye < width);
      gmode = edge_rows ? 2 : 1;
      if (yp != nullptr) { // CTAs of this launch that deliver rows to the lower / upper neighbour
        int rows_lo = 0, rows_hi = 0;
        for (int y = p.ys; y <= p.ye; ++y) {
          rows_lo += (y - p.gf.ys < g->ny - 1 - p.gf.ye) ? 1 : 0;
          rows_hi += (p.gf.ye - y < p.gf.ys) ? 1 : 0;
        }
        p.gf.expect[0] = static_cast<unsigned>(rows_lo) * grid.y;
        p.gf.expect[1] = static_cast<unsigned>(rows_hi) * grid.y;
      }
    }
#define B200_LAUNCH_TMA_M(T, M)                                                                          \
  {                                                                                                      \
    B200_CUDA_OK(cudaFuncSetAttribute(k_hw3d_rhs_tma<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      static_cast<int>(smem)));                                          \
    k_hw3d_rhs_tma<T, M><<<grid, T, smem, s>>>(p);                                                       \
  }
#define B200_LAUNCH_TMA(T)                                                                               \
  {                                                                                                      \
    if (gmode == 2) {                                                                                    \
      B200_LAUNCH_TMA_M(T, 2)                                                                            \
    } else if (gmode == 1) {                                                                             \
      B200_LAUNCH_TMA_M(T, 1)                                                                            \
    } else {                                                                                             \
      B200_LAUNCH_TMA_M(T, 0)                                                                            \
    }                                                                                                    \
  }
    switch (half) {
    case 64: B200_LAUNCH_TMA(64) break;
    case 128: B200_LAUNCH_TMA(128) break;
    case 256: B200_LAUNCH_TMA(256) break;
    default: B200_LAUNCH_TMA(512) break;
    }
#undef B200_LAUNCH_TMA
#undef B200_LAUNCH_TMA_M
    B200_LAUNCH_OK();
    return 0;
  }
  p.chunk = pickChunk(north, len);
  dim3 grid(static_cast<unsigned>((north + 127) / 128), static_cast<unsigned>((len + p.chunk - 1) / p.chunk));
  if (vec2) {
    k_hw3d_rhs<2><<<grid, 128, 0, s>>>(p);
  } else {
    k_hw3d_rhs<1><<<grid, 128, 0, s>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_hw_convective(const bout_b200_grid* g, const double* dx, const double* dz,
                                       double alpha, double kappa, const double* n,
                                       const double* vort, const double* phi, const double* zonal_n,
                                       const double* zonal_phi, double* ddt_n, double* ddt_vort,
                                       const bout_b200_rk_epilogue* epilogue, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(dx && dz && n && vort && phi, BOUT_B200_ERR_ARG, "hw_convective: null argument");
  B200_REQUIRE((ddt_n && ddt_vort) || epilogue, BOUT_B200_ERR_ARG, "hw_convective: no output requested");
  B200_REQUIRE(g->xstart >= 1 && g->xend <= g->nx - 2, BOUT_B200_ERR_GUARDS,
               "hw_convective: needs an x guard cell");
  HwConvParams p{};
  p.n = n;
  p.vort = vort;
  p.phi = phi;
  p.zonal_n = zonal_n;
  p.zonal_phi = zonal_phi;
  p.dx = dx;
  p.dz = dz;
  p.ddt_n = ddt_n;
  p.ddt_vort = ddt_vort;
  p.ep = makeEpilogue(epilogue);
  p.alpha = alpha;
  p.kappa = kappa;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = g->xstart;
  p.xe = g->xend;
  p.ys = g->ystart;
  p.ye = g->yend;
  bool vec2 = (g->nz % 2 == 0) && aligned16(n) && aligned16(vort) && aligned16(phi)
              && (!ddt_n || aligned16(ddt_n)) && (!ddt_vort || aligned16(ddt_vort));
  if (epilogue) {
    for (int i = 0; i < 2; ++i) {
      vec2 = vec2 && aligned16(epilogue->start[i]) && aligned16(epilogue->acc_in[i])
             && aligned16(epilogue->state_out[i]) && aligned16(epilogue->acc_out[i]);
    }
  }
  const int nzv = vec2 ? g->nz / 2 : g->nz;
  const long long north = static_cast<long long>(p.ye - p.ys + 1) * nzv;
  const int len = p.xe - p.xs + 1;
  p.chunk = pickChunk(north, len);
  dim3 grid(static_cast<unsigned>((north + 127) / 128), static_cast<unsigned>((len + p.chunk - 1) / p.chunk));
  if (vec2) {
    k_hw_convective<2><<<grid, 128, 0, asStream(stream)>>>(p);
  } else {
    k_hw_convective<1><<<grid, 128, 0, asStream(stream)>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}
