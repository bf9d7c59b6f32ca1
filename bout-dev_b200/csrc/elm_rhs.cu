// Fused right-hand side of the elm-pb reduced-MHD model (examples/elm-pb/elm_pb.cxx:1674-1955) for the
// shipped switch set: linear, evolve Psi, diamag + diamag_phi0, constn0, bm_exb_flag = bm_mag_flag = 0
// (BRACKET_STD = b0xGrad_dot_Grad, difops.cxx:563-601), hyper-resistivity, filter_z (+ low_pass_z):
//
//   ddt(Psi) = -Grad_par(phi B0)/B0 + eta Jpar - b0xGrad_dot_Grad(phi0, Psi) B0/Bxy - eta hyperresist Delp2(Jpar)
//   ddt(U)   = B0^2 b0xGrad_dot_Grad(Psi, J0) + b0xcv . Grad(P) - B0^2 Grad_par(Jpar) - b0xGrad_dot_Grad(phi0, U)
//   ddt(P)   = -b0xGrad_dot_Grad(phi, P0) - b0xGrad_dot_Grad(phi0, P)
//   ddt(.)   = lowPass(filter(ddt(.), n0), zmax, keep_zonal)
//
// ONE kernel instead of the reference's ~60 full-field passes (5 brackets of ~12 passes each, 5 first
// derivatives, the axpys, 6 FFT filter passes): every input field is read once from HBM (x neighbours
// of the four stencil fields come from L2), the three time derivatives are written once.
// Methods are fixed to the shipped ones: first C4 / C2 / C4, upwind W3 / C2 / W3 in x / y / z.
// Y neighbours come from the parallel slices (yup / ydown) of ShiftedMetric::calcParallelSlices.
//
// The z filter runs inside the kernel: a CTA owns whole z rows, so the one kept Fourier mode n0 is a
// block-wide dot product with the twiddle table and the filtered row is written directly (the
// reference's rfft -> zero all but n0 -> irfft, field3d.cxx:620-661, to round-off).
//
// Three terms are taken through field-aligned coordinates by the reference because their operand
// carries no parallel slices (index_derivs_interface.hxx:205-211): Grad_par(phi*B0) and the y-advection
// of the Field2D operands J0 and P0.  toFieldAligned/fromFieldAligned multiply z mode k by
// exp(+-i k zShift), so on every mode except the Nyquist one the round trip equals the slice form used
// here (yup/ydown hold exactly f(y+-1) shifted by zShift(y) - zShift(y+-1)); the Nyquist mode, whose
// imaginary part c2r drops, is discarded by the filter (n0 != nz/2 is required).  Without filter_z the
// harness uses the operator-by-operator path, which reproduces the Nyquist handling exactly.
#include <cstdlib>

#include "fft.cuh"
#include "laplace_coefs.cuh"
#include "stencil_ops.cuh"

namespace b200 {

const cd* twiddleTable(int nz, cudaStream_t stream, int* status); // zfft_ops.cu

struct ElmParams {
  bout_b200_elmpb_rhs_args a;
  const cd* __restrict__ tw;
  int nx, ny, nz, xs, xe, ys, ye;
  int rows_per_cta;
};

__device__ __forceinline__ double2 ld2(const double* __restrict__ p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

// x-stencil (5 rows) + z-stencil (centre row) + the two slices of a Field3D at this thread's V consecutive z points
// (V = 2: 16-byte transactions; V = 1: half the registers, twice the warps per SM).
template <int V>
struct F3 {
  double xmm[V], xm[V], xp[V], xpp[V], up[V], dn[V];
  double c[4 + V]; // z0-2 .. z0+V+1
  __device__ __forceinline__ St sx(int j) const { return St{xmm[j], xm[j], c[2 + j], xp[j], xpp[j]}; }
  __device__ __forceinline__ St sz(int j) const { return St{c[j], c[j + 1], c[j + 2], c[j + 3], c[j + 4]}; }
  __device__ __forceinline__ double yp(int j) const { return up[j]; }
  __device__ __forceinline__ double ym(int j) const { return dn[j]; }
};
template <int V>
__device__ __forceinline__ void ldv(const double* __restrict__ p, double (&out)[V]) {
  if constexpr (V == 2) {
    const double2 t = ld2(p);
    out[0] = t.x;
    out[1] = t.y;
  } else {
    out[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ F3<V> loadF3(const double* __restrict__ f, const double* __restrict__ fup,
                                        const double* __restrict__ fdn, long long i2, int ny, int nz, int z0) {
  F3<V> r;
  const long long plane = static_cast<long long>(ny) * nz;
  const double* __restrict__ row = f + i2 * nz;
  if constexpr (V == 2) {
    const double2 a = ld2(row + wrapz(z0 - 2, nz)), b = ld2(row + z0), c = ld2(row + wrapz(z0 + 2, nz));
    r.c[0] = a.x, r.c[1] = a.y, r.c[2] = b.x, r.c[3] = b.y, r.c[4] = c.x, r.c[5] = c.y;
  } else {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      r.c[q] = __ldg(row + wrapz(z0 - 2 + q, nz));
    }
  }
  ldv<V>(row - 2 * plane + z0, r.xmm);
  ldv<V>(row - plane + z0, r.xm);
  ldv<V>(row + plane + z0, r.xp);
  ldv<V>(row + 2 * plane + z0, r.xpp);
  if (fup != nullptr) { // (no slices: their y-derivative terms are added in mode space, see RowC::x*)
    ldv<V>(fup + (i2 + 1) * nz + z0, r.up); // yup holds the y+1 neighbour at row y+1
    ldv<V>(fdn + (i2 - 1) * nz + z0, r.dn);
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      r.up[j] = 0.0;
      r.dn[j] = 0.0;
    }
  }
  return r;
}

// Per-(x,y) row constants, computed once per row by a few threads and kept in shared memory (every thread of
// the row reads the same values: broadcast loads) so that they do not occupy registers.
struct RowC {
  double rdx, rdy, rdz, g_22, g_23, g_12, rjs;
  double B0sq, rB0, B0rBxy, B0p, B0m, gpar;
  double vx0, vy0, vz0;                 // velocity of the Field2D potential phi0 (difops.cxx:580-588, DDZ(phi0) = 0)
  double dJ0[2], dP0[2];                // WENO3 x-derivative factor of J0 / P0 for vx <= 0 ([0]) and vx > 0 ([1])
  double dJ0y, dP0y;                    // J0(y+1) - J0(y-1), P0(y+1) - P0(y-1)
  double bcx, bcy, bcz;
  double xU[2], xP[2]; // mode-space additions to the row sums of ddt(U), ddt(P) (re, im), zero on the slice path
  double xPsi[2];      // likewise ddt(Psi): -hyperresist eta Delp2(Jpar) evaluated on the row modes of Jpar
};

// b0xGrad_dot_Grad(f3d, A2d): dAx = the two WENO3 factors of A's x stencil, dAy = A(y+1) - A(y-1)
__device__ __forceinline__ double b0x3d2d(const St& fx, double fyp, double fym, const St& fz, const double (&dAx)[2],
                                          double dAy, const RowC& m) {
  const double dpdx = applyStandard<BOUT_B200_DDX_C4>(fx) * m.rdx;
  const double dpdy = 0.5 * (fyp - fym) * m.rdy;
  const double dpdz = applyStandard<BOUT_B200_DDX_C4>(fz) * m.rdz;
  const double vx = m.g_22 * dpdz - m.g_23 * dpdy;
  const double vy = m.g_23 * dpdx - m.g_12 * dpdz;
  double r = (vx * dAx[vx > 0.0 ? 1 : 0]) * m.rdx;
  r += vy * 0.5 * dAy * m.rdy;
  return r * m.rjs;
}
// b0xGrad_dot_Grad(f2d, A3d): the velocity (vx0, vy0, vz0) is a Field2D quantity
__device__ __forceinline__ double b0x2d3d(const St& ax, double ayp, double aym, const St& az, const RowC& m) {
  double r = weno3(m.vx0, ax) * m.rdx;
  r += m.vy0 * 0.5 * (ayp - aym) * m.rdy;
  r += weno3(m.vz0, az) * m.rdz;
  return r * m.rjs;
}

__device__ __forceinline__ St st2dx(const double* __restrict__ f, long long i2, int ny) {
  return St{__ldg(f + i2 - 2 * ny), __ldg(f + i2 - ny), __ldg(f + i2), __ldg(f + i2 + ny), __ldg(f + i2 + 2 * ny)};
}

// Sum of six doubles per thread over the `nthr` threads of one row group (all threads of the CTA call).
__device__ __forceinline__ void rowReduce3(double (&acc)[6], double* red, int tin, int nthr, int rloc) {
  if ((nthr & 31) == 0) { // whole warps: shuffle tree, then one partial per warp
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      }
    }
    const int nw = nthr >> 5, w = tin >> 5;
    double* r = red + rloc * 6 * 32;
    if ((tin & 31) == 0) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        r[q * 32 + w] = acc[q];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double s = 0.0;
      for (int i = 0; i < nw; ++i) {
        s += r[q * 32 + i];
      }
      acc[q] = s;
    }
  } else { // short or odd rows (one row per CTA): every thread sums the row from shared memory
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      red[q * nthr + tin] = acc[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double s = 0.0;
      for (int i = 0; i < nthr; ++i) {
        s += red[q * nthr + i];
      }
      acc[q] = s;
    }
  }
}

// All per-row constants of row (x, y) (a few dozen dependent loads, two sincos, one tridagCoefs): evaluated by
// k_elm_rowc for every row ahead of the main kernel when the caller provides rowc_ws, else by one thread of the row.
__device__ __forceinline__ RowC elmRowConstants(const ElmParams& p, int x, int y) {
  const bout_b200_elmpb_rhs_args& a = p.a;
  const int ny = p.ny;
  const long long i2 = static_cast<long long>(x) * p.ny + y;
  RowC c;
  c.rdx = 1.0 / __ldg(a.dx + i2);
  c.rdy = 1.0 / __ldg(a.dy + i2);
  c.rdz = 1.0 / __ldg(a.dz + i2);
  c.g_22 = __ldg(a.g_22 + i2);
  c.g_23 = __ldg(a.g_23 + i2);
  c.g_12 = __ldg(a.g_12 + i2);
  const double sg = sqrt(c.g_22);
  c.rjs = 1.0 / (__ldg(a.J + i2) * sg);
  const double B0 = __ldg(a.B0 + i2);
  c.B0sq = B0 * B0;
  c.rB0 = 1.0 / B0;                          // Field3D / Field2D multiplies by the reciprocal, as the reference does
  c.B0rBxy = B0 * (1.0 / __ldg(a.Bxy + i2));
  c.B0p = __ldg(a.B0 + i2 + 1);
  c.B0m = __ldg(a.B0 + i2 - 1);
  c.gpar = c.rdy / sg;                       // Grad_par = DDY / dy * invSg (coordinates.cxx:1529-1535)
  const double d0x = applyStandard<BOUT_B200_DDX_C4>(st2dx(a.phi0, i2, ny)) * c.rdx;
  const double d0y = 0.5 * (__ldg(a.phi0 + i2 + 1) - __ldg(a.phi0 + i2 - 1)) * c.rdy;
  c.vx0 = c.g_22 * 0.0 - c.g_23 * d0y;
  c.vy0 = c.g_23 * d0x - c.g_12 * 0.0;
  c.vz0 = c.g_12 * d0y - c.g_22 * d0x;
  const St J0x = st2dx(a.J0, i2, ny), P0x = st2dx(a.P0, i2, ny);
  c.dJ0[0] = weno3Deriv(false, J0x);
  c.dJ0[1] = weno3Deriv(true, J0x);
  c.dP0[0] = weno3Deriv(false, P0x);
  c.dP0[1] = weno3Deriv(true, P0x);
  c.dJ0y = __ldg(a.J0 + i2 + 1) - __ldg(a.J0 + i2 - 1);
  c.dP0y = __ldg(a.P0 + i2 + 1) - __ldg(a.P0 + i2 - 1);
  c.bcx = __ldg(a.b0xcv_x + i2);
  c.bcy = __ldg(a.b0xcv_y + i2);
  c.bcz = __ldg(a.b0xcv_z + i2);
  c.xU[0] = c.xU[1] = c.xP[0] = c.xP[1] = c.xPsi[0] = c.xPsi[1] = 0.0;
  if (a.U_mode != nullptr && a.d2Jpar == nullptr) {
    // Delp2(Jpar) of mode n0 straight from the row modes (Coordinates::Delp2, coordinates.cxx:1662-1699:
    // a J(x-1) + b J(x) + c J(x+1) with Laplacian::tridagCoefs); eta is a Field2D on this path
    LapCoefs L{};
    L.dx = a.dx;
    L.g11 = a.g11;
    L.g33 = a.g33;
    L.g13 = a.g13;
    L.G1 = a.G1;
    L.G3 = a.G3;
    L.zlength = a.zlength;
    L.nx = p.nx;
    L.ny = ny;
    L.all_terms = a.delp2_all_terms;
    L.nonuniform = a.delp2_nonuniform;
    cd ca, cb, cc;
    tridagCoefs(L, x, y, a.filter_n0 * 2.0 * M_PI / a.zlength, false, false, ca, cb, cc);
    auto md = [&](long long q) { return cd{__ldg(a.J_mode + 2 * q), __ldg(a.J_mode + 2 * q + 1)}; };
    const cd d2 = cadd(cadd(cmul(ca, md(i2 - ny)), cmul(cb, md(i2))), cmul(cc, md(i2 + ny)));
    const double he = a.hyperresist * __ldg(a.eta + i2);
    c.xPsi[0] = -he * d2.re;
    c.xPsi[1] = -he * d2.im;
  }
  if (a.U_mode != nullptr) {
    // mode n0 of the slice pair of a field: M(y+1) exp(-i k (zs(y) - zs(y+1))) - M(y-1) exp(-i k (zs(y) - zs(y-1)))
    const double kwave = a.filter_n0 * 2.0 * M_PI / a.zlength;
    const double zs0 = __ldg(a.zShift + i2);
    double sn, cs;
    sincos(kwave * (zs0 - __ldg(a.zShift + i2 + 1)), &sn, &cs);
    const cd pu = {cs, -sn};
    sincos(kwave * (zs0 - __ldg(a.zShift + i2 - 1)), &sn, &cs);
    const cd pd = {cs, -sn};
    auto diff = [&](const double* __restrict__ mode) {
      const cd mu = {__ldg(mode + 2 * (i2 + 1)), __ldg(mode + 2 * (i2 + 1) + 1)};
      const cd md = {__ldg(mode + 2 * (i2 - 1)), __ldg(mode + 2 * (i2 - 1) + 1)};
      return csub(cmul(mu, pu), cmul(md, pd));
    };
    const cd dU = diff(a.U_mode), dP = diff(a.P_mode), dJ = diff(a.J_mode);
    const double cy = c.vy0 * 0.5 * c.rdy * c.rjs;     // y-advection by the phi0 velocity (b0x2d3d)
    const double cP = 0.5 * c.rdy * c.bcy;             // b0xcv.y DDY(P)
    const double cJ = 0.5 * c.gpar * c.B0sq;           // B0^2 Grad_par(Jpar)
    c.xU[0] = -cy * dU.re + cP * dP.re - cJ * dJ.re;
    c.xU[1] = -cy * dU.im + cP * dP.im - cJ * dJ.im;
    c.xP[0] = -cy * dP.re;
    c.xP[1] = -cy * dP.im;
  }
  return c;
}

constexpr int kRowCDoubles = static_cast<int>(sizeof(RowC) / sizeof(double));

__global__ void __launch_bounds__(128) k_elm_rowc(const ElmParams p) {
  const int nyb = p.ye - p.ys + 1;
  const long long nrows = static_cast<long long>(p.xe - p.xs + 1) * nyb;
  const long long rid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (rid >= nrows) {
    return;
  }
  const int x = p.xs + static_cast<int>(rid / nyb), y = p.ys + static_cast<int>(rid % nyb);
  const RowC c = elmRowConstants(p, x, y);
  double* __restrict__ dst = p.a.rowc_ws + rid * kRowCDoubles;
  const double* src = reinterpret_cast<const double*>(&c);
#pragma unroll
  for (int q = 0; q < kRowCDoubles; ++q) {
    dst[q] = src[q];
  }
}

template <int MINB, int V>
__global__ void __launch_bounds__(256, MINB) k_elmpb_rhs(const ElmParams p) {
  __shared__ double red[8 * 6 * 32];
  const bout_b200_elmpb_rhs_args& a = p.a;
  const int nzp = p.nz / V;                  // threads per row
  const int rloc = threadIdx.x / nzp, tin = threadIdx.x % nzp;
  const int nyb = p.ye - p.ys + 1;
  const long long nrows = static_cast<long long>(p.xe - p.xs + 1) * nyb;
  const long long rowid = static_cast<long long>(blockIdx.x) * p.rows_per_cta + rloc;
  const bool active = rowid < nrows;
  const long long rid = active ? rowid : nrows - 1;
  const int x = p.xs + static_cast<int>(rid / nyb);
  const int y = p.ys + static_cast<int>(rid % nyb);
  const long long i2 = static_cast<long long>(x) * p.ny + y;
  const int z0 = V * tin;
  const int ny = p.ny, nz = p.nz;

  __shared__ RowC rows[8];
  if (a.rowc_ws != nullptr) { // precomputed by k_elm_rowc: one coalesced load per row
    const double* __restrict__ src = a.rowc_ws + rid * kRowCDoubles;
    double* dst = reinterpret_cast<double*>(&rows[rloc]);
    for (int q = tin; q < kRowCDoubles; q += nzp) {
      dst[q] = __ldg(src + q);
    }
  } else if (tin == 0) {
    rows[rloc] = elmRowConstants(p, x, y);
  }
  __syncthreads();
  const RowC& m = rows[rloc];

  // One stencil field at a time (loads next to their use) so that at most one F3 is live.
  double tU[V], tP[V], tPsi[V];
  {
    const F3<V> phi = loadF3<V>(a.phi, a.phi_up, a.phi_dn, i2, ny, nz, z0);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const double gp = 0.5 * (m.B0p * phi.yp(j) - m.B0m * phi.ym(j)) * m.gpar; // Grad_par(phi B0)
      tPsi[j] = -1.0 * gp * m.rB0;
      tP[j] = -1.0 * b0x3d2d(phi.sx(j), phi.yp(j), phi.ym(j), phi.sz(j), m.dP0, m.dP0y, m);
    }
  }
  {
    const F3<V> P = loadF3<V>(a.P, a.P_up, a.P_dn, i2, ny, nz, z0);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const double dPx = applyStandard<BOUT_B200_DDX_C4>(P.sx(j)) * m.rdx;
      const double dPy = 0.5 * (P.yp(j) - P.ym(j)) * m.rdy;
      const double dPz = applyStandard<BOUT_B200_DDX_C4>(P.sz(j)) * m.rdz;
      tU[j] = dPx * m.bcx + dPy * m.bcy + dPz * m.bcz;
      tP[j] -= b0x2d3d(P.sx(j), P.yp(j), P.ym(j), P.sz(j), m);
    }
  }
  {
    const F3<V> Psi = loadF3<V>(a.Psi, a.Psi_up, a.Psi_dn, i2, ny, nz, z0);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const double bg = b0x2d3d(Psi.sx(j), Psi.yp(j), Psi.ym(j), Psi.sz(j), m);
      tPsi[j] -= bg * m.B0rBxy;
      const double t1 = b0x3d2d(Psi.sx(j), Psi.yp(j), Psi.ym(j), Psi.sz(j), m.dJ0, m.dJ0y, m);
      tU[j] = t1 * m.B0sq + tU[j];
    }
  }
  {
    const F3<V> U = loadF3<V>(a.U, a.U_up, a.U_dn, i2, ny, nz, z0);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      tU[j] -= b0x2d3d(U.sx(j), U.yp(j), U.ym(j), U.sz(j), m);
    }
  }
  {
    double J[V], Jup[V], Jdn[V], d2J[V], eta[V];
    ldv<V>(a.Jpar + i2 * nz + z0, J);
    if (a.d2Jpar != nullptr) {
      ldv<V>(a.d2Jpar + i2 * nz + z0, d2J);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        d2J[j] = 0.0;
      }
    }
    if (a.J_up != nullptr) {
      ldv<V>(a.J_up + (i2 + 1) * nz + z0, Jup);
      ldv<V>(a.J_dn + (i2 - 1) * nz + z0, Jdn);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        Jup[j] = Jdn[j] = 0.0;
      }
    }
    if (a.eta_is2d) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        eta[j] = __ldg(a.eta + i2);
      }
    } else {
      ldv<V>(a.eta + i2 * nz + z0, eta);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      tPsi[j] += eta[j] * J[j] - a.hyperresist * eta[j] * d2J[j];
      tU[j] -= 0.5 * (Jup[j] - Jdn[j]) * m.gpar * m.B0sq;
    }
  }

  const long long i3 = i2 * nz + z0;
  auto stv = [&](double* __restrict__ q, const double (&v)[V]) {
    if constexpr (V == 2) {
      *reinterpret_cast<double2*>(q) = make_double2(v[0], v[1]);
    } else {
      *q = v[0];
    }
  };
  if (a.filter_n0 < 0) { // no filter: raw time derivatives
    if (active) {
      stv(a.ddt_U + i3, tU);
      stv(a.ddt_P + i3, tP);
      stv(a.ddt_Psi + i3, tPsi);
    }
    return;
  }
  // ---- filter(., n0): X = (1/nz) sum_j f_j exp(-2 pi i n0 j / nz); f_j <- (k = 0: X) / (else: 2 Re X exp(+...))
  const int n0 = a.filter_n0;
  cd w[V];
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < V; ++j) {
    w[j] = p.tw[(static_cast<long long>(n0) * (z0 + j)) % nz];
    acc[0] += tU[j] * w[j].re;
    acc[1] += tU[j] * w[j].im;
    acc[2] += tP[j] * w[j].re;
    acc[3] += tP[j] * w[j].im;
    acc[4] += tPsi[j] * w[j].re;
    acc[5] += tPsi[j] * w[j].im;
  }
  rowReduce3(acc, red, tin, nzp, rloc);
  if (!active) {
    return;
  }
  acc[0] += m.xU[0];
  acc[1] += m.xU[1];
  acc[2] += m.xP[0];
  acc[3] += m.xP[1];
  acc[4] += m.xPsi[0];
  acc[5] += m.xPsi[1];
  const double inv = 1.0 / nz;
  const double mult = n0 == 0 ? 1.0 : 2.0;
  auto back = [&](double xr, double xi, double keep, double* __restrict__ out, const double* __restrict__ start,
                  double* __restrict__ stage) {
    // Re(X conj(w)) with w = exp(-i theta): w.re = cos(theta), w.im = -sin(theta)
    const double s = keep * mult * inv;
    double v[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = s * (xr * w[j].re + xi * w[j].im);
    }
    stv(out + i3, v);
    if (stage != nullptr) { // explicit RK stage state: start + (c dt) ddt   (rk4.cxx:179-200)
      double st[V];
      ldv<V>(start + i3, st);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        st[j] += a.rk_cdt * v[j];
      }
      stv(stage + i3, st);
    }
  };
  back(acc[0], acc[1], a.keep_U, a.ddt_U, a.rk_start_U, a.rk_stage_U);
  back(acc[2], acc[3], a.keep_P, a.ddt_P, a.rk_start_P, a.rk_stage_P);
  back(acc[4], acc[5], a.keep_Psi, a.ddt_Psi, a.rk_start_Psi, a.rk_stage_Psi);
}

} // namespace b200

using namespace b200;

extern "C" int bout_b200_elmpb_rowc_doubles(void) { return kRowCDoubles; }

extern "C" int bout_b200_elmpb_rhs(const bout_b200_grid* g, const bout_b200_elmpb_rhs_args* args, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(args != nullptr, BOUT_B200_ERR_ARG, "elmpb_rhs: null arguments");
  const bout_b200_elmpb_rhs_args& a = *args;
  const void* ptrs[] = {a.dx, a.dy, a.dz, a.J, a.g_22, a.g_12, a.g_23, a.Bxy, a.B0, a.J0, a.P0, a.phi0, a.b0xcv_x,
                        a.b0xcv_y, a.b0xcv_z, a.U, a.P, a.Psi, a.phi, a.Jpar, a.eta,
                        a.Psi_up, a.Psi_dn, a.phi_up, a.phi_dn, a.ddt_U, a.ddt_P, a.ddt_Psi};
  for (const void* q : ptrs) {
    B200_REQUIRE(q != nullptr, BOUT_B200_ERR_ARG, "elmpb_rhs: null field");
  }
  const bool modes = a.U_mode != nullptr;
  if (a.d2Jpar == nullptr) {
    B200_REQUIRE(modes && a.eta_is2d && a.g11 && a.g33 && a.g13 && (!a.delp2_all_terms || (a.G1 && a.G3)),
                 BOUT_B200_ERR_ARG,
                 "elmpb_rhs: without d2Jpar the mode path, a Field2D eta and the Delp2 metric (g11, g33, g13, G1, G3) are needed");
  }
  if (modes) {
    B200_REQUIRE(a.P_mode && a.J_mode && a.zShift && a.zlength > 0.0 && a.filter_n0 >= 0, BOUT_B200_ERR_ARG,
                 "elmpb_rhs: the mode path needs U_mode, P_mode, J_mode, zShift, zlength and the in-kernel filter");
    B200_REQUIRE(!a.U_up && !a.U_dn && !a.P_up && !a.P_dn && !a.J_up && !a.J_dn, BOUT_B200_ERR_ARG,
                 "elmpb_rhs: give either the U, P, Jpar slices or their modes, not both");
  } else {
    B200_REQUIRE(a.U_up && a.U_dn && a.P_up && a.P_dn && a.J_up && a.J_dn, BOUT_B200_ERR_ARG,
                 "elmpb_rhs: null slice");
  }
  const void* vec[] = {a.U, a.P, a.Psi, a.phi, a.Jpar, a.d2Jpar /* may be NULL */, a.U_up, a.U_dn, a.P_up, a.P_dn, a.Psi_up, a.Psi_dn,
                       a.phi_up, a.phi_dn, a.J_up, a.J_dn, a.ddt_U, a.ddt_P, a.ddt_Psi};
  for (const void* q : vec) {
    B200_REQUIRE(aligned16(q), BOUT_B200_ERR_UNSUPPORTED, "elmpb_rhs: fields must be 16-byte aligned");
  }
  B200_REQUIRE(a.eta_is2d || aligned16(a.eta), BOUT_B200_ERR_UNSUPPORTED, "elmpb_rhs: eta must be 16-byte aligned");
  const void* rk[] = {a.rk_start_U, a.rk_start_P, a.rk_start_Psi, a.rk_stage_U, a.rk_stage_P, a.rk_stage_Psi};
  int nrk = 0;
  for (const void* q : rk) {
    nrk += q != nullptr;
    B200_REQUIRE(aligned16(q), BOUT_B200_ERR_UNSUPPORTED, "elmpb_rhs: RK fields must be 16-byte aligned");
  }
  B200_REQUIRE(nrk == 0 || (nrk == 6 && a.filter_n0 >= 0), BOUT_B200_ERR_ARG,
               "elmpb_rhs: the RK stage epilogue needs all six rk_* fields and the in-kernel filter");
  B200_REQUIRE(a.rk_stage_U != a.U && a.rk_stage_P != a.P && a.rk_stage_Psi != a.Psi, BOUT_B200_ERR_ARG,
               "elmpb_rhs: a stage state must not be a field this launch reads neighbours of");
  B200_REQUIRE(g->nz % 2 == 0 && g->nz >= 4, BOUT_B200_ERR_UNSUPPORTED, "elmpb_rhs: nz must be even and >= 4");
  B200_REQUIRE(g->xstart >= 2 && g->nx - 1 - g->xend >= 2 && g->ystart >= 1 && g->ny - 1 - g->yend >= 1,
               BOUT_B200_ERR_GUARDS, "elmpb_rhs: needs 2 x guard cells and 1 y guard cell");
  // V = 2 z points per thread (16-byte transactions, 80 registers, 24 warps/SM) or V = 1 (8-byte transactions, fewer
  // registers, more warps in flight); BOUT_B200_ELM_VEC overrides the default
  static const int vec_env = getenv("BOUT_B200_ELM_VEC") ? atoi(getenv("BOUT_B200_ELM_VEC")) : 0;
  int V = vec_env == 1 ? 1 : 2;
  if (g->nz / V > 256 && V == 1) {
    V = 2;
  }
  const int nzp = g->nz / V;
  B200_REQUIRE(a.filter_n0 < g->nz / 2, BOUT_B200_ERR_ARG,
               "elmpb_rhs: the kept mode must be below the Nyquist mode (see the note on field-aligned terms)");
  int threads, rows;
  B200_REQUIRE(nzp <= 256, BOUT_B200_ERR_UNSUPPORTED, "elmpb_rhs: nz > 512 not supported");
  rows = (nzp % 32 == 0) ? 256 / nzp : 1; // several short rows per CTA when whole warps divide evenly
  rows = rows > 8 ? 8 : rows;
  threads = rows * nzp;
  ElmParams p{};
  p.a = a;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = g->xstart;
  p.xe = g->xend;
  p.ys = g->ystart;
  p.ye = g->yend;
  p.rows_per_cta = rows;
  cudaStream_t s = asStream(stream);
  int st = 0;
  p.tw = twiddleTable(g->nz, s, &st);
  if (st != 0) {
    return st;
  }
  const long long nrows = static_cast<long long>(p.xe - p.xs + 1) * (p.ye - p.ys + 1);
  const unsigned blocks = static_cast<unsigned>((nrows + rows - 1) / rows);
  if (a.rowc_ws != nullptr) {
    k_elm_rowc<<<static_cast<unsigned>((nrows + 127) / 128), 128, 0, s>>>(p);
    B200_LAUNCH_OK();
  }
  static const int minb = getenv("BOUT_B200_ELM_MINB") ? atoi(getenv("BOUT_B200_ELM_MINB")) : 0;
  if (V == 1) {
    if (minb == 3) {
      k_elmpb_rhs<3, 1><<<blocks, threads, 0, s>>>(p);
    } else {
      k_elmpb_rhs<4, 1><<<blocks, threads, 0, s>>>(p);
    }
  } else if (minb == 2) {
    k_elmpb_rhs<2, 2><<<blocks, threads, 0, s>>>(p);
  } else if (minb == 4) {
    k_elmpb_rhs<4, 2><<<blocks, threads, 0, s>>>(p);
  } else {
    k_elmpb_rhs<3, 2><<<blocks, threads, 0, s>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}
