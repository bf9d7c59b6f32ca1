// Device-side stencil functors: the arithmetic of the reference's REGISTER_*_DERIVATIVE bodies
// (src/mesh/index_derivs.cxx:70-301), written once and used by the stand-alone stencil kernels and
// by the fused RHS-term kernels.  Operation order follows the reference so results agree to
// round-off; fp64 throughout, no fast-math.
#pragma once

#include "common.cuh"

namespace b200 {

struct St {
  double mm, m, c, p, pp;
};

constexpr double kWenoSmall = 1.0e-8; // include/bout/index_derivs.hxx
// The reference divides by 12.; multiplying by the rounded reciprocal differs by <= 1 ulp and
// avoids an fp64 division (~25 DFMA issue slots) per point.
constexpr double kTwelfth = 1.0 / 12.0;

__host__ __device__ constexpr int opGuards(int op) {
  return (op == BOUT_B200_DDX_C2 || op == BOUT_B200_DDX_W2 || op == BOUT_B200_D2DX2_C2
          || op == BOUT_B200_VDDX_C2 || op == BOUT_B200_VDDX_U1 || op == BOUT_B200_FDDX_U1)
             ? 1
             : 2;
}
// 0 = standard (f only), 1 = upwind (v at the point), 2 = flux (stencil of v)
__host__ __device__ constexpr int opKind(int op) { return op < 20 ? 0 : (op < 30 ? 1 : 2); }

__device__ __forceinline__ double sq(double x) { return x * x; }

// VDDX_WENO3, index_derivs.cxx:190-216.  The reference computes r = a/b then w = 1/(1 + 2 r^2);
// here w = b^2 / (b^2 + 2 a^2) -- the same number with ONE fp64 division instead of two (fp64
// division is ~25 DFMA slots on sm_100 and made this stencil compute-bound); the two forms differ
// by a few ulp, far inside the 1e-12 parity tolerance.
// That one division goes through rcpPositive (common.cuh): the fused elm-pb kernel evaluates six of these per point.
// The derivative factor of VDDX_WENO3 for one upwind direction (positive = velocity > 0); weno3(vc, f) = vc * it.
__device__ __forceinline__ double weno3Deriv(bool positive, const St& f) {
  double deriv, a;
  const double b = kWenoSmall + sq(f.p - 2.0 * f.c + f.m);
  if (positive) {
    a = kWenoSmall + sq(f.c - 2.0 * f.m + f.mm);
    deriv = (-f.mm + 3. * f.m - 3. * f.c + f.p);
  } else {
    a = kWenoSmall + sq(f.pp - 2.0 * f.p + f.c);
    deriv = (-f.m + 3. * f.c - 3. * f.p + f.pp);
  }
  const double b2 = b * b;
  const double w = b2 * rcpPositive(b2 + 2.0 * a * a);
  return 0.5 * ((f.p - f.m) - w * deriv);
}
__device__ __forceinline__ double weno3(double vc, const St& f) { return vc * weno3Deriv(vc > 0.0, f); }

template <int OP>
__device__ __forceinline__ double applyStandard(const St& f) {
  if constexpr (OP == BOUT_B200_DDX_C2) {
    return 0.5 * (f.p - f.m);
  } else if constexpr (OP == BOUT_B200_DDX_C4) {
    return (8. * f.p - 8. * f.m + f.mm - f.pp) * kTwelfth;
  } else if constexpr (OP == BOUT_B200_DDX_W2) { // :80-99
    const double dc = 0.5 * (f.p - f.m), dl = f.c - f.m, dr = f.p - f.c;
    const double isl = sq(dl), isr = sq(dr);
    const double isc = (13. / 3.) * sq(f.p - 2. * f.c + f.m) + 0.25 * sq(f.p - f.m);
    const double al = 0.25 / sq(kWenoSmall + isl), ar = 0.25 / sq(kWenoSmall + isr);
    const double ac = 0.5 / sq(kWenoSmall + isc), sa = al + ar + ac;
    return (al * dl + ar * dr + ac * dc) / sa;
  } else if constexpr (OP == BOUT_B200_DDX_S2) { // :102-110, SIGN(0) = +1
    double result = (8. * f.p - 8. * f.m + f.mm - f.pp) * kTwelfth;
    const double sign = f.c < 0.0 ? -1.0 : 1.0;
    result += sign * (f.pp - 4. * f.p + 6. * f.c - 4. * f.m + f.mm) * kTwelfth;
    return result;
  } else if constexpr (OP == BOUT_B200_DDX_W3) { // :219-255 split-flux CWENO3
    double ma = fabs(f.c);
    ma = fmax(ma, fabs(f.m));
    ma = fmax(ma, fabs(f.p));
    ma = fmax(ma, fabs(f.mm));
    ma = fmax(ma, fabs(f.pp));
    const St sp = {f.mm + ma, f.m + ma, f.c + ma, f.p + ma, f.pp + ma};
    const St sm = {ma - f.mm, ma - f.m, ma - f.c, ma - f.p, ma - f.pp};
    return weno3(0.5, sp) + weno3(-0.5, sm);
  } else if constexpr (OP == BOUT_B200_D2DX2_C2) {
    return f.p + f.m - 2. * f.c;
  } else if constexpr (OP == BOUT_B200_D2DX2_C4) {
    return (-f.pp + 16. * f.p - 30. * f.c + 16. * f.m - f.mm) * kTwelfth;
  } else if constexpr (OP == BOUT_B200_D4DX4_C2) {
    return (f.pp - 4. * f.p + 6. * f.c - 4. * f.m + f.mm);
  } else {
    return 0.0;
  }
}

template <int OP>
__device__ __forceinline__ double applyUpwind(double vc, const St& f) {
  if constexpr (OP == BOUT_B200_VDDX_C2) {
    return vc * 0.5 * (f.p - f.m);
  } else if constexpr (OP == BOUT_B200_VDDX_C4) {
    return vc * (8. * f.p - 8. * f.m + f.mm - f.pp) * kTwelfth;
  } else if constexpr (OP == BOUT_B200_VDDX_U1) {
    return vc >= 0.0 ? vc * (f.c - f.m) : vc * (f.p - f.c);
  } else if constexpr (OP == BOUT_B200_VDDX_U2) {
    return vc >= 0.0 ? vc * (1.5 * f.c - 2.0 * f.m + 0.5 * f.mm)
                     : vc * (-0.5 * f.pp + 2.0 * f.p - 1.5 * f.c);
  } else if constexpr (OP == BOUT_B200_VDDX_U3) {
    return vc >= 0.0 ? vc * (4. * f.p - 12. * f.m + 2. * f.mm + 6. * f.c) * kTwelfth
                     : vc * (-4. * f.m + 12. * f.p - 2. * f.pp - 6. * f.c) * kTwelfth;
  } else if constexpr (OP == BOUT_B200_VDDX_W3) {
    return weno3(vc, f);
  } else {
    return 0.0;
  }
}

template <int OP>
__device__ __forceinline__ double applyFlux(const St& v, const St& f) {
  if constexpr (OP == BOUT_B200_FDDX_U1) { // :266-281
    double vs = 0.5 * (v.m + v.c);
    double result = (vs >= 0.0) ? vs * f.m : vs * f.c;
    vs = 0.5 * (v.c + v.p);
    result -= (vs >= 0.0) ? vs * f.c : vs * f.p;
    return -result;
  } else if constexpr (OP == BOUT_B200_FDDX_U2) { // :283-293
    double vs = 0.5 * (v.m + v.c);
    double result =
        (vs >= 0.0) ? vs * (1.5 * f.m - 0.5 * f.mm) : vs * (1.5 * f.c - 0.5 * f.p);
    vs = 0.5 * (v.c + v.p);
    result -= (vs >= 0.0) ? vs * (1.5 * f.c - 0.5 * f.m) : vs * (1.5 * f.p - 0.5 * f.pp);
    return -result;
  } else if constexpr (OP == BOUT_B200_FDDX_C2) {
    return 0.5 * (v.p * f.p - v.m * f.m);
  } else if constexpr (OP == BOUT_B200_FDDX_C4) {
    return (8. * v.p * f.p - 8. * v.m * f.m + v.mm * f.mm - v.pp * f.pp) * kTwelfth;
  } else {
    return 0.0;
  }
}

// Uniform entry: any op kind from (vc, vstencil, fstencil)
template <int OP>
__device__ __forceinline__ double applyOp(double vc, const St& v, const St& f) {
  if constexpr (opKind(OP) == 0) {
    return applyStandard<OP>(f);
  } else if constexpr (opKind(OP) == 1) {
    return applyUpwind<OP>(vc, f);
  } else {
    return applyFlux<OP>(v, f);
  }
}

} // namespace b200
