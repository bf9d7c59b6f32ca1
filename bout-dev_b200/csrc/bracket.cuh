// Arakawa bracket arithmetic (reference src/mesh/difops.cxx:940-1041 and :693-765), shared by the
// stand-alone bracket kernels and the fused RHS kernels.
#pragma once

#include "common.cuh"

namespace b200 {

// Four consecutive z values of one (x,y) row around a z-pair: z-1, z, z+1, z+2 (periodic).
struct Row4 {
  double v[4];
};

template <int VEC>
__device__ __forceinline__ Row4 loadRow4(const double* __restrict__ row, int z, int nz) {
  Row4 r;
  if constexpr (VEC == 2) {
    const double2 c = __ldg(reinterpret_cast<const double2*>(row + z));
    r.v[0] = __ldg(row + wrapz(z - 1, nz));
    r.v[1] = c.x;
    r.v[2] = c.y;
    r.v[3] = __ldg(row + wrapz(z + 2, nz));
  } else {
    r.v[0] = __ldg(row + wrapz(z - 1, nz));
    r.v[1] = __ldg(row + z);
    r.v[2] = __ldg(row + wrapz(z + 1, nz));
    r.v[3] = 0.0;
  }
  return r;
}

// [f,g] at z-lane i (0 or 1) from the three x-planes of f and g; NOT yet multiplied by the
// spacing factor 1/(12 dx dz).
__device__ __forceinline__ double arakawaCore(const Row4& Fxm, const Row4& Fx, const Row4& Fxp,
                                              const Row4& Gxm, const Row4& Gx, const Row4& Gxp,
                                              int i) {
  const int zm = i, z = i + 1, zp = i + 2;
  // J++ = DDZ(f)*DDX(g) - DDX(f)*DDZ(g)
  const double Jpp = ((Fx.v[zp] - Fx.v[zm]) * (Gxp.v[z] - Gxm.v[z])
                      - (Fxp.v[z] - Fxm.v[z]) * (Gx.v[zp] - Gx.v[zm]));
  // J+x
  const double Jpx =
      (Gxp.v[z] * (Fxp.v[zp] - Fxp.v[zm]) - Gxm.v[z] * (Fxm.v[zp] - Fxm.v[zm])
       - Gx.v[zp] * (Fxp.v[zp] - Fxm.v[zp]) + Gx.v[zm] * (Fxp.v[zm] - Fxm.v[zm]));
  // Jx+
  const double Jxp =
      (Gxp.v[zp] * (Fx.v[zp] - Fxp.v[z]) - Gxm.v[zm] * (Fxm.v[z] - Fx.v[zm])
       - Gxm.v[zp] * (Fx.v[zp] - Fxm.v[z]) + Gxp.v[zm] * (Fxp.v[z] - Fx.v[zm]));
  return Jpp + Jpx + Jxp;
}

// [f, g2d] for a Field2D g (difops.cxx:693-765)
__device__ __forceinline__ double arakawaCore2D(const Row4& fxm, const Row4& fc, const Row4& fxp,
                                                double gxm, double gc, double gxp, int i) {
  const int zm = i, zp = i + 2;
  const double Jpp = 2 * (fc.v[zp] - fc.v[zm]) * (gxp - gxm);
  const double Jpx = gxp * (fxp.v[zp] - fxp.v[zm]) - gxm * (fxm.v[zp] - fxm.v[zm])
                     + gc * (fxp.v[zm] - fxp.v[zp] - fxm.v[zm] + fxm.v[zp]);
  return Jpp + Jpx;
}

} // namespace b200
