// Fused Arakawa bracket kernels (subsystem 2): one pass, no temporaries.
// Reference: bracket(Field3D,Field3D,BRACKET_ARAKAWA) src/mesh/difops.cxx:940-1041;
//            bracket(Field3D,Field2D,BRACKET_ARAKAWA) src/mesh/difops.cxx:693-765.
//
// Each thread owns a (y, z-pair) column and marches along x with the three x-planes of f and g
// held in registers (z-1, z, z+1, z+2 of each), so f and g are each read once per chunk; the
// z-neighbour scalars come out of L1.  Algorithmic traffic: 2 reads + 1 write = 24 B/point.
#include <cstdlib>

#include "bracket.cuh"

namespace b200 {

struct BracketParams {
  const double* __restrict__ f;
  const double* __restrict__ g; // Field3D, or Field2D for the 3d2d kernel
  const double* __restrict__ dx;
  const double* __restrict__ dz;
  double* __restrict__ out;
  int nx, ny, nz, xs, xe, ys, ye, chunk;
};

template <int VEC, bool G2D, int MINB>
__global__ void __launch_bounds__(256, MINB) k_bracket_arakawa(const BracketParams p) {
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

  Row4 Fm, F0, Fp, Gm, G0, Gp;
  double gm = 0, g0 = 0, gp = 0;
  F0 = loadRow4<VEC>(p.f + (x0 - 1) * plane + rowoff, z, p.nz);
  Fp = loadRow4<VEC>(p.f + x0 * plane + rowoff, z, p.nz);
  if constexpr (G2D) {
    g0 = __ldg(p.g + static_cast<long long>(x0 - 1) * p.ny + y);
    gp = __ldg(p.g + static_cast<long long>(x0) * p.ny + y);
  } else {
    G0 = loadRow4<VEC>(p.g + (x0 - 1) * plane + rowoff, z, p.nz);
    Gp = loadRow4<VEC>(p.g + x0 * plane + rowoff, z, p.nz);
  }
#pragma unroll 2
  for (int x = x0; x <= x1; ++x) {
    Fm = F0;
    F0 = Fp;
    Fp = loadRow4<VEC>(p.f + (x + 1) * plane + rowoff, z, p.nz);
    if constexpr (G2D) {
      gm = g0;
      g0 = gp;
      gp = __ldg(p.g + static_cast<long long>(x + 1) * p.ny + y);
    } else {
      Gm = G0;
      G0 = Gp;
      Gp = loadRow4<VEC>(p.g + (x + 1) * plane + rowoff, z, p.nz);
    }
    const long long i2 = static_cast<long long>(x) * p.ny + y;
    // 1 / (12 dz dx), difops.cxx:961: dx, dz > 0, so the Newton-refined hardware reciprocal applies (<= 2 ulp)
    const double spacingFactor = rcpPositive(12 * __ldg(p.dz + i2) * __ldg(p.dx + i2));
    Vz<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      double J;
      if constexpr (G2D) {
        J = arakawaCore2D(Fm, F0, Fp, gm, g0, gp, i);
      } else {
        J = arakawaCore(Fm, F0, Fp, Gm, G0, Gp, i);
      }
      r.v[i] = J * spacingFactor;
    }
    r.store(p.out + x * plane + rowoff + z);
  }
}

template <bool G2D>
int launchBracket(const bout_b200_grid* g, const double* f, const double* gg, const double* dx,
                  const double* dz, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && gg && dx && dz && out, BOUT_B200_ERR_ARG, "bracket: null argument");
  B200_REQUIRE(out != f && out != gg, BOUT_B200_ERR_ARG, "bracket: out must not alias an input");
  B200_REQUIRE(g->xstart >= 1 && g->xend <= g->nx - 2, BOUT_B200_ERR_GUARDS,
               "bracket: needs one x guard cell each side");
  BracketParams p{f, gg, dx, dz, out, g->nx, g->ny, g->nz, g->xstart, g->xend, g->ystart, g->yend, 16};
  const bool vec2 = (g->nz % 2 == 0) && aligned16(f) && aligned16(out) && (G2D || aligned16(gg));
  const int nzv = vec2 ? g->nz / 2 : g->nz;
  const long long north = static_cast<long long>(p.ye - p.ys + 1) * nzv;
  const int len = p.xe - p.xs + 1;
  int chunk = 4;
  for (int c : {64, 32, 16, 8}) {
    if (north * ((len + c - 1) / c) >= 148LL * 2048) {
      chunk = c;
      break;
    }
  }
  p.chunk = chunk;
  dim3 grid(static_cast<unsigned>((north + 255) / 256), static_cast<unsigned>((len + chunk - 1) / chunk));
  // 80 registers (3 CTAs per SM) unless asked otherwise: BOUT_B200_BRACKET_MINB=4 caps at 64 with a few spills
  static const int minb = getenv("BOUT_B200_BRACKET_MINB") ? atoi(getenv("BOUT_B200_BRACKET_MINB")) : 3;
  if (vec2 && minb == 4) {
    k_bracket_arakawa<2, G2D, 4><<<grid, 256, 0, asStream(stream)>>>(p);
  } else if (vec2) {
    k_bracket_arakawa<2, G2D, 3><<<grid, 256, 0, asStream(stream)>>>(p);
  } else {
    k_bracket_arakawa<1, G2D, 3><<<grid, 256, 0, asStream(stream)>>>(p);
  }
  B200_LAUNCH_OK();
  return 0;
}

} // namespace b200

extern "C" int bout_b200_bracket_arakawa(const bout_b200_grid* g, const double* f,
                                         const double* gg, const double* dx, const double* dz,
                                         double* out, void* stream) {
  return b200::launchBracket<false>(g, f, gg, dx, dz, out, stream);
}

extern "C" int bout_b200_bracket_arakawa_3d2d(const bout_b200_grid* g, const double* f,
                                              const double* g2d, const double* dx,
                                              const double* dz, double* out, void* stream) {
  return b200::launchBracket<true>(g, f, g2d, dx, dz, out, stream);
}
