// Index-derivative stencil kernels (subsystem 1 of the hot path).
//
// Replaces DerivativeType::standard / upwindOrFlux + populateStencil of the reference
// (include/bout/index_derivs.hxx:84-116, include/bout/stencils.hxx:40-119) and the follow-up
// "/ dx" passes of Coordinates::DDX.. (src/mesh/coordinates.cxx:1467-1514).
//
// Design (HBM/L2-bound fp64):
//  * z is the contiguous index: every thread owns a z-PAIR and moves it with 16-byte double2
//    transactions (VEC=2; VEC=1 scalar path for odd nz / unaligned views).
//  * X and Y stencils "march": a thread keeps the 2*NG+1 neighbour planes of its (orthogonal, z)
//    column in registers and walks a chunk of the stencil direction, so each input element is
//    loaded once per chunk instead of 3-5 times (L2 bandwidth on B200 is not much above HBM
//    bandwidth, so re-reads from L2 would cap the kernel well below the HBM roofline).
//  * Z stencils read the two neighbouring double2 (or the two neighbouring scalars for NG=1);
//    those hit L1, so DRAM/L2 traffic stays at one read + one write per point.
//  * The metric scaling (1/dx, 1/dx^2, d1_dx/dx ...) is applied in the same pass.
#include "stencil_ops.cuh"

namespace b200 {

struct StencilParams {
  const double* __restrict__ v;
  const double* __restrict__ f;
  const double* __restrict__ up;   // plane source for +direction neighbours (Y slices) or f
  const double* __restrict__ down; // plane source for -direction neighbours or f
  const double* __restrict__ up2;  // second parallel slices (y + 2 / y - 2) when the op needs two
  const double* __restrict__ down2;
  const double* __restrict__ scale1;
  const double* __restrict__ scale2;
  double* __restrict__ out;
  int nx, ny, nz;
  int xs, xe, ys, ye; // output box
  int chunk;          // marching chunk length
};

template <int VEC, int NG>
__device__ __forceinline__ St lane(const Vz<VEC> (&w)[2 * NG + 1], int i) {
  // w[0..2NG] window of planes centred at w[NG]; pick z-lane i
  St s;
  s.c = w[NG].v[i];
  s.m = w[NG - 1].v[i];
  s.p = w[NG + 1].v[i];
  if constexpr (NG == 2) {
    s.mm = w[0].v[i];
    s.pp = w[4].v[i];
  } else {
    s.mm = 0.0;
    s.pp = 0.0;
  }
  return s;
}

// Marching kernel for DIR = X (0) or Y (1).
template <int DIR, int OP, int OP2, int VEC>
__global__ void __launch_bounds__(256) k_stencil_march(const StencilParams p) {
  constexpr int NG = (OP2 >= 0 && opGuards(OP2 >= 0 ? OP2 : 0) > opGuards(OP)) ? 2 : opGuards(OP);
  constexpr int W = 2 * NG + 1;
  constexpr int KIND = opKind(OP);
  const int nzv = p.nz / VEC;
  // orthogonal extent: for X-march it is y, for Y-march it is x
  const int os = DIR == 0 ? p.ys : p.xs;
  const int oe = DIR == 0 ? p.ye : p.xe;
  const int ms = DIR == 0 ? p.xs : p.ys;
  const int me = DIR == 0 ? p.xe : p.ye;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(oe - os + 1) * nzv) {
    return;
  }
  const int zv = static_cast<int>(t % nzv);
  const int o = os + static_cast<int>(t / nzv);
  const int m0 = ms + blockIdx.y * p.chunk;
  const int m1 = min(m0 + p.chunk - 1, me);
  const long long stride = DIR == 0 ? static_cast<long long>(p.ny) * p.nz : p.nz;
  // element offset of (m, o, z)
  const long long ortho_off = DIR == 0 ? static_cast<long long>(o) * p.nz
                                       : static_cast<long long>(o) * p.ny * p.nz;
  const long long col = ortho_off + static_cast<long long>(zv) * VEC;
  // 2D (x,y) index stride for the scale arrays
  const long long stride2 = DIR == 0 ? p.ny : 1;
  const long long col2 = DIR == 0 ? o : static_cast<long long>(o) * p.ny;

  Vz<VEC> w[W];
  Vz<VEC> wv[W];
#pragma unroll
  for (int i = 0; i < W - 1; ++i) {
    const int m = m0 - NG + i;
    const double* src = (i < NG) ? p.down : (i == NG ? p.f : p.up);
    w[i + 1] = Vz<VEC>::load(src + col + m * stride);
    if constexpr (KIND == 2) {
      wv[i + 1] = Vz<VEC>::load(p.v + col + m * stride);
    }
  }
#pragma unroll 4
  for (int m = m0; m <= m1; ++m) {
    // shift window and bring in plane m+NG
#pragma unroll
    for (int i = 0; i < W - 1; ++i) {
      w[i] = w[i + 1];
    }
    w[W - 1] = Vz<VEC>::load(p.up + col + (m + NG) * stride);
    if constexpr (KIND == 2) {
#pragma unroll
      for (int i = 0; i < W - 1; ++i) {
        wv[i] = wv[i + 1];
      }
      wv[W - 1] = Vz<VEC>::load(p.v + col + (m + NG) * stride);
    }
    // When separate yup/ydown slices are given, the centre comes from f and the planes that
    // were loaded as "up" for earlier centres must be re-read from the right source.
    if (p.up != p.f) {
      w[NG] = Vz<VEC>::load(p.f + col + m * stride);
      w[NG - 1] = Vz<VEC>::load(p.down + col + (m - 1) * stride);
      w[NG + 1] = Vz<VEC>::load(p.up + col + (m + 1) * stride);
      if constexpr (NG == 2) { // each offset has its own slice field (Field3D::ynext(+-2))
        w[0] = Vz<VEC>::load(p.down2 + col + (m - 2) * stride);
        w[4] = Vz<VEC>::load(p.up2 + col + (m + 2) * stride);
      }
    }
    Vz<VEC> vc;
    if constexpr (KIND == 1) {
      vc = Vz<VEC>::load(p.v + col + m * stride);
    }
    const double s1 = p.scale1 ? __ldg(p.scale1 + col2 + m * stride2) : 1.0;
    double s2 = 0.0;
    if constexpr (OP2 >= 0) {
      s2 = p.scale2 ? __ldg(p.scale2 + col2 + m * stride2) : 1.0;
    }
    Vz<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const St fs = lane<VEC, NG>(w, i);
      St vs = {0, 0, 0, 0, 0};
      double vci = 0.0;
      if constexpr (KIND == 2) {
        vs = lane<VEC, NG>(wv, i);
      }
      if constexpr (KIND == 1) {
        vci = vc.v[i];
      }
      double val = applyOp<OP>(vci, vs, fs) * s1;
      if constexpr (OP2 >= 0) {
        val += (applyStandard<(OP2 >= 0 ? OP2 : 0)>(fs)) * s2;
      }
      r.v[i] = val;
    }
    r.store(p.out + col + m * stride);
  }
}

// Z stencil: one thread per (x, y, z-vector).
template <int OP, int OP2, int VEC>
__global__ void __launch_bounds__(256) k_stencil_z(const StencilParams p) {
  constexpr int NG = (OP2 >= 0 && opGuards(OP2 >= 0 ? OP2 : 0) > opGuards(OP)) ? 2 : opGuards(OP);
  constexpr int KIND = opKind(OP);
  const int nzv = p.nz / VEC;
  const int nyb = p.ye - p.ys + 1;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.xe - p.xs + 1) * nyb * nzv;
  if (t >= total) {
    return;
  }
  int zv, yb_, xb_;
  splitIndex(t, nzv, nyb, zv, yb_, xb_);
  const int y = p.ys + yb_, x = p.xs + xb_;
  const long long row2 = static_cast<long long>(x) * p.ny + y;
  const long long row = row2 * p.nz;
  const int z = zv * VEC;
  const int nz = p.nz;

  // gather f[z-NG .. z+VEC-1+NG] with periodic wrap (include/bout/region.hxx:297-310)
  double fw[VEC + 2 * NG];
  double vw[KIND == 2 ? VEC + 2 * NG : (KIND == 1 ? VEC : 1)];
  auto gather = [&](const double* __restrict__ src, double* dst) {
    if constexpr (VEC == 2) {
      const double2 c = __ldg(reinterpret_cast<const double2*>(src + row + z));
      dst[NG] = c.x;
      dst[NG + 1] = c.y;
      if constexpr (NG == 2) {
        const double2 l = __ldg(reinterpret_cast<const double2*>(src + row + wrapz(z - 2, nz)));
        const double2 r = __ldg(reinterpret_cast<const double2*>(src + row + wrapz(z + 2, nz)));
        dst[0] = l.x;
        dst[1] = l.y;
        dst[4] = r.x;
        dst[5] = r.y;
      } else {
        dst[0] = __ldg(src + row + wrapz(z - 1, nz));
        dst[3] = __ldg(src + row + wrapz(z + 2, nz));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 1 + 2 * NG; ++i) {
        int zz = (z - NG + i) % nz; // nz may be smaller than NG: use a true modulo
        if (zz < 0) {
          zz += nz;
        }
        dst[i] = __ldg(src + row + zz);
      }
    }
  };
  gather(p.f, fw);
  if constexpr (KIND == 2) {
    gather(p.v, vw);
  } else if constexpr (KIND == 1) {
    if constexpr (VEC == 2) {
      const double2 c = __ldg(reinterpret_cast<const double2*>(p.v + row + z));
      vw[0] = c.x;
      vw[1] = c.y;
    } else {
      vw[0] = __ldg(p.v + row + z);
    }
  }
  const double s1 = p.scale1 ? __ldg(p.scale1 + row2) : 1.0;
  double s2 = 0.0;
  if constexpr (OP2 >= 0) {
    s2 = p.scale2 ? __ldg(p.scale2 + row2) : 1.0;
  }
  Vz<VEC> r;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    St fs, vs = {0, 0, 0, 0, 0};
    fs.c = fw[NG + i];
    fs.m = fw[NG + i - 1];
    fs.p = fw[NG + i + 1];
    if constexpr (NG == 2) {
      fs.mm = fw[i];
      fs.pp = fw[i + 4];
    } else {
      fs.mm = 0.0;
      fs.pp = 0.0;
    }
    double vci = 0.0;
    if constexpr (KIND == 2) {
      vs.c = vw[NG + i];
      vs.m = vw[NG + i - 1];
      vs.p = vw[NG + i + 1];
      if constexpr (NG == 2) {
        vs.mm = vw[i];
        vs.pp = vw[i + 4];
      }
    } else if constexpr (KIND == 1) {
      vci = vw[i];
    }
    double val = applyOp<OP>(vci, vs, fs) * s1;
    if constexpr (OP2 >= 0) {
      val += (applyStandard<(OP2 >= 0 ? OP2 : 0)>(fs)) * s2;
    }
    r.v[i] = val;
  }
  r.store(p.out + row + z);
}

template <int OP, int OP2, int VEC>
int launchStencil(const StencilParams& p, int dir, cudaStream_t stream) {
  const int nzv = p.nz / VEC;
  if (dir == BOUT_B200_Z) {
    const long long total = static_cast<long long>(p.xe - p.xs + 1) * (p.ye - p.ys + 1) * nzv;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    k_stencil_z<OP, OP2, VEC><<<blocks, 256, 0, stream>>>(p);
  } else {
    const int len = dir == BOUT_B200_X ? (p.xe - p.xs + 1) : (p.ye - p.ys + 1);
    const long long north =
        static_cast<long long>(dir == BOUT_B200_X ? (p.ye - p.ys + 1) : (p.xe - p.xs + 1)) * nzv;
    // chunk: as long as possible (less halo re-read) while still filling the machine
    // (measured at 256x64x256, profiles/r02ah_micro_chunk.log: chunks of 4 beat 8 -- two waves of CTAs instead of a
    // partial one -- while 16 and 32 lose 20-40 %)
    int chunk = 4;
    for (int c : {64, 32, 16, 8}) {
      if (north * ((len + c - 1) / c) >= 148LL * 2048) {
        chunk = c;
        break;
      }
    }
    static const int chunk_env = getenv("BOUT_B200_MARCH_CHUNK") ? atoi(getenv("BOUT_B200_MARCH_CHUNK")) : 0;
    if (chunk_env > 0) {
      chunk = chunk_env;
    }
    StencilParams q = p;
    q.chunk = chunk;
    dim3 grid(static_cast<unsigned>((north + 255) / 256), static_cast<unsigned>((len + chunk - 1) / chunk));
    if (dir == BOUT_B200_X) {
      k_stencil_march<0, OP, OP2, VEC><<<grid, 256, 0, stream>>>(q);
    } else {
      k_stencil_march<1, OP, OP2, VEC><<<grid, 256, 0, stream>>>(q);
    }
  }
  B200_LAUNCH_OK();
  return 0;
}

template <int OP, int OP2>
int launchStencilVec(const StencilParams& p, int dir, bool vec2, cudaStream_t stream) {
  return vec2 ? launchStencil<OP, OP2, 2>(p, dir, stream) : launchStencil<OP, OP2, 1>(p, dir, stream);
}

} // namespace b200

using namespace b200;

static int stencilImpl(const bout_b200_grid* g, int op, int dir, int region, const double* v, const double* f,
                       const double* yup, const double* ydown, const double* yup2, const double* ydown2,
                       const double* scale1, int op2, const double* scale2, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f != nullptr && out != nullptr, BOUT_B200_ERR_ARG, "stencil: null field");
  B200_REQUIRE(dir >= 0 && dir <= 2, BOUT_B200_ERR_ARG, "stencil: bad direction %d", dir);
  B200_REQUIRE(f != out && v != out, BOUT_B200_ERR_ARG, "stencil: out must not alias an input");
  Box b;
  if (regionBox(g, region, &b) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  const int kind = opKind(op);
  B200_REQUIRE(kind == 0 || v != nullptr, BOUT_B200_ERR_ARG, "stencil: op %d needs a velocity field", op);
  B200_REQUIRE(op2 < 0 || (kind == 0 && (op2 == BOUT_B200_DDX_C2 || op2 == BOUT_B200_DDX_C4)),
               BOUT_B200_ERR_UNSUPPORTED, "stencil: unsupported second op %d", op2);
  cudaStream_t s = asStream(stream);

  // A direction with a single point gives an all-zero field (index_derivs_interface.hxx:152-156)
  const int npoint = dir == BOUT_B200_X ? g->nx : (dir == BOUT_B200_Y ? g->ny : g->nz);
  if (npoint == 1) {
    B200_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * g->nx * g->ny * g->nz, s));
    return 0;
  }
  int ng = opGuards(op);
  if (op2 >= 0 && opGuards(op2) > ng) {
    ng = opGuards(op2);
  }
  if (dir == BOUT_B200_X) {
    B200_REQUIRE(b.xs - ng >= 0 && b.xe + ng <= g->nx - 1, BOUT_B200_ERR_GUARDS,
                 "stencil: X region [%d,%d] needs %d guard cells (nx=%d)", b.xs, b.xe, ng, g->nx);
  } else if (dir == BOUT_B200_Y) {
    B200_REQUIRE(b.ys - ng >= 0 && b.ye + ng <= g->ny - 1, BOUT_B200_ERR_GUARDS,
                 "stencil: Y region [%d,%d] needs %d guard cells (ny=%d)", b.ys, b.ye, ng, g->ny);
    B200_REQUIRE((yup == nullptr) == (ydown == nullptr), BOUT_B200_ERR_ARG,
                 "stencil: yup and ydown must be given together");
    B200_REQUIRE(yup == nullptr || ng == 1 || (yup2 != nullptr && ydown2 != nullptr), BOUT_B200_ERR_ARG,
                 "stencil: op %d needs two parallel slices each way", op);
    B200_REQUIRE(yup == nullptr || kind != 2, BOUT_B200_ERR_UNSUPPORTED,
                 "stencil: flux operators on parallel slices are not supported");
  }
  StencilParams p;
  p.v = v;
  p.f = f;
  p.up = (dir == BOUT_B200_Y && yup != nullptr) ? yup : f;
  p.down = (dir == BOUT_B200_Y && ydown != nullptr) ? ydown : f;
  p.up2 = (dir == BOUT_B200_Y && yup != nullptr && yup2 != nullptr) ? yup2 : p.up;
  p.down2 = (dir == BOUT_B200_Y && ydown != nullptr && ydown2 != nullptr) ? ydown2 : p.down;
  p.scale1 = scale1;
  p.scale2 = scale2;
  p.out = out;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = b.xs;
  p.xe = b.xe;
  p.ys = b.ys;
  p.ye = b.ye;
  p.chunk = 16;
  const bool vec2 = (g->nz % 2 == 0) && aligned16(f) && aligned16(out) && (v == nullptr || aligned16(v))
                    && aligned16(p.up) && aligned16(p.down) && aligned16(p.up2) && aligned16(p.down2);

#define CASE1(OPC)                                                   \
  case OPC:                                                          \
    if (op2 < 0) {                                                   \
      return launchStencilVec<OPC, -1>(p, dir, vec2, s);             \
    }                                                                \
    break;
#define CASE2(OPC)                                                   \
  case OPC:                                                          \
    if (op2 < 0) {                                                   \
      return launchStencilVec<OPC, -1>(p, dir, vec2, s);             \
    } else if (op2 == BOUT_B200_DDX_C2) {                            \
      return launchStencilVec<OPC, BOUT_B200_DDX_C2>(p, dir, vec2, s); \
    } else {                                                         \
      return launchStencilVec<OPC, BOUT_B200_DDX_C4>(p, dir, vec2, s); \
    }
  switch (op) {
    CASE1(BOUT_B200_DDX_C2)
    CASE1(BOUT_B200_DDX_C4)
    CASE1(BOUT_B200_DDX_W2)
    CASE1(BOUT_B200_DDX_S2)
    CASE1(BOUT_B200_DDX_W3)
    CASE2(BOUT_B200_D2DX2_C2)
    CASE2(BOUT_B200_D2DX2_C4)
    CASE1(BOUT_B200_D4DX4_C2)
    CASE1(BOUT_B200_VDDX_C2)
    CASE1(BOUT_B200_VDDX_C4)
    CASE1(BOUT_B200_VDDX_U1)
    CASE1(BOUT_B200_VDDX_U2)
    CASE1(BOUT_B200_VDDX_U3)
    CASE1(BOUT_B200_VDDX_W3)
    CASE1(BOUT_B200_FDDX_U1)
    CASE1(BOUT_B200_FDDX_U2)
    CASE1(BOUT_B200_FDDX_C2)
    CASE1(BOUT_B200_FDDX_C4)
  default:
    break;
  }
#undef CASE1
#undef CASE2
  return fail(BOUT_B200_ERR_UNSUPPORTED, "stencil: unsupported op %d / op2 %d", op, op2);
}

extern "C" int bout_b200_stencil(const bout_b200_grid* g, int op, int dir, int region,
                                 const double* v, const double* f, const double* yup,
                                 const double* ydown, const double* scale1, int op2,
                                 const double* scale2, double* out, void* stream) {
  return stencilImpl(g, op, dir, region, v, f, yup, ydown, nullptr, nullptr, scale1, op2, scale2, out, stream);
}

extern "C" int bout_b200_stencil_yslices(const bout_b200_grid* g, int op, int region, const double* v,
                                         const double* f, int nslices, const double* const* yup,
                                         const double* const* ydown, const double* scale1, int op2,
                                         const double* scale2, double* out, void* stream) {
  B200_REQUIRE(nslices >= 1 && nslices <= 2 && yup && ydown && yup[0] && ydown[0], BOUT_B200_ERR_ARG,
               "stencil_yslices: 1 or 2 slices each way");
  return stencilImpl(g, op, BOUT_B200_Y, region, v, f, yup[0], ydown[0], nslices > 1 ? yup[1] : nullptr,
                     nslices > 1 ? ydown[1] : nullptr, scale1, op2, scale2, out, stream);
}
