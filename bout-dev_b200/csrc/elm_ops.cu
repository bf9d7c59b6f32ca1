// Operators of the elm-pb / reduced-MHD right-hand sides that are not plain one-direction stencils:
//   * b0xGrad_dot_Grad(phi, A)  -- the BRACKET_STD Poisson bracket (src/mesh/difops.cxx:455-601) as ONE
//     kernel: three derivatives of phi, the ExB-like velocity from the covariant metric, and the three
//     upwind advection terms of A, instead of the reference's chain of ~12 full-field passes;
//   * averageY (src/physics/smoothing.cxx:272-323) split into its local mean and its broadcast so that
//     the NYPE > 1 reduction (one all-reduce of an [nx][nz] plane) sits between the two;
//   * smooth_x (src/physics/smoothing.cxx:46-73), the 1-2-1 radial filter.
// Y neighbours come from parallel slices (yup/ydown fields, include/bout/field3d.hxx:265-300) when
// the mesh uses the shifted metric, or from the field itself (identity transform / Field2D).
#include "stencil_ops.cuh"

namespace b200 {

__device__ __forceinline__ double applyStandardRt(int op, const St& f) {
  switch (op) {
  case BOUT_B200_DDX_C2: return applyStandard<BOUT_B200_DDX_C2>(f);
  case BOUT_B200_DDX_C4: return applyStandard<BOUT_B200_DDX_C4>(f);
  case BOUT_B200_DDX_W2: return applyStandard<BOUT_B200_DDX_W2>(f);
  case BOUT_B200_DDX_S2: return applyStandard<BOUT_B200_DDX_S2>(f);
  default: return applyStandard<BOUT_B200_DDX_W3>(f);
  }
}
__device__ __forceinline__ double applyUpwindRt(int op, double vc, const St& f) {
  switch (op) {
  case BOUT_B200_VDDX_C2: return applyUpwind<BOUT_B200_VDDX_C2>(vc, f);
  case BOUT_B200_VDDX_C4: return applyUpwind<BOUT_B200_VDDX_C4>(vc, f);
  case BOUT_B200_VDDX_U1: return applyUpwind<BOUT_B200_VDDX_U1>(vc, f);
  case BOUT_B200_VDDX_U2: return applyUpwind<BOUT_B200_VDDX_U2>(vc, f);
  case BOUT_B200_VDDX_U3: return applyUpwind<BOUT_B200_VDDX_U3>(vc, f);
  default: return applyUpwind<BOUT_B200_VDDX_W3>(vc, f);
  }
}

struct B0xParams {
  bout_b200_b0xgrad c;
  const double* __restrict__ phi;
  const double* __restrict__ phi_up[2];
  const double* __restrict__ phi_dn[2];
  const double* __restrict__ A;
  const double* __restrict__ A_up[2];
  const double* __restrict__ A_dn[2];
  double* __restrict__ out;
  double* __restrict__ yterm; // optional: the VDDY term on its own (already / (J sqrt(g_22))), left out of `out`
  int nx, ny, nz, xs, xe, ys, ye;
  int phi2d, a2d;
};

// Stencil of a field along one direction around (x, y, z); is2d fields have no z index.
struct FieldView {
  const double* __restrict__ f;
  const double* __restrict__ up[2];
  const double* __restrict__ dn[2];
  int ny, nz, is2d;
  __device__ __forceinline__ double at(const double* __restrict__ p, int x, int y, int z) const {
    const long long i2 = static_cast<long long>(x) * ny + y;
    return is2d ? __ldg(p + i2) : __ldg(p + i2 * nz + z);
  }
  __device__ __forceinline__ St sx(int x, int y, int z, bool wide) const {
    St s;
    s.c = at(f, x, y, z);
    s.m = at(f, x - 1, y, z);
    s.p = at(f, x + 1, y, z);
    s.mm = wide ? at(f, x - 2, y, z) : 0.0;
    s.pp = wide ? at(f, x + 2, y, z) : 0.0;
    return s;
  }
  __device__ __forceinline__ St sy(int x, int y, int z, bool wide) const {
    St s;
    s.c = at(f, x, y, z);
    s.m = at(dn[0], x, y - 1, z);
    s.p = at(up[0], x, y + 1, z);
    s.mm = wide ? at(dn[1], x, y - 2, z) : 0.0;
    s.pp = wide ? at(up[1], x, y + 2, z) : 0.0;
    return s;
  }
  __device__ __forceinline__ St sz(int x, int y, int z, bool wide) const {
    St s;
    s.c = at(f, x, y, z);
    s.m = at(f, x, y, wrapz(z - 1, nz));
    s.p = at(f, x, y, wrapz(z + 1, nz));
    int zmm = z - 2, zpp = z + 2; // nz may be 2 or 3: true modulo
    zmm = ((zmm % nz) + nz) % nz;
    zpp = zpp % nz;
    s.mm = wide ? at(f, x, y, zmm) : 0.0;
    s.pp = wide ? at(f, x, y, zpp) : 0.0;
    return s;
  }
};

__global__ void __launch_bounds__(256) k_b0xgrad_dot_grad(const B0xParams p) {
  const int nyb = p.ye - p.ys + 1;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.xe - p.xs + 1) * nyb * p.nz;
  if (t >= total) {
    return;
  }
  const int z = static_cast<int>(t % p.nz);
  const long long xy = t / p.nz;
  const int y = p.ys + static_cast<int>(xy % nyb);
  const int x = p.xs + static_cast<int>(xy / nyb);
  const long long i2 = static_cast<long long>(x) * p.ny + y;
  const bout_b200_b0xgrad& c = p.c;

  FieldView ph = {p.phi, {p.phi_up[0], p.phi_up[1]}, {p.phi_dn[0], p.phi_dn[1]}, p.ny, p.nz, p.phi2d};
  FieldView av = {p.A, {p.A_up[0], p.A_up[1]}, {p.A_dn[0], p.A_dn[1]}, p.ny, p.nz, p.a2d};
  const double rdx = 1.0 / __ldg(c.dx + i2), rdy = 1.0 / __ldg(c.dy + i2), rdz = 1.0 / __ldg(c.dz + i2);
  // derivatives of phi: index derivative / spacing (coordinates.cxx:1467-1514)
  const double dpdx = applyStandardRt(c.ddx_op, ph.sx(x, y, z, opGuards(c.ddx_op) == 2)) * rdx;
  const double dpdy = applyStandardRt(c.ddy_op, ph.sy(x, y, z, opGuards(c.ddy_op) == 2)) * rdy;
  const double dpdz = p.phi2d ? 0.0 : applyStandardRt(c.ddz_op, ph.sz(x, y, z, opGuards(c.ddz_op) == 2)) * rdz;
  const double g_22 = __ldg(c.g_22 + i2), g_23 = __ldg(c.g_23 + i2), g_12 = __ldg(c.g_12 + i2);
  // difops.cxx:580-588
  const double vx = g_22 * dpdz - g_23 * dpdy;
  const double vy = g_23 * dpdx - g_12 * dpdz;
  double vz = g_12 * dpdy - g_22 * dpdx;
  if (c.int_shift_torsion != nullptr) {
    vz += __ldg(c.int_shift_torsion + i2) * vx;
  }
  // difops.cxx:590-592: VDDX + VDDY + VDDZ, each index upwind / spacing, then / (J sqrt(g_22))
  double r = applyUpwindRt(c.vddx_op, vx, av.sx(x, y, z, opGuards(c.vddx_op) == 2)) * rdx;
  const double ry = applyUpwindRt(c.vddy_op, vy, av.sy(x, y, z, opGuards(c.vddy_op) == 2)) * rdy;
  const double rjs = 1.0 / (__ldg(c.J + i2) * sqrt(g_22)); // Field3D /= Field2D multiplies by the reciprocal
  if (p.yterm != nullptr) {
    p.yterm[i2 * p.nz + z] = ry * rjs;
  } else {
    r += ry;
  }
  if (!p.a2d) {
    r += applyUpwindRt(c.vddz_op, vz, av.sz(x, y, z, opGuards(c.vddz_op) == 2)) * rdz;
  }
  p.out[i2 * p.nz + z] = r * rjs;
}

// averageY, local half: xz[x][z] = mean over ystart..yend of f(x, y, z)
__global__ void __launch_bounds__(256) k_ymean(const double* __restrict__ f, double* __restrict__ xz, int nx,
                                               int ny, int nz, int ys, int ye) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(nx) * nz) {
    return;
  }
  const int z = static_cast<int>(t % nz), x = static_cast<int>(t / nz);
  const double* __restrict__ col = f + static_cast<long long>(x) * ny * nz + z;
  double acc = 0.0; // same left-to-right order as the reference's loop
  for (int y = ys; y <= ye; ++y) {
    acc += __ldg(col + static_cast<long long>(y) * nz);
  }
  xz[t] = acc / (ye - ys + 1);
}

// averageY, broadcast half: out(x, y, z) = scale * xz[x][z] for every y (guards included)
__global__ void __launch_bounds__(256) k_yspread(const double* __restrict__ xz, double scale,
                                                 double* __restrict__ out, int nx, int ny, int nz) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(nx) * ny * nz;
  if (t >= total) {
    return;
  }
  const int z = static_cast<int>(t % nz);
  const int x = static_cast<int>(t / (static_cast<long long>(ny) * nz));
  const double v = __ldg(xz + static_cast<long long>(x) * nz + z);
  out[t] = scale == 1.0 ? v : v * scale;
}

// smooth_x: 0.5 f + 0.25 (f(x-1) + f(x+1)) for 0 < x < nx-1, edge planes copied
__global__ void __launch_bounds__(256) k_smooth_x(const double* __restrict__ f, double* __restrict__ out, int nx,
                                                  long long plane) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= plane * nx) {
    return;
  }
  const int x = static_cast<int>(t / plane);
  const double c = __ldg(f + t);
  if (x == 0 || x == nx - 1) {
    out[t] = c;
  } else {
    out[t] = 0.5 * c + 0.25 * (__ldg(f + t - plane) + __ldg(f + t + plane));
  }
}

} // namespace b200

using namespace b200;

static bool firstOp(int op) { return op >= BOUT_B200_DDX_C2 && op <= BOUT_B200_DDX_W3; }
static bool upwindOp(int op) { return op >= BOUT_B200_VDDX_C2 && op <= BOUT_B200_VDDX_W3; }

extern "C" int bout_b200_b0xgrad_dot_grad(const bout_b200_grid* g, const bout_b200_b0xgrad* cfg,
                                          const double* phi, int phi_is2d, int phi_nslices,
                                          const double* const* phi_yup, const double* const* phi_ydown,
                                          const double* A, int a_is2d, int a_nslices,
                                          const double* const* a_yup, const double* const* a_ydown,
                                          double* out, double* yterm_out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(cfg && phi && A && out, BOUT_B200_ERR_ARG, "b0xGrad_dot_Grad: null argument");
  B200_REQUIRE(cfg->dx && cfg->dy && cfg->dz && cfg->J && cfg->g_22 && cfg->g_12 && cfg->g_23, BOUT_B200_ERR_ARG,
               "b0xGrad_dot_Grad: metric missing");
  B200_REQUIRE(firstOp(cfg->ddx_op) && firstOp(cfg->ddy_op) && firstOp(cfg->ddz_op), BOUT_B200_ERR_ARG,
               "b0xGrad_dot_Grad: ddx/ddy/ddz_op must be first-derivative functors");
  B200_REQUIRE(upwindOp(cfg->vddx_op) && upwindOp(cfg->vddy_op) && upwindOp(cfg->vddz_op), BOUT_B200_ERR_ARG,
               "b0xGrad_dot_Grad: vddx/vddy/vddz_op must be upwind functors");
  B200_REQUIRE(!(phi_is2d && a_is2d), BOUT_B200_ERR_UNSUPPORTED, "b0xGrad_dot_Grad: Field2D x Field2D has no 3D result");
  B200_REQUIRE(out != phi && out != A, BOUT_B200_ERR_ARG, "b0xGrad_dot_Grad: out must not alias an input");
  const int ngx = opGuards(cfg->ddx_op) > opGuards(cfg->vddx_op) ? opGuards(cfg->ddx_op) : opGuards(cfg->vddx_op);
  const int ngy_phi = opGuards(cfg->ddy_op), ngy_a = opGuards(cfg->vddy_op);
  const int ngy = ngy_phi > ngy_a ? ngy_phi : ngy_a;
  B200_REQUIRE(g->xstart >= ngx && g->nx - 1 - g->xend >= ngx, BOUT_B200_ERR_GUARDS,
               "b0xGrad_dot_Grad: needs %d x guard cells", ngx);
  B200_REQUIRE(g->ystart >= ngy && g->ny - 1 - g->yend >= ngy, BOUT_B200_ERR_GUARDS,
               "b0xGrad_dot_Grad: needs %d y guard cells", ngy);
  B200_REQUIRE(phi_nslices == 0 || (phi_nslices >= ngy_phi && phi_nslices <= 2 && phi_yup && phi_ydown),
               BOUT_B200_ERR_ARG, "b0xGrad_dot_Grad: phi needs %d parallel slices each way", ngy_phi);
  B200_REQUIRE(a_nslices == 0 || (a_nslices >= ngy_a && a_nslices <= 2 && a_yup && a_ydown), BOUT_B200_ERR_ARG,
               "b0xGrad_dot_Grad: A needs %d parallel slices each way", ngy_a);
  B0xParams p{};
  p.c = *cfg;
  p.phi = phi;
  p.A = A;
  for (int i = 0; i < 2; ++i) { // no slices: the field is its own neighbour source (identity transform, Field2D)
    p.phi_up[i] = (phi_nslices > i && !phi_is2d) ? phi_yup[i] : phi;
    p.phi_dn[i] = (phi_nslices > i && !phi_is2d) ? phi_ydown[i] : phi;
    p.A_up[i] = (a_nslices > i && !a_is2d) ? a_yup[i] : A;
    p.A_dn[i] = (a_nslices > i && !a_is2d) ? a_ydown[i] : A;
  }
  p.out = out;
  p.yterm = yterm_out;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xs = g->xstart;
  p.xe = g->xend;
  p.ys = g->ystart;
  p.ye = g->yend;
  p.phi2d = phi_is2d;
  p.a2d = a_is2d;
  const long long total = static_cast<long long>(p.xe - p.xs + 1) * (p.ye - p.ys + 1) * p.nz;
  k_b0xgrad_dot_grad<<<static_cast<unsigned>((total + 255) / 256), 256, 0, asStream(stream)>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_ymean(const bout_b200_grid* g, const double* f, double* xz, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && xz, BOUT_B200_ERR_ARG, "ymean: null argument");
  const long long n = static_cast<long long>(g->nx) * g->nz;
  k_ymean<<<static_cast<unsigned>((n + 255) / 256), 256, 0, asStream(stream)>>>(f, xz, g->nx, g->ny, g->nz,
                                                                             g->ystart, g->yend);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_yspread(const bout_b200_grid* g, const double* xz, double scale, double* out,
                                 void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(xz && out, BOUT_B200_ERR_ARG, "yspread: null argument");
  const long long n = static_cast<long long>(g->nx) * g->ny * g->nz;
  k_yspread<<<static_cast<unsigned>((n + 255) / 256), 256, 0, asStream(stream)>>>(xz, scale, out, g->nx, g->ny,
                                                                               g->nz);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_smooth_x(const bout_b200_grid* g, const double* f, double* out, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f && out && f != out, BOUT_B200_ERR_ARG, "smooth_x: null or aliased argument");
  const long long plane = static_cast<long long>(g->ny) * g->nz;
  const long long n = plane * g->nx;
  k_smooth_x<<<static_cast<unsigned>((n + 255) / 256), 256, 0, asStream(stream)>>>(f, out, g->nx, plane);
  B200_LAUNCH_OK();
  return 0;
}
