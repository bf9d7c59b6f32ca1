// Boundary conditions on guard cells (subsystem a15) for CELL_CENTRE Field3D with a constant value.
// Reference: BoundaryDirichlet::apply standard case (src/mesh/boundary_standard.cxx:527-581):
//   first guard = 2*val - f_inside, remaining guards = val;
// BoundaryNeumann::apply CELL_CENTRE (:2167-2190): g1 = f_in + delta*val and, for width 2,
//   g2 = f_in2 + 3*delta*val with delta = bx*dx + by*dy at the first guard cell;
// apply_ddt (:591-600, :2201-2210): time derivative set to zero in the boundary cells;
// BoundaryFree_O2::apply / BoundaryFree_O3::apply standard case (:3281-3293, :3545-3558): every guard cell, outwards,
//   g = 2 f(-1) - f(-2)  /  g = 3 f(-1) - 3 f(-2) + f(-3), f(-k) being k cells inwards along the boundary normal.
// X boundaries cover y = ystart..yend (src/mesh/impls/bout/boutmesh.cxx:1028-1037); Y boundaries
// cover x = xlo..xhi as given (corner cells included by default, boutmesh.cxx:1057-1059).
#include "common.cuh"

namespace b200 {

struct BndryParams {
  double* __restrict__ f;
  const double* __restrict__ spacing; // dx (x sides) or dy (y sides), Field2D; only for Neumann
  double val;
  int nx, ny, nz;
  int kind; // 0 dirichlet, 1 neumann, 2 zero (apply_ddt), 3 extend (outer guards copy their inner neighbour),
            // 4 free_o2, 5 free_o3 (extrapolation)
  int side; // 0 x-inner, 1 x-outer, 2 y-lower, 3 y-upper
  int width;
  int lo, hi; // range of the other index (y for x sides, x for y sides), inclusive
  int first;  // first guard index along the boundary normal
  int b;      // outward direction (-1 or +1)
};

__device__ __forceinline__ void boundaryCell(const BndryParams& p, long long t) {
  const long long n = static_cast<long long>(p.hi - p.lo + 1) * p.nz;
  if (t >= n) {
    return;
  }
  const int z = static_cast<int>(t % p.nz);
  const int o = p.lo + static_cast<int>(t / p.nz);
  const bool xside = p.side < 2;
  auto idx3 = [&](int normal) -> long long {
    const int x = xside ? normal : o;
    const int y = xside ? o : normal;
    return (static_cast<long long>(x) * p.ny + y) * p.nz + z;
  };
  if (p.kind == 2) {
    for (int i = 0; i < p.width; ++i) {
      p.f[idx3(p.first + i * p.b)] = 0.0;
    }
  } else if (p.kind == 3) { // examples/elm-pb/elm_pb.cxx:1498-1516: phi(i) = phi(i+-1) beyond the first guard
    for (int i = 1; i < p.width; ++i) {
      p.f[idx3(p.first + i * p.b)] = p.f[idx3(p.first + (i - 1) * p.b)];
    }
  } else if (p.kind == 4) {
    for (int i = 0; i < p.width; ++i) {
      const int n0 = p.first + i * p.b;
      p.f[idx3(n0)] = 2 * p.f[idx3(n0 - p.b)] - p.f[idx3(n0 - 2 * p.b)];
    }
  } else if (p.kind == 5) {
    for (int i = 0; i < p.width; ++i) {
      const int n0 = p.first + i * p.b;
      p.f[idx3(n0)] = 3.0 * p.f[idx3(n0 - p.b)] - 3.0 * p.f[idx3(n0 - 2 * p.b)] + p.f[idx3(n0 - 3 * p.b)];
    }
  } else if (p.kind == 0) {
    p.f[idx3(p.first)] = 2 * p.val - p.f[idx3(p.first - p.b)];
    for (int i = 1; i < p.width; ++i) {
      p.f[idx3(p.first + i * p.b)] = p.val;
    }
  } else {
    const int x = xside ? p.first : o;
    const int y = xside ? o : p.first;
    const double delta = p.b * __ldg(p.spacing + static_cast<long long>(x) * p.ny + y);
    p.f[idx3(p.first)] = p.f[idx3(p.first - p.b)] + delta * p.val;
    if (p.width == 2) {
      p.f[idx3(p.first + p.b)] = p.f[idx3(p.first - 2 * p.b)] + 3.0 * delta * p.val;
    }
  }
}

__global__ void __launch_bounds__(256) k_boundary(const BndryParams p) {
  boundaryCell(p, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x);
}

// Several (field, side, condition) triples in ONE launch: blockIdx.y picks the entry.  The entries are reached
// through a switch, not by indexing the parameter array with a run-time value (that would copy the whole struct to
// local memory in every thread).  Entries do not depend on each other: a caller puts at most one entry per (field,
// side) into a batch and keeps x sides (rows ystart..yend) apart from y sides that include their corner cells.
constexpr int kMaxBndryBatch = 8;
struct BndryBatch {
  BndryParams s[kMaxBndryBatch];
};
__global__ void __launch_bounds__(256) k_boundary_batch(const BndryBatch b) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  switch (blockIdx.y) {
  case 0: boundaryCell(b.s[0], t); break;
  case 1: boundaryCell(b.s[1], t); break;
  case 2: boundaryCell(b.s[2], t); break;
  case 3: boundaryCell(b.s[3], t); break;
  case 4: boundaryCell(b.s[4], t); break;
  case 5: boundaryCell(b.s[5], t); break;
  case 6: boundaryCell(b.s[6], t); break;
  default: boundaryCell(b.s[7], t); break;
  }
}

} // namespace b200

using namespace b200;

// fills p for one (field, side, condition); returns 1 when there is nothing to do on that side
static int boundarySetup(const bout_b200_grid* g, int kind, int side, double value, const double* spacing2d, int lo,
                         int hi, double* f, BndryParams& p) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(f != nullptr && kind >= 0 && kind <= 5 && side >= 0 && side <= 3, BOUT_B200_ERR_ARG,
               "boundary: bad argument");
  B200_REQUIRE(kind != 1 || spacing2d != nullptr, BOUT_B200_ERR_ARG, "boundary: neumann needs the spacing");
  p = BndryParams{};
  p.f = f;
  p.spacing = spacing2d;
  p.val = value;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.kind = kind;
  p.side = side;
  switch (side) {
  case 0: p.width = g->xstart; p.first = g->xstart - 1; p.b = -1; break;             // BoundaryRegionXIn
  case 1: p.width = g->nx - g->xend - 1; p.first = g->nx - p.width; p.b = 1; break; // BoundaryRegionXOut
  case 2: p.width = g->ystart; p.first = g->ystart - 1; p.b = -1; break;             // BoundaryRegionYDown
  default: p.width = g->ny - g->yend - 1; p.first = g->ny - p.width; p.b = 1; break; // BoundaryRegionYUp
  }
  if (p.width <= 0) {
    return 1; // no guard cells on this side
  }
  {
    const int interior = side < 2 ? g->xend - g->xstart + 1 : g->yend - g->ystart + 1;
    B200_REQUIRE(kind < 4 || interior >= (kind == 4 ? 2 : 3), BOUT_B200_ERR_GUARDS,
                 "boundary: free_o%d needs %d interior cells along the normal", kind - 2, kind - 2);
  }
  if (lo < 0 || hi < 0) { // default ranges
    lo = side < 2 ? g->ystart : 0;
    hi = side < 2 ? g->yend : g->nx - 1;
  }
  p.lo = lo;
  p.hi = hi;
  return 0;
}

extern "C" int bout_b200_boundary_batch(const bout_b200_grid* g, int n, const bout_b200_bndry_spec* specs, void* stream) {
  B200_REQUIRE(specs != nullptr && n >= 1, BOUT_B200_ERR_ARG, "boundary_batch: bad argument");
  for (int i0 = 0; i0 < n; i0 += kMaxBndryBatch) {
    BndryBatch b{};
    int m = 0;
    long long longest = 0;
    for (int i = i0; i < n && i < i0 + kMaxBndryBatch; ++i) {
      const bout_b200_bndry_spec& q = specs[i];
      const int rc = boundarySetup(g, q.kind, q.side, q.value, q.spacing2d, q.lo, q.hi, q.f, b.s[m]);
      if (rc < 0) {
        return rc;
      }
      if (rc == 0) {
        const long long cells = static_cast<long long>(b.s[m].hi - b.s[m].lo + 1) * g->nz;
        longest = cells > longest ? cells : longest;
        ++m;
      }
    }
    if (m == 0) {
      continue;
    }
    for (int i = m; i < kMaxBndryBatch; ++i) { // unused entries: empty ranges
      b.s[i] = b.s[0];
      b.s[i].hi = b.s[i].lo - 1;
    }
    const dim3 grid(static_cast<unsigned>((longest + 255) / 256), static_cast<unsigned>(m));
    k_boundary_batch<<<grid, 256, 0, asStream(stream)>>>(b);
    B200_LAUNCH_OK();
  }
  return 0;
}

extern "C" int bout_b200_boundary(const bout_b200_grid* g, int kind, int side, double value,
                                  const double* spacing2d, int lo, int hi, double* f, void* stream) {
  BndryParams p{};
  const int rc = boundarySetup(g, kind, side, value, spacing2d, lo, hi, f, p);
  if (rc != 0) {
    return rc < 0 ? rc : 0;
  }
  const long long n = static_cast<long long>(p.hi - p.lo + 1) * g->nz;
  k_boundary<<<static_cast<unsigned>((n + 255) / 256), 256, 0, asStream(stream)>>>(p);
  B200_LAUNCH_OK();
  return 0;
}
