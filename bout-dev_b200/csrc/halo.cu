// Guard-cell exchange building blocks (subsystem a13, Mesh::communicate).
// Reference: BoutMesh::send/sendX/sendY/wait and pack_data/unpack_data
// (src/mesh/impls/bout/boutmesh.cxx:1221-1555, :2400-2466).
//
//  * bout_b200_halo_self   : the exchange a rank does with ITSELF when it is its own neighbour
//    (periodic Y on one rank in y -- closed field lines with NYPE = 1 -- and periodicX with
//    NXPE = 1).  One kernel moves the guards of up to 8 fields; Y first, then X including the
//    y-guard corners, the order Mesh::communicate uses with include_corner_cells
//    (src/mesh/mesh.cxx:351-362).
//  * bout_b200_halo_pack_y / unpack_y : strided y-guard rows <-> a contiguous message buffer laid
//    out [field][x][y][z] as pack_data does.  X guard planes are already contiguous in memory
//    (x is the slowest index), so X messages are sent straight from / received straight into the
//    field with no pack step.
#include "common.cuh"

namespace b200 {

constexpr int kMaxHaloFields = 8;

struct HaloSelfParams {
  double* f[kMaxHaloFields];
  int nf;
  int nx, ny, nz, xstart, xend, ystart, yend;
  int dir; // 0 = X periodic wrap, 1 = Y periodic wrap
};

// dir = 1: guard row y < ystart  <- interior row y + (yend - ystart + 1); y > yend <- y - ny_int
// dir = 0: same along x, for all y (corners included)
__global__ void __launch_bounds__(256) k_halo_self(const HaloSelfParams p) {
  const int mg_lo = p.dir == 0 ? p.xstart : p.ystart;
  const int mg_hi = p.dir == 0 ? p.nx - p.xend - 1 : p.ny - p.yend - 1;
  const int nguard = mg_lo + mg_hi;
  const int nint = p.dir == 0 ? (p.xend - p.xstart + 1) : (p.yend - p.ystart + 1);
  const int northo = p.dir == 0 ? p.ny : p.nx; // the other (x or y) extent: everything
  const long long per_field = static_cast<long long>(nguard) * northo * p.nz;
  const long long total = per_field * p.nf;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int fi = static_cast<int>(t / per_field);
    long long r = t % per_field;
    const int z = static_cast<int>(r % p.nz);
    r /= p.nz;
    const int o = static_cast<int>(r % northo);
    const int gi = static_cast<int>(r / northo); // 0..nguard-1
    const int m_dst = gi < mg_lo ? gi : (p.dir == 0 ? p.xend : p.yend) + 1 + (gi - mg_lo);
    const int m_src = gi < mg_lo ? m_dst + nint : m_dst - nint;
    const int xd = p.dir == 0 ? m_dst : o, yd = p.dir == 0 ? o : m_dst;
    const int xs = p.dir == 0 ? m_src : o, ys = p.dir == 0 ? o : m_src;
    double* f = p.f[fi];
    f[(static_cast<long long>(xd) * p.ny + yd) * p.nz + z] = f[(static_cast<long long>(xs) * p.ny + ys) * p.nz + z];
  }
}

struct PackYParams {
  double* f[kMaxHaloFields];
  double* buf;
  int nf, nx, ny, nz;
  int y0, nrows; // rows y0 .. y0+nrows-1
  int xlo, xhi;
  int unpack;
};

__global__ void __launch_bounds__(256) k_pack_y(const PackYParams p) {
  const int nxb = p.xhi - p.xlo + 1;
  const long long per_field = static_cast<long long>(nxb) * p.nrows * p.nz;
  const long long total = per_field * p.nf;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int fi = static_cast<int>(t / per_field);
    long long r = t % per_field;
    const int z = static_cast<int>(r % p.nz);
    r /= p.nz;
    const int yy = static_cast<int>(r % p.nrows);
    const int x = p.xlo + static_cast<int>(r / p.nrows);
    double* cell = p.f[fi] + (static_cast<long long>(x) * p.ny + (p.y0 + yy)) * p.nz + z;
    if (p.unpack) {
      *cell = p.buf[t];
    } else {
      p.buf[t] = *cell;
    }
  }
}

} // namespace b200

using namespace b200;

extern "C" int bout_b200_halo_self(const bout_b200_grid* g, int nfields, double* const* fields,
                                   int periodic_x, int periodic_y, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(fields != nullptr && nfields >= 1 && nfields <= kMaxHaloFields, BOUT_B200_ERR_ARG,
               "halo_self: 1..%d fields", kMaxHaloFields);
  HaloSelfParams p{};
  for (int i = 0; i < nfields; ++i) {
    B200_REQUIRE(fields[i] != nullptr, BOUT_B200_ERR_ARG, "halo_self: null field");
    p.f[i] = fields[i];
  }
  p.nf = nfields;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.xstart = g->xstart;
  p.xend = g->xend;
  p.ystart = g->ystart;
  p.yend = g->yend;
  for (int dir = 1; dir >= 0; --dir) { // Y first, then X with corners
    if (!(dir == 1 ? periodic_y : periodic_x)) {
      continue;
    }
    const int nguard = dir == 0 ? g->xstart + (g->nx - g->xend - 1) : g->ystart + (g->ny - g->yend - 1);
    if (nguard == 0) {
      continue;
    }
    const int nint = dir == 0 ? (g->xend - g->xstart + 1) : (g->yend - g->ystart + 1);
    B200_REQUIRE(nint >= (dir == 0 ? g->xstart : g->ystart), BOUT_B200_ERR_ARG,
                 "halo_self: fewer interior points than guard cells");
    p.dir = dir;
    const long long total = static_cast<long long>(nguard) * (dir == 0 ? g->ny : g->nx) * g->nz * nfields;
    const long long want = (total + 255) / 256;
    const unsigned blocks = static_cast<unsigned>(want < 148LL * 32 ? want : 148LL * 32);
    k_halo_self<<<blocks, 256, 0, asStream(stream)>>>(p);
    B200_LAUNCH_OK();
  }
  return 0;
}

static int packY(const bout_b200_grid* g, int nfields, double* const* fields, int y0, int nrows,
                 int xlo, int xhi, double* buf, int unpack, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(fields && buf && nfields >= 1 && nfields <= kMaxHaloFields, BOUT_B200_ERR_ARG,
               "halo_pack_y: bad argument");
  B200_REQUIRE(y0 >= 0 && nrows >= 1 && y0 + nrows <= g->ny && xlo >= 0 && xhi < g->nx && xlo <= xhi,
               BOUT_B200_ERR_ARG, "halo_pack_y: bad range");
  PackYParams p{};
  for (int i = 0; i < nfields; ++i) {
    p.f[i] = fields[i];
  }
  p.buf = buf;
  p.nf = nfields;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.y0 = y0;
  p.nrows = nrows;
  p.xlo = xlo;
  p.xhi = xhi;
  p.unpack = unpack;
  const long long total = static_cast<long long>(xhi - xlo + 1) * nrows * g->nz * nfields;
  const long long want = (total + 255) / 256;
  const unsigned blocks = static_cast<unsigned>(want < 148LL * 32 ? want : 148LL * 32);
  k_pack_y<<<blocks, 256, 0, asStream(stream)>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int bout_b200_halo_pack_y(const bout_b200_grid* g, int nfields, double* const* fields,
                                     int y0, int nrows, int xlo, int xhi, double* buf, void* stream) {
  return packY(g, nfields, fields, y0, nrows, xlo, xhi, buf, 0, stream);
}
extern "C" int bout_b200_halo_unpack_y(const bout_b200_grid* g, int nfields, double* const* fields,
                                       int y0, int nrows, int xlo, int xhi, const double* buf,
                                       void* stream) {
  return packY(g, nfields, fields, y0, nrows, xlo, xhi, const_cast<double*>(buf), 1, stream);
}
