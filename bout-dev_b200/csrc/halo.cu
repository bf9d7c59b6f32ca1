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
//  * bout_b200_peer_* / bout_b200_halo_push_y / bout_b200_halo_wait : the same Y exchange WITHOUT a
//    communication library -- the pack kernel stores the rows straight into the neighbour GPU's receive
//    buffer over NVLink (CUDA IPC peer mapping) and its last block raises a flag in the neighbour's memory
//    (writes -> __threadfence_system -> block counter -> flag, the threadfence-reduction pattern); the
//    receiver spins on its own flag, then unpacks.  One hop per message, no rendezvous kernel on either
//    side; receive buffers are double-buffered by the parity of the exchange counter, which is sufficient
//    because a rank cannot start exchange c + 1 before it has waited for both neighbours' exchange c.
#include <cstring>

#include "common.cuh"

#include <cstdlib>

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
// One z row of one guard cell line per group of threads: blockIdx.y = field, the row index decoded once per thread
// with 32-bit arithmetic, VEC doubles per access (the element-per-thread form with 64-bit index chains and a run-time
// index into the pointer array -- a local-memory copy of the parameters -- took 11 us for 4 MB).
template <int VEC>
__global__ void __launch_bounds__(256) k_halo_self(const HaloSelfParams p) {
  const int mg_lo = p.dir == 0 ? p.xstart : p.ystart;
  const int mg_hi = p.dir == 0 ? p.nx - p.xend - 1 : p.ny - p.yend - 1;
  const int nguard = mg_lo + mg_hi;
  const int nint = p.dir == 0 ? (p.xend - p.xstart + 1) : (p.yend - p.ystart + 1);
  const int northo = p.dir == 0 ? p.ny : p.nx; // the other (x or y) extent: everything
  const int nzv = p.nz / VEC;
  const int fi = blockIdx.y;
  double* __restrict__ f = fi == 0 ? p.f[0] : fi == 1 ? p.f[1] : fi == 2 ? p.f[2] : fi == 3 ? p.f[3]
                           : fi == 4 ? p.f[4] : fi == 5 ? p.f[5] : fi == 6 ? p.f[6] : p.f[7];
  const unsigned total = static_cast<unsigned>(nguard) * northo * nzv; // per field (< 2^32: checked by the host)
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int zv = static_cast<int>(t % nzv);
    const unsigned r = t / nzv;
    const int o = static_cast<int>(r % northo);
    const int gi = static_cast<int>(r / northo); // 0..nguard-1
    const int m_dst = gi < mg_lo ? gi : (p.dir == 0 ? p.xend : p.yend) + 1 + (gi - mg_lo);
    const int m_src = gi < mg_lo ? m_dst + nint : m_dst - nint;
    const int xd = p.dir == 0 ? m_dst : o, yd = p.dir == 0 ? o : m_dst;
    const int xs = p.dir == 0 ? m_src : o, ys = p.dir == 0 ? o : m_src;
    const long long dst = (static_cast<long long>(xd) * p.ny + yd) * p.nz + zv * VEC;
    const long long src = (static_cast<long long>(xs) * p.ny + ys) * p.nz + zv * VEC;
    if constexpr (VEC == 2) {
      *reinterpret_cast<double2*>(f + dst) = *reinterpret_cast<const double2*>(f + src);
    } else {
      f[dst] = f[src];
    }
  }
}

struct PackYParams {
  double* f[kMaxHaloFields];
  double* buf;
  int nf, nx, ny, nz;
  int y0, nrows; // rows y0 .. y0+nrows-1
  int xlo, xhi;
  int unpack;
  // push mode (buf is peer memory): the last block to finish stores `value` to *flag
  unsigned long long* flag;
  unsigned long long value;
  unsigned int* done; // local block counter, reset by the last block
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
  if (p.flag != nullptr) {
    __threadfence_system(); // this thread's stores to the peer buffer are visible system-wide
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(p.done, 1u);
      if (prev == gridDim.x - 1) { // every block has fenced its stores
        *p.done = 0;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.flag), "l"(p.value) : "memory");
      }
    }
  }
}

// Both directions of one Y exchange in ONE launch each way (blockIdx.y = direction), 16-byte transactions.
struct PushY2Params {
  double* f[kMaxHaloFields];
  double* buf[2];               // peer receive buffers (push) / local receive buffers (unpack); NULL = no neighbour
  unsigned long long* flag[2];  // peer flags (push) / local flags (unpack)
  unsigned int* done;           // two local block counters (push)
  int* status;                  // time-out word (unpack)
  long long timeout;            // cycles the receiver waits for its neighbour before giving up (unpack)
  unsigned long long value;
  int y0[2];
  int nf, nx, ny, nz, nrows;
};

template <int VEC, bool UNPACK>
__global__ void __launch_bounds__(256) k_xchg_y2(const PushY2Params p) {
  const int d = blockIdx.y;
  if (p.buf[d] == nullptr) {
    return;
  }
  if (UNPACK) { // wait for the neighbour's rows to have landed in the local receive buffer
    __shared__ int ok;
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      unsigned long long v;
      ok = 0;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p.flag[d]) : "memory");
        if (v >= p.value) {
          ok = 1;
          break;
        }
        __nanosleep(100);
      } while (clock64() - t0 < p.timeout);
      if (!ok) {
        *p.status = 1;
      }
    }
    __syncthreads();
    if (!ok) {
      // The neighbour never signalled: the receive buffer holds stale or partial rows.  Nothing is unpacked, the
      // status word says why, and the trap makes every later call on this context fail (BOUT_B200_ERR_CUDA) --
      // a run must not continue on guard cells that never arrived.
      __threadfence_system();
      __trap();
    }
  }
  const int nzv = p.nz / VEC;
  const long long per_field = static_cast<long long>(p.nx) * p.nrows * nzv;
  const long long total = per_field * p.nf;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int fi = static_cast<int>(t / per_field);
    long long r = t % per_field;
    const int zv = static_cast<int>(r % nzv);
    r /= nzv;
    const int yy = static_cast<int>(r % p.nrows);
    const int x = static_cast<int>(r / p.nrows);
    double* cell = p.f[fi] + (static_cast<long long>(x) * p.ny + (p.y0[d] + yy)) * p.nz + zv * VEC;
    double* slot = p.buf[d] + t * VEC;
    if (VEC == 2) {
      if (UNPACK) {
        *reinterpret_cast<double2*>(cell) = *reinterpret_cast<const double2*>(slot);
      } else {
        *reinterpret_cast<double2*>(slot) = *reinterpret_cast<const double2*>(cell);
      }
    } else {
      if (UNPACK) {
        *cell = *slot;
      } else {
        *slot = *cell;
      }
    }
  }
  if (!UNPACK) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(p.done + d, 1u);
      if (prev == gridDim.x - 1) {
        p.done[d] = 0;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.flag[d]), "l"(p.value) : "memory");
      }
    }
  }
}

// Receiver side: spin until the neighbour's flag reaches `value` (bounded; then status = 1 and the context traps).
__global__ void k_flag_wait(const unsigned long long* flag, unsigned long long value, int* status, long long timeout) {
  const long long t0 = clock64();
  unsigned long long v;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= value) {
      return;
    }
    __nanosleep(200);
  } while (clock64() - t0 < timeout);
  *status = 1;
  __threadfence_system();
  __trap();
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
    bool vec2 = g->nz % 2 == 0;
    for (int i = 0; i < nfields; ++i) {
      vec2 = vec2 && aligned16(fields[i]);
    }
    const long long total = static_cast<long long>(nguard) * (dir == 0 ? g->ny : g->nx) * (vec2 ? g->nz / 2 : g->nz);
    B200_REQUIRE(total < (1LL << 31), BOUT_B200_ERR_UNSUPPORTED, "halo_self: guard region too large");
    const long long want = (total + 255) / 256;
    const dim3 grid(static_cast<unsigned>(want < 148LL * 16 ? want : 148LL * 16), static_cast<unsigned>(nfields));
    if (vec2) {
      k_halo_self<2><<<grid, 256, 0, asStream(stream)>>>(p);
    } else {
      k_halo_self<1><<<grid, 256, 0, asStream(stream)>>>(p);
    }
    B200_LAUNCH_OK();
  }
  return 0;
}

static int packY(const bout_b200_grid* g, int nfields, double* const* fields, int y0, int nrows,
                 int xlo, int xhi, double* buf, int unpack, void* stream, unsigned long long* flag = nullptr,
                 unsigned long long value = 0, unsigned int* done = nullptr) {
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
  p.flag = flag;
  p.value = value;
  p.done = done;
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

// ---- peer-memory exchange -------------------------------------------------------------------------------
extern "C" int bout_b200_peer_alloc(size_t bytes, void** ptr) {
  B200_REQUIRE(ptr != nullptr && bytes > 0, BOUT_B200_ERR_ARG, "peer_alloc: bad argument");
  B200_CUDA_OK(cudaMalloc(ptr, bytes));
  B200_CUDA_OK(cudaMemset(*ptr, 0, bytes));
  B200_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}
extern "C" int bout_b200_peer_free(void* ptr) {
  B200_CUDA_OK(cudaFree(ptr));
  return 0;
}
extern "C" int bout_b200_peer_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  B200_REQUIRE(ptr && handle64, BOUT_B200_ERR_ARG, "peer_export: null argument");
  cudaIpcMemHandle_t h;
  B200_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
extern "C" int bout_b200_peer_open(const unsigned char* handle64, void** ptr) {
  B200_REQUIRE(ptr && handle64, BOUT_B200_ERR_ARG, "peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  B200_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int bout_b200_peer_close(void* ptr) {
  B200_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return 0;
}
extern "C" int bout_b200_halo_push_y(const bout_b200_grid* g, int nfields, double* const* fields, int y0,
                                     int nrows, int xlo, int xhi, double* peer_buf,
                                     unsigned long long* peer_flag, unsigned long long value,
                                     unsigned int* local_counter, void* stream) {
  B200_REQUIRE(peer_flag && local_counter, BOUT_B200_ERR_ARG, "halo_push_y: null flag");
  return packY(g, nfields, fields, y0, nrows, xlo, xhi, peer_buf, 0, stream, peer_flag, value, local_counter);
}
extern "C" int bout_b200_halo_wait(const unsigned long long* local_flag, unsigned long long value, int* status,
                                   void* stream) {
  B200_REQUIRE(local_flag && status, BOUT_B200_ERR_ARG, "halo_wait: null argument");
  k_flag_wait<<<1, 1, 0, asStream(stream)>>>(local_flag, value, status, haloTimeoutCycles());
  B200_LAUNCH_OK();
  return 0;
}

static int xchgY2(const bout_b200_grid* g, int nfields, double* const* fields, int nrows, const int* y0,
                  double* const* buf, unsigned long long* const* flag, unsigned long long value,
                  unsigned int* counters, int* status, bool unpack, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(fields && y0 && buf && flag && nfields >= 1 && nfields <= kMaxHaloFields && nrows >= 1,
               BOUT_B200_ERR_ARG, "halo exchange: bad argument");
  PushY2Params p{};
  bool vec2 = g->nz % 2 == 0;
  for (int i = 0; i < nfields; ++i) {
    B200_REQUIRE(fields[i] != nullptr, BOUT_B200_ERR_ARG, "halo exchange: null field");
    p.f[i] = fields[i];
    vec2 = vec2 && aligned16(fields[i]);
  }
  for (int d = 0; d < 2; ++d) {
    p.buf[d] = buf[d];
    p.flag[d] = flag[d];
    p.y0[d] = y0[d];
    B200_REQUIRE(buf[d] == nullptr || (flag[d] != nullptr && y0[d] >= 0 && y0[d] + nrows <= g->ny), BOUT_B200_ERR_ARG,
                 "halo exchange: bad rows or missing flag");
    vec2 = vec2 && (buf[d] == nullptr || aligned16(buf[d]));
  }
  p.done = counters;
  p.status = status;
  p.timeout = haloTimeoutCycles();
  p.value = value;
  p.nf = nfields;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  p.nrows = nrows;
  const long long total = static_cast<long long>(g->nx) * nrows * (vec2 ? g->nz / 2 : g->nz) * nfields;
  const long long want = (total + 255) / 256;
  const dim3 grid(static_cast<unsigned>(want < 148LL * 4 ? want : 148LL * 4), 2);
  cudaStream_t s = asStream(stream);
  if (unpack) {
    if (vec2) {
      k_xchg_y2<2, true><<<grid, 256, 0, s>>>(p);
    } else {
      k_xchg_y2<1, true><<<grid, 256, 0, s>>>(p);
    }
  } else {
    if (vec2) {
      k_xchg_y2<2, false><<<grid, 256, 0, s>>>(p);
    } else {
      k_xchg_y2<1, false><<<grid, 256, 0, s>>>(p);
    }
  }
  B200_LAUNCH_OK();
  return 0;
}
extern "C" int bout_b200_halo_push_y2(const bout_b200_grid* g, int nfields, double* const* fields, int nrows,
                                      const int* y0, double* const* peer_buf, unsigned long long* const* peer_flag,
                                      unsigned long long value, unsigned int* local_counters, void* stream) {
  B200_REQUIRE(local_counters != nullptr, BOUT_B200_ERR_ARG, "halo_push_y2: null counters");
  return xchgY2(g, nfields, fields, nrows, y0, peer_buf, peer_flag, value, local_counters, nullptr, false, stream);
}
extern "C" int bout_b200_halo_wait_unpack_y2(const bout_b200_grid* g, int nfields, double* const* fields, int nrows,
                                             const int* y0, double* const* local_buf,
                                             unsigned long long* const* local_flag, unsigned long long value,
                                             int* status, void* stream) {
  B200_REQUIRE(status != nullptr, BOUT_B200_ERR_ARG, "halo_wait_unpack_y2: null status");
  return xchgY2(g, nfields, fields, nrows, y0, local_buf, local_flag, value, nullptr, status, true, stream);
}
