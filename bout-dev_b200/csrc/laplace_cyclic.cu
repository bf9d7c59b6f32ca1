// LaplaceCyclic (subsystem 3): batched z-FFT in shared memory -> batched complex tridiagonal solve
// per (y, kz) with coefficients regenerated on chip -> inverse z-FFT.
//
// Replaces LaplaceCyclic::solve(Field3D, Field3D)
// (reference src/invert/laplace/impls/cyclic/cyclic_laplace.cxx:187-290), FFTTransform::forward /
// backward (src/invert/laplace/common_transform.cxx:29-115), Laplacian::tridagMatrix
// (src/invert/laplace/invert_laplace.cxx:416-768) and the single-rank path of CyclicReduce
// (include/bout/cyclic_reduction.hxx:169-637).
//
// Data layout: the spectral work array is [x - xs][y - ys][kz] complex (kz fastest), i.e. the
// natural output order of the row FFTs.  The reference transposes to [y*nmode+kz][x] so that each
// system is contiguous for its serial CPU solver; on the GPU one thread owns one (y,kz) system and
// walks x, so at every x the 32 lanes of a warp read 32 consecutive kz -- fully coalesced with no
// transpose at all.
//
// Tridiagonal algorithm: Thomas (forward elimination storing c', d'; back substitution).  For
// periodic X the cyclic system is solved with the Sherman-Morrison correction (two right-hand
// sides through the same sweeps).  The reference reduces to a 2x2 interface system first
// (needed for its MPI decomposition); both are direct solvers for the same matrix and agree to
// round-off (parity tolerance for this path: 1e-10 relative).
#include "laplace_coefs.cuh"

namespace b200 {

struct LapParams {
  LapCoefs m;
  LapBndry bd;
  const double* __restrict__ rhs;
  const double* __restrict__ x0;
  double* __restrict__ out;
  cd* __restrict__ spec;  // [nxs][nys][nmode]
  cd* __restrict__ cprime; // [nxs][nys][nmode]
  cd* __restrict__ qwork;  // periodic only: [nxs][nys][nmode]
  int* __restrict__ status;
  const cd* __restrict__ tw;
  int nx, ny, nz;
  int ys, ye;     // y range solved
  int nmode;
  int zero_dc;
  int in_set, out_set;
  int tpr, rows, pow2;
};

__device__ __forceinline__ RowFFT makeRow(const LapParams& p, unsigned char* smem_raw, int& rloc) {
  const int nbuf = fftBufElems(p.nz);
  rloc = threadIdx.x / p.tpr;
  RowFFT r;
  r.nz = p.nz;
  r.N = p.nz / 2;
  r.pow2 = p.pow2 != 0;
  r.tpr = p.tpr;
  r.lane = threadIdx.x % p.tpr;
  r.A = reinterpret_cast<cd*>(smem_raw) + static_cast<size_t>(rloc) * 2 * nbuf;
  r.B = r.A + nbuf;
  r.tw = p.tw;
  return r;
}

__device__ __forceinline__ void loadRealRow(const RowFFT& r, const double* __restrict__ src) {
  double* a = reinterpret_cast<double*>(r.A);
  if (r.nz % 2 == 0 && aligned16(src)) {
    for (int j = r.lane; j < r.nz / 2; j += r.tpr) {
      reinterpret_cast<double2*>(a)[j] = __ldg(reinterpret_cast<const double2*>(src) + j);
    }
  } else {
    for (int j = r.lane; j < r.nz; j += r.tpr) {
      a[j] = __ldg(src + j);
    }
  }
}

// K1: rfft of every (x,y) row of rhs (or x0 in boundary rows when INVERT_SET) -> spec
__global__ void __launch_bounds__(128) k_lap_forward(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int rloc;
  const RowFFT r = makeRow(p, smem_raw, rloc);
  const int nys = p.ye - p.ys + 1;
  const int nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  const long long rowid = static_cast<long long>(blockIdx.x) * p.rows + rloc;
  const bool active = rowid < nrows;
  const long long rid = active ? rowid : nrows - 1;
  const int ixl = static_cast<int>(rid / nys), iyl = static_cast<int>(rid % nys);
  const int x = p.bd.xs + ixl, y = p.ys + iyl;
  // common_transform.cxx:49-57
  const bool from_x0 = ((x < p.bd.inbndry) && p.in_set) || ((p.nx - x - 1 < p.bd.outbndry) && p.out_set);
  const double* __restrict__ src = (from_x0 ? p.x0 : p.rhs) + (static_cast<long long>(x) * p.ny + y) * p.nz;
  loadRealRow(r, src);
  __syncthreads();
  const cd* X = rfftForward(r);
  if (active) {
    cd* __restrict__ dst = p.spec + rid * p.nmode;
    for (int k = r.lane; k < p.nmode; k += r.tpr) {
      dst[k] = X[k];
    }
  }
}

__device__ __forceinline__ cd crecip(cd z) {
  const double d = 1.0 / (z.re * z.re + z.im * z.im);
  return {z.re * d, -z.im * d};
}

// K2: one thread per (y, kz) system, Thomas along x.
template <bool PERIODIC>
__global__ void __launch_bounds__(128) k_lap_tridiag(const LapParams p) {
  const int nys = p.ye - p.ys + 1;
  const int nsys = nys * p.nmode;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsys) {
    return;
  }
  const int iyl = t / p.nmode, kz = t % p.nmode;
  const int jy = p.ys + iyl;
  const int n = p.bd.xe - p.bd.xs + 1;
  const long long xstride = static_cast<long long>(nys) * p.nmode;
  cd* __restrict__ d = p.spec + t;
  cd* __restrict__ cp = p.cprime + t;
  const double kwave = kz * 2.0 * M_PI / p.m.zlength; // common_transform.cxx:73
  bool bad = false;

  if constexpr (!PERIODIC) {
    cd cprev = {0, 0}, dprev = {0, 0};
    for (int ix = 0; ix < n; ++ix) {
      cd a, b, c, r = d[ix * xstride];
      tridagRow(p.m, p.bd, p.bd.xs + ix, jy, kz, kwave, a, b, c, r);
      const cd denom = csub(b, cmul(a, cprev));
      bad = bad || !((denom.re * denom.re + denom.im * denom.im) > 1e-300);
      const cd inv = crecip(denom);
      cprev = cmul(c, inv);
      dprev = cmul(csub(r, cmul(a, dprev)), inv);
      cp[ix * xstride] = cprev;
      d[ix * xstride] = dprev;
    }
    cd x = dprev;
    for (int ix = n - 2; ix >= 0; --ix) {
      x = csub(d[ix * xstride], cmul(cp[ix * xstride], x));
      d[ix * xstride] = x;
    }
  } else {
    // cyclic system: row 0 couples to x[n-1] through a0, row n-1 to x[0] through c_{n-1}
    cd* __restrict__ q = p.qwork + t;
    cd a0, b0, c0, r0 = d[0];
    tridagRow(p.m, p.bd, p.bd.xs, jy, kz, kwave, a0, b0, c0, r0);
    cd an, bn, cn, rn = d[(n - 1) * xstride];
    tridagRow(p.m, p.bd, p.bd.xs + n - 1, jy, kz, kwave, an, bn, cn, rn);
    const cd gamma = {-b0.re, -b0.im};
    const cd vn = cmul(a0, crecip(gamma)); // v = (1, 0, ..., 0, a0/gamma)
    cd cprev = {0, 0}, dprev = {0, 0}, qprev = {0, 0};
    for (int ix = 0; ix < n; ++ix) {
      cd a, b, c, r = d[ix * xstride];
      tridagRow(p.m, p.bd, p.bd.xs + ix, jy, kz, kwave, a, b, c, r);
      cd u = {0, 0};
      if (ix == 0) {
        b = csub(b, gamma);
        u = gamma;
        a = {0, 0};
      }
      if (ix == n - 1) {
        b = csub(b, cmul(cn, vn));
        u = cadd(u, cn);
        c = {0, 0};
      }
      const cd denom = csub(b, cmul(a, cprev));
      bad = bad || !((denom.re * denom.re + denom.im * denom.im) > 1e-300);
      const cd inv = crecip(denom);
      cprev = cmul(c, inv);
      dprev = cmul(csub(r, cmul(a, dprev)), inv);
      qprev = cmul(csub(u, cmul(a, qprev)), inv);
      cp[ix * xstride] = cprev;
      d[ix * xstride] = dprev;
      q[ix * xstride] = qprev;
    }
    cd y = dprev, qq = qprev;
    const cd yn = y, qn = qq;
    for (int ix = n - 2; ix >= 0; --ix) {
      const cd cpi = cp[ix * xstride];
      y = csub(d[ix * xstride], cmul(cpi, y));
      qq = csub(q[ix * xstride], cmul(cpi, qq));
      d[ix * xstride] = y;
      q[ix * xstride] = qq;
    }
    // y, qq now hold element 0
    const cd num = cadd(y, cmul(vn, yn));
    const cd den = cadd(cadd(cd{1.0, 0.0}, qq), cmul(vn, qn));
    bad = bad || !((den.re * den.re + den.im * den.im) > 1e-300);
    const cd fact = cmul(num, crecip(den));
    double sum = 0.0;
    for (int ix = 0; ix < n; ++ix) {
      const cd x = csub(d[ix * xstride], cmul(fact, q[ix * xstride]));
      d[ix * xstride] = x;
      sum += x.re;
    }
    if (kz == 0) { // subtract the x-average of the kz=0 mode (cyclic_laplace.cxx:266-287)
      const double avg = sum / static_cast<double>(n);
      for (int ix = 0; ix < n; ++ix) {
        d[ix * xstride].re -= avg;
      }
    }
  }
  if (bad) {
    atomicExch(p.status, 1);
  }
}

// K3: spec -> irfft -> out rows (common_transform.cxx:82-115)
__global__ void __launch_bounds__(128) k_lap_backward(const LapParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int rloc;
  const RowFFT r = makeRow(p, smem_raw, rloc);
  const int nys = p.ye - p.ys + 1;
  const int nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  const long long rowid = static_cast<long long>(blockIdx.x) * p.rows + rloc;
  const bool active = rowid < nrows;
  const long long rid = active ? rowid : nrows - 1;
  const int ixl = static_cast<int>(rid / nys), iyl = static_cast<int>(rid % nys);
  const int x = p.bd.xs + ixl, y = p.ys + iyl;
  const cd* __restrict__ src = p.spec + rid * p.nmode;
  for (int k = r.lane; k <= p.nz / 2; k += r.tpr) {
    cd v = {0.0, 0.0};
    if (k < p.nmode && !(k == 0 && p.zero_dc)) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(src) + k);
      v = {t.x, t.y};
    }
    r.A[k] = v;
  }
  __syncthreads();
  const double* res = rfftBackward(r, r.A);
  if (active) {
    double* __restrict__ dst = p.out + (static_cast<long long>(x) * p.ny + y) * p.nz;
    if (p.nz % 2 == 0 && aligned16(dst)) {
      for (int j = r.lane; j < p.nz / 2; j += r.tpr) {
        reinterpret_cast<double2*>(dst)[j] = reinterpret_cast<const double2*>(res)[j];
      }
    } else {
      for (int j = r.lane; j < p.nz; j += r.tpr) {
        dst[j] = res[j];
      }
    }
  }
}

} // namespace b200

using namespace b200;

static void lapBndryWidths(const bout_b200_grid* g, const bout_b200_laplace* L, int* inb, int* outb) {
  *inb = *outb = g->xstart;
  if ((L->global_flags & INV_BOTH_BNDRY_ONE) || g->xstart < 2) {
    *inb = *outb = 1;
  }
  if (L->inner_boundary_flags & INV_BNDRY_ONE) {
    *inb = 1;
  }
  if (L->outer_boundary_flags & INV_BNDRY_ONE) {
    *outb = 1;
  }
}

extern "C" size_t bout_b200_laplace_cyclic_workspace(const bout_b200_grid* g, const bout_b200_laplace* L) {
  if (g == nullptr || L == nullptr) {
    return 0;
  }
  const size_t n = static_cast<size_t>(g->nx) * g->ny * (L->maxmode + 1) * sizeof(cd);
  return 256 + n * (L->periodic_x ? 3 : 2);
}

extern "C" int bout_b200_laplace_cyclic_solve(const bout_b200_grid* g, const bout_b200_laplace* L,
                                              const double* rhs, const double* x0, double* out,
                                              void* workspace, size_t workspace_bytes, void* stream) {
  if (checkGrid(g) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  B200_REQUIRE(L && rhs && out && workspace, BOUT_B200_ERR_ARG, "laplace: null argument");
  B200_REQUIRE(L->dx && L->g11 && L->g33 && L->g13, BOUT_B200_ERR_ARG, "laplace: metric missing");
  B200_REQUIRE(!L->all_terms || (L->G1 && L->G3), BOUT_B200_ERR_ARG, "laplace: all_terms needs G1, G3");
  B200_REQUIRE(L->maxmode >= 0 && L->maxmode <= g->nz / 2, BOUT_B200_ERR_ARG, "laplace: bad maxmode");
  B200_REQUIRE(workspace_bytes >= bout_b200_laplace_cyclic_workspace(g, L), BOUT_B200_ERR_ARG,
               "laplace: workspace too small");
  B200_REQUIRE(aligned16(workspace), BOUT_B200_ERR_ARG, "laplace: workspace must be 16-byte aligned");
  B200_REQUIRE(L->first_x && L->last_x, BOUT_B200_ERR_UNSUPPORTED,
               "laplace: X-decomposed solves go through bout_b200_laplace_cyclic_reduce/_finish");
  const int setflags = INV_SET;
  B200_REQUIRE(!((L->inner_boundary_flags | L->outer_boundary_flags) & setflags) || x0 != nullptr,
               BOUT_B200_ERR_ARG, "laplace: INVERT_SET needs x0");
  const int gradflags = INV_DC_GRAD | INV_AC_GRAD;
  if (((L->inner_boundary_flags | L->outer_boundary_flags) & gradflags)
      && ((L->inner_boundary_flags | L->outer_boundary_flags) & (INV_SET | INV_RHS))) {
    B200_REQUIRE(L->g_11 != nullptr, BOUT_B200_ERR_ARG, "laplace: gradient-with-value boundary needs g_11");
  }
  if ((L->inner_boundary_flags | L->outer_boundary_flags) & (INV_DC_GRADPAR | INV_DC_GRADPARINV)) {
    B200_REQUIRE(L->g_22 != nullptr, BOUT_B200_ERR_ARG, "laplace: GRADPAR boundary needs g_22");
  }
  cudaStream_t s = asStream(stream);

  LapParams p{};
  p.m.dx = L->dx;
  p.m.g11 = L->g11;
  p.m.g33 = L->g33;
  p.m.g13 = L->g13;
  p.m.G1 = L->G1;
  p.m.G3 = L->G3;
  p.m.g_11 = L->g_11;
  p.m.g_22 = L->g_22;
  p.m.A = L->A;
  p.m.C1 = L->C1;
  p.m.C2 = L->C2;
  p.m.D = L->D;
  p.m.zlength = L->zlength;
  p.m.nx = g->nx;
  p.m.ny = g->ny;
  p.m.all_terms = L->all_terms;
  p.m.nonuniform = L->nonuniform;
  p.bd.inner_flags = L->inner_boundary_flags;
  p.bd.outer_flags = L->outer_boundary_flags;
  p.bd.global_flags = L->global_flags;
  lapBndryWidths(g, L, &p.bd.inbndry, &p.bd.outbndry);
  p.bd.periodic_x = L->periodic_x;
  p.bd.first_x = L->first_x;
  p.bd.last_x = L->last_x;
  // cyclic_laplace.cxx:88-101
  p.bd.xs = (L->first_x && !L->periodic_x) ? 0 : g->xstart;
  p.bd.xe = (L->last_x && !L->periodic_x) ? g->nx - 1 : g->xend;
  p.bd.nx_local = g->nx;
  p.rhs = rhs;
  p.x0 = x0;
  p.out = out;
  p.nx = g->nx;
  p.ny = g->ny;
  p.nz = g->nz;
  // cyclic_laplace.cxx:216-232 (include_yguards only matters where a y boundary exists)
  p.ys = L->include_yguards_lower ? 0 : g->ystart;
  p.ye = L->include_yguards_upper ? g->ny - 1 : g->yend;
  p.nmode = L->maxmode + 1;
  p.zero_dc = (L->global_flags & INV_ZERO_DC) ? 1 : 0;
  p.in_set = (L->first_x && (L->inner_boundary_flags & INV_SET)) ? 1 : 0;
  p.out_set = (L->last_x && (L->outer_boundary_flags & INV_SET)) ? 1 : 0;

  const size_t nspec = static_cast<size_t>(g->nx) * g->ny * p.nmode;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  p.status = reinterpret_cast<int*>(ws);
  p.spec = reinterpret_cast<cd*>(ws + 256);
  p.cprime = p.spec + nspec;
  p.qwork = L->periodic_x ? p.cprime + nspec : nullptr;
  int st = 0;
  p.tw = twiddleTable(g->nz, s, &st);
  if (st != 0) {
    return st;
  }
  const RowGeom geo = rowGeometry(g->nz, 128);
  p.tpr = geo.tpr;
  p.rows = geo.rows;
  p.pow2 = geo.pow2 ? 1 : 0;
  B200_REQUIRE(geo.smemBytes <= 200 * 1024, BOUT_B200_ERR_UNSUPPORTED, "laplace: nz=%d too large", g->nz);
  if (geo.smemBytes > 48 * 1024) {
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_forward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(geo.smemBytes)));
    B200_CUDA_OK(cudaFuncSetAttribute(k_lap_backward, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(geo.smemBytes)));
  }
  B200_CUDA_OK(cudaMemsetAsync(p.status, 0, sizeof(int), s));
  const int nys = p.ye - p.ys + 1, nxs = p.bd.xe - p.bd.xs + 1;
  const long long nrows = static_cast<long long>(nxs) * nys;
  const unsigned rblocks = static_cast<unsigned>((nrows + geo.rows - 1) / geo.rows);
  k_lap_forward<<<rblocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  const int nsys = nys * p.nmode;
  if (L->periodic_x) {
    k_lap_tridiag<true><<<(nsys + 127) / 128, 128, 0, s>>>(p);
  } else {
    k_lap_tridiag<false><<<(nsys + 127) / 128, 128, 0, s>>>(p);
  }
  B200_LAUNCH_OK();
  k_lap_backward<<<rblocks, 128, geo.smemBytes, s>>>(p);
  B200_LAUNCH_OK();
  return 0;
}

// Synchronises the stream and reports a zero pivot seen by the last solve that used `workspace`
// (reference: BoutException "Zero pivot in CyclicReduce::reduce", cyclic_reduction.hxx:553).
extern "C" int bout_b200_laplace_cyclic_status(void* workspace, void* stream) {
  B200_REQUIRE(workspace != nullptr, BOUT_B200_ERR_ARG, "laplace status: null workspace");
  int flag = 0;
  B200_CUDA_OK(cudaMemcpyAsync(&flag, workspace, sizeof(int), cudaMemcpyDeviceToHost, asStream(stream)));
  B200_CUDA_OK(cudaStreamSynchronize(asStream(stream)));
  if (flag != 0) {
    return fail(BOUT_B200_ERR_PIVOT, "Zero pivot in the cyclic tridiagonal solve");
  }
  return 0;
}
