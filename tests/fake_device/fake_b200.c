/* TEST DOUBLE, never shipped and never loaded by the product: a host-memory stand-in for the handful of C-ABI
 * entry points (include/bout_b200.h) that the C++ host layer's explicit integrators go through.  tests/
 * LD_PRELOADs it in a child process so that the HOST LOGIC of bout-dev_b200/host (Solver::create, the euler /
 * rkgeneric / splitrk control flow, copy-on-write states) can be compared with the reference build on a box
 * without a GPU.  Stencils and the self halo are forwarded to the oracle's C restatement when
 * $B200_FAKE_ORACLE_PORT names it.  Every other entry point still resolves to the real library and fails with
 * BOUT_B200_ERR_CUDA.
 * The GPU tests run the same cases through the real kernels. */
#define _GNU_SOURCE
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "bout_b200.h"

const char* bout_b200_last_error(void) { return "fake device (tests/fake_device/fake_b200.c)"; }
int bout_b200_malloc(void** ptr, size_t bytes) {
  *ptr = calloc(bytes ? bytes : 1, 1);
  return *ptr ? 0 : BOUT_B200_ERR_CUDA;
}
int bout_b200_free(void* ptr) {
  free(ptr);
  return 0;
}
int bout_b200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  (void)stream;
  memmove(dst, src, bytes);
  return 0;
}
int bout_b200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  (void)stream;
  memmove(dst, src, bytes);
  return 0;
}
int bout_b200_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  (void)stream;
  memmove(dst, src, bytes);
  return 0;
}
int bout_b200_stream_sync(void* stream) {
  (void)stream;
  return 0;
}

static int box(const bout_b200_grid* g, int region, int* xs, int* xe, int* ys, int* ye) {
  switch (region) {
  case BOUT_B200_RGN_NOBNDRY: *xs = g->xstart, *xe = g->xend, *ys = g->ystart, *ye = g->yend; return 0;
  case BOUT_B200_RGN_ALL: *xs = 0, *xe = g->nx - 1, *ys = 0, *ye = g->ny - 1; return 0;
  case BOUT_B200_RGN_NOX: *xs = g->xstart, *xe = g->xend, *ys = 0, *ye = g->ny - 1; return 0;
  case BOUT_B200_RGN_NOY: *xs = 0, *xe = g->nx - 1, *ys = g->ystart, *ye = g->yend; return 0;
  default: return BOUT_B200_ERR_ARG;
  }
}

int bout_b200_elementwise(const bout_b200_grid* g, int region, int nterms, const bout_b200_ewterm* t, double* out,
                          void* stream) {
  int xs, xe, ys, ye;
  (void)stream;
  if (nterms < 1 || nterms > 8 || box(g, region, &xs, &xe, &ys, &ye) != 0) {
    if (getenv("B200_FAKE_TRACE")) {
      fprintf(stderr, "fake bout_b200_elementwise: nterms = %d, region = %d\n", nterms, region);
    }
    return BOUT_B200_ERR_ARG;
  }
  for (int x = xs; x <= xe; ++x) {
    for (int y = ys; y <= ye; ++y) {
      const long long i2 = (long long)x * g->ny + y;
      for (int z = 0; z < g->nz; ++z) {
        const long long i3 = i2 * g->nz + z;
        double acc = 0.0;
        for (int k = 0; k < nterms; ++k) {
          double v = t[k].coef;
          if (t[k].f1) {
            v *= t[k].f1_is2d ? t[k].f1[i2] : t[k].f1[i3];
          }
          if (t[k].f2) {
            v *= t[k].f2_is2d ? t[k].f2[i2] : t[k].f2[i3];
          }
          if (t[k].div) {
            v = t[k].div_is2d ? v * (1.0 / t[k].div[i2]) : v / t[k].div[i3];
          }
          acc = (k == 0) ? v : acc + v;
        }
        out[i3] = acc;
      }
    }
  }
  return 0;
}

int bout_b200_lincomb(const bout_b200_grid* g, int region, int n, const double* coef, const double* const* x,
                      double* out, void* stream) {
  bout_b200_ewterm t[8];
  if (n < 1 || n > 8) {
    return BOUT_B200_ERR_ARG;
  }
  for (int i = 0; i < n; ++i) {
    if (x[i] == NULL || x[i] == out) {
      if (getenv("B200_FAKE_TRACE")) {
        fprintf(stderr, "fake bout_b200_lincomb: operand %d of %d is %s\n", i, n, x[i] == NULL ? "NULL" : "the output");
      }
      return BOUT_B200_ERR_ARG;
    }
    memset(&t[i], 0, sizeof(t[i]));
    t[i].coef = coef[i];
    t[i].f1 = x[i];
  }
  return bout_b200_elementwise(g, region, n, t, out, stream);
}

int bout_b200_unary(const bout_b200_grid* g, int op, const double* f, double arg, double* out, void* stream) {
  const long long n = (long long)g->nx * g->ny * g->nz;
  (void)stream;
  for (long long i = 0; i < n; ++i) {
    const double v = f[i];
    double r;
    switch (op) {
    case 0: r = -1.0 * v; break;
    case 1: r = sqrt(v); break;
    case 2: r = fabs(v); break;
    case 3: r = exp(v); break;
    case 4: r = log(v); break;
    case 5: r = sin(v); break;
    case 6: r = cos(v); break;
    case 7: r = tanh(v); break;
    case 8: r = v * v; break;
    case 9: r = pow(v, arg); break;
    case 10: r = v; break;
    default: return BOUT_B200_ERR_ARG;
    }
    out[i] = r;
  }
  return 0;
}

int bout_b200_rk_stage(const bout_b200_grid* g, int region, int mode, double dt, double c, const double* start,
                       const double* k1, const double* k2, const double* k3, const double* k4, double* out,
                       void* stream) {
  int xs, xe, ys, ye;
  (void)stream;
  if (box(g, region, &xs, &xe, &ys, &ye) != 0 || mode < 0 || mode > 3) {
    return BOUT_B200_ERR_ARG;
  }
  for (int x = xs; x <= xe; ++x) {
    for (int y = ys; y <= ye; ++y) {
      for (int z = 0; z < g->nz; ++z) {
        const long long i = ((long long)x * g->ny + y) * g->nz + z;
        switch (mode) {
        case 0: out[i] = start[i] + c * dt * k1[i]; break;
        case 1: out[i] = start[i] + (1. / 6.) * dt * (k1[i] + 2. * k2[i] + 2. * k3[i] + k4[i]); break;
        case 2: out[i] = 0.75 * start[i] + 0.25 * k1[i] + 0.25 * dt * k2[i]; break;
        default: out[i] = (1. / 3) * start[i] + (2. / 3.) * (k1[i] + dt * k2[i]); break;
        }
      }
    }
  }
  return 0;
}

size_t bout_b200_err_sum_scratch_doubles(void) { return 16; }
int bout_b200_err_sum(const bout_b200_grid* g, int region, const double* a, const double* b, double atol,
                      double* scratch, int accumulate, double* result, void* stream) {
  int xs, xe, ys, ye;
  double s = 0.0;
  (void)stream;
  (void)scratch;
  if (box(g, region, &xs, &xe, &ys, &ye) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  for (int x = xs; x <= xe; ++x) {
    for (int y = ys; y <= ye; ++y) {
      for (int z = 0; z < g->nz; ++z) {
        const long long i = ((long long)x * g->ny + y) * g->nz + z;
        s += fabs(a[i] - b[i]) / (fabs(a[i]) + fabs(b[i]) + atol);
      }
    }
  }
  result[0] = (accumulate ? result[0] : 0.0) + s;
  return 0;
}

int bout_b200_err_max(const bout_b200_grid* g, int region, const double* a, const double* b, double* scratch,
                      int accumulate, double* result, void* stream) {
  int xs, xe, ys, ye;
  double s = 0.0;
  (void)stream;
  (void)scratch;
  if (box(g, region, &xs, &xe, &ys, &ye) != 0) {
    return BOUT_B200_ERR_ARG;
  }
  for (int x = xs; x <= xe; ++x) {
    for (int y = ys; y <= ye; ++y) {
      for (int z = 0; z < g->nz; ++z) {
        const long long i = ((long long)x * g->ny + y) * g->nz + z;
        s = fmax(s, fabs(a[i] - b[i]));
      }
    }
  }
  result[0] = accumulate ? fmax(result[0], s) : s;
  return 0;
}

/* output step: streams are tokens, page-locked memory is plain memory, the "asynchronous" copy happens at once */
int bout_b200_stream_create(void** stream) {
  *stream = malloc(1);
  return *stream ? 0 : BOUT_B200_ERR_CUDA;
}
int bout_b200_stream_destroy(void* stream) {
  free(stream);
  return 0;
}
int bout_b200_stream_wait(void* waiter, void* signaller) {
  (void)waiter;
  (void)signaller;
  return 0;
}
int bout_b200_host_alloc(void** ptr, size_t bytes) {
  *ptr = malloc(bytes ? bytes : 1);
  return *ptr ? 0 : BOUT_B200_ERR_CUDA;
}
int bout_b200_host_free(void* ptr) {
  free(ptr);
  return 0;
}
int bout_b200_memcpy_d2h_async(void* dst, const void* src, size_t bytes, void* stream) {
  (void)stream;
  memmove(dst, src, bytes);
  return 0;
}

int bout_b200_boundary(const bout_b200_grid* g, int kind, int side, double value, const double* spacing2d, int lo,
                       int hi, double* f, void* stream);
int bout_b200_boundary_batch(const bout_b200_grid* g, int n, const bout_b200_bndry_spec* specs, void* stream) {
  for (int i = 0; i < n; ++i) { /* one after the other: the entries of a batch are independent */
    const int rc = bout_b200_boundary(g, specs[i].kind, specs[i].side, specs[i].value, specs[i].spacing2d, specs[i].lo,
                                      specs[i].hi, specs[i].f, stream);
    if (rc != 0) {
      return rc;
    }
  }
  return 0;
}

int bout_b200_where(const bout_b200_grid* g, const double* test, int test_is2d, const double* gt0, double gt0_value,
                    const double* le0, double le0_value, double* out, void* stream) {
  const long long n = (long long)g->nx * g->ny * g->nz;
  (void)stream;
  for (long long i = 0; i < n; ++i) {
    const double t = test[test_is2d ? i / g->nz : i];
    out[i] = t > 0.0 ? (gt0 ? gt0[i] : gt0_value) : (le0 ? le0[i] : le0_value);
  }
  return 0;
}

/* boundary conditions (kinds as in include/bout_b200.h; 3 "extend" is elm-pb specific and not needed by the host logic) */
int bout_b200_boundary(const bout_b200_grid* g, int kind, int side, double value, const double* spacing2d, int lo,
                       int hi, double* f, void* stream) {
  int width, first, b;
  (void)stream;
  if (kind < 0 || kind > 5 || kind == 3 || side < 0 || side > 3) {
    return BOUT_B200_ERR_UNSUPPORTED;
  }
  switch (side) {
  case 0: width = g->xstart; first = g->xstart - 1; b = -1; break;
  case 1: width = g->nx - g->xend - 1; first = g->nx - width; b = 1; break;
  case 2: width = g->ystart; first = g->ystart - 1; b = -1; break;
  default: width = g->ny - g->yend - 1; first = g->ny - width; b = 1; break;
  }
  if (width <= 0) {
    return 0;
  }
  if (lo < 0 || hi < 0) {
    lo = side < 2 ? g->ystart : 0;
    hi = side < 2 ? g->yend : g->nx - 1;
  }
#define AT(n, o, z) f[(((long long)(side < 2 ? (n) : (o))) * g->ny + (side < 2 ? (o) : (n))) * g->nz + (z)]
  for (int o = lo; o <= hi; ++o) {
    for (int z = 0; z < g->nz; ++z) {
      for (int i = 0; i < width; ++i) {
        const int n = first + i * b;
        if (kind == 2) {
          AT(n, o, z) = 0.0;
        } else if (kind == 4) {
          AT(n, o, z) = 2 * AT(n - b, o, z) - AT(n - 2 * b, o, z);
        } else if (kind == 5) {
          AT(n, o, z) = 3.0 * AT(n - b, o, z) - 3.0 * AT(n - 2 * b, o, z) + AT(n - 3 * b, o, z);
        } else if (kind == 0) {
          AT(n, o, z) = (i == 0) ? 2 * value - AT(n - b, o, z) : value;
        } else { /* neumann */
          const int x = side < 2 ? first : o, y = side < 2 ? o : first;
          const double delta = b * spacing2d[(long long)x * g->ny + y];
          if (i == 0) {
            AT(n, o, z) = AT(n - b, o, z) + delta * value;
          } else if (i == 1 && width == 2) {
            AT(n, o, z) = AT(first - 2 * b, o, z) + 3.0 * delta * value;
          }
        }
      }
    }
  }
#undef AT
  return 0;
}

/* stencils and the self halo: forwarded to the oracle's plain-C restatement (oracle/_ref/liboracle_port.so, path in
 * $B200_FAKE_ORACLE_PORT) -- tests may use the oracle, the product never does.  Lets the CPU suite run whole
 * derivative-based models (examples/IMEX/advection-diffusion) through the host layer to check its dispatch
 * (method names -> operator ids, 1/dx factors, communicate) against the reference build. */
#include <dlfcn.h>
typedef int (*bo_index_deriv_fn)(const void* g, int op, int dir, const double* v, const double* f, const double* yup,
                                 const double* ydown, double* out);
static bo_index_deriv_fn oracle_index_deriv(void) {
  static bo_index_deriv_fn fn = NULL;
  if (fn == NULL) {
    const char* path = getenv("B200_FAKE_ORACLE_PORT");
    void* lib = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : NULL;
    fn = lib ? (bo_index_deriv_fn)dlsym(lib, "bo_index_deriv") : NULL;
  }
  return fn;
}

int bout_b200_stencil(const bout_b200_grid* g, int op, int dir, int region, const double* v, const double* f,
                      const double* yup, const double* ydown, const double* scale1, int op2, const double* scale2,
                      double* out, void* stream) {
  const long long n = (long long)g->nx * g->ny * g->nz;
  bo_index_deriv_fn deriv = oracle_index_deriv();
  (void)stream;
  if (deriv == NULL || region != BOUT_B200_RGN_NOBNDRY) {
    return BOUT_B200_ERR_UNSUPPORTED;
  }
  double* t1 = (double*)calloc((size_t)n, sizeof(double));
  double* t2 = (double*)calloc((size_t)n, sizeof(double));
  int rc = deriv(g, op, dir, v, f, yup, ydown, t1); /* bo_grid has the layout of bout_b200_grid */
  if (rc == 0 && op2 >= 0) {
    rc = deriv(g, op2, dir, v, f, yup, ydown, t2);
  }
  for (int x = g->xstart; rc == 0 && x <= g->xend; ++x) {
    for (int y = g->ystart; y <= g->yend; ++y) {
      const long long i2 = (long long)x * g->ny + y;
      for (int z = 0; z < g->nz; ++z) {
        const long long i = i2 * g->nz + z;
        double r = (scale1 ? scale1[i2] : 1.0) * t1[i];
        if (op2 >= 0) {
          r += (scale2 ? scale2[i2] : 1.0) * t2[i];
        }
        out[i] = r;
      }
    }
  }
  free(t1);
  free(t2);
  if (rc != 0 && getenv("B200_FAKE_TRACE")) {
    fprintf(stderr, "fake bout_b200_stencil: oracle index derivative op=%d dir=%d op2=%d failed (%d) on %dx%dx%d\n", op, dir, op2,
            rc, g->nx, g->ny, g->nz);
  }
  return rc == 0 ? 0 : BOUT_B200_ERR_ARG;
}

/* smooth_x (smoothing.cxx:46-73) without its closing communicate */
int bout_b200_smooth_x(const bout_b200_grid* g, const double* f, double* out, void* stream) {
  (void)stream;
  const size_t row = (size_t)g->ny * g->nz;
  memcpy(out, f, row * sizeof(double));
  memcpy(out + (size_t)(g->nx - 1) * row, f + (size_t)(g->nx - 1) * row, row * sizeof(double));
  for (int x = 1; x < g->nx - 1; ++x) {
    for (size_t i = 0; i < row; ++i) {
      out[x * row + i] = 0.5 * f[x * row + i] + 0.25 * (f[(x - 1) * row + i] + f[(x + 1) * row + i]);
    }
  }
  return 0;
}

/* averageY(Field3D) in its two halves (smoothing.cxx:272-323): mean over ystart..yend per (x, z); every y row set to it */
int bout_b200_ymean(const bout_b200_grid* g, const double* f, double* xz, void* stream) {
  (void)stream;
  for (int x = 0; x < g->nx; ++x) {
    for (int z = 0; z < g->nz; ++z) {
      double acc = 0.0;
      for (int y = g->ystart; y <= g->yend; ++y) {
        acc += f[(((long long)x) * g->ny + y) * g->nz + z];
      }
      xz[(long long)x * g->nz + z] = acc / (g->yend - g->ystart + 1);
    }
  }
  return 0;
}
int bout_b200_yspread(const bout_b200_grid* g, const double* xz, double scale, double* out, void* stream) {
  (void)stream;
  for (int x = 0; x < g->nx; ++x) {
    for (int y = 0; y < g->ny; ++y) {
      for (int z = 0; z < g->nz; ++z) {
        out[(((long long)x) * g->ny + y) * g->nz + z] = scale * xz[(long long)x * g->nz + z];
      }
    }
  }
  return 0;
}

/* pack_data / unpack_data (boutmesh.cxx:2400-2466): rows y0..y0+nrows-1, x in [xlo,xhi] <-> [field][x][y][z] */
int bout_b200_halo_pack_y(const bout_b200_grid* g, int nfields, double* const* fields, int y0, int nrows, int xlo, int xhi,
                          double* buf, void* stream) {
  (void)stream;
  size_t o = 0;
  for (int k = 0; k < nfields; ++k) {
    for (int x = xlo; x <= xhi; ++x) {
      for (int y = y0; y < y0 + nrows; ++y) {
        memcpy(buf + o, fields[k] + (((long long)x) * g->ny + y) * g->nz, sizeof(double) * g->nz);
        o += g->nz;
      }
    }
  }
  return 0;
}

int bout_b200_halo_unpack_y(const bout_b200_grid* g, int nfields, double* const* fields, int y0, int nrows, int xlo, int xhi,
                            const double* buf, void* stream) {
  (void)stream;
  size_t o = 0;
  for (int k = 0; k < nfields; ++k) {
    for (int x = xlo; x <= xhi; ++x) {
      for (int y = y0; y < y0 + nrows; ++y) {
        memcpy(fields[k] + (((long long)x) * g->ny + y) * g->nz, buf + o, sizeof(double) * g->nz);
        o += g->nz;
      }
    }
  }
  return 0;
}

int bout_b200_halo_self(const bout_b200_grid* g, int nfields, double* const* fields, int periodic_x, int periodic_y,
                        void* stream) {
  const int mx = g->xend - g->xstart + 1, my = g->yend - g->ystart + 1;
  (void)stream;
  for (int k = 0; k < nfields; ++k) {
    double* f = fields[k];
#define F3(x, y, z) f[(((long long)(x)) * g->ny + (y)) * g->nz + (z)]
    if (periodic_y) { /* Y first, then X with the y guards (src/mesh/mesh.cxx:351-362) */
      for (int x = 0; x < g->nx; ++x) {
        for (int z = 0; z < g->nz; ++z) {
          for (int i = 0; i < g->ystart; ++i) {
            F3(x, i, z) = F3(x, i + my, z);
          }
          for (int i = 0; i < g->ny - g->yend - 1; ++i) {
            F3(x, g->yend + 1 + i, z) = F3(x, g->ystart + i, z);
          }
        }
      }
    }
    if (periodic_x) {
      for (int y = 0; y < g->ny; ++y) {
        for (int z = 0; z < g->nz; ++z) {
          for (int i = 0; i < g->xstart; ++i) {
            F3(i, y, z) = F3(i + mx, y, z);
          }
          for (int i = 0; i < g->nx - g->xend - 1; ++i) {
            F3(g->xend + 1 + i, y, z) = F3(g->xstart + i, y, z);
          }
        }
      }
    }
#undef F3
  }
  return 0;
}

/* Arakawa brackets, Delp2 (FFT branch), LaplaceCyclic and DC forwarded to the oracle port as well: the unchanged
 * examples/hasegawa-wakatani/hw.cxx then runs on CPU through the host layer (Laplacian options and flags,
 * split-operator RHS, the integrators) against the reference build. */
typedef struct {
  const double *dx, *dz, *g11, *g33, *g13, *g_11, *g_22, *G1, *G3, *IntShiftTorsion;
  double zlength;
  int all_terms, nonuniform, inc_int_shear, periodic_x, first_x, last_x;
} fake_lapl_metric; /* layout of bo_lapl_metric (oracle/bout_oracle.c) */
typedef struct {
  int global, inner, outer;
} fake_lapl_flags;

static void* oracle_symbol(const char* name) {
  static void* lib = NULL;
  if (lib == NULL) {
    const char* path = getenv("B200_FAKE_ORACLE_PORT");
    lib = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : NULL;
  }
  return lib ? dlsym(lib, name) : NULL;
}

static double* constant2d(const bout_b200_grid* g, double v) {
  const long long n = (long long)g->nx * g->ny;
  double* a = (double*)malloc((size_t)n * sizeof(double));
  for (long long i = 0; i < n; ++i) {
    a[i] = v;
  }
  return a;
}

static fake_lapl_metric to_metric(const bout_b200_grid* g, const bout_b200_laplace* L, double* dz) {
  fake_lapl_metric m;
  memset(&m, 0, sizeof(m));
  m.dx = L->dx, m.dz = dz, m.g11 = L->g11, m.g33 = L->g33, m.g13 = L->g13, m.g_11 = L->g_11, m.g_22 = L->g_22;
  m.G1 = L->G1, m.G3 = L->G3, m.zlength = L->zlength, m.all_terms = L->all_terms, m.nonuniform = L->nonuniform;
  m.periodic_x = L->periodic_x, m.first_x = L->first_x, m.last_x = L->last_x;
  (void)g;
  return m;
}

int bout_b200_bracket_arakawa(const bout_b200_grid* g, const double* f, const double* gg, const double* dx,
                              const double* dz, double* out, void* stream) {
  typedef int (*fn_t)(const void*, const double*, const double*, const double*, const double*, double*);
  fn_t fn = (fn_t)oracle_symbol("bo_bracket_arakawa");
  (void)stream;
  return fn ? fn(g, f, gg, dx, dz, out) : BOUT_B200_ERR_UNSUPPORTED;
}
int bout_b200_bracket_arakawa_3d2d(const bout_b200_grid* g, const double* f, const double* g2d, const double* dx,
                                   const double* dz, double* out, void* stream) {
  typedef int (*fn_t)(const void*, const double*, const double*, const double*, const double*, double*);
  fn_t fn = (fn_t)oracle_symbol("bo_bracket_arakawa_3d2d");
  (void)stream;
  return fn ? fn(g, f, g2d, dx, dz, out) : BOUT_B200_ERR_UNSUPPORTED;
}
int bout_b200_dc(const bout_b200_grid* g, const double* f, double* out2d, void* stream) {
  typedef int (*fn_t)(const void*, const double*, double*);
  fn_t fn = (fn_t)oracle_symbol("bo_dc");
  (void)stream;
  return fn ? fn(g, f, out2d) : BOUT_B200_ERR_UNSUPPORTED;
}
/* Coordinates::Delp2(f, useFFT = false) with C2 methods (coordinates.cxx:1701-1702):
 * G1 DDX + G3 DDZ + g11 D2DX2 (+ the non-uniform d1_dx term of D2DX2) + g33 D2DZ2 + 2 g13 D2DXDZ over RGN_NOBNDRY */
int bout_b200_delp2_fd(const bout_b200_grid* g, const double* f, const double* dx, const double* dz, const double* g11,
                       const double* g33, const double* g13, const double* G1, const double* G3, const double* d1_dx,
                       double* out, void* stream) {
  const int ny = g->ny, nz = g->nz;
  (void)stream;
#define F3(x, y, z) f[((size_t)(x) * ny + (y)) * nz + (((z) % nz + nz) % nz)]
  for (int x = g->xstart; x <= g->xend; ++x) {
    for (int y = g->ystart; y <= g->yend; ++y) {
      const size_t i2 = (size_t)x * ny + y;
      const double rdx = 1.0 / dx[i2], rdz = 1.0 / dz[i2];
      for (int z = 0; z < nz; ++z) {
        const double ddx = (0.5 * (F3(x + 1, y, z) - F3(x - 1, y, z))) * rdx;
        const double ddz = (0.5 * (F3(x, y, z + 1) - F3(x, y, z - 1))) * rdz;
        double d2dx2 = (F3(x + 1, y, z) + F3(x - 1, y, z) - 2. * F3(x, y, z)) * (rdx * rdx);
        if (d1_dx != NULL) {
          d2dx2 += (d1_dx[i2] * (0.5 * (F3(x + 1, y, z) - F3(x - 1, y, z)))) * rdx;
        }
        const double d2dz2 = (F3(x, y, z + 1) + F3(x, y, z - 1) - 2. * F3(x, y, z)) * (rdz * rdz);
        const double ddx_zp = (0.5 * (F3(x + 1, y, z + 1) - F3(x - 1, y, z + 1))) * rdx;
        const double ddx_zm = (0.5 * (F3(x + 1, y, z - 1) - F3(x - 1, y, z - 1))) * rdx;
        const double d2dxdz = (0.5 * (ddx_zp - ddx_zm)) * rdz;
        out[i2 * nz + z] = G1[i2] * ddx + G3[i2] * ddz + g11[i2] * d2dx2 + g33[i2] * d2dz2 + 2 * g13[i2] * d2dxdz;
      }
    }
  }
#undef F3
  return 0;
}

int bout_b200_delp2_fft(const bout_b200_grid* g, const bout_b200_laplace* L, const double* f, double* out,
                        void* stream) {
  typedef int (*fn_t)(const void*, const void*, const double*, double*);
  fn_t fn = (fn_t)oracle_symbol("bo_delp2_fft");
  (void)stream;
  if (fn == NULL) {
    return BOUT_B200_ERR_UNSUPPORTED;
  }
  double* dz = constant2d(g, L->zlength / g->nz);
  fake_lapl_metric m = to_metric(g, L, dz);
  const int rc = fn(g, &m, f, out);
  free(dz);
  return rc;
}
size_t bout_b200_laplace_cyclic_workspace(const bout_b200_grid* g, const bout_b200_laplace* L) {
  (void)g;
  (void)L;
  return 64;
}
int bout_b200_laplace_cyclic_solve(const bout_b200_grid* g, const bout_b200_laplace* L, const double* rhs,
                                   const double* x0, double* out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  typedef int (*fn_t)(const void*, const void*, const void*, int, const double*, const double*, const double*,
                      const double*, const double*, const double*, double*);
  fn_t fn = (fn_t)oracle_symbol("bo_laplace_cyclic");
  (void)workspace;
  (void)workspace_bytes;
  (void)stream;
  if (fn == NULL || L->include_yguards_lower || L->include_yguards_upper) {
    return BOUT_B200_ERR_UNSUPPORTED;
  }
  double* dz = constant2d(g, L->zlength / g->nz);
  double* zeros = constant2d(g, 0.0);
  double* ones = constant2d(g, 1.0);
  fake_lapl_metric m = to_metric(g, L, dz);
  fake_lapl_flags fl = {L->global_flags, L->inner_boundary_flags, L->outer_boundary_flags};
  const int rc = fn(g, &m, &fl, L->maxmode, L->A ? L->A : zeros, L->C1 ? L->C1 : ones, L->C2 ? L->C2 : ones,
                    L->D ? L->D : ones, rhs, x0 ? x0 : rhs, out);
  free(dz);
  free(zeros);
  free(ones);
  return rc == 0 ? 0 : BOUT_B200_ERR_PIVOT;
}

/* ---- X-decomposed solve (bout_b200_laplace_cyclic_reduce / _finish), emulated: the test double has no distributed
 * elimination.  The "interface record" a rank contributes is its whole slab -- right-hand side, initial guess, metric
 * and coefficient arrays -- and _finish stitches the records of the X row into the global problem (interior columns from
 * their owner, the boundary guard columns from the first / last rank), solves that with the single-rank routine above
 * and keeps its own columns.  What this pins is the HOST side: record size, all-gather order (PE_XIND), first_x /
 * last_x, which cells a rank receives.  The real kernels are checked against the oracle by tests/multigpu_check.py. */
#define FAKE_IFACE_ARRAYS 12
static const double* const* fake_lapl_arrays(const bout_b200_laplace* L, const double** a) {
  a[0] = L->dx, a[1] = L->g11, a[2] = L->g33, a[3] = L->g13, a[4] = L->G1, a[5] = L->G3, a[6] = L->g_11, a[7] = L->g_22;
  a[8] = L->A, a[9] = L->C1, a[10] = L->C2, a[11] = L->D;
  return a;
}
size_t bout_b200_laplace_cyclic_iface_doubles(const bout_b200_grid* g, const bout_b200_laplace* L) {
  (void)L;
  const size_t n3 = (size_t)g->nx * g->ny * g->nz, n2 = (size_t)g->nx * g->ny;
  return 2 * n3 + FAKE_IFACE_ARRAYS * n2 + FAKE_IFACE_ARRAYS + 1;
}
int bout_b200_laplace_cyclic_reduce(const bout_b200_grid* g, const bout_b200_laplace* L, const double* rhs, const double* x0,
                                    void* workspace, size_t workspace_bytes, double* iface_out, void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  (void)stream;
  const size_t n3 = (size_t)g->nx * g->ny * g->nz, n2 = (size_t)g->nx * g->ny;
  const double* a[FAKE_IFACE_ARRAYS];
  fake_lapl_arrays(L, a);
  double* p = iface_out;
  memcpy(p, rhs, n3 * sizeof(double));
  p += n3;
  if (x0) {
    memcpy(p, x0, n3 * sizeof(double));
  } else {
    memset(p, 0, n3 * sizeof(double));
  }
  p += n3;
  for (int k = 0; k < FAKE_IFACE_ARRAYS; ++k, p += n2) {
    if (a[k]) {
      memcpy(p, a[k], n2 * sizeof(double));
    } else {
      memset(p, 0, n2 * sizeof(double));
    }
  }
  for (int k = 0; k < FAKE_IFACE_ARRAYS; ++k) {
    *p++ = a[k] ? 1.0 : 0.0;
  }
  *p = x0 ? 1.0 : 0.0;
  return 0;
}
int bout_b200_laplace_cyclic_finish(const bout_b200_grid* g, const bout_b200_laplace* L, const double* iface_all, int nxpe,
                                    int pe_xind, double* out, void* workspace, size_t workspace_bytes, void* stream) {
  const int mxg = g->xstart, mxsub = g->nx - 2 * mxg, NX = nxpe * mxsub + 2 * mxg;
  const size_t n3 = (size_t)g->nx * g->ny * g->nz, n2 = (size_t)g->nx * g->ny, rec = bout_b200_laplace_cyclic_iface_doubles(g, L);
  const size_t row3 = (size_t)g->ny * g->nz, row2 = (size_t)g->ny;
  bout_b200_grid G = *g;
  G.nx = NX, G.xend = NX - mxg - 1;
  double* f3[2];
  double* f2[FAKE_IFACE_ARRAYS];
  for (int k = 0; k < 2; ++k) {
    f3[k] = (double*)calloc((size_t)NX * row3, sizeof(double));
  }
  for (int k = 0; k < FAKE_IFACE_ARRAYS; ++k) {
    f2[k] = (double*)calloc((size_t)NX * row2, sizeof(double));
  }
  double* outG = (double*)calloc((size_t)NX * row3, sizeof(double));
  for (int r = 0; r < nxpe; ++r) {
    const double* R = iface_all + (size_t)r * rec;
    const int lo = r == 0 ? 0 : mxg, hi = r == nxpe - 1 ? g->nx - 1 : g->nx - mxg - 1;
    for (int lx = lo; lx <= hi; ++lx) {
      const size_t gx = (size_t)r * mxsub + lx;
      memcpy(f3[0] + gx * row3, R + (size_t)lx * row3, row3 * sizeof(double));
      memcpy(f3[1] + gx * row3, R + n3 + (size_t)lx * row3, row3 * sizeof(double));
      for (int k = 0; k < FAKE_IFACE_ARRAYS; ++k) {
        memcpy(f2[k] + gx * row2, R + 2 * n3 + (size_t)k * n2 + (size_t)lx * row2, row2 * sizeof(double));
      }
    }
  }
  const double* flags = iface_all + (size_t)pe_xind * rec + 2 * n3 + FAKE_IFACE_ARRAYS * n2;
  bout_b200_laplace LG = *L;
#define PICK(k) (flags[k] != 0.0 ? f2[k] : NULL)
  LG.dx = PICK(0), LG.g11 = PICK(1), LG.g33 = PICK(2), LG.g13 = PICK(3), LG.G1 = PICK(4), LG.G3 = PICK(5), LG.g_11 = PICK(6);
  LG.g_22 = PICK(7), LG.A = PICK(8), LG.C1 = PICK(9), LG.C2 = PICK(10), LG.D = PICK(11);
#undef PICK
  LG.first_x = 1, LG.last_x = 1;
  const int rc = bout_b200_laplace_cyclic_solve(&G, &LG, f3[0], flags[FAKE_IFACE_ARRAYS] != 0.0 ? f3[1] : NULL, outG, workspace,
                                                workspace_bytes, stream);
  if (rc == 0) {
    const int lo = pe_xind == 0 ? 0 : mxg, hi = pe_xind == nxpe - 1 ? g->nx - 1 : g->nx - mxg - 1;
    for (int lx = lo; lx <= hi; ++lx) {
      memcpy(out + (size_t)lx * row3, outG + ((size_t)pe_xind * mxsub + lx) * row3, row3 * sizeof(double));
    }
  }
  for (int k = 0; k < 2; ++k) {
    free(f3[k]);
  }
  for (int k = 0; k < FAKE_IFACE_ARRAYS; ++k) {
    free(f2[k]);
  }
  free(outG);
  return rc;
}

/* ---- FieldPerp: plane copies, and the inversion of one plane through the full solve above ---------------------- */
int bout_b200_copy_xz(const bout_b200_grid* g, int jy, double* field3d, double* perp, int to_perp, void* stream) {
  (void)stream;
  if (jy < 0 || jy >= g->ny) {
    return BOUT_B200_ERR_ARG;
  }
  for (int x = 0; x < g->nx; ++x) {
    double* row3 = field3d + ((size_t)x * g->ny + jy) * g->nz;
    double* rowp = perp + (size_t)x * g->nz;
    memcpy(to_perp ? rowp : row3, to_perp ? row3 : rowp, (size_t)g->nz * sizeof(double));
  }
  return 0;
}
size_t bout_b200_laplace_cyclic_workspace_perp(const bout_b200_grid* g, const bout_b200_laplace* L) {
  (void)g;
  (void)L;
  return 64;
}
int bout_b200_laplace_cyclic_solve_perp(const bout_b200_grid* g, const bout_b200_laplace* L, int jy, const double* rhs,
                                        const double* x0, double* out, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  const size_t n3 = (size_t)g->nx * g->ny * g->nz;
  double* b3 = (double*)calloc(n3, sizeof(double));
  double* x3 = (double*)calloc(n3, sizeof(double));
  double* o3 = (double*)calloc(n3, sizeof(double));
  bout_b200_copy_xz(g, jy, b3, (double*)rhs, 0, NULL);
  bout_b200_copy_xz(g, jy, x3, (double*)(x0 ? x0 : rhs), 0, NULL);
  /* every y row of the full solve is an independent set of systems: the other rows solve a zero right-hand side */
  const int rc = bout_b200_laplace_cyclic_solve(g, L, b3, x3, o3, workspace, workspace_bytes, stream);
  if (rc == 0) {
    bout_b200_copy_xz(g, jy, o3, out, 1, NULL);
  }
  free(b3);
  free(x3);
  free(o3);
  return rc;
}

/* ---- fused hasegawa-wakatani-3d right-hand side (host-logic tests of the single-index loop dispatcher) --------
 * Brackets through the oracle port; DDZ, Delp2 and Div_par_Grad_par restated from the reference's
 * include/bout/single_index_ops.hxx:201-209, :222-246, :255-278 (identity parallel transform: yup / ydown are the
 * field itself), as oracle/oracle.py::hw3d_rhs does in numpy. */
size_t bout_b200_fused_coef_doubles(const bout_b200_grid* g) { return (size_t)g->nx * g->ny * 16; }
int bout_b200_fused_coefs(const bout_b200_grid* g, const bout_b200_metric* m, double* ws, void* stream) {
  (void)g, (void)m, (void)ws, (void)stream;
  return 0;
}
int bout_b200_hw3d_rhs(const bout_b200_grid* g, const bout_b200_metric* m, double alpha, double kappa, double Dn,
                       double Dvort, const double* n, const double* vort, const double* phi, double* ddt_n,
                       double* ddt_vort, const bout_b200_rk_epilogue* epilogue, double* coef_workspace, void* stream) {
  (void)coef_workspace, (void)stream;
  if (m == NULL || (epilogue == NULL && (ddt_n == NULL || ddt_vort == NULL))) {
    return BOUT_B200_ERR_UNSUPPORTED;
  }
  const size_t n3 = (size_t)g->nx * g->ny * g->nz;
  /* with an epilogue the time derivatives may be omitted: they are formed in scratch arrays and the stage update
   * state_out = start + a ddt, acc_out = (acc_in ? acc_in : start) + b ddt (include/bout_b200.h) follows below */
  double* own[2] = {NULL, NULL};
  if (ddt_n == NULL) {
    ddt_n = own[0] = (double*)calloc(n3, sizeof(double));
  }
  if (ddt_vort == NULL) {
    ddt_vort = own[1] = (double*)calloc(n3, sizeof(double));
  }
  double* bn = (double*)malloc(n3 * sizeof(double));
  double* bv = (double*)malloc(n3 * sizeof(double));
  int rc = bout_b200_bracket_arakawa(g, phi, n, m->dx, m->dz, bn, NULL);
  if (rc == 0) {
    rc = bout_b200_bracket_arakawa(g, phi, vort, m->dx, m->dz, bv, NULL);
  }
  if (rc != 0) {
    free(bn);
    free(bv);
    return rc;
  }
  const int ny = g->ny, nz = g->nz;
#define AT(f, x, y, z) (f)[((size_t)(x) * ny + (y)) * nz + (((z) % nz + nz) % nz)]
#define M2(a, x, y) (a)[(size_t)(x) * ny + (y)]
  for (int x = g->xstart; x <= g->xend; ++x) {
    for (int y = g->ystart; y <= g->yend; ++y) {
      const double dx = M2(m->dx, x, y), dz = M2(m->dz, x, y), dy = M2(m->dy, x, y), J = M2(m->J, x, y);
      const double d1 = m->d1_dx ? M2(m->d1_dx, x, y) : 0.0;
      for (int z = 0; z < nz; ++z) {
        double lap[2];
        const double* fl[2] = {n, vort};
        for (int q = 0; q < 2; ++q) {
          const double* f = fl[q];
          lap[q] = (M2(m->G1, x, y) + d1 * M2(m->g11, x, y)) * (AT(f, x + 1, y, z) - AT(f, x - 1, y, z)) / (2.0 * dx)
                   + M2(m->G3, x, y) * (AT(f, x, y, z + 1) - AT(f, x, y, z - 1)) / (2.0 * dz)
                   + M2(m->g11, x, y) * (AT(f, x + 1, y, z) - 2.0 * AT(f, x, y, z) + AT(f, x - 1, y, z)) / (dx * dx)
                   + M2(m->g33, x, y) * (AT(f, x, y, z + 1) - 2.0 * AT(f, x, y, z) + AT(f, x, y, z - 1)) / (dz * dz)
                   + 2 * M2(m->g13, x, y)
                         * ((AT(f, x + 1, y, z + 1) - AT(f, x - 1, y, z + 1))
                            - (AT(f, x + 1, y, z - 1) - AT(f, x - 1, y, z - 1)))
                         / (4. * dz * dx);
        }
        double div = 0.0;
        if (ny > 1) {
#define PMN(x, y, z) (AT(phi, x, y, z) - AT(n, x, y, z))
          const double gu = 2. * (PMN(x, y + 1, z) - PMN(x, y, z)) / (dy + M2(m->dy, x, y + 1));
          const double fu = gu * (J + M2(m->J, x, y + 1)) / (M2(m->g_22, x, y) + M2(m->g_22, x, y + 1));
          const double gl = 2. * (PMN(x, y, z) - PMN(x, y - 1, z)) / (dy + M2(m->dy, x, y - 1));
          const double fd = gl * (J + M2(m->J, x, y - 1)) / (M2(m->g_22, x, y) + M2(m->g_22, x, y - 1));
          div = alpha * (fu - fd) / (dy * J);
#undef PMN
        }
        const double ddz = 0.5 * (AT(phi, x, y, z + 1) - AT(phi, x, y, z - 1)) / dz;
        AT(ddt_n, x, y, z) = -AT(bn, x, y, z) - div - kappa * ddz + Dn * lap[0];
        AT(ddt_vort, x, y, z) = -AT(bv, x, y, z) - div + Dvort * lap[1];
      }
    }
  }
  if (epilogue != NULL) {
    double* k[2] = {ddt_n, ddt_vort};
    for (int e = 0; e < 2; ++e) {
      const double* base = epilogue->acc_in[e] ? epilogue->acc_in[e] : epilogue->start[e];
      for (int x = g->xstart; x <= g->xend; ++x) {
        for (int y = g->ystart; y <= g->yend; ++y) {
          for (int z = 0; z < nz; ++z) {
            const double d = AT(k[e], x, y, z);
            const double st = AT(epilogue->start[e], x, y, z), ac = AT(base, x, y, z);
            if (epilogue->state_out[e]) {
              AT(epilogue->state_out[e], x, y, z) = st + epilogue->a * d;
            }
            if (epilogue->acc_out[e]) {
              AT(epilogue->acc_out[e], x, y, z) = ac + epilogue->b * d;
            }
          }
        }
      }
    }
  }
#undef AT
#undef M2
  free(bn);
  free(bv);
  free(own[0]);
  free(own[1]);
  return 0;
}
