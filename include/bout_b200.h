/* bout_b200.h -- C-ABI of the B200-native BOUT++ RHS hot path (libbout_b200.so).
 *
 * Every entry point takes plain device pointers and sizes: no C++ types, no torch types.  Fields are
 * device-resident fp64 arrays in BOUT++'s own layout -- Field3D [nx][ny][nz] with z fastest and guard
 * cells included (reference include/bout/field3d.hxx:382), Field2D [nx][ny]
 * (include/bout/field2d.hxx:209).  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * All functions return 0 on success and a negative bout_b200_status otherwise; the message of the last
 * failure on the calling thread is available from bout_b200_last_error().  Nothing here falls back to
 * the CPU: without a CUDA device every compute call fails with BOUT_B200_ERR_CUDA.
 *
 * The reference has no FFI for this path: its "operator API" is the set of C++ functions physics models
 * compile against.  Each entry point below names the reference function it
 * replaces; the C++ header layer in bout-dev_b200/host/ maps the reference's signatures onto these calls,
 * and INTEGRATION.md shows the binding a BOUT++ maintainer would add.
 */
#ifndef BOUT_B200_H
#define BOUT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bout_b200_status {
  BOUT_B200_OK = 0,
  BOUT_B200_ERR_ARG = -1,      /* bad argument (maps to BoutException from ASSERTn) */
  BOUT_B200_ERR_CUDA = -2,     /* CUDA runtime failure / no device */
  BOUT_B200_ERR_GUARDS = -3,   /* stencil needs more guard cells than the grid has
                                  (reference ASSERT2 in include/bout/index_derivs.hxx:89) */
  BOUT_B200_ERR_PIVOT = -4,    /* zero pivot in the tridiagonal solve
                                  (reference include/bout/cyclic_reduction.hxx:553) */
  BOUT_B200_ERR_UNSUPPORTED = -5
} bout_b200_status;

/* Local grid, guard cells included (reference src/mesh/impls/bout/boutmesh.cxx:391-404). */
typedef struct bout_b200_grid {
  int nx, ny, nz;     /* LocalNx, LocalNy, LocalNz */
  int xstart, xend;   /* interior x range, inclusive */
  int ystart, yend;   /* interior y range, inclusive */
} bout_b200_grid;

/* Iteration regions (reference src/mesh/mesh.cxx:650-665). */
typedef enum bout_b200_region {
  BOUT_B200_RGN_NOBNDRY = 0, /* xstart..xend, ystart..yend, all z (default of every derivative) */
  BOUT_B200_RGN_ALL = 1,
  BOUT_B200_RGN_NOX = 2,     /* xstart..xend, all y */
  BOUT_B200_RGN_NOY = 3      /* all x, ystart..yend */
} bout_b200_region;

typedef enum bout_b200_dir { BOUT_B200_X = 0, BOUT_B200_Y = 1, BOUT_B200_Z = 2 } bout_b200_dir;

/* Stencil functors (reference src/mesh/index_derivs.cxx, line of each REGISTER_* in brackets). */
typedef enum bout_b200_stencil_op {
  BOUT_B200_DDX_C2 = 0,    /* [70]  */
  BOUT_B200_DDX_C4 = 1,    /* [75]  */
  BOUT_B200_DDX_W2 = 2,    /* [80]  CWENO2 */
  BOUT_B200_DDX_S2 = 3,    /* [102] */
  BOUT_B200_DDX_W3 = 4,    /* [219] CWENO3 */
  BOUT_B200_D2DX2_C2 = 10, /* [119] */
  BOUT_B200_D2DX2_C4 = 11, /* [124] */
  BOUT_B200_D4DX4_C2 = 12, /* [131] */
  BOUT_B200_VDDX_C2 = 20,  /* [148] */
  BOUT_B200_VDDX_C4 = 21,  /* [153] */
  BOUT_B200_VDDX_U1 = 22,  /* [158] */
  BOUT_B200_VDDX_U2 = 23,  /* [167] */
  BOUT_B200_VDDX_U3 = 24,  /* [178] */
  BOUT_B200_VDDX_W3 = 25,  /* [190] WENO3 */
  BOUT_B200_FDDX_U1 = 30,  /* [266] */
  BOUT_B200_FDDX_U2 = 31,  /* [283] */
  BOUT_B200_FDDX_C2 = 32,  /* [295] */
  BOUT_B200_FDDX_C4 = 33   /* [299] */
} bout_b200_stencil_op;

const char* bout_b200_last_error(void);
int bout_b200_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches counter). */
unsigned long long bout_b200_launch_count(void);

/* ---- (1) index_derivs DerivativeStore stencils ------------------------------------------------
 * out[region] = scale1[x,y] * op(stencil of f along dir) (+ scale2[x,y] * op2(stencil))
 * Replaces DerivativeType::standard / upwindOrFlux (include/bout/index_derivs.hxx:84-116) over
 * populateStencil (include/bout/stencils.hxx:40-119) and the division by dx/dy/dz that follows in
 * Coordinates::DDX/DDY/DDZ (src/mesh/coordinates.cxx:1467-1514) and D2DX2.. (src/sys/derivs.cxx:149-352).
 *  - v: velocity field for VDDX (value at the point) / FDDX (stencil) ops, NULL otherwise.
 *  - yup/ydown: parallel slices for dir == Y (DIRECTION::YOrthogonal); NULL = use f itself.
 *  - scale1/scale2: Field2D multipliers (e.g. 1/dx, 1/dx^2, d1_dx/dx) or NULL for 1.0; op2 < 0 = none.
 *  Cells outside the region are left untouched, as emptyFrom() leaves them in the reference. */
int bout_b200_stencil(const bout_b200_grid* g, int op, int dir, int region, const double* v,
                      const double* f, const double* yup, const double* ydown,
                      const double* scale1, int op2, const double* scale2, double* out,
                      void* stream);

/* FFT-in-Z derivative: order 1 (FFTDerivativeType, src/mesh/index_derivs.cxx:418-481, with
 * mesh:ddz:fft_filter) or order 2 (FFT2ndDerivativeType :483-541), times scale[x,y]. */
int bout_b200_ddz_fft(const bout_b200_grid* g, int order, double fft_filter, int region,
                      const double* f, const double* scale, double* out, void* stream);

/* ---- (2) fused Arakawa bracket -----------------------------------------------------------------
 * bracket(Field3D f, Field3D g, BRACKET_ARAKAWA) (src/mesh/difops.cxx:940-1041);
 * dx, dz Field2D spacings.  out = [f,g] on RGN_NOBNDRY. */
int bout_b200_bracket_arakawa(const bout_b200_grid* g, const double* f, const double* gg,
                              const double* dx, const double* dz, double* out, void* stream);
/* bracket(Field3D f, Field2D g, BRACKET_ARAKAWA) (src/mesh/difops.cxx:693-765). */
int bout_b200_bracket_arakawa_3d2d(const bout_b200_grid* g, const double* f, const double* g2d,
                                   const double* dx, const double* dz, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BOUT_B200_H */
