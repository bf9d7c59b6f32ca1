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

/* filter(f, N0) (src/field/field3d.cxx:620-661): keep only z-mode n0. */
int bout_b200_zfilter(const bout_b200_grid* g, int region, int n0, const double* f, double* out,
                      void* stream);
/* lowPass(f, zmax, keep_zonal) (src/field/field3d.cxx:664-710). */
int bout_b200_zlowpass(const bout_b200_grid* g, int region, int zmax, int keep_zonal,
                       const double* f, double* out, void* stream);
/* shiftZ (src/field/field3d.cxx:715-741; ShiftedMetric::shiftZ src/mesh/parallel/shiftedmetric.cxx:200-224):
 * multiply mode k by exp(-i k 2pi/zlength * angle); angle/zlength per (x,y) when the Field2D
 * pointers are non-NULL, else the scalars.  out may alias f. */
int bout_b200_zshift(const bout_b200_grid* g, int region, const double* f, const double* angle2d,
                     double angle0, const double* zlength2d, double zlength0, double* out,
                     void* stream);
/* DC(f): z-mean of every (x,y) row, RGN_ALL (src/field/field3d.cxx:785-803). out2d is a Field2D. */
int bout_b200_dc(const bout_b200_grid* g, const double* f, double* out2d, void* stream);

/* ---- (2) fused Arakawa bracket -----------------------------------------------------------------
 * bracket(Field3D f, Field3D g, BRACKET_ARAKAWA) (src/mesh/difops.cxx:940-1041);
 * dx, dz Field2D spacings.  out = [f,g] on RGN_NOBNDRY. */
int bout_b200_bracket_arakawa(const bout_b200_grid* g, const double* f, const double* gg,
                              const double* dx, const double* dz, double* out, void* stream);
/* bracket(Field3D f, Field2D g, BRACKET_ARAKAWA) (src/mesh/difops.cxx:693-765). */
int bout_b200_bracket_arakawa_3d2d(const bout_b200_grid* g, const double* f, const double* g2d,
                                   const double* dx, const double* dz, double* out, void* stream);

/* ---- (2b) fused Delp2 and (3) LaplaceCyclic -------------------------------------------------------
 * Laplacian settings + the Field2D metric the coefficients are generated from on chip
 * (reference Laplacian::Laplacian options src/invert/laplace/invert_laplace.cxx:68-138 and
 * Laplacian::tridagCoefs :280-375).  All pointers are device Field2D arrays [nx][ny]; optional ones
 * may be NULL. */
typedef struct bout_b200_laplace {
  const double* dx;
  const double* g11;
  const double* g33;
  const double* g13;
  const double* G1;   /* needed when all_terms */
  const double* G3;   /* needed when all_terms */
  const double* g_11; /* only for INVERT_{DC,AC}_GRAD combined with INVERT_SET / INVERT_RHS */
  const double* g_22; /* only for INVERT_DC_GRADPAR(INV) */
  const double* A;    /* setCoefA; NULL = 0 */
  const double* C1;   /* setCoefC1 / setCoefC; NULL = 1 */
  const double* C2;   /* setCoefC2 / setCoefC; NULL = 1 */
  const double* D;    /* setCoefD; NULL = 1 */
  double zlength;     /* Coordinates::zlength(), = dz * nz */
  int all_terms;      /* laplace:all_terms (default true) */
  int nonuniform;     /* laplace:nonuniform (default coords->non_uniform) */
  int global_flags, inner_boundary_flags, outer_boundary_flags; /* invert_laplace.hxx:72-104 */
  int maxmode;        /* laplace:maxmode / filter, clamped to [0, nz/2] */
  int periodic_x;     /* mesh->periodicX */
  int first_x, last_x;/* mesh->firstX(), lastX() (1,1 on a single rank in X) */
  int include_yguards_lower, include_yguards_upper; /* laplace:include_yguards where a y boundary exists */
} bout_b200_laplace;

/* Coordinates::Delp2(Field3D, useFFT=true) (src/mesh/coordinates.cxx:1662-1699); uses
 * dx,g11,g33,g13,G1,G3,zlength,all_terms,nonuniform of L (the default Laplacian instance). */
int bout_b200_delp2_fft(const bout_b200_grid* g, const bout_b200_laplace* L, const double* f,
                        double* out, void* stream);
/* Coordinates::Delp2(Field3D, useFFT=false) with C2 methods (coordinates.cxx:1701-1702);
 * d1_dx = NULL when coords->non_uniform is false. */
int bout_b200_delp2_fd(const bout_b200_grid* g, const double* f, const double* dx, const double* dz,
                       const double* g11, const double* g33, const double* g13, const double* G1,
                       const double* G3, const double* d1_dx, double* out, void* stream);

/* LaplaceCyclic::solve(Field3D rhs, Field3D x0)
 * (src/invert/laplace/impls/cyclic/cyclic_laplace.cxx:187-290).  `workspace` is caller-owned device
 * memory of at least bout_b200_laplace_cyclic_workspace() bytes, 16-byte aligned, reusable across
 * calls (no allocation happens inside a solve).  x0 may be NULL unless INVERT_SET is used.
 * out is written for x in [xs,xe] (boundary cells included on edge ranks unless periodic),
 * y in the solved range; other cells are left untouched. */
size_t bout_b200_laplace_cyclic_workspace(const bout_b200_grid* g, const bout_b200_laplace* L);
int bout_b200_laplace_cyclic_solve(const bout_b200_grid* g, const bout_b200_laplace* L,
                                   const double* rhs, const double* x0, double* out,
                                   void* workspace, size_t workspace_bytes, void* stream);
/* Synchronises `stream`; BOUT_B200_ERR_PIVOT if the last solve on this workspace met a zero pivot
 * (reference throws BoutException, include/bout/cyclic_reduction.hxx:553). */
int bout_b200_laplace_cyclic_status(void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BOUT_B200_H */
