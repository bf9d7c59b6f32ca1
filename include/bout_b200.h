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

/* ---- device memory (Array<T> backing store, reference include/bout/array.hxx:61-139) ----------------
 * For host layers that do not link the CUDA runtime: cudaMalloc (zero-filled) / cudaFree / copies.
 * The host-side pool (bout-dev_b200/host) calls malloc only on a pool miss, so a steady-state time
 * step allocates nothing.  h2d / d2h copies synchronise `stream` before returning. */
int bout_b200_malloc(void** ptr, size_t bytes);
int bout_b200_free(void* ptr);
int bout_b200_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int bout_b200_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int bout_b200_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int bout_b200_stream_sync(void* stream);
/* Output step (reference PhysicsModelMonitor::call, src/physics/physicsmodel.cxx:241-268): a second,
 * non-blocking stream and page-locked staging so that a dump snapshot leaves the device while the next steps compute.
 * _stream_wait: work queued on `waiter` afterwards starts once everything queued on `signaller` so far is done.
 * _memcpy_d2h_async only queues the copy; synchronise `stream` before reading dst. */
int bout_b200_stream_create(void** stream);
int bout_b200_stream_destroy(void* stream);
int bout_b200_stream_wait(void* waiter, void* signaller);
int bout_b200_host_alloc(void** ptr, size_t bytes);
int bout_b200_host_free(void* ptr);
int bout_b200_memcpy_d2h_async(void* dst, const void* src, size_t bytes, void* stream);

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

/* Y stencils on a mesh with a non-identity parallel transform: neighbour y +- k comes from the parallel
 * slice field yup[k-1] / ydown[k-1] (Field3D::ynext(+-k), include/bout/field3d.hxx:265-300; stencil
 * population include/bout/stencils.hxx:73-100), nslices = 1 or 2 as the operator needs.  Standard and
 * upwind functors (the velocity is taken at the point). */
int bout_b200_stencil_yslices(const bout_b200_grid* g, int op, int region, const double* v,
                              const double* f, int nslices, const double* const* yup,
                              const double* const* ydown, const double* scale1, int op2,
                              const double* scale2, double* out, void* stream);

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
/* filter(f, n0) followed by lowPass(., zmax, keep_zonal) in ONE pass over the field (the ddt post-processing
 * of examples/elm-pb/elm_pb.cxx:1931-1955: src/field/field3d.cxx:620-661 then :664-710). */
int bout_b200_zfilter_lowpass(const bout_b200_grid* g, int region, int n0, int zmax, int keep_zonal,
                              const double* f, double* out, void* stream);
/* shiftZ (src/field/field3d.cxx:715-741; ShiftedMetric::shiftZ src/mesh/parallel/shiftedmetric.cxx:200-224):
 * multiply mode k by exp(-i k 2pi/zlength * angle); angle/zlength per (x,y) when the Field2D
 * pointers are non-NULL, else the scalars.  out may alias f. */
int bout_b200_zshift(const bout_b200_grid* g, int region, const double* f, const double* angle2d,
                     double angle0, const double* zlength2d, double zlength0, double* out,
                     void* stream);
/* ShiftedMetric::calcParallelSlices (src/mesh/parallel/shiftedmetric.cxx:54-124, 226-246): for k = 1..nslices
 * yup[k-1](x, y+k, :) = f(x, y+k, :) shifted in z by zShift(x,y) - zShift(x,y+k) (mode jz multiplied by
 * exp(-i jz 2pi/zlength * shift)), for y in ystart..yend and all x; ydown[k-1] likewise with y-k.
 * Other cells of the slice fields are left untouched.  toFieldAligned / fromFieldAligned are
 * bout_b200_zshift with angle2d = -zShift / +zShift. */
int bout_b200_parallel_slices(const bout_b200_grid* g, const double* f, const double* zshift2d,
                              double zlength, int nslices, double* const* yup, double* const* ydown,
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

/* b0xGrad_dot_Grad(phi, A) = bracket(phi, A, BRACKET_STD) * B (src/mesh/difops.cxx:455-601) in one pass:
 * dphi/dx,y,z with the mesh's first-derivative functors, v = (g_22 dz - g_23 dy, g_23 dx - g_12 dz,
 * g_12 dy - g_22 dx) [+ IntShiftTorsion vx on vz when mesh->IncIntShear], VDDX(vx,A) + VDDY(vy,A) +
 * VDDZ(vz,A) with the mesh's upwind functors, divided by J sqrt(g_22).  Either operand may be a Field2D
 * (*_is2d; its z derivative / advection term vanishes).  Y neighbours of a Field3D come from
 * *_nslices (0, 1 or 2) parallel slices each way; 0 = the field itself (identity transform).
 * yterm_out (optional): the VDDY(vy, A) / (J sqrt(g_22)) term is written there instead of being added
 * to out.  With a Field2D A on a shifted-metric mesh the reference evaluates that term in field-aligned
 * coordinates (the converted A carries no slices, include/bout/index_derivs_interface.hxx:360-390), i.e.
 * out += fromFieldAligned(toFieldAligned(yterm)): two bout_b200_zshift passes and one _elementwise. */
typedef struct bout_b200_b0xgrad {
  const double *dx, *dy, *dz, *J, *g_22, *g_12, *g_23; /* Field2D */
  const double* int_shift_torsion;                     /* NULL unless mesh->IncIntShear */
  int ddx_op, ddy_op, ddz_op;                          /* BOUT_B200_DDX_* (mesh:ddx:first ...) */
  int vddx_op, vddy_op, vddz_op;                       /* BOUT_B200_VDDX_* (mesh:ddx:upwind ...) */
} bout_b200_b0xgrad;
int bout_b200_b0xgrad_dot_grad(const bout_b200_grid* g, const bout_b200_b0xgrad* cfg,
                               const double* phi, int phi_is2d, int phi_nslices,
                               const double* const* phi_yup, const double* const* phi_ydown,
                               const double* A, int a_is2d, int a_nslices,
                               const double* const* a_yup, const double* const* a_ydown,
                               double* out, double* yterm_out, void* stream);

/* One Fourier coefficient per (x,y) row of up to 4 fields: out[f][x][y] = sum_j f(x,y,j) exp(-2 pi i n0 j / nz)
 * (complex, interleaved re/im, unnormalised), RGN_ALL. */
int bout_b200_row_mode(const bout_b200_grid* g, int nfields, const double* const* fields, int n0,
                       double* const* out_complex2d, void* stream);

/* ---- fused elm-pb right-hand side (a18: examples/elm-pb/elm_pb.cxx:1674-1955, shipped switch set) --------
 * ONE kernel for ddt(Psi), ddt(U), ddt(P) of the linear 3-field model incl. the final filter(., n0) +
 * lowPass: five b0xGrad_dot_Grad brackets (3D x 2D and 2D x 3D, difops.cxx:563-601), b0xcv . Grad(P),
 * Grad_par(phi B0) and Grad_par(Jpar), resistive + hyper-resistive terms.  Methods fixed to the shipped
 * ones (first C4/C2/C4, upwind W3/C2/W3).  All Field3D inputs carry valid guard cells (2 in x, 1 in y);
 * *_up / *_dn are the yup / ydown parallel slices (bout_b200_parallel_slices).  d2Jpar = Delp2(Jpar).
 * filter_n0 >= 0 (and below the Nyquist mode): the kept z mode, filtered in-kernel; keep_* = 1.0, or 0.0 if
 * the following lowPass removes that mode for the field (n0 > low_pass_z, or n0 = 0 without keep_zonal).
 * filter_n0 < 0 writes the unfiltered sums (valid only up to the Nyquist-mode handling of the reference's
 * field-aligned terms, see csrc/elm_rhs.cu).  Outputs cover RGN_NOBNDRY.  With the rk_* members set the kernel
 * also writes the next explicit-RK stage state, so a stage costs no extra pass. */
typedef struct bout_b200_elmpb_rhs_args {
  const double *dx, *dy, *dz, *J, *g_22, *g_12, *g_23, *Bxy;                 /* Field2D metric */
  const double *B0, *J0, *P0, *phi0, *b0xcv_x, *b0xcv_y, *b0xcv_z;           /* Field2D equilibrium */
  const double *U, *P, *Psi, *phi, *Jpar, *d2Jpar;                           /* Field3D */
  const double* eta;                                                         /* Field3D, or Field2D if eta_is2d */
  const double *U_up, *U_dn, *P_up, *P_dn, *Psi_up, *Psi_dn, *phi_up, *phi_dn, *J_up, *J_dn;
  /* Alternative to the U, P, Jpar slices (all three or none; needs filter_n0 >= 0): their y-derivative terms are
   * linear with z-independent coefficients, so after the filter only mode filter_n0 of each neighbour row
   * matters: U_mode / P_mode / J_mode = bout_b200_row_mode(., filter_n0) of the communicated fields (complex
   * [nx][ny]), shifted between the rows' frames with zShift / zlength as ShiftedMetric does.  With these set
   * the six *_up / *_dn pointers of U, P, Jpar may be NULL. */
  const double *U_mode, *P_mode, *J_mode, *zShift;
  /* With the modes AND a Field2D eta, d2Jpar may be NULL: Delp2(Jpar) of mode filter_n0 is then formed from
   * J_mode inside the kernel with Laplacian::tridagCoefs of the default Laplacian (its all_terms / nonuniform
   * settings below and the Field2D metric g11, g33, g13, G1, G3), and no Delp2(Jpar) pass is needed. */
  const double *g11, *g33, *g13, *G1, *G3;
  int delp2_all_terms, delp2_nonuniform;
  /* optional scratch of bout_b200_elmpb_rowc_doubles() * nx * ny doubles: the per-row constants are then evaluated
   * by a small kernel ahead of the main one instead of by one thread of every row */
  double* rowc_ws;
  double zlength;
  double hyperresist;
  double keep_U, keep_P, keep_Psi;
  int eta_is2d;
  int filter_n0;
  double *ddt_U, *ddt_P, *ddt_Psi;
  /* optional explicit-RK epilogue (all six or none): rk_stage_X = rk_start_X + rk_cdt * ddt_X on RGN_NOBNDRY, the
   * `k5 = start + c dt k` loops of RK4Solver::take_step (src/solver/impls/rk4/rk4.cxx:179-200) */
  const double *rk_start_U, *rk_start_P, *rk_start_Psi;
  double *rk_stage_U, *rk_stage_P, *rk_stage_Psi;
  double rk_cdt;
} bout_b200_elmpb_rhs_args;
int bout_b200_elmpb_rhs(const bout_b200_grid* g, const bout_b200_elmpb_rhs_args* args, void* stream);
int bout_b200_elmpb_rowc_doubles(void);

/* averageY(Field3D) (src/physics/smoothing.cxx:272-323) in two halves so that the all-reduce over the
 * ranks of a Y communicator sits between them: xz[x][z] = mean over ystart..yend of f;
 * out(x, y, z) = scale * xz[x][z] for every y (scale = 1/NYPE after a SUM all-reduce, else 1). */
int bout_b200_ymean(const bout_b200_grid* g, const double* f, double* xz, void* stream);
int bout_b200_yspread(const bout_b200_grid* g, const double* xz, double scale, double* out,
                      void* stream);
/* smooth_x (src/physics/smoothing.cxx:46-73): 1-2-1 filter in x over 0 < x < nx-1, edge planes copied;
 * the caller communicates the result as the reference does. */
int bout_b200_smooth_x(const bout_b200_grid* g, const double* f, double* out, void* stream);

/* ---- (2b) fused Delp2 and (3) LaplaceCyclic -------------------------------------------------------
 * Laplacian settings + the Field2D metric the coefficients are generated from on chip
 * (reference Laplacian::Laplacian options src/invert/laplace/invert_laplace.cxx:68-138 and
 * Laplacian::tridagCoefs :280-375).  All pointers are device Field2D arrays [nx][ny]; optional ones
 * may be NULL. */
/* ---- Y guard rows delivered by the kernel that produces them, over NVLink peer memory --------------------------------
 * Mesh::communicate's Y exchange (sendY / wait, src/mesh/impls/bout/boutmesh.cxx:1245-1555) without pack, send or
 * unpack passes: every field that takes part lives at the same offset of a per-rank arena (bout_b200_peer_alloc) which
 * the two Y neighbours map with CUDA IPC; the CTAs that produce the top / bottom MYG interior rows of a field also store
 * them straight into the neighbour's guard rows (x guard cells of those rows included), and the last of them raises a
 * flag in the neighbour's arena.  The CTAs that read (or remotely write) next to side d first wait for up to two local
 * flags -- e.g. "the neighbour's rows of the state have arrived" and "its rows of phi have arrived".  Used by
 * bout_b200_hw3d_rhs_guards (through bout_b200_guard_fill) and by the block path of bout_b200_laplace_cyclic_solve
 * (bout_b200_laplace.ypeer; no waits there).  A wait that times out (BOUT_B200_HALO_TIMEOUT_S) sets *status and traps. */
typedef struct bout_b200_ypeer {
  long long delta[2];               /* [0] lower, [1] upper neighbour: (its mapped arena - this rank's arena) in doubles */
  unsigned long long* peer_flag[2]; /* flag in that neighbour's arena, raised to `signal`; NULL = no neighbour there */
  unsigned long long signal;
  unsigned int* counters;           /* two local block counters, zero between launches */
  const unsigned long long* wait_flag[2][2]; /* local flags the CTAs next to side d wait for (NULL = none) ... */
  unsigned long long wait_value[2][2];       /* ... to reach this value */
  int* status;
} bout_b200_ypeer;

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
  int wrap_yguards;   /* single rank, periodic y: the solve also writes each solution row into the y guard row it
                         wraps to -- what the Mesh::communicate(phi) that follows the solve in a model's rhs() would
                         copy (boutmesh.cxx:1221-1555); 0 = the reference's behaviour (guard rows untouched) */
  const bout_b200_ypeer* ypeer; /* Y-decomposed ranks: the solve delivers the top / bottom MYG solution rows into the
                         neighbours' y guard rows and raises their flags (block path only); NULL = off */
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
/* Factorisation reuse.  The tridiagonal matrix of a solve depends on the metric, the A/C/D coefficients
 * and the flags, not on the right-hand side; the reference rebuilds and re-eliminates it in every call
 * (cyclic_laplace.cxx:236-250, cyclic_reduction.hxx:119-167).  _factor computes the coefficient table and
 * the pivots' reciprocals once into `workspace`; every later _solve with the same grid, the same
 * settings (all members of L equal, pointers compared by value) and the same workspace skips that half
 * of the work.  The caller must call _factor again (or _invalidate) whenever it changes the CONTENTS of
 * the metric / coefficient arrays (the host layer does so from setCoefA/C/D and the flag setters).
 * Without a valid factorisation _solve computes one as part of the call (one-shot, reference semantics).
 * Configurations that do not use the fused-sweep kernels ignore _factor (returns 0). */
int bout_b200_laplace_cyclic_factor(const bout_b200_grid* g, const bout_b200_laplace* L,
                                    void* workspace, size_t workspace_bytes, void* stream);
int bout_b200_laplace_cyclic_invalidate(void* workspace);
/* X-decomposed solves (NXPE > 1; replaces the distributed part of CyclicReduce::solve,
 * include/bout/cyclic_reduction.hxx:169-440).  Each rank calls _reduce on its slab (L->first_x /
 * L->last_x say whether it owns a physical boundary): forward z-FFT, local elimination, and
 * bout_b200_laplace_cyclic_iface_doubles() doubles of interface coefficients written to iface_out
 * (device).  The ranks of one X row all-gather those records in PE_XIND order (NCCL; the only
 * exchange of the solve) into iface_all = [nxpe][iface doubles], then each calls _finish: interface
 * system, back substitution, inverse z-FFT into out.  The workspace carries state from _reduce to
 * _finish. */
size_t bout_b200_laplace_cyclic_iface_doubles(const bout_b200_grid* g, const bout_b200_laplace* L);
int bout_b200_laplace_cyclic_reduce(const bout_b200_grid* g, const bout_b200_laplace* L,
                                    const double* rhs, const double* x0, void* workspace,
                                    size_t workspace_bytes, double* iface_out, void* stream);
int bout_b200_laplace_cyclic_finish(const bout_b200_grid* g, const bout_b200_laplace* L,
                                    const double* iface_all, int nxpe, int pe_xind, double* out,
                                    void* workspace, size_t workspace_bytes, void* stream);
/* Synchronises `stream`; BOUT_B200_ERR_PIVOT if the last solve on this workspace met a zero pivot
 * (reference throws BoutException, include/bout/cyclic_reduction.hxx:553). */
int bout_b200_laplace_cyclic_status(void* workspace, void* stream);

/* ---- field algebra (a12) -----------------------------------------------------------------------------
 * out[region] = sum_i coef_i * f1_i * f2_i [/ div_i]  in one pass (1..8 terms).  f1/f2/div may be NULL
 * (factor 1); *_is2d marks a Field2D operand broadcast over z.  Every operator of
 * src/field/generated_fieldops.cxx is one or two such terms: a*b -> {1,a,b}; a+b -> {1,a},{1,b};
 * a-s -> {1,a},{-s}; a/b2d -> {1,a,div=b2d} (multiplies by 1/b2d as the reference, :253-274);
 * 2.*a + b*c (examples/performance/arithmetic/arithmetic.cxx:52) -> {2,a},{1,b,c}. */
typedef struct bout_b200_ewterm {
  double coef;
  const double* f1;
  const double* f2;
  const double* div;
  int f1_is2d, f2_is2d, div_is2d;
} bout_b200_ewterm;
int bout_b200_elementwise(const bout_b200_grid* g, int region, int nterms,
                          const bout_b200_ewterm* terms, double* out, void* stream);
/* Unary functions over RGN_ALL (include/bout/field.hxx:527-634 FIELD_FUNC; field3d.cxx:581 unary minus):
 * 0 neg, 1 sqrt, 2 abs, 3 exp, 4 log, 5 sin, 6 cos, 7 tanh, 8 SQ, 9 pow(f, arg), 10 copy. */
int bout_b200_unary(const bout_b200_grid* g, int op, const double* f, double arg, double* out,
                    void* stream);
/* where(test, gt0, le0) over RGN_ALL (reference include/bout/where.hxx:41-91): out = test > 0 ? gt0 : le0 element by
 * element.  test is a Field3D, or a Field2D broadcast over z (test_is2d); gt0 / le0 are Field3D arrays, or NULL for
 * the constants gt0_value / le0_value. */
int bout_b200_where(const bout_b200_grid* g, const double* test, int test_is2d, const double* gt0, double gt0_value,
                    const double* le0, double le0_value, double* out, void* stream);

/* ---- boundary conditions (a15) ---------------------------------------------------------------------
 * kind: 0 dirichlet / dirichlet_o2 (src/mesh/boundary_standard.cxx:527-581), 1 neumann (:2167-2190),
 * 2 zero the cells (apply_ddt :591-600), 3 extend: every guard cell beyond the first copies its inner
 * neighbour (the phi guard loops of examples/elm-pb/elm_pb.cxx:1498-1516), 4 free_o2 / 5 free_o3: every guard cell,
 * outwards, extrapolated from the 2 / 3 cells inside it (boundary_standard.cxx:3281-3293, :3545-3558; elm-pb's shipped
 * bndry_yup/ydown = free_o3).  side: 0 x-inner, 1 x-outer, 2 y-lower, 3 y-upper.
 * spacing2d: dx for x sides, dy for y sides (Neumann only).  lo/hi: range of the other index
 * (y for x sides, x for y sides); pass -1,-1 for the reference's default (ystart..yend / all x). */
int bout_b200_boundary(const bout_b200_grid* g, int kind, int side, double value,
                       const double* spacing2d, int lo, int hi, double* f, void* stream);
/* Several boundary conditions in ONE launch (Solver::pre_rhs applies every evolving field's conditions on every side,
 * src/solver/solver.cxx:1486-1500: up to 4 * nfields small passes in the reference).  Each entry is what one
 * bout_b200_boundary call takes; entries must not depend on each other (at most one per field and side; x sides and
 * y sides of one field only together when no entry's cells are another's inputs). */
typedef struct bout_b200_bndry_spec {
  double* f;
  int kind, side;
  double value;
  const double* spacing2d;
  int lo, hi;
} bout_b200_bndry_spec;
int bout_b200_boundary_batch(const bout_b200_grid* g, int n, const bout_b200_bndry_spec* specs, void* stream);
/* zerolaplace (BoundaryZeroLaplace::apply(Field3D), src/mesh/boundary_standard.cxx:2630-2686): guard rows
 * continue the last interior row with the decaying solution of g11 d2f/dx2 - g33 kz^2 f = 0 per z mode and
 * a straight line for kz = 0.  X sides only (side 0 or 1; 2 = both sides in one launch); dx2d, g11, g33 Field2D;
 * lo/hi as above. */
int bout_b200_boundary_zerolaplace(const bout_b200_grid* g, int side, const double* dx2d,
                                   const double* g11, const double* g33, double zlength, int lo, int hi,
                                   double* f, void* stream);

/* ---- guard-cell exchange (a13, Mesh::communicate) ------------------------------------------------------
 * Self exchange when the rank is its own neighbour (periodic Y with NYPE=1, periodicX with NXPE=1):
 * Y first, then X with corners (src/mesh/mesh.cxx:351-362, boutmesh.cxx:1221-1555).  Up to 8 fields. */
int bout_b200_halo_self(const bout_b200_grid* g, int nfields, double* const* fields, int periodic_x,
                        int periodic_y, void* stream);
/* y-guard rows y0..y0+nrows-1, x in [xlo,xhi] <-> contiguous buffer [field][x][y][z]
 * (pack_data / unpack_data, boutmesh.cxx:2400-2466).  X guard planes need no packing. */
int bout_b200_halo_pack_y(const bout_b200_grid* g, int nfields, double* const* fields, int y0,
                          int nrows, int xlo, int xhi, double* buf, void* stream);
int bout_b200_halo_unpack_y(const bout_b200_grid* g, int nfields, double* const* fields, int y0,
                            int nrows, int xlo, int xhi, const double* buf, void* stream);
/* The Y exchange over NVLink peer memory, no communication library on the data path (src/mesh/impls/bout/
 * boutmesh.cxx:1245-1555 sendY/wait): buffers from bout_b200_peer_alloc are exported as 64-byte CUDA IPC handles
 * (exchanged by the host out of band) and opened by the neighbour rank.  bout_b200_halo_push_y packs rows
 * y0..y0+nrows-1 straight into the neighbour's receive buffer and then stores `value` to the neighbour's flag;
 * bout_b200_halo_wait blocks the stream until this rank's own flag reaches `value` (bounded spin; *status = 1
 * on time-out), after which bout_b200_halo_unpack_y reads the local receive buffer.  local_counter: one zeroed
 * unsigned int per concurrent push. */
int bout_b200_peer_alloc(size_t bytes, void** ptr);
int bout_b200_peer_free(void* ptr);
int bout_b200_peer_export(void* ptr, unsigned char* handle64);
int bout_b200_peer_open(const unsigned char* handle64, void** ptr);
int bout_b200_peer_close(void* ptr);
int bout_b200_halo_push_y(const bout_b200_grid* g, int nfields, double* const* fields, int y0, int nrows,
                          int xlo, int xhi, double* peer_buf, unsigned long long* peer_flag,
                          unsigned long long value, unsigned int* local_counter, void* stream);
int bout_b200_halo_wait(const unsigned long long* local_flag, unsigned long long value, int* status,
                        void* stream);
/* The same, both directions per launch (index 0 / 1 of y0, buf, flag = the two neighbours; a NULL buffer skips
 * that side): one push launch (pack + flag) and one receive launch (wait on the local flags + unpack) per exchange. */
int bout_b200_halo_push_y2(const bout_b200_grid* g, int nfields, double* const* fields, int nrows, const int* y0,
                           double* const* peer_buf, unsigned long long* const* peer_flag,
                           unsigned long long value, unsigned int* local_counters, void* stream);
int bout_b200_halo_wait_unpack_y2(const bout_b200_grid* g, int nfields, double* const* fields, int nrows,
                                  const int* y0, double* const* local_buf, unsigned long long* const* local_flag,
                                  unsigned long long value, int* status, void* stream);

/* ---- (4) explicit integrator stage updates --------------------------------------------------------
 * mode 0: out = start + c*dt*k1                      (rk4.cxx:179-200; rk3-ssp.cxx:107 with c=1)
 * mode 1: out = start + (1/6)*dt*(k1+2k2+2k3+k4)     (rk4.cxx:206-209)
 * mode 2: out = 0.75*start + 0.25*k1 + 0.25*dt*k2    (rk3-ssp.cxx:115-117; k1=u1, k2=L)
 * mode 3: out = (1/3)*start + (2/3)*(k1 + dt*k2)     (rk3-ssp.cxx:125-127; k1=u2, k2=L) */
int bout_b200_rk_stage(const bout_b200_grid* g, int region, int mode, double dt, double c,
                       const double* start, const double* k1, const double* k2, const double* k3,
                       const double* k4, double* out, void* stream);

/* out[region] = sum_{j<n} coef[j] * x[j]  (1 <= n <= 8, terms added left to right; out must not alias an x[j]):
 * the state and output loops of the other explicit integrators, one pass instead of one per term --
 * EulerSolver::take_step (src/solver/impls/euler/euler.cxx:178-181), RKScheme::setCurState / constructOutput
 * (src/solver/impls/rkgeneric/rkscheme.cxx:57-84, :173-190), SplitRK::take_diffusion_step / take_advection_step
 * (src/solver/impls/split-rk/split-rk.cxx:226-300). */
int bout_b200_lincomb(const bout_b200_grid* g, int region, int n, const double* coef,
                      const double* const* x, double* out, void* stream);
/* result[0] (+)= sum over region of |a-b| / (|a| + |b| + atol): the local error sum of the adaptive integrators
 * (rk4.cxx:103-106, rkscheme.cxx:150-156, split-rk.cxx:128-133); the caller divides by the number of evolving
 * values.  scratch: bout_b200_err_sum_scratch_doubles() doubles of device memory; result: one device double
 * (accumulate != 0 adds to it, so the fields of a state sum into one number).  The summation order is fixed. */
size_t bout_b200_err_sum_scratch_doubles(void);
int bout_b200_err_sum(const bout_b200_grid* g, int region, const double* a, const double* b, double atol,
                      double* scratch, int accumulate, double* result, void* stream);
/* result[0] = max over region of |a-b| (accumulate != 0: max with the value already there): get_error of the
 * Adams-Bashforth solver (src/solver/impls/adams_bashforth/adams_bashforth.cxx:229-253).  Same scratch as above. */
int bout_b200_err_max(const bout_b200_grid* g, int region, const double* a, const double* b,
                      double* scratch, int accumulate, double* result, void* stream);

/* ---- fused RHS kernels (2)+(4): whole model right-hand sides in one pass ---------------------------------
 * Field2D metric bundle (reference include/bout/coordinates.hxx:83-116); unused members may be NULL. */
typedef struct bout_b200_metric {
  const double *dx, *dy, *dz;
  const double *J, *g_22;
  const double *g11, *g33, *g13;
  const double *G1, *G3;
  const double* d1_dx; /* NULL when coords->non_uniform is false */
} bout_b200_metric;

/* Optional stage update written by the fused kernel for its two evolving fields e = 0,1:
 *   state_out[e] = start[e] + a*ddt_e      acc_out[e] = (acc_in[e] ? acc_in[e] : start[e]) + b*ddt_e
 * over RGN_NOBNDRY; NULL outputs are skipped; acc_out may alias acc_in.  With
 * (a,b) = (dt/2,dt/6), (dt/2,dt/3), (dt,dt/3), (-,dt/6) this is RK4 (rk4.cxx:172-210) with the final
 * combination accumulated stage by stage. */
typedef struct bout_b200_rk_epilogue {
  const double* start[2];
  const double* acc_in[2];
  double* state_out[2];
  double* acc_out[2];
  double a, b;
} bout_b200_rk_epilogue;

size_t bout_b200_fused_coef_doubles(const bout_b200_grid* g);
/* Builds the per-(x,y) coefficient table the fused kernels use (spacing factors, Delp2 and
 * Div_par_Grad_par coefficients) once; pass m = NULL to the fused kernels afterwards. */
int bout_b200_fused_coefs(const bout_b200_grid* g, const bout_b200_metric* m, double* coef_workspace,
                          void* stream);
/* examples/hasegawa-wakatani-3d/hw.cxx:58-74 (single-index operators,
 * include/bout/single_index_ops.hxx:93-128 bracket, :201-209 DDZ, :222-246 Delp2, :255-278
 * Div_par_Grad_par): ddt(n), ddt(vort) from n, vort, phi.  ddt_* may be NULL when an epilogue is given.
 * coef_workspace: bout_b200_fused_coef_doubles(g) doubles of device scratch, rebuilt from m by this
 * call when m != NULL, used as already built by bout_b200_fused_coefs when m == NULL. */
int bout_b200_hw3d_rhs(const bout_b200_grid* g, const bout_b200_metric* m, double alpha, double kappa,
                       double Dn, double Dvort, const double* n, const double* vort,
                       const double* phi, double* ddt_n, double* ddt_vort,
                       const bout_b200_rk_epilogue* epilogue, double* coef_workspace, void* stream);
/* The same with the guard cells of the epilogue's output written by the kernel that produces the interior values
 * (the field the next right-hand side reads: state_out, or acc_out where state_out is NULL), so that no
 * separate boundary / exchange pass runs between two stages:
 *   x_inner / x_outer: the constant-value Dirichlet condition Solver::pre_rhs would apply to the evolving fields
 *     before the next rhs() (src/solver/solver.cxx:1380-1400, BoundaryDirichlet::apply src/mesh/
 *     boundary_standard.cxx:527-581: first guard 2 value - f_inside, the others value; rows ystart..yend);
 *   wrap_y: the y guard rows Mesh::communicate copies on a single rank with periodic y (boutmesh.cxx:1221-1555),
 *     x guard cells of those rows included.
 * ystart / yend are the MESH's interior rows; g may select fewer rows (g->ystart..g->yend inside them), e.g. the
 * interior rows while an exchange of phi is in flight and the edge rows after it. */
typedef struct bout_b200_guard_fill {
  double value;
  int x_inner, x_outer, wrap_y;
  int ystart, yend;
  const struct bout_b200_ypeer* ypeer; /* Y-decomposed ranks: deliver the rows to the neighbours instead (wrap_y = 0) */
} bout_b200_guard_fill;
int bout_b200_hw3d_rhs_guards(const bout_b200_grid* g, const bout_b200_metric* m, double alpha, double kappa,
                              double Dn, double Dvort, const double* n, const double* vort,
                              const double* phi, double* ddt_n, double* ddt_vort,
                              const bout_b200_rk_epilogue* epilogue, const bout_b200_guard_fill* guards,
                              double* coef_workspace, void* stream);
/* examples/hasegawa-wakatani/hw.cxx:94-118 convective(), Arakawa bracket; zonal_* = averageY(DC(f))
 * as Field2D when hw:modified, else NULL. */
int bout_b200_hw_convective(const bout_b200_grid* g, const double* dx, const double* dz, double alpha,
                            double kappa, const double* n, const double* vort, const double* phi,
                            const double* zonal_n, const double* zonal_phi, double* ddt_n,
                            double* ddt_vort, const bout_b200_rk_epilogue* epilogue, void* stream);

/* ---- the sibling inversions on the batched tridiagonal core ---------------------------------------
 * Laplacian types "pcr" and "pcr_thomas" (src/invert/laplace/impls/pcr/pcr.cxx, impls/pcr_thomas/pcr_thomas.cxx) solve
 * the systems of LaplaceCyclic -- same FFTTransform, same tridagMatrix -- with a different MPI elimination order: their
 * factory entries bind bout_b200_laplace_cyclic_solve.  (tests/golden/invert.npz pins that against each solver's OWN
 * reference output; the reference's pcr_thomas with INVERT_BNDRY_ONE and two guard cells leaves the outer guard row
 * out of its boundary elimination and differs from cyclic / pcr there: not reproduced.)
 *
 * LaplaceXZ "cyclic" (src/invert/laplacexz/impls/cyclic/laplacexz-cyclic.cxx:63-277): Div(A Grad_perp f) + B f = rhs
 * in X-Z with g13 = 0; rows x = xstart-1 .. xend+1 of every interior y are solved and written, other cells of `out`
 * are left as they are.  A, B and the metric are Field2D arrays. */
typedef struct bout_b200_laplacexz {
  const double *A, *B;
  const double *J, *g11, *g33, *dx;
  double zlength;
  int inner_boundary_flags, outer_boundary_flags; /* INVERT_DC_GRAD | INVERT_AC_GRAD | INVERT_SET | INVERT_RHS */
} bout_b200_laplacexz;
size_t bout_b200_laplacexz_workspace(const bout_b200_grid* g);
int bout_b200_laplacexz_solve(const bout_b200_grid* g, const bout_b200_laplacexz* xz, const double* rhs,
                              const double* x0, double* out, void* workspace, size_t workspace_bytes, void* stream);
/* InvertPar "cyclic" (src/invert/parderiv/impls/cyclic/cyclic.cxx:48-245):
 *   A f + B d2f/dy2 + C d2f/dydz + D d2f/dz2 + E df/dy = rhs   along y for every interior x and kz,
 * on a FIELD-ALIGNED input (the caller applies toFieldAligned / fromFieldAligned; identity for the identity
 * transform).  Coefficient pointers may be NULL (A = 1, the others 0).  sg = DDY(1/sqrt(g_22)) / sqrt(g_22) as the
 * reference's constructor forms it (:55-56).  closed != 0: periodic in y with the twist-shift phase per x
 * (twist_shift NULL = 0), rows ystart..yend written; closed == 0: open field lines, boundary rows in the y guard
 * cells (:208-229), every y written.  d1_dy NULL when coords->non_uniform is false. */
typedef struct bout_b200_invert_par {
  const double *A, *B, *C, *D, *E;
  const double *sg, *g_22, *dy, *d1_dy;
  const double* twist_shift; /* [nx] */
  double zlength;
  int closed;
} bout_b200_invert_par;
size_t bout_b200_invert_par_workspace(const bout_b200_grid* g);
int bout_b200_invert_par_solve(const bout_b200_grid* g, const bout_b200_invert_par* ip, const double* f, double* out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- FieldPerp: one x-z plane [nx][nz] at a y index (include/bout/fieldperp.hxx:54-330) ---------------------------
 * sliceXZ (src/field/fieldperp.cxx) and Field3D::operator=(const FieldPerp&): to_perp != 0 copies plane jy of field3d
 * into perp, to_perp == 0 copies perp into plane jy of field3d.  Element-wise FieldPerp algebra is
 * bout_b200_elementwise on a grid with ny = 1. */
int bout_b200_copy_xz(const bout_b200_grid* g, int jy, double* field3d, double* perp, int to_perp, void* stream);
/* LaplaceCyclic::solve(const FieldPerp& rhs, const FieldPerp& x0) (cyclic_laplace.cxx:120-185): the systems of row jy
 * only; rhs, x0, out are [nx][nz] planes, the Field2D inputs of L stay full [nx][ny] arrays (row jy is used). */
size_t bout_b200_laplace_cyclic_workspace_perp(const bout_b200_grid* g, const bout_b200_laplace* L);
int bout_b200_laplace_cyclic_solve_perp(const bout_b200_grid* g, const bout_b200_laplace* L, int jy, const double* rhs,
                                        const double* x0, double* out, void* workspace, size_t workspace_bytes,
                                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BOUT_B200_H */
