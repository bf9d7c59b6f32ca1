// C-ABI harness around the BOUT++ reference's own C++ API (TEST INFRASTRUCTURE ONLY).
//
// Linked with the reference sources compiled in place from /root/reference (see Makefile) into
// oracle/_ref/libboutref.so.  It lets tests/ and bench.py's cpu_baseline leg feed raw arrays to the
// reference's operators (DDX.., bracket, Delp2, Laplacian::solve, Mesh::communicate, boundary
// conditions, RK solvers) and read raw arrays back.  Nothing in bout-dev_b200/ (the product) links,
// imports or calls this file.
//
// All Field3D arrays are [LocalNx][LocalNy][LocalNz] doubles (z fastest) INCLUDING guard cells;
// Field2D arrays are [LocalNx][LocalNy].

#include <bout/bout.hxx>
#include <bout/boundary_factory.hxx>
#include <bout/boutcomm.hxx>
#include <bout/coordinates_accessor.hxx>
#include <bout/derivs.hxx>
#include <bout/difops.hxx>
#include <bout/fft.hxx>
#include <bout/field_factory.hxx>
#include <bout/invert_laplace.hxx>
#include <bout/invert/laplacexz.hxx>
#include <bout/invert_parderiv.hxx>
#include <bout/mpi_wrapper.hxx>
#include <bout/optionsreader.hxx>
#include <bout/physicsmodel.hxx>
#include <bout/smoothing.hxx>
#include <bout/solver.hxx>
#include <bout/sys/timer.hxx>
#include <bout/cyclic_reduction.hxx>
#include "solver/impls/rk4/rk4.hxx"
#include "solver/impls/rk3-ssp/rk3-ssp.hxx"
#include "solver/impls/euler/euler.hxx"
#include "solver/impls/rkgeneric/rkgeneric.hxx"
#include "solver/impls/split-rk/split-rk.hxx"
#include "solver/impls/adams_bashforth/adams_bashforth.hxx"

#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <unistd.h>

using bout::globals::mesh;

std::unique_ptr<PhysicsModel> boutref_make_hw();
std::unique_ptr<PhysicsModel> boutref_make_hw3d();
std::unique_ptr<PhysicsModel> boutref_make_blob2d();
std::unique_ptr<PhysicsModel> boutref_make_elmpb();
std::unique_ptr<PhysicsModel> boutref_make_imex(); // examples/IMEX/advection-diffusion/imex.cxx
std::unique_ptr<PhysicsModel> boutref_make_eigenbox(); // examples/eigen-box/eigen-box.cxx
std::unique_ptr<PhysicsModel> boutref_make_advreact(); // examples/IMEX/advection-reaction/split_operator.cxx
std::unique_ptr<PhysicsModel> boutref_make_ode_pair(); // tests/models/ode_pair.cxx (this repository's test model)
std::unique_ptr<PhysicsModel> boutref_make_elem_access(); // tests/models/elem_access.cxx
std::unique_ptr<PhysicsModel> boutref_make_perp_ops();    // tests/models/perp_ops.cxx
std::unique_ptr<PhysicsModel> boutref_make_vec_ops();     // tests/models/vec_ops.cxx
std::unique_ptr<PhysicsModel> boutref_make_par_ops();     // tests/models/par_ops.cxx
std::unique_ptr<PhysicsModel> boutref_make_avg_ops();     // tests/models/avg_ops.cxx
std::unique_ptr<PhysicsModel> boutref_make_conduction();  // examples/conduction/conduction.cxx (unchanged)
std::unique_ptr<PhysicsModel> boutref_make_blob2d_outer(); // examples/blob2d-outerloop/blob2d.cxx (unchanged)
int boutref_elmpb_get_impl(PhysicsModel* pm, const char* name, double* out, int* kind);

namespace {
// The reference's own solvers with their (protected) variable registry made reachable, so the
// harness can load states into an unchanged model and read ddt() back.
struct DriverIface {
  virtual ~DriverIface() = default;
  virtual int nvars() const = 0;
  virtual const std::string& varName(int i) const = 0;
  virtual Field3D& var(int i) = 0;
  virtual Field3D& ddtVar(int i) = 0;
  virtual int rhs(BoutReal t) = 0;
  virtual int rhsCalls() = 0;
  virtual Solver* solver() = 0;
};
template <class Base>
class DriverSolver : public Base, public DriverIface {
public:
  explicit DriverSolver(Options* opts) : Base(opts) {}
  int nvars() const override { return static_cast<int>(this->f3d.size()); }
  const std::string& varName(int i) const override { return this->f3d[i].name; }
  Field3D& var(int i) override { return *this->f3d[i].var; }
  Field3D& ddtVar(int i) override { return *this->f3d[i].F_var; }
  int rhs(BoutReal t) override { return this->run_rhs(t); }
  int rhsCalls() override { return this->resetRHSCounter(); } // calls since the last query
  Solver* solver() override { return this; }
};
std::unique_ptr<PhysicsModel> model;
std::unique_ptr<DriverIface> drv;
Solver* activeSolver() { return drv ? drv->solver() : nullptr; }

std::string last_error;
std::unique_ptr<Laplacian> lap;
bool initialised = false;

Field3D toField3D(const double* data) {
  Field3D f(mesh);
  f.allocate();
  std::memcpy(&f(0, 0, 0), data,
              sizeof(double) * mesh->LocalNx * mesh->LocalNy * mesh->LocalNz);
  return f;
}
Field2D toField2D(const double* data) {
  Field2D f(mesh);
  f.allocate();
  std::memcpy(&f(0, 0), data, sizeof(double) * mesh->LocalNx * mesh->LocalNy);
  return f;
}
void fromField3D(const Field3D& f, double* out) {
  std::memcpy(out, &f(0, 0, 0),
              sizeof(double) * mesh->LocalNx * mesh->LocalNy * mesh->LocalNz);
}
void fromField2D(const Field2D& f, double* out) {
  std::memcpy(out, &f(0, 0), sizeof(double) * mesh->LocalNx * mesh->LocalNy);
}

Coordinates::FieldMetric* metricByName(Coordinates* c, const std::string& n) {
#define M(x)   \
  if (n == #x) \
    return &c->x;
  M(dx) M(dy) M(dz) M(J) M(Bxy) M(g11) M(g22) M(g33) M(g12) M(g13) M(g23) M(g_11) M(g_22)
      M(g_33) M(g_12) M(g_13) M(g_23) M(G1) M(G2) M(G3) M(d1_dx) M(d1_dy) M(d1_dz)
          M(ShiftTorsion) M(IntShiftTorsion) M(G1_11) M(G1_22) M(G1_33) M(G1_12) M(G1_13)
              M(G1_23) M(G2_11) M(G2_22) M(G2_33) M(G2_12) M(G2_13) M(G2_23) M(G3_11) M(G3_22)
                  M(G3_33) M(G3_12) M(G3_13) M(G3_23)
#undef M
                      return nullptr;
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    last_error = e.what();
    return 1;
  }
}

BRACKET_METHOD bracketMethod(const std::string& m) {
  if (m == "ARAKAWA") {
    return BRACKET_ARAKAWA;
  }
  if (m == "SIMPLE") {
    return BRACKET_SIMPLE;
  }
  if (m == "CTU") {
    return BRACKET_CTU;
  }
  return BRACKET_STD;
}

// One operator application. in[] are raw arrays (3D unless the op name says otherwise).
Field3D applyOp(const std::string& op_in, const std::string& method, int nin, const double** in,
                double scalar) {
  const std::string m = method.empty() ? "DEFAULT" : method;
  // "name@comm": Mesh::communicate the Field3D inputs first, so that they carry parallel slices
  // (the shifted-metric path of the Y operators); plain names take fields without slices.
  std::string op = op_in;
  bool comm = false;
  if (op.size() > 5 && op.compare(op.size() - 5, 5, "@comm") == 0) {
    comm = true;
    op.resize(op.size() - 5);
  }
  auto mk = [&](int i) {
    Field3D f = toField3D(in[i]);
    if (comm) {
      mesh->communicate(f);
    }
    return f;
  };
  auto f0 = [&]() { return mk(0); };
  auto f1 = [&]() { return mk(1); };
  auto f2 = [&]() { return mk(2); };
  // --- index-space derivatives (index_derivs_interface.hxx:172-270)
  if (op == "index_DDX") return bout::derivatives::index::DDX(f0(), CELL_DEFAULT, m);
  if (op == "index_DDY") return bout::derivatives::index::DDY(f0(), CELL_DEFAULT, m);
  if (op == "index_DDZ") return bout::derivatives::index::DDZ(f0(), CELL_DEFAULT, m);
  if (op == "index_D2DX2") return bout::derivatives::index::D2DX2(f0(), CELL_DEFAULT, m);
  if (op == "index_D2DY2") return bout::derivatives::index::D2DY2(f0(), CELL_DEFAULT, m);
  if (op == "index_D2DZ2") return bout::derivatives::index::D2DZ2(f0(), CELL_DEFAULT, m);
  if (op == "index_D4DX4") return bout::derivatives::index::D4DX4(f0(), CELL_DEFAULT, m);
  if (op == "index_D4DY4") return bout::derivatives::index::D4DY4(f0(), CELL_DEFAULT, m);
  if (op == "index_D4DZ4") return bout::derivatives::index::D4DZ4(f0(), CELL_DEFAULT, m);
  if (op == "index_VDDX") return bout::derivatives::index::VDDX(f0(), f1(), CELL_DEFAULT, m);
  if (op == "index_VDDY") return bout::derivatives::index::VDDY(f0(), f1(), CELL_DEFAULT, m);
  if (op == "index_VDDZ") return bout::derivatives::index::VDDZ(f0(), f1(), CELL_DEFAULT, m);
  if (op == "index_FDDX") return bout::derivatives::index::FDDX(f0(), f1(), CELL_DEFAULT, m);
  if (op == "index_FDDY") return bout::derivatives::index::FDDY(f0(), f1(), CELL_DEFAULT, m);
  if (op == "index_FDDZ") return bout::derivatives::index::FDDZ(f0(), f1(), CELL_DEFAULT, m);
  // --- coordinate derivatives (src/sys/derivs.cxx:57-352)
  if (op == "DDX") return DDX(f0(), CELL_DEFAULT, m);
  if (op == "DDY") return DDY(f0(), CELL_DEFAULT, m);
  if (op == "DDZ") return DDZ(f0(), CELL_DEFAULT, m);
  if (op == "D2DX2") return D2DX2(f0(), CELL_DEFAULT, m);
  if (op == "D2DY2") return D2DY2(f0(), CELL_DEFAULT, m);
  if (op == "D2DZ2") return D2DZ2(f0(), CELL_DEFAULT, m);
  if (op == "D4DX4") return D4DX4(f0(), CELL_DEFAULT, m);
  if (op == "D4DZ4") return D4DZ4(f0(), CELL_DEFAULT, m);
  if (op == "D2DXDZ") return D2DXDZ(f0(), CELL_DEFAULT, m);
  if (op == "VDDX") return VDDX(f0(), f1(), CELL_DEFAULT, m);
  if (op == "VDDY") return VDDY(f0(), f1(), CELL_DEFAULT, m);
  if (op == "VDDZ") return VDDZ(f0(), f1(), CELL_DEFAULT, m);
  if (op == "FDDX") return FDDX(f0(), f1(), CELL_DEFAULT, m);
  if (op == "FDDZ") return FDDZ(f0(), f1(), CELL_DEFAULT, m);
  // --- differential operators (src/mesh/difops.cxx)
  if (op == "bracket") return bracket(f0(), f1(), bracketMethod(m));
  if (op == "bracket_3d2d") return bracket(f0(), toField2D(in[1]), bracketMethod(m));
  if (op == "bracket_2d3d") return bracket(toField2D(in[0]), f1(), bracketMethod(m));
  if (op == "Delp2") return Delp2(f0(), CELL_DEFAULT, true);
  if (op == "Delp2_fd") return Delp2(f0(), CELL_DEFAULT, false);
  if (op == "toFieldAligned") return toFieldAligned(f0());
  if (op == "fromFieldAligned") {
    Field3D a = f0();
    a.setDirectionY(YDirectionType::Aligned);
    return fromFieldAligned(a);
  }
  if (op == "Vpar_Grad_par") return Vpar_Grad_par(f0(), f1());
  if (op == "Grad_par") return Grad_par(f0());
  if (op == "Div_par") return Div_par(f0());
  if (op == "Grad2_par2") return Grad2_par2(f0());
  if (op == "b0xGrad_dot_Grad") return b0xGrad_dot_Grad(f0(), f1());
  if (op == "b0xGrad_dot_Grad_3d2d") return b0xGrad_dot_Grad(f0(), toField2D(in[1]));
  if (op == "b0xGrad_dot_Grad_2d3d") return b0xGrad_dot_Grad(toField2D(in[0]), f1());
  // --- field algebra (src/field/generated_fieldops.cxx)
  if (op == "add") return f0() + f1();
  if (op == "sub") return f0() - f1();
  if (op == "mul") return f0() * f1();
  if (op == "div") return f0() / f1();
  if (op == "mul_2d") return f0() * toField2D(in[1]);
  if (op == "div_2d") return f0() / toField2D(in[1]);
  if (op == "add_2d") return f0() + toField2D(in[1]);
  if (op == "sub_2d") return f0() - toField2D(in[1]);
  if (op == "mul_s") return f0() * scalar;
  if (op == "div_s") return f0() / scalar;
  if (op == "add_s") return f0() + scalar;
  if (op == "sub_s") return f0() - scalar;
  if (op == "s_sub") return scalar - f0();
  if (op == "s_div") return scalar / f0();
  if (op == "neg") return -f0();
  if (op == "arith_2a_bc") return 2. * f0() + f1() * f2(); // examples/performance/arithmetic.cxx:52
  if (op == "DC3d") {
    Field3D r{emptyFrom(f0())};
    r = DC(f0());
    return r;
  }
  if (op == "averageY") return averageY(f0());
  if (op == "averageY_DC") {
    Field3D r{emptyFrom(f0())};
    r = averageY(DC(f0()));
    return r;
  }
  if (op == "smooth_x") return smooth_x(f0());
  if (op == "filter") return filter(f0(), static_cast<int>(scalar));
  if (op == "lowPass") return lowPass(f0(), static_cast<int>(scalar));
  if (op == "shiftZ") {
    Field3D a = f0();
    shiftZ(a, scalar);
    return a;
  }
  if (op == "sqrt") return sqrt(f0());
  if (op == "exp") return exp(f0());
  if (op == "abs") {
    Field3D a = f0();
    return abs(a);
  }
  throw BoutException("ref_driver: unknown op '{}' (nin={})", op, nin);
}
} // namespace

extern "C" {

const char* boutref_last_error() { return last_error.c_str(); }

// Initialise mesh etc. from the text of a BOUT.inp file.
int boutref_init(const char* ini_text) {
  return guarded([&]() {
    if (initialised) {
      throw BoutException("boutref_init: already initialised; call boutref_finalise first");
    }
    static int argc = 1;
    static char arg0[] = "boutref";
    static char* argv_store[] = {arg0, nullptr};
    static char** argv = argv_store;
    BoutComm::setArgs(argc, argv);
    (void)BoutComm::rank(); // triggers MPI_Init in the shim
    output.disable();
    output_info.disable();
    output_progress.disable();
    output_warn.disable();
    output_verbose.disable();
    output_debug.disable();

    char tmpl[] = "/tmp/boutref_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd < 0) {
      throw BoutException("boutref_init: mkstemp failed");
    }
    const size_t n = std::strlen(ini_text);
    if (write(fd, ini_text, n) != static_cast<ssize_t>(n)) {
      close(fd);
      throw BoutException("boutref_init: short write");
    }
    close(fd);
    OptionsReader::getInstance()->read(Options::getRoot(), "{}", tmpl);
    unlink(tmpl);
    Options::root()["output"]["enabled"].force(false);
    Options::root()["restart_files"]["enabled"].force(false);

    bout::globals::mpi = new MpiWrapper();
    bout::globals::mesh = Mesh::create();
    bout::globals::mesh->load();
    initialised = true;
  });
}

void boutref_finalise() {
  if (!initialised) {
    return;
  }
  drv.reset();
  model.reset();
  lap.reset();
  delete bout::globals::mesh;
  bout::globals::mesh = nullptr;
  Laplacian::cleanup();
  Array<BoutReal>::cleanup();
  Array<dcomplex>::cleanup();
  Array<fcmplx>::cleanup();
  Array<int>::cleanup();
  Array<unsigned long>::cleanup();
  BoundaryFactory::cleanup();
  CoordinatesAccessor::clear();
  Timer::cleanup();
  Options::cleanup();
  OptionsReader::cleanup();
  delete bout::globals::mpi;
  bout::globals::mpi = nullptr;
  initialised = false;
}

// v[0..11] = LocalNx, LocalNy, LocalNz, xstart, xend, ystart, yend, GlobalNx, GlobalNy, GlobalNz,
//            periodicX, (unused)
int boutref_mesh_info(int* v) {
  return guarded([&]() {
    v[0] = mesh->LocalNx;
    v[1] = mesh->LocalNy;
    v[2] = mesh->LocalNz;
    v[3] = mesh->xstart;
    v[4] = mesh->xend;
    v[5] = mesh->ystart;
    v[6] = mesh->yend;
    v[7] = mesh->GlobalNx;
    v[8] = mesh->GlobalNy;
    v[9] = mesh->GlobalNz;
    v[10] = mesh->periodicX ? 1 : 0;
    v[11] = 0;
  });
}

int boutref_get_metric(const char* name, double* out2d) {
  return guarded([&]() {
    auto* m = metricByName(mesh->getCoordinates(), name);
    if (m == nullptr) {
      throw BoutException("unknown metric '{}'", name);
    }
    fromField2D(*m, out2d);
  });
}

int boutref_set_metric(const char* name, const double* in2d) {
  return guarded([&]() {
    auto* m = metricByName(mesh->getCoordinates(), name);
    if (m == nullptr) {
      throw BoutException("unknown metric '{}'", name);
    }
    *m = toField2D(in2d);
  });
}

// Recompute Christoffel symbols, G1..G3, d1_dx.. from the current metric (coordinates.cxx:932).
int boutref_geometry() {
  return guarded([&]() { mesh->getCoordinates()->geometry(); });
}

// Evaluate a BOUT++ expression (FieldFactory) into a Field3D / Field2D array.
int boutref_expr3d(const char* expr, double* out) {
  return guarded([&]() { FieldFactory ff{mesh};
    fromField3D(ff.create3D(expr), out); });
}
int boutref_expr2d(const char* expr, double* out) {
  return guarded([&]() { FieldFactory ff{mesh};
    fromField2D(ff.create2D(expr), out); });
}

int boutref_op(const char* op, const char* method, int nin, const double** in, double scalar,
               double* out) {
  return guarded([&]() {
    // Pre-fill so cells the reference leaves unwritten are recognisable by the caller
    Field3D r = applyOp(op, method, nin, in, scalar);
    fromField3D(r, out);
  });
}

// Seconds per call (mean over reps, after 1 warm-up), results discarded.
double boutref_time_op(const char* op, const char* method, int nin, const double** in,
                       double scalar, int reps) {
  double result = -1.0;
  guarded([&]() {
    // Fields are built once outside the timed loop, as examples/performance/*/ do
    std::vector<Field3D> f;
    const std::string sop = op;
    {
      Field3D r = applyOp(sop, method, nin, in, scalar);
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) {
      Field3D r = applyOp(sop, method, nin, in, scalar);
    }
    const auto t1 = std::chrono::steady_clock::now();
    result = std::chrono::duration<double>(t1 - t0).count() / reps;
  });
  return result;
}

// rfft / irfft (src/invert/fft_fftw.cxx:116-394); out has n/2+1 complex (interleaved re,im)
int boutref_rfft(const double* in, int n, double* out) {
  return guarded([&]() { bout::fft::rfft(in, n, reinterpret_cast<dcomplex*>(out)); });
}
int boutref_irfft(const double* in, int n, double* out) {
  return guarded([&]() { bout::fft::irfft(reinterpret_cast<const dcomplex*>(in), n, out); });
}

// CyclicReduce<dcomplex> on nsys systems of length n (cyclic_reduction.hxx:119-637); arrays are
// [nsys][n] interleaved complex.
int boutref_cyclic_reduce(int nsys, int n, int periodic, const double* a, const double* b,
                          const double* c, const double* rhs, double* x) {
  return guarded([&]() {
    CyclicReduce<dcomplex> cr(BoutComm::get(), n);
    cr.setPeriodic(periodic != 0);
    Matrix<dcomplex> A(nsys, n), B(nsys, n), C(nsys, n), R(nsys, n), X(nsys, n);
    const auto* ca = reinterpret_cast<const dcomplex*>(a);
    const auto* cb = reinterpret_cast<const dcomplex*>(b);
    const auto* cc = reinterpret_cast<const dcomplex*>(c);
    const auto* crhs = reinterpret_cast<const dcomplex*>(rhs);
    for (int s = 0; s < nsys; ++s) {
      for (int i = 0; i < n; ++i) {
        A(s, i) = ca[s * n + i];
        B(s, i) = cb[s * n + i];
        C(s, i) = cc[s * n + i];
        R(s, i) = crhs[s * n + i];
      }
    }
    cr.setCoefs(A, B, C);
    cr.solve(R, X);
    auto* cx = reinterpret_cast<dcomplex*>(x);
    for (int s = 0; s < nsys; ++s) {
      for (int i = 0; i < n; ++i) {
        cx[s * n + i] = X(s, i);
      }
    }
  });
}

// Laplacian::tridagCoefs (invert_laplace.cxx:280-375) using the default instance (as Delp2 does);
// abc = {a.re,a.im,b.re,b.im,c.re,c.im}. ccoef/d may be NULL.
int boutref_tridag_coefs(int jx, int jy, int jz, const double* ccoef2d, const double* d2d,
                         double* abc) {
  return guarded([&]() {
    dcomplex a, b, c;
    Field2D cc, dd;
    if (ccoef2d != nullptr) {
      cc = toField2D(ccoef2d);
    }
    if (d2d != nullptr) {
      dd = toField2D(d2d);
    }
    Laplacian::defaultInstance()->tridagCoefs(jx, jy, jz, a, b, c,
                                              ccoef2d != nullptr ? &cc : nullptr,
                                              d2d != nullptr ? &dd : nullptr);
    abc[0] = a.real();
    abc[1] = a.imag();
    abc[2] = b.real();
    abc[3] = b.imag();
    abc[4] = c.real();
    abc[5] = c.imag();
  });
}

// --- Laplacian::create()->solve()
int boutref_laplace_create(const char* section) {
  return guarded([&]() {
    lap.reset();
    lap = Laplacian::create(&Options::root()[section]);
  });
}
// which: 'A','C','D','1' (C1),'2' (C2); data is a Field2D array
int boutref_laplace_set_coef(char which, const double* in2d) {
  return guarded([&]() {
    const Field2D f = toField2D(in2d);
    switch (which) {
    case 'A': lap->setCoefA(f); break;
    case 'C': lap->setCoefC(f); break;
    case 'D': lap->setCoefD(f); break;
    case '1': lap->setCoefC1(f); break;
    case '2': lap->setCoefC2(f); break;
    default: throw BoutException("bad coefficient selector");
    }
  });
}
int boutref_laplace_set_flags(int global, int inner, int outer) {
  return guarded([&]() {
    lap->setGlobalFlags(global);
    lap->setInnerBoundaryFlags(inner);
    lap->setOuterBoundaryFlags(outer);
  });
}
int boutref_laplace_solve(const double* b, const double* x0, double* out) {
  return guarded([&]() {
    Field3D r = (x0 != nullptr) ? lap->solve(toField3D(b), toField3D(x0))
                                : lap->solve(toField3D(b));
    fromField3D(r, out);
  });
}
// LaplaceCyclic::solve(const FieldPerp&, const FieldPerp&) (cyclic_laplace.cxx:120-185) on the y = jy slices of b, x0
int boutref_laplace_solve_perp(int jy, const double* b, const double* x0, double* out_perp) {
  return guarded([&]() {
    const FieldPerp r = lap->solve(sliceXZ(toField3D(b), jy), sliceXZ(toField3D(x0), jy));
    for (int x = 0; x < mesh->LocalNx; ++x) {
      for (int z = 0; z < mesh->LocalNz; ++z) {
        out_perp[x * mesh->LocalNz + z] = r(x, z);
      }
    }
  });
}
double boutref_time_laplace_solve(const double* b, int reps) {
  double result = -1.0;
  guarded([&]() {
    const Field3D rhs = toField3D(b);
    Field3D r = lap->solve(rhs);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) {
      r = lap->solve(rhs);
    }
    const auto t1 = std::chrono::steady_clock::now();
    result = std::chrono::duration<double>(t1 - t0).count() / reps;
  });
  return result;
}

// --- LaplaceXZ::create(mesh, &options[section])->setCoefs(A, B); solve(b, x0)
//     (src/invert/laplacexz/impls/cyclic/laplacexz-cyclic.cxx)
int boutref_laplacexz_solve(const char* section, const double* A2d, const double* B2d, const double* b,
                            const double* x0, double* out) {
  return guarded([&]() {
    auto xz = LaplaceXZ::create(bout::globals::mesh, &Options::root()[section]);
    xz->setCoefs(toField2D(A2d), toField2D(B2d));
    Field3D r = xz->solve(toField3D(b), toField3D(x0));
    // cells the solver does not write are uninitialised: report them as 0
    Field3D full = 0.0;
    const int xs = bout::globals::mesh->xstart - 1, xe = bout::globals::mesh->xend + 1;
    for (int x = xs; x <= xe; ++x) {
      for (int y = bout::globals::mesh->ystart; y <= bout::globals::mesh->yend; ++y) {
        for (int z = 0; z < bout::globals::mesh->LocalNz; ++z) {
          full(x, y, z) = r(x, y, z);
        }
      }
    }
    fromField3D(full, out);
  });
}
// --- InvertPar::create(&options[section])->setCoefA..E; solve(f)
//     (src/invert/parderiv/impls/cyclic/cyclic.cxx); coefficient pointers may be null (left at their defaults)
int boutref_invert_par_solve(const char* section, const double* A, const double* B, const double* C,
                             const double* D, const double* E, const double* f, double* out) {
  return guarded([&]() {
    auto inv = InvertPar::create(&Options::root()[section]);
    if (A != nullptr) { inv->setCoefA(toField2D(A)); }
    if (B != nullptr) { inv->setCoefB(toField2D(B)); }
    if (C != nullptr) { inv->setCoefC(toField2D(C)); }
    if (D != nullptr) { inv->setCoefD(toField2D(D)); }
    if (E != nullptr) { inv->setCoefE(toField2D(E)); }
    const Field3D r = inv->solve(toField3D(f));
    Field3D full = 0.0;
    Mesh* m = bout::globals::mesh;
    for (int x = m->xstart; x <= m->xend; ++x) {
      for (int y = m->ystart; y <= m->yend; ++y) {
        for (int z = 0; z < m->LocalNz; ++z) {
          full(x, y, z) = r(x, y, z);
        }
      }
    }
    fromField3D(full, out);
  });
}

// the same, every y of the interior x range (open field lines: the y guard cells are part of the solve)
int boutref_invert_par_solve_all_y(const char* section, const double* A, const double* B, const double* C,
                                   const double* D, const double* E, const double* f, double* out) {
  return guarded([&]() {
    auto inv = InvertPar::create(&Options::root()[section]);
    if (A != nullptr) { inv->setCoefA(toField2D(A)); }
    if (B != nullptr) { inv->setCoefB(toField2D(B)); }
    if (C != nullptr) { inv->setCoefC(toField2D(C)); }
    if (D != nullptr) { inv->setCoefD(toField2D(D)); }
    if (E != nullptr) { inv->setCoefE(toField2D(E)); }
    const Field3D r = inv->solve(toField3D(f));
    Field3D full = 0.0;
    Mesh* m = bout::globals::mesh;
    for (int x = m->xstart; x <= m->xend; ++x) {
      for (int y = 0; y < m->LocalNy; ++y) {
        for (int z = 0; z < m->LocalNz; ++z) {
          full(x, y, z) = r(x, y, z);
        }
      }
    }
    fromField3D(full, out);
  });
}

// --- Mesh::communicate on nf fields in place (mesh.hxx:260, boutmesh.cxx:1221-1555)
int boutref_communicate(int nf, double** fields) {
  return guarded([&]() {
    std::vector<Field3D> f;
    f.reserve(nf);
    FieldGroup g;
    for (int i = 0; i < nf; ++i) {
      f.push_back(toField3D(fields[i]));
    }
    for (int i = 0; i < nf; ++i) {
      g.add(f[i]);
    }
    mesh->communicate(g);
    for (int i = 0; i < nf; ++i) {
      fromField3D(f[i], fields[i]);
    }
  });
}

// --- Field3D::applyBoundary("condition") in place (field3d.cxx, boundary_standard.cxx)
int boutref_apply_boundary(double* field, const char* condition) {
  return guarded([&]() {
    Field3D f = toField3D(field);
    f.applyBoundary(condition);
    fromField3D(f, field);
  });
}

// yup/ydown slices after communicate (ParallelTransform::calcParallelSlices)
int boutref_parallel_slices(const double* field, double* yup, double* ydown) {
  return guarded([&]() {
    Field3D f = toField3D(field);
    mesh->communicate(f);
    fromField3D(f.yup(), yup);
    fromField3D(f.ydown(), ydown);
  });
}

// one parallel slice (Field3D::ynext(offset)) after communicate
int boutref_parallel_slice(const double* field, int offset, double* out) {
  return guarded([&]() {
    Field3D f = toField3D(field);
    mesh->communicate(f);
    fromField3D(f.ynext(offset), out);
  });
}

// zShift as the ShiftedMetric holds it (guard cells fixed up, coordinates.cxx:1346-1375) and zlength
int boutref_zshift(double* out2d, double* zlength) {
  return guarded([&]() {
    Options dump;
    mesh->getCoordinates()->getParallelTransform().outputVars(dump);
    fromField2D(dump["zShift"].as<Field2D>(), out2d);
    *zlength = mesh->getCoordinates()->zlength()(0, 0);
  });
}

// ---- unchanged example models driven through the reference's own RK solvers --------------------
// name: "hw" (examples/hasegawa-wakatani), "hw3d" (examples/hasegawa-wakatani-3d), "blob2d";
// solver: "rk4", "rk3ssp", "euler", "rkgeneric", "splitrk" or "adams-bashforth" (options section [solver] of the ini text applies)
int boutref_model_create(const char* name, const char* solver) {
  return guarded([&]() {
    drv.reset();
    model.reset();
    const std::string n = name, s = solver;
    Options* sopts = &Options::root()["solver"];
    if (s == "rk3ssp") {
      drv = std::make_unique<DriverSolver<RK3SSP>>(sopts);
    } else if (s == "euler") {
      drv = std::make_unique<DriverSolver<EulerSolver>>(sopts);
    } else if (s == "rkgeneric") {
      drv = std::make_unique<DriverSolver<RKGenericSolver>>(sopts);
    } else if (s == "splitrk") {
      drv = std::make_unique<DriverSolver<SplitRK>>(sopts);
    } else if (s == "adams-bashforth") {
      drv = std::make_unique<DriverSolver<AdamsBashforthSolver>>(sopts);
    } else if (s == "rk4") {
      drv = std::make_unique<DriverSolver<RK4Solver>>(sopts);
    } else {
      throw BoutException("ref_driver: unknown solver '{}'", s);
    }
    if (n == "hw") {
      model = boutref_make_hw();
    } else if (n == "hw3d") {
      model = boutref_make_hw3d();
    } else if (n == "blob2d") {
      model = boutref_make_blob2d();
    } else if (n == "elmpb") {
      model = boutref_make_elmpb();
    } else if (n == "imex") {
      model = boutref_make_imex();
    } else if (n == "eigenbox") {
      model = boutref_make_eigenbox();
    } else if (n == "advreact") {
      model = boutref_make_advreact();
    } else if (n == "ode_pair") {
      model = boutref_make_ode_pair();
    } else if (n == "elem_access") {
      model = boutref_make_elem_access();
    } else if (n == "perp_ops") {
      model = boutref_make_perp_ops();
    } else if (n == "vec_ops") {
      model = boutref_make_vec_ops();
    } else if (n == "par_ops") {
      model = boutref_make_par_ops();
    } else if (n == "avg_ops") {
      model = boutref_make_avg_ops();
    } else if (n == "conduction") {
      model = boutref_make_conduction();
    } else if (n == "blob2d_outer") {
      model = boutref_make_blob2d_outer();
    } else {
      throw BoutException("unknown model '{}'", n);
    }
    activeSolver()->setModel(model.get());
  });
}
// examples/elm-pb only: equilibrium fields / scalars / per-RHS auxiliaries by name (elmpb_export.cxx)
int boutref_elmpb_get(const char* name, double* out, int* kind) {
  int rc = 0;
  const int g = guarded([&]() { rc = boutref_elmpb_get_impl(model.get(), name, out, kind); });
  return g != 0 ? 100 : rc;
}
// OpenMP threads the reference's BOUT_FOR loops run on in this process (bench.py reports what was USED)
int boutref_omp_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
int boutref_model_nvars() { return drv ? drv->nvars() : 0; }
int boutref_model_rhs_calls() { return drv ? drv->rhsCalls() : 0; }
const char* boutref_model_varname(int i) {
  return drv->varName(i).c_str();
}
int boutref_model_get(int i, double* out) {
  return guarded([&]() { fromField3D(drv->var(i), out); });
}
int boutref_model_set(int i, const double* in) {
  return guarded([&]() {
    Field3D& f = drv->var(i);
    f.allocate();
    std::memcpy(&f(0, 0, 0), in, sizeof(double) * mesh->LocalNx * mesh->LocalNy * mesh->LocalNz);
  });
}
int boutref_model_get_ddt(int i, double* out) {
  return guarded([&]() { fromField3D(drv->ddtVar(i), out); });
}
// Solver::run_rhs (src/solver/solver.cxx:1380-1427): pre_rhs boundary conditions, the model's rhs
// (or convective + diffusive for split-operator models), post_rhs.
int boutref_model_rhs(double t) {
  return guarded([&]() {
    const int rc = drv->rhs(t);
    if (rc != 0) {
      throw BoutException("run_rhs returned {}", rc);
    }
  });
}
double boutref_model_time_rhs(int reps) {
  double result = -1.0;
  guarded([&]() {
    drv->rhs(0.0);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) {
      drv->rhs(0.0);
    }
    const auto t1 = std::chrono::steady_clock::now();
    result = std::chrono::duration<double>(t1 - t0).count() / reps;
  });
  return result;
}
// Solver::solve (src/solver/solver.cxx:510-623) -> RK4Solver::run / RK3SSP::run for nout outputs of
// `output_step` each, internal step solver:timestep.
int boutref_model_solve(int nout, double output_step) {
  return guarded([&]() {
    const int rc = activeSolver()->solve(nout, output_step);
    if (rc != 0) {
      throw BoutException("solve returned {}", rc);
    }
  });
}
} // extern "C"
