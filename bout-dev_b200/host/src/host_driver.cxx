// C-ABI harness over the C++ host layer, for tests/ and bench.py (ctypes).  It plays the role of
// BOUTMAIN's main() (reference include/bout/physicsmodel.hxx:429-452: BoutInitialise -> model -> solver
// -> solve) with raw arrays in and out, and mirrors the calls of the reference-build test driver so that the same
// test code can drive the reference build and this library side by side.
#include "bout/b200_host.hxx"

#include <dlfcn.h>

#include <cstring>
#include <memory>
#include <string>

namespace {
std::string last_error;
std::unique_ptr<Mesh> mesh_owner;
std::unique_ptr<Solver> solver;
std::unique_ptr<PhysicsModel> model;
void* model_lib = nullptr;

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
} // namespace

extern "C" {
const char* b200host_last_error() { return last_error.c_str(); }

void b200host_finalise() {
  solver.reset();
  model.reset();
  Laplacian::cleanup();
  mesh_owner.reset();
  Options::cleanup();
  if (model_lib != nullptr) {
    dlclose(model_lib);
    model_lib = nullptr;
  }
}

// Decomposition over ranks (bout::b200::YDecomposition): called before b200host_init; nxpe = nype = 1 switches it off.
// The buffers are device memory (`capacity` doubles each, gather_recv nxpe * capacity), owned by the caller for the
// life of the mesh.
int b200host_set_decomposition(int nxpe, int xind, int nype, int yind, int (*exchange)(size_t, int, void*),
                               int (*allgather)(size_t, void*), int (*allreduce_y)(size_t, int, void*),
                               double (*allreduce)(double, int, void*), void* user,
                               double* send_lo, double* send_hi, double* recv_lo, double* recv_hi, size_t capacity,
                               double* gather_send, double* gather_recv) {
  return guarded([&]() {
    bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
    yd = bout::b200::YDecomposition{};
    if (nxpe > 1 || nype > 1) {
      yd.nxpe = nxpe;
      yd.xind = xind;
      yd.nype = nype;
      yd.yind = yind;
      yd.exchange = exchange;
      yd.allgather = allgather;
      yd.allreduce_y = allreduce_y;
      yd.allreduce = allreduce;
      yd.user = user;
      yd.send_lo = send_lo;
      yd.send_hi = send_hi;
      yd.recv_lo = recv_lo;
      yd.recv_hi = recv_hi;
      yd.capacity = capacity;
      yd.gather_send = gather_send;
      yd.gather_recv = gather_recv;
    }
  });
}

// BoutInitialise: options from BOUT.inp text, mesh from options
int b200host_init(const char* ini_text) {
  return guarded([&]() {
    b200host_finalise();
    readOptionsText(ini_text, Options::root());
    mesh_owner.reset(Mesh::create());
  });
}

// v[0..10] = LocalNx, LocalNy, LocalNz, xstart, xend, ystart, yend, GlobalNx, GlobalNy, GlobalNz, periodicX
int b200host_mesh_info(int* v) {
  return guarded([&]() {
    Mesh* m = bout::globals::mesh;
    const int vals[11] = {m->LocalNx, m->LocalNy, m->LocalNz, m->xstart, m->xend, m->ystart, m->yend,
                          m->GlobalNx, m->GlobalNy, m->GlobalNz, m->periodicX ? 1 : 0};
    std::memcpy(v, vals, sizeof(vals));
  });
}

static Field2D* metricByName(const std::string& n) {
  Coordinates* c = bout::globals::mesh->getCoordinates();
#define M(x)   \
  if (n == #x) \
    return &c->x;
  M(dx) M(dy) M(dz) M(J) M(Bxy) M(g11) M(g22) M(g33) M(g12) M(g13) M(g23) M(g_11) M(g_22) M(g_33) M(g_12) M(g_13)
      M(g_23) M(G1) M(G2) M(G3) M(d1_dx) M(d1_dy)
#undef M
          return nullptr;
}
int b200host_get_metric(const char* name, double* out2d) {
  return guarded([&]() {
    Field2D* f = metricByName(name);
    if (f == nullptr) {
      throw BoutException("unknown metric '{}'", name);
    }
    const auto v = f->getHost();
    std::memcpy(out2d, v.data(), v.size() * sizeof(double));
  });
}
int b200host_set_metric(const char* name, const double* in2d) {
  return guarded([&]() {
    Field2D* f = metricByName(name);
    if (f == nullptr) {
      throw BoutException("unknown metric '{}'", name);
    }
    f->setHost(std::vector<double>(in2d, in2d + f->size()));
  });
}

// Loads a model shared object built from an UNCHANGED BOUT++ model source (BOUTMAIN -> factory) and
// attaches it to a solver (any solver:type of Solver::create: rk4, rk3ssp, euler, rkgeneric, splitrk, adams-bashforth).
int b200host_model_create(const char* model_so, const char* solver_type) {
  return guarded([&]() {
    solver.reset();
    model.reset();
    if (model_lib != nullptr) {
      dlclose(model_lib);
    }
    model_lib = dlopen(model_so, RTLD_NOW | RTLD_LOCAL);
    if (model_lib == nullptr) {
      throw BoutException("cannot load model '{}': {}", model_so, dlerror());
    }
    using Factory = PhysicsModel* (*)();
    auto make = reinterpret_cast<Factory>(dlsym(model_lib, "bout_b200_make_model"));
    if (make == nullptr) {
      throw BoutException("'{}' does not export bout_b200_make_model (was it built with BOUTMAIN?)", model_so);
    }
    Options::root()["solver"]["type"].force(std::string(solver_type));
    solver = Solver::create();
    model.reset(make());
    solver->setModel(model.get());
  });
}
int b200host_model_nvars() { return solver ? static_cast<int>(solver->f3d.size()) : 0; }
const char* b200host_model_varname(int i) { return solver->f3d[i].name.c_str(); }
int b200host_model_get(int i, double* out) {
  return guarded([&]() { solver->f3d.at(i).var->getHost(out); });
}
int b200host_model_set(int i, const double* in) {
  return guarded([&]() { solver->f3d.at(i).var->setHost(in); });
}
int b200host_model_get_ddt(int i, double* out) {
  return guarded([&]() { solver->f3d.at(i).F_var->getHost(out); });
}
int b200host_model_rhs(double t) {
  return guarded([&]() {
    if (solver->run_rhs(t) != 0) {
      throw BoutException("run_rhs failed");
    }
    bout::b200::check(bout_b200_stream_sync(nullptr));
  });
}
int b200host_model_solve(int nout, double output_step) {
  return guarded([&]() {
    if (solver->solve(nout, output_step) != 0) {
      throw BoutException("solve failed");
    }
  });
}
/// Waits for the device work queued so far (timing brackets around b200host_model_solve)
int b200host_stream_sync() {
  return guarded([&]() { bout::b200::check(bout_b200_stream_sync(nullptr)); });
}
int b200host_rhs_calls() { return solver ? solver->rhs_calls : 0; }
unsigned long long b200host_pool_allocations() { return bout::b200::poolAllocations(); }
unsigned long long b200host_cow_copies() { return bout::b200::cowCopies(); }

// Diagnostic field registered by the model's outputVars() (e.g. "phi")
int b200host_model_output(const char* name, double* out) {
  return guarded([&]() {
    Options state;
    model->outputVars(state);
    const Field3D* f = state[name].field3d;
    if (f == nullptr) {
      throw BoutException("model does not output '{}'", name);
    }
    f->getHost(out);
  });
}

// Expression -> Field3D array (FieldFactory::create3D) for initial-condition parity checks
int b200host_expr3d(const char* expr, double* out) {
  return guarded([&]() {
    Options& o = Options::root()["b200host_tmp"];
    o["function"].setText(expr);
    o["scale"].setText("1.0");
    Field3D f(bout::globals::mesh);
    f.name = "b200host_tmp";
    initial_profile("b200host_tmp", f); // the initial-profile path without Solver::add's boundary conditions
    f.getHost(out);
  });
}
} // extern "C"
