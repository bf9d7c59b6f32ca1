// C++17 host layer: BOUT++'s public API for the physics-model RHS path on top of libbout_b200.so.
//
// Unchanged physics models (examples/hasegawa-wakatani/hw.cxx, examples/blob2d/blob2d.cxx) compile
// against these headers: Field3D / Field2D value types with device-resident, pooled, copy-on-write
// storage (reference include/bout/array.hxx:173-473, src/field/field3d.cxx:74-131), element-wise
// operators (src/field/generated_fieldops.cxx), DDX.. (src/sys/derivs.cxx), bracket / Delp2
// (src/mesh/difops.cxx), Laplacian::create()->solve() (include/bout/invert_laplace.hxx:177-281),
// Mesh::communicate (include/bout/mesh.hxx:260), PhysicsModel / Solver with the explicit rk4 and
// rk3ssp integrators (src/solver/**).  Every numerical operation is a call through the C-ABI of
// include/bout_b200.h; nothing is computed on the host after set-up (options, metrics and initial
// profiles are evaluated once on the host, as in the reference, and uploaded).
#pragma once

#include <cmath>
#include <cstdio>
#include <exception>
#include <functional>
#include <initializer_list>
#include <map>
#include <future>
#include <memory>
#include <sstream>
#include <type_traits>
#include <string>
#include <utility>
#include <vector>

#include "bout_b200.h"

namespace bout {
namespace b200 {
// Minimal "{}" / "{:e}" / "{:d}" / "{:s}" / "{:.3f}" formatter (the reference formats with the third-party fmt
// library; only the replacement-field forms its example models and error messages use are needed here).
inline void formatOne(std::ostringstream& out, const std::string& spec, const std::string& v) { (void)spec; out << v; }
inline void formatOne(std::ostringstream& out, const std::string& spec, const char* v) { (void)spec; out << v; }
template <class T>
inline typename std::enable_if<std::is_arithmetic<T>::value>::type formatOne(std::ostringstream& out,
                                                                              const std::string& spec, T v) {
  if (spec.empty() || spec == "d" || spec == "s") {
    out << v;
    return;
  }
  char buf[64];
  const std::string f = "%" + spec;
  std::snprintf(buf, sizeof(buf), f.c_str(), static_cast<double>(v));
  out << buf;
}
inline void formatRest(std::ostringstream& out, const char* f) {
  for (; *f != 0; ++f) {
    if ((f[0] == '{' && f[1] == '{') || (f[0] == '}' && f[1] == '}')) {
      ++f;
    }
    out << *f;
  }
}
template <class T, class... Rest>
void formatRest(std::ostringstream& out, const char* f, const T& v, const Rest&... rest) {
  for (; *f != 0; ++f) {
    if (f[0] == '{' && f[1] == '{') {
      out << '{';
      ++f;
    } else if (f[0] == '}' && f[1] == '}') {
      out << '}';
      ++f;
    } else if (f[0] == '{') {
      const char* e = f;
      while (*e != 0 && *e != '}') {
        ++e;
      }
      std::string spec(f + 1, e);
      if (!spec.empty() && spec[0] == ':') {
        spec.erase(0, 1);
      }
      formatOne(out, spec, v);
      formatRest(out, *e != 0 ? e + 1 : e, rest...);
      return;
    } else {
      out << *f;
    }
  }
}
template <class... Args>
std::string format(const char* f, const Args&... args) {
  std::ostringstream out;
  formatRest(out, f, args...);
  return out.str();
}
} // namespace b200
} // namespace bout

using BoutReal = double;
#define UNUSED(x)
constexpr BoutReal PI = 3.141592653589793238462643383279502884;
constexpr BoutReal TWOPI = 2 * PI;
/// Physical constants in SI units, the values models normalise with (reference include/bout/constants.hxx:17-30)
namespace SI {
constexpr BoutReal c = 299792458;
constexpr BoutReal mu0 = 4.e-7 * PI;
constexpr BoutReal e0 = 1 / (c * c * mu0);
constexpr BoutReal qe = 1.602176634e-19;
constexpr BoutReal Me = 9.10938356e-31;
constexpr BoutReal Mp = 1.672621898e-27;
constexpr BoutReal kb = 1.38064852e-23;
constexpr BoutReal amu = 1.660539040e-27;
constexpr BoutReal M_Hydrogen = 1.008 * amu;
constexpr BoutReal M_Deuterium = 2.01410178 * amu;
constexpr BoutReal M_Tritium = 3.0160492 * amu;
} // namespace SI

enum CELL_LOC { CELL_DEFAULT = 0, CELL_CENTRE = 1, CELL_CENTER = 1, CELL_XLOW = 2, CELL_YLOW = 3, CELL_ZLOW = 4 };
enum BRACKET_METHOD { BRACKET_STD = 0, BRACKET_SIMPLE = 1, BRACKET_ARAKAWA = 2, BRACKET_CTU = 3, BRACKET_ARAKAWA_OLD = 4 };

// Laplacian flags (reference include/bout/invert_laplace.hxx:72-104)
constexpr int INVERT_DC_GRAD = 1, INVERT_AC_GRAD = 2, INVERT_AC_LAP = 4, INVERT_SYM = 8, INVERT_SET = 16,
              INVERT_RHS = 32, INVERT_DC_LAP = 64, INVERT_BNDRY_ONE = 128, INVERT_DC_GRADPAR = 256,
              INVERT_DC_GRADPARINV = 512, INVERT_IN_CYLINDER = 1024;
constexpr int INVERT_ZERO_DC = 1, INVERT_START_NEW = 2, INVERT_BOTH_BNDRY_ONE = 4, INVERT_4TH_ORDER = 8,
              INVERT_KX_ZERO = 16;

// ---- errors (reference include/bout/boutexception.hxx) ---------------------------------------------
class BoutException : public std::exception {
public:
  explicit BoutException(std::string msg) : message(std::move(msg)) {}
  template <class... Args>
  BoutException(const char* format, const Args&... args) : message(bout::b200::format(format, args...)) {}
  const char* what() const noexcept override { return message.c_str(); }

private:
  std::string message;
};

namespace bout {
namespace b200 {
/// Converts a C-ABI status into the reference's error convention (BoutException)
void check(int status);
} // namespace b200
} // namespace bout

// ---- output (reference include/bout/output.hxx): formatted logging, off by default in library use ----
class Output {
public:
  bool enabled = false;
  template <class T>
  Output& operator<<(const T& v) {
    if (enabled) {
      std::ostringstream s;
      s << v;
      std::fputs(s.str().c_str(), stdout);
    }
    return *this;
  }
  template <class... Args>
  void write(const char* format, const Args&... args) {
    if (enabled) {
      std::fputs(bout::b200::format(format, args...).c_str(), stdout);
    }
  }
  void enable() { enabled = true; }
  void disable() { enabled = false; }
};
extern Output output, output_info, output_warn, output_progress, output_error;

// ---- Options (reference include/bout/options.hxx): tree of sections with string values ---------------
class Field2D;
class Field3D;
class Options {
public:
  struct AttributeType {
    std::string text;
    AttributeType() = default;
    AttributeType(const char* s) : text(s) {}
    AttributeType(const std::string& s) : text(s) {}
    AttributeType(int v) : text(std::to_string(v)) {}
    AttributeType(double v) : text(std::to_string(v)) {}
    AttributeType(bool v) : text(v ? "true" : "false") {}
  };
  static Options& root();
  static Options* getRoot() { return &root(); }
  static void cleanup();

  Options& operator[](const std::string& key);
  Options* getSection(const std::string& key) { return &(*this)[key]; }
  bool isSet(const std::string& key) const;
  bool isSet() const { return has_value; }
  bool isSection() const { return !children.empty(); }
  const std::string& name() const { return full_name; }
  Options& doc(const std::string&) { return *this; }
  void setConditionallyUsed() {}

  template <class T>
  T withDefault(T def) {
    if (!has_value) {
      assignValue(def);
      return def;
    }
    return as<T>();
  }
  std::string withDefault(const char* def) { return withDefault(std::string(def)); }
  template <class T>
  T as() const;
  template <class T>
  void get(const std::string& key, T& var, T def) { var = (*this)[key].withDefault(def); }
  template <class T>
  Options& force(const T& v, const std::string& = "") {
    assignValue(v);
    return *this;
  }
  template <class T>
  Options& operator=(const T& v) {
    assignValue(v);
    return *this;
  }
  /// Diagnostics registration (PhysicsModel::outputVars): kept as a reference to the field
  Options& assignRepeat(const Field3D& f, const std::string& = "t", bool = true, const std::string& = "");
  Options& assignRepeat(const Field2D& f, const std::string& = "t", bool = true, const std::string& = "");
  Options& setAttributes(std::initializer_list<std::pair<std::string, AttributeType>> attrs);

  void setText(const std::string& text) {
    value = text;
    has_value = true;
  }
  const std::string& text() const { return value; }
  std::map<std::string, Options>& sections() { return children; }
  Options* parent() const { return parent_; }
  std::map<std::string, AttributeType> attributes;
  const Field3D* field3d = nullptr;

private:
  template <class T>
  void assignValue(const T& v);
  std::map<std::string, Options> children;
  std::string value, full_name;
  bool has_value = false;
  Options* parent_ = nullptr;
};

template <> std::string Options::as<std::string>() const;
template <> double Options::as<double>() const;
template <> int Options::as<int>() const;
template <> bool Options::as<bool>() const;
template <> void Options::assignValue<std::string>(const std::string& v);
template <> void Options::assignValue<double>(const double& v);
template <> void Options::assignValue<int>(const int& v);
template <> void Options::assignValue<bool>(const bool& v);
template <> void Options::assignValue<Field3D>(const Field3D& v);

/// Reads BOUT.inp text ("key = value", "[section]", "[a:b]" nesting, '#' / ';' comments) into root
void readOptionsText(const std::string& ini_text, Options& root);

#define OPTION(options, var, def) (var) = (*(options))[#var].withDefault(def)

// ---- device storage: pooled, reference-counted blocks (reference include/bout/array.hxx) ---------------
namespace bout {
namespace b200 {
struct Block {
  double* ptr = nullptr;
  size_t count = 0;
  /// Host mirror behind element access from host code (f(x,y,z), f[i]): filled on first use, uploaded again at the
  /// next device access of the block (FieldBase::dptr / wptr are the synchronisation points).
  std::vector<double> host;
  enum Mirror { none = 0, synced = 1, host_newer = 2 } mirror = none;
  /// Stamped from a global counter at creation and at every write intent (FieldBase::makeUnique): (address, version)
  /// names one state of one block's contents.
  unsigned long long version = 0;
  /// How the contents were made, kept while it stays true: op = 1 means contents == a - b element-wise over the whole
  /// array, a and b being the blocks that carried versions a_ver / b_ver.  The single-index loop dispatcher
  /// (single_index_ops.hxx) uses it to see that hw.cxx's phi_minus_n is phi - n without reading any data.
  struct Recipe {
    int op = 0;
    const Block* a = nullptr;
    const Block* b = nullptr;
    unsigned long long a_ver = 0, b_ver = 0;
  } recipe;
  /// Expression fusion: the result of Field3D algebra is not computed when the operator returns.  The block keeps the
  /// combination it stands for, contents == sum_i coef_i * f1_i * f2_i / dv_i over the whole array (the term form of
  /// bout_b200_elementwise), with handles on the operand blocks; a following operator that can absorb those terms
  /// builds a longer combination instead of reading this block, and the ONE kernel launch happens when the device
  /// pointer of a block is first asked for (FieldBase::dptr / wptr, the host mirror).  `ddt(n) = -bracket(phi, n)
  /// + 2 * DDZ(n) * c + D * Delp2(n)` is then one pass over its three operator results instead of five passes with
  /// four temporaries (the reference: one loop per operator, src/field/generated_fieldops.cxx).
  struct LazyTerm {
    double coef = 0.0;
    std::shared_ptr<Block> f1, f2, dv;
    bool f1_2d = false, f2_2d = false, dv_2d = false;
  };
  std::vector<LazyTerm> pending;
  /// Blocks whose pending combination reads this one.  They are launched before this block is written (makeUnique), so
  /// that they see the contents they were built from and their handles do not force a copy-on-write of a block that
  /// is about to be updated in place (Mesh::communicate, boundary conditions, the integrators' stage updates).
  std::vector<std::weak_ptr<Block>> dependents;
  void launchDependents();
  int pgrid[7] = {0, 0, 0, 0, 0, 0, 0}; ///< bout_b200_grid of the pending launch
  void ensure();       ///< launch the pending combination, if any (operands first)
  void syncToHost();   ///< device -> host if the mirror is not current
  void syncToDevice(); ///< host -> device if host code wrote through the mirror
  ~Block();
};
/// Stage update a solver asks the next right-hand side to carry out itself ("axpys fused into the last RHS
/// kernel"): for every evolving variable e,  state_out = start + a ddt_e,  acc_out = (acc_in ? acc_in : start) + b ddt_e
/// over RGN_NOBNDRY.  A fused whole-model kernel that recognises the variables (the hasegawa-wakatani-3d dispatch)
/// passes it on as its bout_b200_rk_epilogue and sets `consumed`; otherwise the solver applies it with its own passes.
struct StageFusion {
  bool active = false, consumed = false;
  std::vector<class ::Field3D*> vars, start, acc_in, state_out, acc_out; ///< per variable; null entries = absent
  double a = 0.0, b = 0.0;
  /// set by the consumer: the time-derivative blocks (not written by the fused kernel) as they were at that moment; a
  /// model that goes on to modify ddt() after its fused loop is detected by the solver through them
  std::vector<const Block*> ddt_blocks;
  std::vector<unsigned long long> ddt_versions;
};
StageFusion& stageFusion();
/// Block of `count` doubles from the size-keyed pool (cudaMalloc only on a pool miss; zero-filled then)
std::shared_ptr<Block> acquire(size_t count);
void releasePool();
size_t poolAllocations(); ///< number of cudaMalloc calls so far (steady state: constant)
size_t cowCopies();       ///< whole-block device copies made by copy-on-write so far (FieldBase::makeUnique)
unsigned long long nextVersion();

/// Decomposition of the host layer's Mesh over NXPE x NYPE ranks (rank = PE_YIND * NXPE + PE_XIND), along y on closed
/// field lines and along x when x is not periodic: BoutMesh::setYDecompositionIndices / setXDecompositionIndices on an all-core topology
/// (src/mesh/impls/bout/boutmesh.cxx:255-341).  One process per GPU; the process that owns the communicator (the
/// Python harness with torch.distributed / NCCL, or any MPI-like host) fills this in BEFORE the mesh is created.
/// Mesh::communicate copies the guard-cell sources of up to `capacity` doubles into send_lo / send_hi (y:
/// bout_b200_halo_pack_y = pack_data, boutmesh.cxx:2400-2466; x: whole planes, contiguous), calls
/// exchange(count, dir) -- which must deliver send_lo to the recv_hi of the rank below and send_hi to the recv_lo of
/// the rank above (dir 0 = y, periodic; dir 1 = x, no wrap: the first rank has nothing below, the last nothing above),
/// ordered after the work already enqueued on the default stream and before work enqueued once it returns (sendY /
/// sendX / wait, boutmesh.cxx:1245-1555) -- and copies recv_lo / recv_hi into the guard cells.
/// An X-decomposed LaplaceCyclic solve (CyclicReduce's distributed part, cyclic_reduction.hxx:169-440) is
/// bout_b200_laplace_cyclic_reduce -> allgather(count): gather_send of every rank into gather_recv [nxpe][count] in
/// PE_XIND order -> bout_b200_laplace_cyclic_finish.
struct YDecomposition {
  int nype = 1, yind = 0;
  int nxpe = 1, xind = 0;
  int (*exchange)(size_t count, int dir, void* user) = nullptr;
  int (*allgather)(size_t count, void* user) = nullptr;
  /// all-reduce (op 0: sum, 1: max) of gather_send[0 .. count) in place over the ranks of one Y column (same PE_XIND):
  /// averageY (smoothing.cxx:238-323, MPI_Allreduce over the y communicator)
  int (*allreduce_y)(size_t count, int op, void* user) = nullptr;
  /// global reduction of one host number over the ranks (op 0: sum, 1: max): the adaptive integrators' error norms
  double (*allreduce)(double value, int op, void* user) = nullptr;
  void* user = nullptr;
  double *send_lo = nullptr, *send_hi = nullptr, *recv_lo = nullptr, *recv_hi = nullptr; ///< device buffers
  size_t capacity = 0;                                                                   ///< doubles in each of them
  double *gather_send = nullptr, *gather_recv = nullptr; ///< device; capacity and nxpe * capacity doubles
  int ranks() const { return nype * nxpe; }
};
YDecomposition& yDecomposition();
} // namespace b200
} // namespace bout

class Mesh;
class Coordinates;
class Solver;
namespace bout {
namespace globals {
extern Mesh* mesh;
}
} // namespace bout

// ---- indices and regions for host-side loops (reference include/bout/region.hxx:168-326, 491-) -------------------
struct Ind3D {
  int ind = 0, ny = 1, nz = 1;
  int x() const { return ind / (ny * nz); }
  int y() const { return (ind / nz) % ny; }
  int z() const { return ind % nz; }
  Ind3D offset(int dx, int dy, int dz) const { // z wraps (region.hxx:297-310); x, y must stay inside the local array
    const int zz = ((z() + dz) % nz + nz) % nz;
    return {((x() + dx) * ny + (y() + dy)) * nz + zz, ny, nz};
  }
  Ind3D xp(int d = 1) const { return offset(d, 0, 0); }
  Ind3D xm(int d = 1) const { return offset(-d, 0, 0); }
  Ind3D yp(int d = 1) const { return offset(0, d, 0); }
  Ind3D ym(int d = 1) const { return offset(0, -d, 0); }
  Ind3D zp(int d = 1) const { return offset(0, 0, d); }
  Ind3D zm(int d = 1) const { return offset(0, 0, -d); }
};
struct Ind2D {
  int ind = 0, ny = 1;
  int x() const { return ind / ny; }
  int y() const { return ind % ny; }
};
template <class I>
class Region {
public:
  std::vector<I> indices;
  typename std::vector<I>::const_iterator begin() const { return indices.begin(); }
  typename std::vector<I>::const_iterator end() const { return indices.end(); }
  size_t size() const { return indices.size(); }
};
/// Host loops over a region (reference BOUT_FOR / BOUT_FOR_SERIAL, region.hxx:124-139): element access goes through the
/// host mirror of the fields touched
#define BOUT_FOR(index, region) for (const auto& index : (region))
#define BOUT_FOR_SERIAL(index, region) for (const auto& index : (region))

// ---- Field2D / Field3D ------------------------------------------------------------------------------
class FieldBase {
public:
  Mesh* getMesh() const;
  CELL_LOC getLocation() const { return CELL_CENTRE; }
  bool isAllocated() const { return static_cast<bool>(data); }
  const double* dptr() const { ///< device pointer for reading (uploads pending host-side element writes first)
    if (!data) {
      return nullptr;
    }
    if (!data->pending.empty()) {
      data->ensure();
    }
    if (data->mirror == bout::b200::Block::host_newer) {
      data->syncToDevice();
    }
    return data->ptr;
  }
  std::string name;
  bout::b200::Block* block() const { return data.get(); }
  const std::shared_ptr<bout::b200::Block>& shared() const { return data; }

protected:
  std::shared_ptr<bout::b200::Block> data;
  Mesh* fieldmesh = nullptr;
  void makeUnique(size_t count); ///< copy-on-write: un-share before writing (field3d.cxx:112-131)
};

class Field2D : public FieldBase {
public:
  Field2D(Mesh* m = nullptr);
  Field2D(BoutReal v, Mesh* m = nullptr);
  Field2D& allocate();
  Field2D& operator=(BoutReal v);
  double* wptr(); ///< device pointer for writing (un-shares)
  size_t size() const;
  void setHost(const std::vector<double>& v); ///< upload [nx][ny]
  std::vector<double> getHost() const;
  /// element access from host code (field2d.hxx:186-258); the z argument of the 3-index form is ignored as there
  BoutReal& operator()(int x, int y);
  const BoutReal& operator()(int x, int y) const;
  BoutReal& operator()(int x, int y, int) { return (*this)(x, y); }
  const BoutReal& operator()(int x, int y, int) const { return (*this)(x, y); }
  BoutReal& operator[](const Ind2D& i);
  const BoutReal& operator[](const Ind2D& i) const;
  BoutReal& operator[](const Ind3D& i) { return (*this)(i.x(), i.y()); }
  const BoutReal& operator[](const Ind3D& i) const { return (*this)(i.x(), i.y()); }
  Field2D& operator+=(const Field2D& r);
  Field2D& operator-=(const Field2D& r);
  Field2D& operator*=(BoutReal s);
};

class FieldPerp;
class Field3D : public FieldBase {
public:
  Field3D(Mesh* m = nullptr);
  Field3D(const Field3D& f);
  Field3D(BoutReal v, Mesh* m = nullptr);
  ~Field3D();
  Field3D& allocate();
  Field3D& operator=(const Field3D& f);
  Field3D& operator=(Field3D&& f) noexcept;
  Field3D& operator=(const Field2D& f);
  Field3D& operator=(const FieldPerp& p); ///< sets plane y = p.getIndex() only (field3d.cxx:293-310)
  Field3D& operator=(BoutReal v);
  double* wptr();
  size_t size() const;
  void setHost(const std::vector<double>& v); ///< upload [nx][ny][nz]
  void setHost(const double* v);
  std::vector<double> getHost() const;
  void getHost(double* out) const;
  /// element access from host code (field3d.hxx:349-420): served from a host mirror of the block; the next device
  /// operation on the field uploads what was written (no arithmetic is done on the host by the library)
  BoutReal& operator()(int x, int y, int z);
  const BoutReal& operator()(int x, int y, int z) const;
  BoutReal& operator[](const Ind3D& i);
  const BoutReal& operator[](const Ind3D& i) const;
  const Region<Ind3D>& getRegion(const std::string& name) const; ///< RGN_ALL | RGN_NOBNDRY | RGN_NOX | RGN_NOY

  Field3D* timeDeriv(); ///< ddt(), created on first use and owned by this field (field3d.cxx:133-138)
  void applyBoundary(bool init = false);          ///< the field's own boundary ops (set by Solver::add)
  void applyBoundary(const std::string& condition);
  void applyTDerivBoundary();
  void setBoundary(const std::string& name); ///< reads [name] bndry_* / [all] bndry_* options

  Field3D& operator+=(const Field3D& r);
  Field3D& operator+=(const Field2D& r);
  Field3D& operator+=(BoutReal r);
  Field3D& operator-=(const Field3D& r);
  Field3D& operator-=(const Field2D& r);
  Field3D& operator-=(BoutReal r);
  Field3D& operator*=(const Field3D& r);
  Field3D& operator*=(const Field2D& r);
  Field3D& operator*=(BoutReal r);
  Field3D& operator/=(const Field3D& r);
  Field3D& operator/=(const Field2D& r);
  Field3D& operator/=(BoutReal r);

  std::vector<std::string> bndry_conditions; ///< [xin, xout, ydown, yup]
private:
  Field3D* deriv = nullptr;
};

inline Field3D& ddt(Field3D& f) { return *f.timeDeriv(); }
template <> Field3D Options::as<Field3D>() const; ///< the option's expression evaluated over the mesh

// ---- FieldPerp: one x-z plane [nx][nz] at a y index (reference include/bout/fieldperp.hxx:54-330) ---------------------
// Device-resident like the other fields (a pooled block of LocalNx * LocalNz doubles); algebra is bout_b200_elementwise
// on a grid with ny = 1, slicing and Field3D = FieldPerp are bout_b200_copy_xz, the Laplacian inversion of one plane is
// bout_b200_laplace_cyclic_solve_perp.
class FieldPerp : public FieldBase {
public:
  FieldPerp(Mesh* m = nullptr, CELL_LOC loc = CELL_CENTRE, int yindex_in = -1);
  FieldPerp(BoutReal v, Mesh* m = nullptr);
  FieldPerp& allocate();
  FieldPerp& operator=(BoutReal v);
  int getIndex() const { return yindex; }          ///< fieldperp.hxx:118
  FieldPerp& setIndex(int y) { yindex = y; return *this; }
  int getNx() const;
  int getNy() const { return 1; }
  int getNz() const;
  double* wptr();
  size_t size() const;
  void setHost(const double* v);                   ///< upload [nx][nz]
  std::vector<double> getHost() const;
  /// element access from host code (fieldperp.hxx:166-216) through the block's host mirror, as for Field3D
  BoutReal& operator()(int x, int z);
  const BoutReal& operator()(int x, int z) const;
  BoutReal& operator()(int x, int, int z) { return (*this)(x, z); }
  const BoutReal& operator()(int x, int, int z) const { return (*this)(x, z); }
  FieldPerp& operator+=(const FieldPerp& r);
  FieldPerp& operator-=(const FieldPerp& r);
  FieldPerp& operator*=(const FieldPerp& r);
  FieldPerp& operator/=(const FieldPerp& r);
  FieldPerp& operator+=(BoutReal r);
  FieldPerp& operator-=(BoutReal r);
  FieldPerp& operator*=(BoutReal r);
  FieldPerp& operator/=(BoutReal r);

private:
  int yindex = -1;
};
FieldPerp sliceXZ(const Field3D& f, int y);        ///< src/field/fieldperp.cxx:150-165
FieldPerp operator+(const FieldPerp& a, const FieldPerp& b);
FieldPerp operator-(const FieldPerp& a, const FieldPerp& b);
FieldPerp operator*(const FieldPerp& a, const FieldPerp& b);
FieldPerp operator/(const FieldPerp& a, const FieldPerp& b);
// FieldPerp (op) Field3D / Field2D use the operand's values at the FieldPerp's y index (generated_fieldops.cxx)
FieldPerp operator+(const FieldPerp& a, const Field3D& b);
FieldPerp operator-(const FieldPerp& a, const Field3D& b);
FieldPerp operator*(const FieldPerp& a, const Field3D& b);
FieldPerp operator/(const FieldPerp& a, const Field3D& b);
FieldPerp operator+(const FieldPerp& a, BoutReal b);
FieldPerp operator-(const FieldPerp& a, BoutReal b);
FieldPerp operator*(const FieldPerp& a, BoutReal b);
FieldPerp operator/(const FieldPerp& a, BoutReal b);
FieldPerp operator+(BoutReal a, const FieldPerp& b);
FieldPerp operator-(BoutReal a, const FieldPerp& b);
FieldPerp operator*(BoutReal a, const FieldPerp& b);
FieldPerp operator/(BoutReal a, const FieldPerp& b);
FieldPerp operator-(const FieldPerp& a);
FieldPerp sqrt(const FieldPerp& f);
FieldPerp abs(const FieldPerp& f);
FieldPerp exp(const FieldPerp& f);
FieldPerp SQ(const FieldPerp& f);

// element-wise algebra (src/field/generated_fieldops.cxx); one kernel per operator
Field3D operator+(const Field3D& a, const Field3D& b);
Field3D operator-(const Field3D& a, const Field3D& b);
Field3D operator*(const Field3D& a, const Field3D& b);
Field3D operator/(const Field3D& a, const Field3D& b);
Field3D operator+(const Field3D& a, const Field2D& b);
Field3D operator-(const Field3D& a, const Field2D& b);
Field3D operator*(const Field3D& a, const Field2D& b);
Field3D operator/(const Field3D& a, const Field2D& b);
Field3D operator+(const Field2D& a, const Field3D& b);
Field3D operator-(const Field2D& a, const Field3D& b);
Field3D operator*(const Field2D& a, const Field3D& b);
Field3D operator/(const Field2D& a, const Field3D& b);
Field3D operator+(const Field3D& a, BoutReal b);
Field3D operator-(const Field3D& a, BoutReal b);
Field3D operator*(const Field3D& a, BoutReal b);
Field3D operator/(const Field3D& a, BoutReal b);
Field3D operator+(BoutReal a, const Field3D& b);
Field3D operator-(BoutReal a, const Field3D& b);
Field3D operator*(BoutReal a, const Field3D& b);
Field3D operator/(BoutReal a, const Field3D& b);
Field3D operator-(const Field3D& a);
Field2D operator+(const Field2D& a, const Field2D& b);
Field2D operator-(const Field2D& a, const Field2D& b);
Field2D operator*(const Field2D& a, const Field2D& b);
Field2D operator/(const Field2D& a, const Field2D& b);
Field2D operator*(const Field2D& a, BoutReal b);
Field2D operator*(BoutReal a, const Field2D& b);
Field2D operator/(const Field2D& a, BoutReal b);
Field2D operator/(BoutReal a, const Field2D& b);
Field2D operator+(const Field2D& a, BoutReal b);
Field2D operator-(const Field2D& a);
Field3D sqrt(const Field3D& f);
Field3D abs(const Field3D& f);
Field3D exp(const Field3D& f);
Field3D log(const Field3D& f);
Field3D sin(const Field3D& f);
Field3D cos(const Field3D& f);
Field3D tanh(const Field3D& f);
Field3D SQ(const Field3D& f);
Field3D pow(const Field3D& f, BoutReal e);
inline BoutReal SQ(BoutReal x) { return x * x; }
Field2D DC(const Field3D& f);                   ///< z-mean (field3d.cxx:785-803)
Field3D filter(const Field3D& f, int N0);       ///< field3d.cxx:620-661
Field3D lowPass(const Field3D& f, int zmax, bool keep_zonal = true); ///< field3d.cxx:664-710
Field2D averageY(const Field2D& f);             ///< smoothing.cxx:238-270
Field3D averageY(const Field3D& f);             ///< smoothing.cxx:272-323
Field3D smooth_x(const Field3D& f);             ///< smoothing.cxx:46-73 (1-2-1 filter in x, then communicate)
inline Field3D emptyFrom(const Field3D& f) { return Field3D(f.getMesh()).allocate(); }
/// where(test, gt0, le0): test > 0 ? gt0 : le0, element by element (reference include/bout/where.hxx:41-91)
Field3D where(const Field3D& test, const Field3D& gt0, const Field3D& le0);
Field3D where(const Field3D& test, const Field3D& gt0, BoutReal le0);
Field3D where(const Field3D& test, BoutReal gt0, const Field3D& le0);
Field3D where(const Field3D& test, BoutReal gt0, BoutReal le0);
Field3D where(const Field2D& test, const Field3D& gt0, const Field3D& le0);
Field3D where(const Field2D& test, const Field3D& gt0, BoutReal le0);
Field3D where(const Field2D& test, BoutReal gt0, const Field3D& le0);

// ---- Coordinates / Mesh ---------------------------------------------------------------------------
class Coordinates {
public:
  explicit Coordinates(Mesh* mesh);
  using FieldMetric = Field2D;
  Field2D dx, dy, dz;
  Field2D J, Bxy;
  Field2D g11, g22, g33, g12, g13, g23;
  Field2D g_11, g_22, g_33, g_12, g_13, g_23;
  Field2D G1, G2, G3;
  Field2D d1_dx, d1_dy;
  bool non_uniform = true;
  BoutReal zlength() const { return zlength_; }
  int geometry(); ///< g_ij, J, d1_dx from the contravariant metric (coordinates.cxx:932-1201; G1..G3 as given)
  /// cached reciprocal spacings used to fuse the "/dx" passes into the stencil kernels
  const Field2D& recip(const std::string& which);
  const Field2D& invSg();                                         ///< 1 / sqrt(g_22) (coordinates.cxx:1879-1886)
  const Field2D& Grad2_par2_DDY_invSg(const std::string& method); ///< DDY(invSg) invSg (coordinates.cxx:1888-1905)
  BoutReal zlength_ = 0.0;

private:
  Mesh* localmesh;
  std::map<std::string, Field2D> cache;
  /// (block, version) of the two source fields each cached entry was formed from: an entry is rebuilt when the model
  /// assigned or modified dx / dy / dz / d1_dx / d1_dy since (no geometry() call needed)
  struct CacheKey {
    const bout::b200::Block* blk[2] = {nullptr, nullptr};
    unsigned long long ver[2] = {0, 0};
  };
  std::map<std::string, CacheKey> cache_key;
};

class Mesh {
public:
  explicit Mesh(Options& opts); ///< BoutMesh::load from options (boutmesh.cxx:456-760, gridfromoptions.cxx)
  ~Mesh();
  static Mesh* create(Options* opts = nullptr);
  int LocalNx, LocalNy, LocalNz, GlobalNx, GlobalNy, GlobalNz;
  int xstart, xend, ystart, yend, zstart = 0, zend;
  int MXG, MYG;
  bool periodicX = false, periodicY = true; ///< periodicY: ixseps >= nx (all core)
  bool symmetricGlobalX = true, symmetricGlobalY = true; ///< mesh:symmetricGlobalX / Y (boutmesh.cxx:72-75)
  bool firstX() const { return PE_XIND == 0; }
  bool lastX() const { return PE_XIND == NXPE - 1; }
  int NYPE = 1, PE_YIND = 0; ///< bout::b200::yDecomposition() at construction
  int NXPE = 1, PE_XIND = 0;
  int OffsetY = 0;           ///< global index of this rank's first interior row minus MYG (boutmesh.cxx:316-341)
  int OffsetX = 0;
  int getNXPE() const { return NXPE; }
  int getXProcIndex() const { return PE_XIND; }
  int getGlobalXIndex(int jx) const { return jx + OffsetX; }
  void exchangeX2D(std::vector<double>& f); ///< guard columns of a host-side [nx][ny] array (NXPE > 1)
  int getNYPE() const { return NYPE; }
  int getYProcIndex() const { return PE_YIND; }
  int getGlobalYIndex(int jy) const { return jy + OffsetY; }
  /// guard rows of a host-side [nx][ny] array from the neighbouring ranks (NYPE > 1)
  void exchangeY2D(std::vector<double>& f);
  BoutReal GlobalX(int jx) const;
  BoutReal GlobalY(int jy) const;
  BoutReal GlobalZ(int jz) const;
  Coordinates* getCoordinates(CELL_LOC = CELL_CENTRE) { return coords.get(); }
  bout_b200_grid grid() const { return {LocalNx, LocalNy, LocalNz, xstart, xend, ystart, yend}; }
  template <class... Ts>
  void communicate(Ts&... fields) {
    std::vector<Field3D*> v{&fields...};
    communicateFields(v);
  }
  void communicateFields(const std::vector<Field3D*>& fields); ///< mesh.cxx:349-378
  const Region<Ind3D>& getRegion3D(const std::string& name); ///< mesh.cxx:650-665
  const Region<Ind3D>& getRegion(const std::string& name) { return getRegion3D(name); }
  /// default derivative methods per direction: mesh:ddx:first etc. (deriv_store.hxx:308-389, 441-478)
  std::string method(char dir, const std::string& kind) const;
  Options* options;

private:
  std::unique_ptr<Coordinates> coords;
  std::map<std::string, std::string> methods;
  std::map<std::string, Region<Ind3D>> regions3d;
};

// ---- derivatives and differential operators -------------------------------------------------------------
Field3D DDX(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D DDY(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D DDZ(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D D2DX2(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D D2DY2(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D D2DZ2(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D VDDX(const Field3D& v, const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D VDDZ(const Field3D& v, const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D VDDY(const Field3D& v, const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
/// flux derivatives d/dx (v f) (derivs.cxx:501-600: the index operator divided by the spacing)
Field3D FDDX(const Field3D& v, const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D FDDY(const Field3D& v, const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
Field3D FDDZ(const Field3D& v, const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT", const std::string& region = "RGN_NOBNDRY");
/// parallel operators with the identity parallel transform (coordinates.cxx:1517-1625, difops.cxx:46-180):
/// Grad_par = DDY / sqrt(g_22), Vpar_Grad_par = VDDY / sqrt(g_22), Div_par(f) = Bxy Grad_par(f / Bxy),
/// Grad2_par2 = DDY(1/sqrt(g_22))/sqrt(g_22) DDY(f) + D2DY2(f) / g_22
Field3D Grad_par(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT");
Field3D Vpar_Grad_par(const Field3D& v, const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT");
Field3D Div_par(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT");
Field3D Grad2_par2(const Field3D& f, CELL_LOC = CELL_DEFAULT, const std::string& method = "DEFAULT");
/// Div_par(K Grad_par f) (difops.cxx:327-367): K Grad2_par2(f) [+ Div_par(K) Grad_par(f) for a field K]
Field3D Div_par_K_Grad_par(BoutReal kY, const Field3D& f, CELL_LOC = CELL_DEFAULT);
Field3D Div_par_K_Grad_par(const Field3D& kY, const Field3D& f, CELL_LOC = CELL_DEFAULT);
Field3D bracket(const Field3D& f, const Field3D& g, BRACKET_METHOD method = BRACKET_STD, CELL_LOC = CELL_DEFAULT, Solver* = nullptr);
Field3D bracket(const Field3D& f, const Field2D& g, BRACKET_METHOD method = BRACKET_STD, CELL_LOC = CELL_DEFAULT, Solver* = nullptr);
Field3D bracket(const Field2D& f, const Field3D& g, BRACKET_METHOD method = BRACKET_STD, CELL_LOC = CELL_DEFAULT, Solver* = nullptr);
Field3D Delp2(const Field3D& f, CELL_LOC = CELL_DEFAULT, bool useFFT = true);

// ---- Laplacian (LaplaceCyclic) ------------------------------------------------------------------------------
class Laplacian {
public:
  Laplacian(Options* options, CELL_LOC loc, Mesh* mesh);
  ~Laplacian();
  static std::unique_ptr<Laplacian> create(Options* opts = nullptr, CELL_LOC loc = CELL_CENTRE, Mesh* mesh = nullptr, Solver* = nullptr);
  static Laplacian* defaultInstance();
  static void cleanup();
  void setCoefA(const Field2D& v) { A = v; hasA = true; }
  void setCoefA(BoutReal v) { A = v; hasA = (v != 0.0); }
  void setCoefC(const Field2D& v) { C1 = v; C2 = v; hasC = true; }
  void setCoefC(BoutReal v) { C1 = v; C2 = v; hasC = (v != 1.0); }
  void setCoefC(const Field3D& v) { setCoefC(DC(v)); }
  void setCoefC1(const Field2D& v) { C1 = v; hasC = true; }
  void setCoefC2(const Field2D& v) { C2 = v; hasC = true; }
  void setCoefD(const Field2D& v) { D = v; hasD = true; }
  void setCoefD(BoutReal v) { D = v; hasD = (v != 1.0); }
  void setGlobalFlags(int f) { global_flags = f; }
  void setInnerBoundaryFlags(int f) { inner_boundary_flags = f; }
  void setOuterBoundaryFlags(int f) { outer_boundary_flags = f; }
  Field3D solve(const Field3D& b) { return solve(b, b); }
  Field3D solve(const Field3D& b, const Field3D& x0);
  FieldPerp solve(const FieldPerp& b) { return solve(b, b); }   ///< invert_laplace.hxx:246
  FieldPerp solve(const FieldPerp& b, const FieldPerp& x0);     ///< cyclic_laplace.cxx:120-185
  void fill(bout_b200_laplace& L) const;

private:
  Mesh* localmesh;
  Coordinates* coords;
  Field2D A, C1, C2, D;
  bool hasA = false, hasC = false, hasD = false;
  int global_flags, inner_boundary_flags, outer_boundary_flags, maxmode;
  bool all_terms, nonuniform, include_yguards;
  bool pcr_thomas = false;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  void* workspace_perp = nullptr; ///< FieldPerp solves keep their own (the factorisation cache is keyed by workspace)
  size_t workspace_perp_bytes = 0;
  static std::unique_ptr<Laplacian> instance;
};

// ---- LaplaceXZ (reference include/bout/invert/laplacexz.hxx, impls/cyclic) ------------------------------------------
class LaplaceXZ {
public:
  LaplaceXZ(Mesh* m, Options* options, CELL_LOC loc);
  ~LaplaceXZ();
  static std::unique_ptr<LaplaceXZ> create(Mesh* m = nullptr, Options* opt = nullptr, CELL_LOC loc = CELL_CENTRE);
  void setCoefs(const Field2D& A, const Field2D& B);
  void setCoefs(const Field3D& A, const Field3D& B) { setCoefs(DC(A), DC(B)); }
  Field3D solve(const Field3D& b, const Field3D& x0);

private:
  Mesh* localmesh;
  Field2D A, B;
  int inner_boundary_flags = 0, outer_boundary_flags = 0;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
};

// ---- InvertPar (reference include/bout/invert_parderiv.hxx, parderiv/impls/cyclic) ---------------------------------
class InvertPar {
public:
  InvertPar(Options* options, CELL_LOC loc, Mesh* m);
  ~InvertPar();
  static std::unique_ptr<InvertPar> create(Options* opt = nullptr, CELL_LOC loc = CELL_CENTRE, Mesh* m = nullptr);
  Field3D solve(const Field3D& f);
  Field3D solve(const Field3D& f, const Field3D&) { return solve(f); }
  void setCoefA(const Field2D& f) { A = f; hasA = true; }
  void setCoefB(const Field2D& f) { B = f; hasB = true; }
  void setCoefC(const Field2D& f) { C = f; hasC = true; }
  void setCoefD(const Field2D& f) { D = f; hasD = true; }
  void setCoefE(const Field2D& f) { E = f; hasE = true; }
  void setCoefA(const Field3D& f) { setCoefA(DC(f)); }
  void setCoefB(const Field3D& f) { setCoefB(DC(f)); }
  void setCoefC(const Field3D& f) { setCoefC(DC(f)); }
  void setCoefD(const Field3D& f) { setCoefD(DC(f)); }
  void setCoefE(const Field3D& f) { setCoefE(DC(f)); }
  void setCoefA(BoutReal v) { setCoefA(Field2D(v, localmesh)); }
  void setCoefB(BoutReal v) { setCoefB(Field2D(v, localmesh)); }
  void setCoefC(BoutReal v) { setCoefC(Field2D(v, localmesh)); }
  void setCoefD(BoutReal v) { setCoefD(Field2D(v, localmesh)); }
  void setCoefE(BoutReal v) { setCoefE(Field2D(v, localmesh)); }

private:
  Mesh* localmesh;
  Field2D A, B, C, D, E, sg;
  bool hasA = false, hasB = false, hasC = false, hasD = false, hasE = false;
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
};

// ---- output step: dump file + device -> host -> disk pipeline (src/dump_writer.cxx) ------------------
namespace bout {
namespace b200 {
/// NetCDF classic (CDF-2) dump file: dims t (record), x, y, z; integer scalars; t_array, iteration, one (t,x,y,z)
/// double variable per field -- readable with scipy.io.netcdf_file / boutdata.collect like BOUT.dmp.<rank>.nc
struct DumpScalar { ///< written once: int or double
  std::string name;
  bool is_double;
  int i;
  double d;
};
struct DumpOnceField { ///< written once: (x, y) or (x, y, z) doubles, host data
  std::string name;
  bool is3d;
  std::vector<double> data;
};
class DumpFile {
public:
  DumpFile(const std::string& path, int nx, int ny, int nz, const std::vector<DumpScalar>& scalars,
           const std::vector<DumpOnceField>& once, const std::vector<std::string>& fields);
  ~DumpFile();
  void append(double t, int iteration, const std::vector<const double*>& host_fields); ///< one record, host data
  int records() const;

private:
  struct Impl;
  std::unique_ptr<Impl> impl;
};

/// Snapshots leave the device on a second stream while the solver keeps stepping; records reach the file in order
class OutputPipeline {
public:
  /// path: dump file ("" = none); restart_path: restart file rewritten at every snapshot from its first n_evolving
  /// fields ("" = none)
  OutputPipeline(const std::string& path, Mesh* mesh, const std::vector<std::string>& names,
                 std::vector<DumpScalar> extra_scalars = {}, const std::vector<DumpOnceField>& once = {},
                 const std::string& restart_path = "", size_t n_evolving = 0);
  ~OutputPipeline();
  void snapshot(double time, int iteration, const std::vector<const Field3D*>& fields);
  void finish(); ///< drains both stages; every snapshot is on disk afterwards
  int recordsWritten() const;

private:
  struct Slot {
    double* staging = nullptr;  ///< page-locked, nfields * count doubles
    std::vector<Field3D> held;  ///< handles that keep the snapshot's device blocks alive during the copy
    std::future<void> writing;  ///< host thread appending this slot's record
    bool copying = false;
    double time = 0.0;
    int iteration = 0;
  };
  void retire(Slot& s);
  void handOver(Slot& s);
  std::unique_ptr<DumpFile> file;
  void* copy_stream = nullptr;
  Slot slots[2];
  size_t nfields, count = 0;
  unsigned next = 0;
  std::string restart_path;
  size_t n_evolving = 0;
  std::vector<std::string> field_names;
  int dims[3] = {0, 0, 0};
};

/// Variables without a record dimension of a NetCDF classic file (CDF-1 / CDF-2), ints widened to double
struct NcVariable {
  std::vector<int> shape;
  std::vector<double> data;
  bool is_int = false;
};
std::map<std::string, NcVariable> readNcFixedVariables(const std::string& path);
void writeRestartFile(const std::string& path, int nx, int ny, int nz, double time, int iteration,
                      const std::vector<std::string>& names, const std::vector<const double*>& host_fields);
} // namespace b200
} // namespace bout

/// Sets a variable from the [name] (or [all]) `function` / `scale` options (reference src/field/initialprofiles.cxx:33-48)
void initial_profile(const std::string& name, Field3D& var);

/// OpenMP wrappers of the reference (include/bout/openmpwrap.hxx): host loops of a model are serial here
#define BOUT_OMP(...)
#define BOUT_OMP_SAFE(...)
#define BOUT_OMP_PERF(...)

// ---- Solver / PhysicsModel ---------------------------------------------------------------------------------
class PhysicsModel;
class Solver;
/// Output monitor (reference include/bout/monitor.hxx:28-75).  Only the default period (every output step) exists here.
class Monitor {
public:
  explicit Monitor(BoutReal timestep_ = -1) : timestep(timestep_) {}
  virtual ~Monitor() = default;
  virtual int call(Solver* solver, BoutReal time, int iter, int nout) = 0;
  virtual void cleanup() {}
  BoutReal getTimestep() const { return timestep; }

private:
  BoutReal timestep;
};
enum class MonitorPosition { BACK, FRONT };

class Solver {
public:
  explicit Solver(Options* opts);
  virtual ~Solver() = default;
  /// solver:type = rk4 | rk3ssp | euler | rkgeneric (solver:scheme = rkf45 | cashkarp | rk4 | rkf34) | splitrk |
  /// adams-bashforth
  static std::unique_ptr<Solver> create(Options* opts = nullptr);
  void add(Field3D& v, const std::string& name); ///< solver.cxx:210: initial profile + boundary set-up
  void setModel(PhysicsModel* m);
  int solve(int nout = -1, BoutReal output_step = 0.0); ///< solver.cxx:510-623
  int run_rhs(BoutReal t); ///< solver.cxx:1380-1427 incl. the split-operator path
  int run_convective(BoutReal t); ///< solver.cxx:1430-1458
  int run_diffusive(BoutReal t);  ///< solver.cxx:1460-1484
  /// One fixed step of the non-adaptive integrators (rk4, rk3ssp); the others override run_output_step
  virtual void take_step(BoutReal t, BoutReal dt);
  /// Advances simtime to `target` (one output step): the body of the reference solvers' run() loops
  virtual void run_output_step(BoutReal target);
  /// Called once from solve() before the first step (the reference solvers' init())
  virtual void init_solver() {}
  virtual void setMaxTimestep(BoutReal) {} ///< solver.hxx:369; EulerSolver / RKGenericSolver override
  /// solver.cxx:840-867; called at the start (iter = -1) and after every output step, in order; a non-zero return
  /// stops the run with a BoutException as in Solver::call_monitors (solver.cxx:874-900)
  void addMonitor(Monitor* monitor, MonitorPosition pos = MonitorPosition::BACK);
  void removeMonitor(Monitor* monitor);
  struct Var {
    Field3D* var;
    Field3D* F_var;
    std::string name;
  };
  std::vector<Var> f3d;
  BoutReal simtime = 0.0;
  int rhs_calls = 0;   ///< run_rhs + run_convective (rhs_ncalls, solver.cxx:1425,1455)
  int rhs_calls_i = 0; ///< run_diffusive (rhs_ncalls_i, solver.cxx:1482)
  BoutReal getCurrentTimestep() const { return timestep; }

protected:
  PhysicsModel* model = nullptr;
  Options* options;
  BoutReal timestep, output_timestep;
  int nout;
  bool initialised = false;
  bool is_nonsplit_model_diffusive = true; ///< solver.cxx:97-101
public:
  int stage_fusion = -1; ///< rk4: -1 not tried yet, 1 the model's fused kernel takes the stage updates, 0 it does not
protected:
  std::vector<Monitor*> monitors;
  void call_monitors(BoutReal t, int iter, int nout_total);
  void pre_rhs();
  void post_rhs();
};

/// `dump` member of PhysicsModel behind SAVE_REPEAT / SAVE_ONCE (reference bout::DataFileFacade,
/// include/bout/physicsmodel.hxx:91-137): registers variables for the output step; nothing is written unless
/// `[output] path` is set.
class DataFileFacade {
public:
  void addRepeat(Field3D& f, const std::string& name) { repeat3d.push_back({name, &f}); }
  void addRepeat(Field2D&, const std::string& name) { unsupported.push_back(name); }
  void addRepeat(BoutReal&, const std::string& name) { unsupported.push_back(name); }
  void addRepeat(int&, const std::string& name) { unsupported.push_back(name); }
  void addOnce(Field3D& f, const std::string& name) { once3d.push_back({name, &f}); }
  void addOnce(Field2D& f, const std::string& name) { once2d.push_back({name, &f}); }
  void addOnce(BoutReal& v, const std::string& name) { once_real.push_back({name, &v}); }
  void addOnce(int& v, const std::string& name) { once_int.push_back({name, &v}); }
  std::vector<std::pair<std::string, Field3D*>> repeat3d, once3d;
  std::vector<std::pair<std::string, Field2D*>> once2d;
  std::vector<std::pair<std::string, BoutReal*>> once_real;
  std::vector<std::pair<std::string, int*>> once_int;
  std::vector<std::string> unsupported; ///< repeated scalars / Field2D: rejected when the output step is enabled
};

class PhysicsModel {
public:
  virtual ~PhysicsModel() = default;
  Mesh* mesh = nullptr;
  Solver* solver = nullptr;
  DataFileFacade dump;
  void initialise(Solver* s);
  int runRHS(BoutReal t) { return rhs(t); }
  int runConvective(BoutReal t) { return convective(t); }
  int runDiffusive(BoutReal t) { return diffusive(t); }
  bool splitOperator() const { return splitop; }
  virtual void outputVars(Options&) {}
  /// physicsmodel.hxx:299-308: called by the solver at every output (after the model's variables are up to date)
  virtual int outputMonitor(BoutReal, int, int) { return 0; }
  virtual int timestepMonitor(BoutReal, BoutReal) { return 0; }

protected:
  virtual int init(bool restarting) = 0;
  virtual int rhs(BoutReal) { return 1; }
  virtual int convective(BoutReal) { return 1; }
  virtual int diffusive(BoutReal) { return 1; }
  void setSplitOperator(bool split = true) { splitop = split; }
  bool splitop = false;
};

#define B200_SOLVE_FOR1(var) solver->add(var, #var);
#define B200_FE_1(m, a) m(a)
#define B200_FE_2(m, a, ...) m(a) B200_FE_1(m, __VA_ARGS__)
#define B200_FE_3(m, a, ...) m(a) B200_FE_2(m, __VA_ARGS__)
#define B200_FE_4(m, a, ...) m(a) B200_FE_3(m, __VA_ARGS__)
#define B200_FE_5(m, a, ...) m(a) B200_FE_4(m, __VA_ARGS__)
#define B200_FE_6(m, a, ...) m(a) B200_FE_5(m, __VA_ARGS__)
#define B200_GET_FE(_1, _2, _3, _4, _5, _6, NAME, ...) NAME
#define B200_FOR_EACH(m, ...) \
  B200_GET_FE(__VA_ARGS__, B200_FE_6, B200_FE_5, B200_FE_4, B200_FE_3, B200_FE_2, B200_FE_1)(m, __VA_ARGS__)
#define SOLVE_FOR(...) { B200_FOR_EACH(B200_SOLVE_FOR1, __VA_ARGS__) }
#define SOLVE_FOR2(a, b) SOLVE_FOR(a, b)
#define SOLVE_FOR3(a, b, c) SOLVE_FOR(a, b, c)
#define B200_SAVE_REPEAT1(var) dump.addRepeat(var, #var);
#define B200_SAVE_ONCE1(var) dump.addOnce(var, #var);
#define SAVE_REPEAT(...) { B200_FOR_EACH(B200_SAVE_REPEAT1, __VA_ARGS__) }
#define SAVE_ONCE(...) { B200_FOR_EACH(B200_SAVE_ONCE1, __VA_ARGS__) }
#define SAVE_REPEAT2(a, b) SAVE_REPEAT(a, b)
#define SAVE_REPEAT3(a, b, c) SAVE_REPEAT(a, b, c)
#define SAVE_ONCE2(a, b) SAVE_ONCE(a, b)
#define SAVE_ONCE3(a, b, c) SAVE_ONCE(a, b, c)

/// Model factory used by the harness instead of main(): BOUTMAIN(Model) defines it
#ifndef B200_MODEL_FACTORY
#define B200_MODEL_FACTORY bout_b200_make_model
#endif
#define BOUTMAIN(ModelClass) \
  extern "C" PhysicsModel* B200_MODEL_FACTORY() { return new ModelClass(); }
