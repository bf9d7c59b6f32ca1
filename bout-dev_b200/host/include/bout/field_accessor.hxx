// FieldAccessor and the symbolic side of single-index loops (reference include/bout/field_accessor.hxx:20-136,
// include/bout/rajalib.hxx:40-130).
//
// In the reference a FieldAccessor hands raw pointers to a loop body that is compiled for the device by RAJA.  Here
// the model is compiled by the host compiler only and the data lives in HBM, so the body of a BOUT_FOR_RAJA loop is
// run ONCE with a symbolic index: the single-index operators (single_index_ops.hxx) return expression handles
// instead of numbers, `ddt(acc)[i] = ...` records an assignment, and when the body returns the recorded assignments
// are handed to the dispatcher in b200_host.cxx, which launches the fused kernel whose formula they spell
// (hasegawa-wakatani-3d/hw.cxx:58-74 -> bout_b200_hw3d_rhs) or, failing that, one kernel per operator.
//
// A handle converts to BoutReal -- `BoutReal div_current = alpha * Div_par_Grad_par(acc, i);` must keep compiling --
// as a quiet NaN whose payload is the node number; every operator below looks for that payload in its double
// operands.  Arithmetic between two plain doubles of which one is such a NaN would lose the node: the dispatcher
// cannot see it, and the loop's result would be NaN, not a wrong number.
#pragma once

#include "b200_host.hxx"

#include <cstdint>
#include <cstring>

namespace bout {
namespace b200 {
namespace si {
enum Kind { CONST = 0, FIELD, BRACKET, DDX, DDY, DDZ, DELP2, DIV_PAR_GRAD_PAR, ADD, SUB, MUL, DIV, NEG };
struct Node {
  Kind kind;
  double value;      ///< CONST
  const Field3D* f;  ///< first field operand of an operator node
  const Field3D* g;  ///< second field operand (BRACKET)
  int a, b;          ///< child nodes (ADD .. NEG)
};
struct Assignment {
  Field3D* target; ///< the field whose ddt() is assigned
  int node;
  bool accumulate; ///< +=
};
/// The recording of one loop body
struct Recorder {
  std::vector<Node> nodes;
  std::vector<Assignment> assigns;
  bool active = false;
  static Recorder& get();
  int add(const Node& n) {
    if (!active) {
      throw BoutException("single-index operator used outside a BOUT_FOR_RAJA loop");
    }
    nodes.push_back(n);
    return static_cast<int>(nodes.size()) - 1;
  }
};
/// Runs what the recorder holds over `region` (b200_host.cxx)
void dispatch(const Region<Ind3D>& region);

constexpr std::uint64_t kBoxTag = 0x7FF8B20000000000ull, kBoxMask = 0xFFFFFF0000000000ull;
inline double box(int id) {
  const std::uint64_t bits = kBoxTag | static_cast<std::uint32_t>(id);
  double d;
  std::memcpy(&d, &bits, sizeof(d));
  return d;
}
inline bool unbox(double d, int* id) {
  std::uint64_t bits;
  std::memcpy(&bits, &d, sizeof(bits));
  if ((bits & kBoxMask) != kBoxTag) {
    return false;
  }
  *id = static_cast<int>(bits & 0xFFFFFFFFull);
  return true;
}
} // namespace si
} // namespace b200
} // namespace bout

/// The loop index of a recorded body
struct SIndex {};

/// Handle of a recorded expression
class SIExpr {
public:
  explicit SIExpr(int id_) : id(id_) {}
  /// a double operand: a boxed handle or a constant
  static SIExpr of(double v) {
    int n = 0;
    if (bout::b200::si::unbox(v, &n)) {
      return SIExpr(n);
    }
    return SIExpr(bout::b200::si::Recorder::get().add({bout::b200::si::CONST, v, nullptr, nullptr, -1, -1}));
  }
  static SIExpr op(bout::b200::si::Kind k, SIExpr a, SIExpr b) {
    return SIExpr(bout::b200::si::Recorder::get().add({k, 0.0, nullptr, nullptr, a.id, b.id}));
  }
  operator BoutReal() const { return bout::b200::si::box(id); }
  int id;
};
inline SIExpr operator+(SIExpr a, SIExpr b) { return SIExpr::op(bout::b200::si::ADD, a, b); }
inline SIExpr operator-(SIExpr a, SIExpr b) { return SIExpr::op(bout::b200::si::SUB, a, b); }
inline SIExpr operator*(SIExpr a, SIExpr b) { return SIExpr::op(bout::b200::si::MUL, a, b); }
inline SIExpr operator/(SIExpr a, SIExpr b) { return SIExpr::op(bout::b200::si::DIV, a, b); }
// a number on either side (any arithmetic type, so that `expr / 4` does not compete with the built-in operators
// through the conversion to BoutReal)
#define B200_SI_SCALAR_OPS(OP)                                                                   \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>        \
  inline SIExpr operator OP(SIExpr a, T b) {                                                     \
    return a OP SIExpr::of(static_cast<double>(b));                                              \
  }                                                                                              \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>        \
  inline SIExpr operator OP(T a, SIExpr b) {                                                     \
    return SIExpr::of(static_cast<double>(a)) OP b;                                              \
  }
B200_SI_SCALAR_OPS(+)
B200_SI_SCALAR_OPS(-)
B200_SI_SCALAR_OPS(*)
B200_SI_SCALAR_OPS(/)
#undef B200_SI_SCALAR_OPS
inline SIExpr operator-(SIExpr a) { return SIExpr::op(bout::b200::si::NEG, a, SIExpr(-1)); }
inline SIExpr operator+(SIExpr a) { return a; }

/// `ddt(acc)[i]` and `acc.ddt[i]`: the left-hand side of a recorded assignment
struct SIAssignTarget {
  Field3D* field;
  void operator=(SIExpr e) const { bout::b200::si::Recorder::get().assigns.push_back({field, e.id, false}); }
  void operator=(double v) const { *this = SIExpr::of(v); } // a boxed handle or a constant
  void operator+=(SIExpr e) const { bout::b200::si::Recorder::get().assigns.push_back({field, e.id, true}); }
  void operator+=(double v) const { *this += SIExpr::of(v); }
  void operator-=(SIExpr e) const { *this += -e; }
  void operator-=(double v) const { *this += SIExpr::of(-v); }
};
struct BoutRealArray {
  Field3D* field;
  SIAssignTarget operator[](SIndex) const { return SIAssignTarget{field}; }
};

template <CELL_LOC location = CELL_CENTRE, class FieldType = Field3D>
struct FieldAccessor {
  static_assert(std::is_same<FieldType, Field3D>::value, "FieldAccessor: Field3D accessors are implemented on the B200 path");
  FieldAccessor() = delete;
  explicit FieldAccessor(FieldType& f) : field(&f), ddt{&f} {
    if (!f.isAllocated()) {
      throw BoutException("FieldAccessor: field is not allocated");
    }
    Mesh* m = f.getMesh();
    nx = m->LocalNx;
    ny = m->LocalNy;
    nz = m->LocalNz;
    mesh_nz = m->LocalNz;
  }
  /// the field's value at the loop index
  SIExpr operator[](SIndex) const {
    return SIExpr(bout::b200::si::Recorder::get().add({bout::b200::si::FIELD, 0.0, field, nullptr, -1, -1}));
  }
  FieldType* field;
  BoutRealArray ddt; ///< time-derivative data
  int nx = 0, ny = 0, nz = 0, mesh_nz = 0;
};

template <CELL_LOC location, class FieldType>
inline const BoutRealArray& ddt(const FieldAccessor<location, FieldType>& fa) {
  return fa.ddt;
}
