// Single-index operators on FieldAccessors (reference include/bout/single_index_ops.hxx:53-416), recorded rather than
// evaluated: see field_accessor.hxx.  What each node means numerically is defined by the kernel the dispatcher
// launches for it (bout_b200.h cites the reference lines): bracket = Arakawa (:93-128), DDX / DDY / DDZ = 2nd-order
// central (:152-212), Delp2 = the finite-difference form with G1, G3 and the non-uniform correction (:222-246),
// Div_par_Grad_par on the y neighbours (:258-278).
#pragma once

#include "field_accessor.hxx"

namespace bout {
namespace b200 {
namespace si {
inline SIExpr fieldOp(Kind k, const Field3D* f, const Field3D* g = nullptr) {
  return SIExpr(Recorder::get().add({k, 0.0, f, g, -1, -1}));
}
} // namespace si
} // namespace b200
} // namespace bout

template <CELL_LOC location>
inline SIExpr bracket(const FieldAccessor<location>& f, const FieldAccessor<location>& g, SIndex) {
  return bout::b200::si::fieldOp(bout::b200::si::BRACKET, f.field, g.field);
}
template <CELL_LOC location>
inline SIExpr DDX(const FieldAccessor<location>& f, SIndex) {
  return bout::b200::si::fieldOp(bout::b200::si::DDX, f.field);
}
template <CELL_LOC location>
inline SIExpr DDY(const FieldAccessor<location>& f, SIndex) {
  return bout::b200::si::fieldOp(bout::b200::si::DDY, f.field);
}
template <CELL_LOC location>
inline SIExpr DDZ(const FieldAccessor<location>& f, SIndex) {
  return bout::b200::si::fieldOp(bout::b200::si::DDZ, f.field);
}
template <CELL_LOC location>
inline SIExpr Delp2(const FieldAccessor<location>& f, SIndex) {
  return bout::b200::si::fieldOp(bout::b200::si::DELP2, f.field);
}
template <CELL_LOC location>
inline SIExpr Div_par_Grad_par(const FieldAccessor<location>& f, SIndex) {
  return bout::b200::si::fieldOp(bout::b200::si::DIV_PAR_GRAD_PAR, f.field);
}
