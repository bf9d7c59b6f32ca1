// Vector2D / Vector3D, Grad, where, interp_to for the B200 host layer (reference include/bout/vector2d.hxx,
// vector3d.hxx, vecops.hxx:47-50, where.hxx, interpolation.hxx:60-206).
//
// A vector is three component fields and a covariant flag; every operation below is ordinary whole-field algebra on the
// components, so with the host layer's deferred, fused evaluation (bout::b200::Block::pending) a dot product such as
// elm_pb.cxx:1760 `b0xcv * Grad(P)` is ONE element-wise pass over the three derivative fields.  Cell locations: the
// host mesh is unstaggered (StaggerGrids = false), every field lives at CELL_CENTRE; interp_to is the identity for
// CELL_CENTRE / CELL_DEFAULT and refuses anything else, as the reference does without staggered grids.
#pragma once

#include "b200_host.hxx"

class Vector3D;
class Vector2D {
public:
  explicit Vector2D(Mesh* m = nullptr) : x(m), y(m), z(m) {}
  Field2D x, y, z;
  bool covariant = true; ///< components are covariant (default, vector2d.hxx:71)
  Mesh* getMesh() const { return x.getMesh(); }
  CELL_LOC getLocation() const { return CELL_CENTRE; }
  Vector2D& setLocation(CELL_LOC loc) {
    if (loc != CELL_CENTRE && loc != CELL_DEFAULT) {
      throw BoutException("Vector2D::setLocation: the B200 host mesh is unstaggered (CELL_CENTRE only)");
    }
    return *this;
  }
  Vector2D& operator=(BoutReal v) {
    x = v;
    y = v;
    z = v;
    return *this;
  }
  void toCovariant();     ///< multiply by g_ij (vector2d.cxx:55-110)
  void toContravariant(); ///< multiply by g^ij
  Vector2D& operator*=(BoutReal s) {
    x *= s;
    y *= s;
    z *= s;
    return *this;
  }
  const Field3D operator*(const Vector3D& rhs) const; ///< dot product
  const Field2D operator*(const Vector2D& rhs) const; ///< dot product
};

class Vector3D {
public:
  explicit Vector3D(Mesh* m = nullptr) : x(m), y(m), z(m) {}
  Field3D x, y, z;
  bool covariant = true; ///< vector3d.hxx:94
  Mesh* getMesh() const { return x.getMesh(); }
  CELL_LOC getLocation() const { return CELL_CENTRE; }
  Vector3D& setLocation(CELL_LOC loc) {
    if (loc != CELL_CENTRE && loc != CELL_DEFAULT) {
      throw BoutException("Vector3D::setLocation: the B200 host mesh is unstaggered (CELL_CENTRE only)");
    }
    return *this;
  }
  Vector3D& operator=(BoutReal v) {
    x = v;
    y = v;
    z = v;
    return *this;
  }
  Vector3D& operator=(const Vector2D& v) {
    x = v.x;
    y = v.y;
    z = v.z;
    covariant = v.covariant;
    return *this;
  }
  void toCovariant();     ///< vector3d.cxx:60-116 (one location)
  void toContravariant(); ///< vector3d.cxx:117-170
  Vector3D& operator+=(const Vector3D& r);
  Vector3D& operator-=(const Vector3D& r);
  Vector3D& operator*=(BoutReal s) {
    x *= s;
    y *= s;
    z *= s;
    return *this;
  }
  Vector3D& operator*=(const Field3D& s) {
    x *= s;
    y *= s;
    z *= s;
    return *this;
  }
  Vector3D& operator*=(const Field2D& s) {
    x *= s;
    y *= s;
    z *= s;
    return *this;
  }
  Vector3D& operator/=(BoutReal s) {
    x /= s;
    y /= s;
    z /= s;
    return *this;
  }
  const Vector3D operator-() const {
    Vector3D r(*this);
    r.x = -x;
    r.y = -y;
    r.z = -z;
    return r;
  }
  const Field3D operator*(const Vector3D& rhs) const; ///< dot product (vector3d.cxx:469-501)
  const Field3D operator*(const Vector2D& rhs) const; ///< dot product (vector3d.cxx:503-537)
};

namespace bout {
namespace b200 {
/// the other vector's components in the same (co / contra) form as `like`
template <class V>
inline V sameForm(const V& v, bool covariant) {
  V r(v);
  if (covariant) {
    r.toCovariant();
  } else {
    r.toContravariant();
  }
  return r;
}
/// a . b for two vectors of the SAME form through the metric (g^ij for covariant pairs, g_ij for contravariant ones),
/// term by term as the reference writes it
template <class R, class A, class B>
inline R metricDot(const A& a, const B& b, bool covariant, Mesh* mesh) {
  Coordinates* m = mesh->getCoordinates();
  if (covariant) {
    R r = a.x * b.x * m->g11 + a.y * b.y * m->g22 + a.z * b.z * m->g33;
    r += (a.x * b.y + a.y * b.x) * m->g12 + (a.x * b.z + a.z * b.x) * m->g13 + (a.y * b.z + a.z * b.y) * m->g23;
    return r;
  }
  R r = a.x * b.x * m->g_11 + a.y * b.y * m->g_22 + a.z * b.z * m->g_33;
  r += (a.x * b.y + a.y * b.x) * m->g_12 + (a.x * b.z + a.z * b.x) * m->g_13 + (a.y * b.z + a.z * b.y) * m->g_23;
  return r;
}
} // namespace b200
} // namespace bout

inline void Vector3D::toCovariant() {
  if (!covariant) {
    Coordinates* m = getMesh()->getCoordinates();
    const Field3D gx = m->g_11 * x + m->g_12 * y + m->g_13 * z;
    const Field3D gy = m->g_22 * y + m->g_12 * x + m->g_23 * z;
    const Field3D gz = m->g_33 * z + m->g_13 * x + m->g_23 * y;
    x = gx;
    y = gy;
    z = gz;
    covariant = true;
  }
}
inline void Vector3D::toContravariant() {
  if (covariant) {
    Coordinates* m = getMesh()->getCoordinates();
    const Field3D gx = m->g11 * x + m->g12 * y + m->g13 * z;
    const Field3D gy = m->g22 * y + m->g12 * x + m->g23 * z;
    const Field3D gz = m->g33 * z + m->g13 * x + m->g23 * y;
    x = gx;
    y = gy;
    z = gz;
    covariant = false;
  }
}
inline void Vector2D::toCovariant() {
  if (!covariant) {
    Coordinates* m = getMesh()->getCoordinates();
    const Field2D gx = m->g_11 * x + m->g_12 * y + m->g_13 * z;
    const Field2D gy = m->g_22 * y + m->g_12 * x + m->g_23 * z;
    const Field2D gz = m->g_33 * z + m->g_13 * x + m->g_23 * y;
    x = gx;
    y = gy;
    z = gz;
    covariant = true;
  }
}
inline void Vector2D::toContravariant() {
  if (covariant) {
    Coordinates* m = getMesh()->getCoordinates();
    const Field2D gx = m->g11 * x + m->g12 * y + m->g13 * z;
    const Field2D gy = m->g22 * y + m->g12 * x + m->g23 * z;
    const Field2D gz = m->g33 * z + m->g13 * x + m->g23 * y;
    x = gx;
    y = gy;
    z = gz;
    covariant = false;
  }
}
inline Vector3D& Vector3D::operator+=(const Vector3D& r) { // vector3d.cxx:283-297: the right-hand side in this form
  const Vector3D o = bout::b200::sameForm(r, covariant);
  x += o.x;
  y += o.y;
  z += o.z;
  return *this;
}
inline Vector3D& Vector3D::operator-=(const Vector3D& r) {
  const Vector3D o = bout::b200::sameForm(r, covariant);
  x -= o.x;
  y -= o.y;
  z -= o.z;
  return *this;
}
inline const Field3D Vector3D::operator*(const Vector3D& rhs) const {
  if (rhs.covariant != covariant) { // one of each: the components multiply (vector3d.cxx:475-477)
    return x * rhs.x + y * rhs.y + z * rhs.z;
  }
  return bout::b200::metricDot<Field3D>(*this, rhs, covariant, getMesh());
}
inline const Field3D Vector3D::operator*(const Vector2D& rhs) const {
  if (rhs.covariant != covariant) {
    return x * rhs.x + y * rhs.y + z * rhs.z;
  }
  return bout::b200::metricDot<Field3D>(*this, rhs, covariant, getMesh());
}
inline const Field3D Vector2D::operator*(const Vector3D& rhs) const { return rhs * (*this); } // vector2d.cxx:343
inline const Field2D Vector2D::operator*(const Vector2D& rhs) const {
  if (rhs.covariant != covariant) {
    return x * rhs.x + y * rhs.y + z * rhs.z;
  }
  return bout::b200::metricDot<Field2D>(*this, rhs, covariant, getMesh());
}
inline const Vector3D operator+(Vector3D a, const Vector3D& b) { return a += b; }
inline const Vector3D operator-(Vector3D a, const Vector3D& b) { return a -= b; }
inline const Vector3D operator*(Vector3D a, BoutReal s) { return a *= s; }
inline const Vector3D operator*(BoutReal s, Vector3D a) { return a *= s; }
inline const Vector3D operator*(Vector3D a, const Field3D& s) { return a *= s; }
inline const Vector3D operator*(const Field3D& s, Vector3D a) { return a *= s; }
inline const Vector3D operator*(Vector3D a, const Field2D& s) { return a *= s; }
inline const Vector3D operator*(const Field2D& s, Vector3D a) { return a *= s; }
inline const Vector3D operator/(Vector3D a, BoutReal s) { return a /= s; }
inline const Vector2D operator*(Vector2D a, BoutReal s) { return a *= s; }
inline const Vector2D operator*(BoutReal s, Vector2D a) { return a *= s; }
/// |v| = sqrt(v . v) (vector3d.cxx:617-619)
inline const Field3D abs(const Vector3D& v) { return sqrt(v * v); }

/// Grad(f): covariant (DDX, DDY, DDZ) (vecops.cxx:69-96)
inline Vector3D Grad(const Field3D& f, CELL_LOC outloc = CELL_DEFAULT, const std::string& method = "DEFAULT") {
  if (outloc != CELL_DEFAULT && outloc != CELL_CENTRE) {
    throw BoutException("Grad: the B200 host mesh is unstaggered (CELL_CENTRE only)");
  }
  Vector3D r(f.getMesh());
  r.x = DDX(f, outloc, method);
  r.y = DDY(f, outloc, method);
  r.z = DDZ(f, outloc, method);
  r.covariant = true;
  return r;
}
/// v . Grad(f) for a contravariant or covariant v, the form vecops.cxx:310-330 reduces to on unstaggered meshes with
/// central differences: the dot product with the covariant gradient
inline Field3D V_dot_Grad_central(const Vector3D& v, const Field3D& f) { return v * Grad(f); }

/// interp_to (interpolation.hxx:60-206): the identity on an unstaggered mesh
template <class T>
inline const T interp_to(const T& var, CELL_LOC loc, const std::string& = "RGN_ALL") {
  if (loc != CELL_DEFAULT && loc != CELL_CENTRE) {
    throw BoutException("interp_to: asked to interpolate to a staggered location, but the B200 host mesh has "
                        "StaggerGrids = false");
  }
  return var;
}
