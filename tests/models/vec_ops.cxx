// Test model of this repository (NOT from the reference): Vector2D / Vector3D, Grad, dot products in the three
// co / contravariant pairings, toContravariant, vector algebra, where and interp_to -- the pieces of
// examples/elm-pb/elm_pb.cxx's operator set (b0xcv * Grad(P), elm_pb.cxx:1760) outside the brackets.  The same source
// is compiled into the reference build (oracle/Makefile; golden values: tests/golden/make_golden_invert.py) and
// against the B200 host layer.
#include <bout/physicsmodel.hxx>
#include <bout/interpolation.hxx>
#include <bout/vecops.hxx>
#include <bout/vector2d.hxx>
#include <bout/vector3d.hxx>
#include <bout/where.hxx>

class VecOps : public PhysicsModel {
private:
  Field3D f, g;
  Vector2D b0;
  Field2D s2;

protected:
  int init(bool) override {
    SOLVE_FOR(f, g);
    b0.covariant = false; // contravariant components, as elm-pb reads bxcv (elm_pb.cxx:321)
    b0.x = 0.0;
    b0.y = 0.0;
    b0.z = 0.0;
    s2 = 0.0;
    for (int x = 0; x < mesh->LocalNx; x++) {
      for (int y = 0; y < mesh->LocalNy; y++) {
        b0.x(x, y) = 0.3 + 0.02 * x - 0.01 * y;
        b0.y(x, y) = -0.2 + 0.015 * y;
        b0.z(x, y) = 0.1 * x - 0.05 * y + 0.4;
        s2(x, y) = (x + 2 * y) % 3 == 0 ? -1.0 : 0.5 * (x - y);
      }
    }
    return 0;
  }
  int rhs(BoutReal) override {
    mesh->communicate(f, g);
    Vector3D gf = Grad(f); // covariant
    Vector3D gg = Grad(g);
    // contravariant . covariant: components multiply
    ddt(f) = b0 * gf;
    // covariant . covariant: through g^ij
    Field3D dot = gf * gg;
    // vector algebra and a change of form
    Vector3D v = 2.0 * gf - gg * f;
    v.toContravariant();
    Field3D mixed = v * gg + abs(gg);
    ddt(g) = where(s2, dot, -2.0 * dot) + where(f, 1.5, mixed) + interp_to(f, CELL_CENTRE) * 0.5;
    return 0;
  }
};

BOUTMAIN(VecOps)
