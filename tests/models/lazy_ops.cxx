// Test model of this repository (NOT from the reference): whole-field algebra arranged to exercise the host layer's
// deferred, fused evaluation (bout::b200::Block::pending) -- a combination absorbed by the next operator, an operand
// written in place after an expression that reads it was built, a pending combination as multiplier and as divisor,
// more than eight terms, Field2D operands, constants, and a time derivative read while still pending.  The expected
// values are the plain element-wise formulas (tests/lazy_ops_check.py evaluates them with numpy).
#include <bout/physicsmodel.hxx>

class LazyOps : public PhysicsModel {
private:
  Field3D a, b, c;
  Field2D m2;

protected:
  int init(bool) override {
    SOLVE_FOR(a, b);
    c = 0.0;
    m2 = 0.0;
    for (int x = 0; x < mesh->LocalNx; x++) {
      for (int y = 0; y < mesh->LocalNy; y++) {
        m2(x, y) = 1.0 + 0.01 * x + 0.02 * y;
        for (int z = 0; z < mesh->LocalNz; z++) {
          c(x, y, z) = 0.1 * x + 0.2 * y + 0.05 * z + 1.0;
        }
      }
    }
    return 0;
  }
  int rhs(BoutReal) override {
    Field3D t = a * b;
    Field3D u = t + 2.0 * a - b / m2 + 1.5;
    a(3, 2, 1) = 7.5; // in place, after t and u were built from the old value
    Field3D v = u * c / (a + 4.0);
    Field3D w = v - u * u;
    Field3D s = a + b + c + t + u + v + w + a * c + b * c + 2.0;
    ddt(a) = s / 3.0 - (1.0 - w);
    ddt(b) = -ddt(a) + m2 * u;
    return 0;
  }
};

BOUTMAIN(LazyOps)
