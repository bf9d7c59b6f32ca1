// Test model of this repository (NOT from the reference): a BOUT_FOR_RAJA single-index loop that is NOT the
// hasegawa-wakatani-3d pattern, so the host layer's loop dispatcher takes its one-kernel-per-operator path, next to
// the same expression written with whole-field operators.  ddt(a) (loop) must equal ddt(b) (field operators):
//   - bracket(g, a) * 1.5 + kz * DDZ(g) - DDX(a) / 4 + a[i] * 0.25 + 2
// with a boxed intermediate (`BoutReal t = ...;`) and a `+=` assignment, as loop bodies of the reference's
// *-outerloop examples use them.
#include <bout/physicsmodel.hxx>
#include <bout/single_index_ops.hxx>
#include <bout/rajalib.hxx>

class SiLoop : public PhysicsModel {
private:
  Field3D a, b, g;
  BoutReal kz = 0.75;

protected:
  int init(bool) override {
    SOLVE_FOR(a, b);
    g = 0.0;
    for (int x = 0; x < mesh->LocalNx; x++) {
      for (int y = 0; y < mesh->LocalNy; y++) {
        for (int z = 0; z < mesh->LocalNz; z++) {
          g(x, y, z) = 0.3 * x - 0.01 * x * x + 0.2 * y + 0.05 * z * (mesh->LocalNz - z);
        }
      }
    }
    return 0;
  }
  int rhs(BoutReal) override {
    b = a; // same state in both
    mesh->communicate(a, b, g);
    auto a_acc = FieldAccessor<>(a);
    auto g_acc = FieldAccessor<>(g);
    BOUT_FOR_RAJA(i, a.getRegion("RGN_NOBNDRY"), CAPTURE(kz)) {
      BoutReal t = kz * DDZ(g_acc, i);
      ddt(a_acc)[i] = -bracket(g_acc, a_acc, i) * 1.5 + t - DDX(a_acc, i) / 4;
      ddt(a_acc)[i] += a_acc[i] * 0.25 + 2;
    };
    ddt(b) = -1.5 * bracket(g, b, BRACKET_ARAKAWA) + kz * DDZ(g, CELL_DEFAULT, "C2") - DDX(b, CELL_DEFAULT, "C2") / 4.0
             + 0.25 * b + 2.0;
    return 0;
  }
};

BOUTMAIN(SiLoop)
