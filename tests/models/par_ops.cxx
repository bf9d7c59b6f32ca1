// Test model of this repository (NOT from the reference): the parallel operators and the upwind / flux derivatives of
// the host layer that the three headline models do not use -- Grad_par, Vpar_Grad_par, Div_par, Grad2_par2 (identity
// parallel transform), VDDY, FDDX / FDDY / FDDZ -- on a mesh whose g_22, dy and Bxy vary.  The same source is compiled
// into the reference build (golden values: tests/golden/make_golden_vec.py) and against the B200 host layer.
#include <bout/derivs.hxx>
#include <bout/difops.hxx>
#include <bout/physicsmodel.hxx>

class ParOps : public PhysicsModel {
private:
  Field3D f, g, v;

protected:
  int init(bool) override {
    SOLVE_FOR(f, g);
    v = 0.0;
    for (int x = 0; x < mesh->LocalNx; x++) {
      for (int y = 0; y < mesh->LocalNy; y++) {
        for (int z = 0; z < mesh->LocalNz; z++) {
          v(x, y, z) = 0.4 * ((x + 2 * y + z) % 5) - 0.7 + 0.01 * x; // both signs: both upwind branches
        }
      }
    }
    return 0;
  }
  int rhs(BoutReal) override {
    mesh->communicate(f, g, v);
    ddt(f) = Grad_par(f) + Vpar_Grad_par(v, g) + Div_par(g) + Grad2_par2(f);
    ddt(g) = VDDY(v, f) + FDDX(v, g) + FDDY(v, f) + FDDZ(v, g);
    return 0;
  }
};

BOUTMAIN(ParOps)
