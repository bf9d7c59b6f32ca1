// Test model of this repository (NOT from the reference): averageY of a Field3D and of a Field2D (elm-pb's zonal
// averages, src/physics/smoothing.cxx:238-323) and smooth_x through the host layer, on a mesh with several y rows.  The
// same source is compiled into the reference build (golden values: tests/golden/make_golden_avg.py) and against the
// B200 host layer.
#include <bout/physicsmodel.hxx>
#include <bout/smoothing.hxx>

class AvgOps : public PhysicsModel {
private:
  Field3D f, g;

protected:
  int init(bool) override {
    SOLVE_FOR(f, g);
    return 0;
  }
  int rhs(BoutReal) override {
    mesh->communicate(f, g);
    ddt(f) = averageY(f) + 0.5 * averageY(f * g);
    ddt(g) = g - averageY(g) + averageY(DC(f)) + smooth_x(f);
    return 0;
  }
};

BOUTMAIN(AvgOps)
