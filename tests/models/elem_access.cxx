// Test model of this repository (NOT from the reference): host-side element access mixed with field operations,
// written against the public BOUT++ API only.  Compiled against the reference build and against the B200 host layer
// from this one source (see ode_pair.cxx).  What it exercises: Field3D::operator()(x,y,z) for writing and
// read-modify-write in init(), and in rhs() a BOUT_FOR loop with operator[](Ind3D) and a z-offset index that updates
// ddt(f) AFTER a field (device) operation produced it, every evaluation.
//   ddt(f) = -w f + 0.5 w f(z+1),   w(x,y,z) = 1 + 0.1 x + 0.01 z
#include <bout/physicsmodel.hxx>

class ElemAccess : public PhysicsModel {
private:
  Field3D f, w;

protected:
  int init(bool) override {
    SOLVE_FOR(f);
    w = 0.0;
    for (int x = 0; x < mesh->LocalNx; x++) {
      for (int y = 0; y < mesh->LocalNy; y++) {
        for (int z = 0; z < mesh->LocalNz; z++) {
          w(x, y, z) = 1.0 + 0.1 * x + 0.01 * z;
        }
      }
    }
    for (int x = mesh->xstart; x <= mesh->xend; x++) {
      for (int y = mesh->ystart; y <= mesh->yend; y++) {
        for (int z = 0; z < mesh->LocalNz; z++) {
          f(x, y, z) *= (1.0 + 0.05 * z);
        }
      }
    }
    return 0;
  }
  int rhs(BoutReal) override {
    ddt(f) = -1.0 * w * f;
    BOUT_FOR(i, f.getRegion("RGN_NOBNDRY")) { ddt(f)[i] += 0.5 * w[i] * f[i.zp()]; }
    return 0;
  }
};

BOUTMAIN(ElemAccess)
