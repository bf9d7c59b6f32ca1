// Test model of this repository (NOT from the reference): FieldPerp through the public BOUT++ API -- sliceXZ, FieldPerp
// algebra with FieldPerp / BoutReal / Field3D operands, compound assignment, host-side element access,
// Laplacian::solve(const FieldPerp&, const FieldPerp&) and Field3D = FieldPerp (include/bout/fieldperp.hxx:54-330,
// src/field/fieldperp.cxx, generated_fieldops.cxx, cyclic_laplace.cxx:120-185).  ddt(f) carries the algebra result in
// plane [test] jy (zero elsewhere), ddt(g) the solve; the reference build runs the same source for the golden values
// (tests/golden/make_golden_invert.py).
#include <bout/fieldperp.hxx>
#include <bout/invert_laplace.hxx>
#include <bout/physicsmodel.hxx>

class PerpOps : public PhysicsModel {
private:
  Field3D f, g;
  int jy;
  std::unique_ptr<Laplacian> lap;

protected:
  int init(bool) override {
    SOLVE_FOR(f, g);
    jy = Options::root()["test"]["jy"].withDefault(3);
    lap = Laplacian::create(&Options::root()["lap_cyclic"]);
    return 0;
  }
  int rhs(BoutReal) override {
    const FieldPerp p = sliceXZ(f, jy), q = sliceXZ(g, jy);
    FieldPerp r = 2.0 * p + p * q - q / 1.5;
    r += p;
    r *= 0.5;
    r = r - 0.25;
    FieldPerp s = r * f + q / (3.0 + g * g); // FieldPerp (op) Field3D -> FieldPerp on the plane of the FieldPerp
    s = -s;
    s(2, 1) += 0.125;                        // element access from host code
    s(mesh->xend, mesh->LocalNz - 1) = s(mesh->xstart, 0) * 2.0;
    Field3D a = 0.0;
    a = s;                                   // Field3D = FieldPerp: plane jy only
    ddt(f) = a;
    const FieldPerp sol = lap->solve(p, q);
    Field3D b = 0.0;
    b = sol;
    ddt(g) = b;
    return 0;
  }
};

BOUTMAIN(PerpOps)
