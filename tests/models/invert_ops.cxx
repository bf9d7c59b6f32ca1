// Test model of this repository (NOT from the reference): the sibling inversions of SURVEY 8(f)4 through the public
// BOUT++ API -- Laplacian::create with type = pcr, LaplaceXZ::create()->setCoefs()->solve(), InvertPar::create()->
// setCoefA/B()->solve() -- with coefficients taken from the mesh metric so that the reference build can run the same
// calls (tests/golden/make_golden_invert.py).  ddt(b) = the selected solver applied to b (x0 as the initial guess /
// boundary values); [test] which = pcr | xz | par.
#include <bout/invert/laplacexz.hxx>
#include <bout/invert_laplace.hxx>
#include <bout/invert_parderiv.hxx>
#include <bout/physicsmodel.hxx>

class InvertOps : public PhysicsModel {
private:
  Field3D b, x0;
  std::string which;
  std::unique_ptr<Laplacian> lap;
  std::unique_ptr<LaplaceXZ> xz;
  std::unique_ptr<InvertPar> par;

protected:
  int init(bool) override {
    SOLVE_FOR(b, x0);
    which = Options::root()["test"]["which"].withDefault(std::string("xz"));
    Coordinates* c = mesh->getCoordinates();
    if (which == "pcr") {
      lap = Laplacian::create(&Options::root()["lap_pcr"]);
      lap->setInnerBoundaryFlags(INVERT_AC_GRAD);
      lap->setCoefA(c->g33 * 0.3);
    } else if (which == "xz") {
      xz = LaplaceXZ::create(mesh, &Options::root()["xz1"]);
      xz->setCoefs(c->g11, -c->g33);
    } else {
      par = InvertPar::create(&Options::root()["parderiv"]);
      par->setCoefA(c->g33);
      par->setCoefB(c->g11 * (-0.4));
    }
    return 0;
  }
  int rhs(BoutReal) override {
    if (which == "pcr") {
      ddt(b) = lap->solve(b, x0);
    } else if (which == "xz") {
      ddt(b) = xz->solve(b, x0);
    } else {
      ddt(b) = par->solve(b);
    }
    ddt(x0) = 0.0;
    return 0;
  }
};

BOUTMAIN(InvertOps)
