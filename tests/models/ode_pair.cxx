// Test model of this repository (NOT from the reference): two coupled nonlinear ODEs in every cell, written
// against the public BOUT++ API only (PhysicsModel, SOLVE_FOR, ddt, Field3D algebra).  It is compiled twice from
// this one source -- against the reference build (oracle/Makefile -> oracle/_ref/libboutref.so) and against the
// B200 host layer (bout-dev_b200/host/Makefile -> host/models/ode_pair.so) -- so that every explicit integrator of
// the host layer can be checked against the reference's own integrator on the same model:
//   rhs:        ddt(f) = g - damp f^3,   ddt(g) = -f - damp g
//   convective: ddt(f) = g,              ddt(g) = -f
//   diffusive:  ddt(f) = -damp f^3,      ddt(g) = -damp g         ([ode] split = true)
// It also registers output variables the way models do: SAVE_REPEAT(energy), SAVE_ONCE(damp, answer, f_initial).
#include <bout/physicsmodel.hxx>

class OdePair : public PhysicsModel {
private:
  Field3D f, g;
  Field3D energy, f_initial; // diagnostics: written every output / once
  BoutReal damp;
  int answer = 42;
  int quit_at = -1; // [ode] quit_at = <iteration>: outputMonitor asks the solver to stop there

protected:
  int init(bool) override {
    auto& options = Options::root()["ode"];
    damp = options["damp"].withDefault(0.3);
    const bool split = options["split"].withDefault(false);
    quit_at = options["quit_at"].withDefault(-1);
    SOLVE_FOR(f, g);
    f_initial = f;
    energy = 0.5 * (f * f + g * g);
    SAVE_REPEAT(energy);
    SAVE_ONCE(damp, answer, f_initial);
    if (split) {
      setSplitOperator();
    }
    return 0;
  }
  int rhs(BoutReal) override {
    ddt(f) = g - damp * f * f * f;
    ddt(g) = -1.0 * f - damp * g;
    energy = 0.5 * (f * f + g * g);
    return 0;
  }
  int convective(BoutReal) override {
    ddt(f) = g;
    ddt(g) = -1.0 * f;
    energy = 0.5 * (f * f + g * g);
    return 0;
  }
  int outputMonitor(BoutReal, int iter, int) override { return (quit_at >= 0 && iter == quit_at) ? 7 : 0; }
  int diffusive(BoutReal) override {
    ddt(f) = -damp * f * f * f;
    ddt(g) = -damp * g;
    return 0;
  }
};

BOUTMAIN(OdePair)
