// Force-included when compiling a reference example model (examples/*/ *.cxx) into the oracle
// library: replaces the BOUTMAIN(Model) main() by a factory function named MODEL_FACTORY so several
// unchanged model sources can live in one shared library.
#include <bout/physicsmodel.hxx>
#include <memory>
#undef BOUTMAIN
#define BOUTMAIN(ModelClass)                                      \
  std::unique_ptr<PhysicsModel> MODEL_FACTORY() {                 \
    return std::unique_ptr<PhysicsModel>(new ModelClass());       \
  }
