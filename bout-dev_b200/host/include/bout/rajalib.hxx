// BOUT_FOR_RAJA (reference include/bout/rajalib.hxx:40-130): the loop body is a lambda taking the index.  The
// reference gives it to RAJA::forall, which runs it for every index of the region on the device; here it is called
// once with a symbolic index, and what it recorded (field_accessor.hxx) is dispatched to device kernels over the
// region.  The extra macro arguments (CAPTURE(...)) go into the lambda capture exactly as in the reference.
#pragma once

#include "field_accessor.hxx"

namespace bout {
namespace b200 {
namespace si {
class SymbolicLoop {
public:
  explicit SymbolicLoop(const Region<Ind3D>& region_) : region(region_) {}
  template <class F>
  const SymbolicLoop& operator<<(F body) const {
    Recorder& r = Recorder::get();
    if (r.active) {
      throw BoutException("BOUT_FOR_RAJA: loops cannot be nested");
    }
    r.nodes.clear();
    r.assigns.clear();
    r.active = true;
    try {
      body(SIndex{});
    } catch (...) {
      r.active = false;
      throw;
    }
    r.active = false;
    dispatch(region);
    return *this;
  }

private:
  const Region<Ind3D>& region;
};
} // namespace si
} // namespace b200
} // namespace bout

#define B200_SHADOW_ARG(var) , var = var
#define B200_FE0_1(m, a) m(a)
#define B200_FE0_2(m, a, ...) m(a) B200_FE0_1(m, __VA_ARGS__)
#define B200_FE0_3(m, a, ...) m(a) B200_FE0_2(m, __VA_ARGS__)
#define B200_FE0_4(m, a, ...) m(a) B200_FE0_3(m, __VA_ARGS__)
#define B200_FE0_5(m, a, ...) m(a) B200_FE0_4(m, __VA_ARGS__)
#define B200_FE0_6(m, a, ...) m(a) B200_FE0_5(m, __VA_ARGS__)
#define B200_FE0_7(m, a, ...) m(a) B200_FE0_6(m, __VA_ARGS__)
#define B200_FE0_8(m, a, ...) m(a) B200_FE0_7(m, __VA_ARGS__)
#define B200_GET_FE0(_1, _2, _3, _4, _5, _6, _7, _8, NAME, ...) NAME
/// CAPTURE(a, b) -> `, a = a, b = b`: shadows class members inside the lambda (rajalib.hxx:98-103)
#define CAPTURE(...)                                                                                             \
  B200_GET_FE0(__VA_ARGS__, B200_FE0_8, B200_FE0_7, B200_FE0_6, B200_FE0_5, B200_FE0_4, B200_FE0_3, B200_FE0_2, \
               B200_FE0_1)(B200_SHADOW_ARG, __VA_ARGS__)
#define BOUT_FOR_RAJA(index, region, ...) \
  bout::b200::si::SymbolicLoop(region) << [= __VA_ARGS__](SIndex index)
