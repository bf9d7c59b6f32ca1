// Stand-in for the CMake-generated bout/revision.hxx.
#ifndef BOUT_REVISION_H
#define BOUT_REVISION_H
namespace bout {
namespace version {
inline constexpr auto revision = "8397daab";
} // namespace version
} // namespace bout
#endif
