// Stand-in for the CMake-generated bout/version.hxx (reference commit 8397daab).
#ifndef BOUT_VERSION_H
#define BOUT_VERSION_H
namespace bout {
namespace version {
inline constexpr auto full = "5.2.0-oracle";
inline constexpr int major = 5;
inline constexpr int minor = 2;
inline constexpr int patch = 0;
inline constexpr auto prerelease = "oracle";
inline constexpr double as_double = 5.20;
} // namespace version
} // namespace bout
#endif
