// Hand-written build configuration for compiling the BOUT++ reference sources as a
// single-rank CPU oracle (stands in for the file CMake generates from
// cmake_build_defines.hxx.in).  All optional third-party features are off; FFTW is provided by
// oracle/shim/fftw3.h.  BOUT_CHECK_LEVEL / BOUT_USE_OPENMP may be overridden on the command line.
#pragma once
#ifndef BOUT_BUILD_CONFIG_HXX
#define BOUT_BUILD_CONFIG_HXX

#ifndef BOUT_CHECK_LEVEL
#define BOUT_CHECK_LEVEL 0
#endif
#define BOUT_OPENMP_SCHEDULE static
#define BOUT_HAS_ARKODE 0
#define BOUT_HAS_CVODE 0
#define BOUT_HAS_FFTW 1
#define BOUT_HAS_GETTEXT 0
#define BOUT_HAS_HYPRE 0
#define BOUT_HAS_IDA 0
#define BOUT_HAS_LAPACK 0
#define BOUT_HAS_NETCDF 0
#define BOUT_HAS_ADIOS2 0
#define BOUT_HAS_PETSC 0
#define BOUT_HAS_PRETTY_FUNCTION 1
#define BOUT_HAS_PVODE 0
#define BOUT_HAS_SCOREP 0
#define BOUT_HAS_SLEPC 0
#define BOUT_HAS_SUNDIALS 0
#define BOUT_HAS_UUID_SYSTEM_GENERATOR 0
#define BOUT_USE_COLOR 0
#ifndef BOUT_USE_OPENMP
#define BOUT_USE_OPENMP 0
#endif
#define BOUT_USE_OUTPUT_DEBUG 0
#define BOUT_USE_SIGFPE 0
#define BOUT_USE_SIGNAL 0
#define BOUT_USE_TRACK 0
#define BOUT_HAS_UMPIRE 0
#define BOUT_HAS_CALIPER 0
#define BOUT_HAS_RAJA 0
#define BOUT_HAS_CUDA 0
#define BOUT_METRIC_TYPE 2D
#define BOUT_USE_METRIC_3D 0
#define BOUT_USE_MSGSTACK 0
#define BOUT_HAS_LEGACY_NETCDF 0
#define BOUT_USE_BACKTRACE 0
#define BOUT_FLAGS_STRING oracle
#endif // BOUT_BUILD_CONFIG_HXX
