// Force-included when compiling the reference's src/sys/options/options_io.cxx only:
// with BOUT_HAS_NETCDF=0 the reference's options_netcdf.hxx:15 names
// RegisterUnavailableOptionsIO without its bout:: namespace. Bring it into scope.
#include "bout/options_io.hxx"
using bout::RegisterUnavailableOptionsIO;
