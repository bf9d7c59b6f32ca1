#include "cpptrace.hpp"
