// LaplaceXZ (reference include/bout/invert/laplacexz.hxx): declared in b200_host.hxx
#pragma once
#include "../b200_host.hxx"
