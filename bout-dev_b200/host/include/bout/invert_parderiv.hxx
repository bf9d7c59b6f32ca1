// InvertPar (reference include/bout/invert_parderiv.hxx): declared in b200_host.hxx
#pragma once
#include "b200_host.hxx"
