#pragma once
#include "b200_vectors.hxx"
