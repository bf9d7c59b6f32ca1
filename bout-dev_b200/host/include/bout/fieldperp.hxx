#pragma once
#include "b200_host.hxx"
