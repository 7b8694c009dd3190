// Lattice.h -- include name of the reference kept for source compatibility; the interface lives in lbm_reference_api.hpp
#pragma once
#include "lbm_reference_api.hpp"
