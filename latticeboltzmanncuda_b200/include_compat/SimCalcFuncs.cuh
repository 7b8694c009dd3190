// SimCalcFuncs.cuh -- the reference's launch wrappers (simulation/include/SimCalcFuncs.cuh:13-26).
// `LatticeData::latticePtr.ptr` names a lattice created by `Lattice` (see Lattice::getCudaDataPointer); the
// launch shape is ignored and `templattice` is not needed (the library keeps both population buffers).
#pragma once

#include <cuda_runtime.h>

#include "Lattice.h"

namespace RunCudaFunctions
{
    void run_calculate_streaming(dim3 blocks, dim3 threads, LatticeData lattice, LatticeData templattice);

    void run_calculate_collision(dim3 blocks, dim3 threads, LatticeData lattice, float timescale);

    void run_calculate_reflections(dim3 blocks, dim3 threads, LatticeData lattice, LatticeData templattice);

    void run_calculate_reflections_data(dim3 blocks, dim3 threads, LatticeData lattice, LatticeData templattice, ReflectionData* reflections);

    void run_prime_points(dim3 blocks, dim3 threads, LatticeData lattice);

    void run_generate_flow(dim3 blocks, dim3 threads, LatticeData lattice, FlowData* flowData);
}
