// lbm_cli.cpp -- the reference's command line (LatticeBoltzmannCuda.cpp:11-96) on the compatibility layer:
//   lbm_cli <model.obj> <threads x y z> <blocks x y z> <fluid density> <tau> <lattice spacing> <steps>
// The lattice is threads*blocks cells per axis; every population starts at 0.001 with the +x population at
// 0.002; the same values are injected each step on the plane x == 0; the accumulated reflection force is
// printed before every step, as the reference does.  Optional 12th argument: print the force every K steps only
// (the reference blocks on a device-to-host copy every step).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "Lattice.h"
#include "LatticePoint.h"
#include "FlowSurface.h"
#include "FlowCriterion.cuh"
#include "SimCalcFuncs.cuh"

int main(int argc, char* argv[])
{
    if (argc < 12) {
        fprintf(stderr, "usage: %s model.obj tx ty tz bx by bz density tau spacing steps [print_every]\n", argv[0]);
        return -1;
    }
    dim3 threads(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
    dim3 blocks(atoi(argv[5]), atoi(argv[6]), atoi(argv[7]));
    FluidData fluid(atof(argv[8]), atof(argv[9]));
    const int nx = threads.x * blocks.x, ny = threads.y * blocks.y, nz = threads.z * blocks.z;
    const int steps = atoi(argv[11]);
    const int every = argc > 12 ? atoi(argv[12]) : 1;

    Lattice lattice(nx, ny, nz, blocks, threads, fluid, atof(argv[10]));

    LatticePoint flowRef;
    std::memset((void*)&flowRef, 0, sizeof flowRef);
    for (int i = 0; i < 27; ++i) flowRef.particle_distribution[i] = 0.001;
    flowRef.particle_distribution[6] = 0.002;

    const size_t n = (size_t)nx * ny * nz;
    LatticePoint* data = new LatticePoint[n];
    std::memset((void*)data, 0, sizeof(LatticePoint) * n);
    for (size_t k = 0; k < n; ++k)
        for (int i = 0; i < 27; ++i) data[k].particle_distribution[i] = flowRef.particle_distribution[i];
    lattice.load_data(data);
    delete[] data;

    lattice.insertModel(argv[1]);
    lattice.preProcessModel();

    FlowData* inducedFlow = new FlowData(&FlowCriterion::ptestPlaneCriteria, flowRef);
    lattice.setFlowData(inducedFlow);

    ReflectionData* reflection = new ReflectionData;
    lattice.setReflectionData(reflection);

    LatticeData latticeData(lattice.getCudaDataPointer(), lattice.getDimensions());
    RunCudaFunctions::run_prime_points(blocks, threads, latticeData);

    for (int step = 0; step < steps; ++step) {
        if (every > 0 && step % every == 0) {
            reflection = lattice.retrieve_reflection_data();
            std::cout << reflection->x << " " << reflection->y << " " << reflection->z << std::endl;
        }
        lattice.simulateStreaming();
        lattice.simulateReflections();
        lattice.simulateCollision();
        lattice.simulateFlow();
    }
    lattice.synchronize();
    return 0;
}
