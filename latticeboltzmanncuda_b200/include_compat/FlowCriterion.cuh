// FlowCriterion.cuh -- the two source-cell predicates of the reference (simulation/source/FlowCriterion.cu:5-16).
// The reference keeps them as __device__ functions reached through a device function pointer on every thread;
// here they are ordinary host functions that Lattice::setFlowData turns into an index box for the kernel.
#pragma once

namespace FlowCriterion
{
    bool testCriteria(int x, int y, int z);         // x == 1 && y == 1 && z == 1
    bool testPlaneCriteria(int x, int y, int z);    // x == 0

    extern bool (*ptestCriteria)(int, int, int);
    extern bool (*ptestPlaneCriteria)(int, int, int);
}
