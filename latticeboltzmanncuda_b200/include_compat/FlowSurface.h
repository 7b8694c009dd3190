// FlowSurface.h -- FlowData with the reference's constructor signature (simulation/include/FlowSurface.h:9-47)
#pragma once

#include "LatticePoint.h"
#include "FlowCriterion.cuh"

struct FlowData
{
    LatticePoint inducedFlow;
    bool (*pointCriterion)(int x, int y, int z);
    int val = 1;

    FlowData(bool (**criterion)(int, int, int), LatticePoint flow) : inducedFlow(flow), pointCriterion(*criterion) {}

    void setCriterion(bool (**criterion)(int, int, int)) { pointCriterion = *criterion; }

    void setInducedFlow(LatticePoint flow) { inducedFlow = flow; }
};
