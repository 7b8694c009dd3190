"""latticeboltzmanncuda_b200 -- B200-native (sm_100a) per-timestep lattice Boltzmann path.

A drop-in for the collide-stream hot path of CharlieNicholls/LatticeBoltzmannCuda: hand-written
CUDA kernels behind a C ABI (include/lbm_b200.h), a C++ shim with the reference's class
signatures (latticeboltzmanncuda_b200/include_compat/), and this Python mirror of the
reference's `Lattice` interface used by the tests and the benchmark.
"""
from . import capi
from .capi import (ENGINE_AUTO, ENGINE_INPLACE, ENGINE_PULL2, BC_HALO, BC_PERIODIC, BC_WALL, BC_ZERO_INFLOW, BOUNDARY_BOUNCE_BACK, BOUNDARY_REFLECT,
                   EQ_FP32, EQ_FP64, FORCE_ACCUM_F64_PARTIALS, FORCE_ACCUM_REFERENCE_FP32, FORCE_PHYSICAL, FORCE_REFERENCE_SLOT_INDEX, HALO_BLOB_BYTES, LATTICE_POINT,
                   MOMENT_REFERENCE, MOMENT_STANDARD, PHASE_ALL, PHASE_COLLIDE, PHASE_REFLECT, PHASE_SOURCE,
                   PHASE_STREAM, REFLECTION_VALUES, LbmError)
from .lattice import FlowCriterion, FlowData, FluidData, Lattice, ReflectionData
from .model import Model, compressInts, distributeVector, getReflectionDirection, rotateVectorAroundAxis

__all__ = [n for n in dir() if not n.startswith("_")]
