// ref_harness.cpp -- extern "C" doors into the reference's own `Lattice` class and CUDA
// kernels, compiled UNMODIFIED from /root/reference into oracle/_ref/libref_gpu.so
// (SimCalcFuncs.cu, FlowCriterion.cu, Lattice.cpp, Geometry.cpp, Constants.cpp, Utils.cpp).
// Only Model.cpp is replaced (CGAL is absent: ref_model_stub.cpp + cgal_stub/).
// Used by tests/ as the strongest oracle and by bench.py --impl reference as the reference
// arm.  TEST INFRASTRUCTURE ONLY -- the product never links this.
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "Lattice.h"
#include "LatticePoint.h"
#include "FlowSurface.h"
#include "FlowCriterion.cuh"
#include "SimCalcFuncs.cuh"

// a flow criterion of our own that is not an index box (oracle/ref_extra_criterion.cu), C++ linkage
namespace LbmExtraCriterion { extern __device__ bool (*platticePattern)(int, int, int); }

struct RefHandle {
    Lattice* lattice;
    void* storage;
    dim3 blocks, threads;
    int X, Y, Z;
    ReflectionData* reflection;
    FlowData* flow;
};

extern "C" {

int ref_sizeof_lattice_point() { return (int)sizeof(LatticePoint); }
int ref_sizeof_reflection_values() { return (int)sizeof(ReflectionValues); }
int ref_sizeof_flow_data() { return (int)sizeof(FlowData); }

// Lattice has members it reads before writing (Lattice.h:85-102); it is
// constructed in zeroed storage so that those reads see null pointers.
void* ref_create(int x, int y, int z, int bx, int by, int bz, int tx, int ty, int tz,
                 float density, float tau, float spacing)
{
    RefHandle* h = new RefHandle();
    h->storage = std::calloc(1, sizeof(Lattice));
    h->blocks = dim3(bx, by, bz);
    h->threads = dim3(tx, ty, tz);
    h->X = x; h->Y = y; h->Z = z;
    h->lattice = new (h->storage) Lattice(x, y, z, h->blocks, h->threads, FluidData(density, tau), spacing);
    h->reflection = nullptr;
    h->flow = nullptr;
    // cudaMalloc3D memory is not zeroed; the reference's tests assume it is (Q10)
    cudaPitchedPtr p = h->lattice->getCudaDataPointer();
    cudaMemset(p.ptr, 0, p.pitch * (size_t)y * (size_t)z);
    return h;
}

void ref_destroy(void* hv)
{
    RefHandle* h = (RefHandle*)hv;
    cudaDeviceSynchronize();
    h->lattice->~Lattice();
    std::free(h->storage);
    delete h->reflection;
    delete h;
}

void ref_load(void* hv, LatticePoint* host) { ((RefHandle*)hv)->lattice->load_data(host); }

void ref_retrieve(void* hv, LatticePoint* out)
{
    RefHandle* h = (RefHandle*)hv;
    LatticePoint* a = h->lattice->retrieve_data();
    if (!a) return;
    std::memcpy((void*)out, (void*)a, sizeof(LatticePoint) * (size_t)h->X * h->Y * h->Z);
    delete[] a;
}

void ref_prime(void* hv)
{
    RefHandle* h = (RefHandle*)hv;
    LatticeData d(h->lattice->getCudaDataPointer(), h->lattice->getDimensions());
    RunCudaFunctions::run_prime_points(h->blocks, h->threads, d);
}

void ref_stream(void* hv)  { ((RefHandle*)hv)->lattice->simulateStreaming(); }
void ref_reflect(void* hv) { ((RefHandle*)hv)->lattice->simulateReflections(); }
void ref_collide(void* hv) { ((RefHandle*)hv)->lattice->simulateCollision(); }
void ref_flow(void* hv)    { ((RefHandle*)hv)->lattice->simulateFlow(); }
void ref_lattice_step(void* hv) { ((RefHandle*)hv)->lattice->simulateLattice(); }
void ref_sync() { cudaDeviceSynchronize(); }

void ref_insert_model(void* hv, const char* path) { ((RefHandle*)hv)->lattice->insertModel(std::string(path)); }
void ref_preprocess(void* hv) { ((RefHandle*)hv)->lattice->preProcessModel(); }

// kind 1: FlowCriterion::ptestCriteria (cell 1,1,1); kind 2: ptestPlaneCriteria (x == 0); kind 3: a criterion of our own
// that is not a box (oracle/ref_extra_criterion.cu), to exercise the reference's per-cell predicate call
void ref_set_flow(void* hv, int kind, const float* f27)
{
    RefHandle* h = (RefHandle*)hv;
    LatticePoint flowRef;
    std::memset((void*)&flowRef, 0, sizeof(flowRef));
    for (int i = 0; i < 27; ++i) flowRef.particle_distribution[i] = f27[i];
    h->flow = new FlowData(kind == 1 ? &FlowCriterion::ptestCriteria
                           : kind == 3 ? &LbmExtraCriterion::platticePattern : &FlowCriterion::ptestPlaneCriteria, flowRef);
    h->lattice->setFlowData(h->flow);
}

void ref_set_reflection_data(void* hv, float x, float y, float z)
{
    RefHandle* h = (RefHandle*)hv;
    if (!h->reflection) h->reflection = new ReflectionData;
    h->reflection->x = x; h->reflection->y = y; h->reflection->z = z;
    h->lattice->setReflectionData(h->reflection);
}

void ref_get_reflection_data(void* hv, float* out)
{
    ReflectionData* r = ((RefHandle*)hv)->lattice->retrieve_reflection_data();
    out[0] = r->x; out[1] = r->y; out[2] = r->z;
}

// fetch the device-side table of one reflecting cell (LatticePoint::d_reflections)
void ref_fetch_table(const LatticePoint* cell, ReflectionValues* out)
{
    cudaMemcpy(out, cell->d_reflections, sizeof(ReflectionValues), cudaMemcpyDeviceToHost);
}

// the CLI's loop body, LatticeBoltzmannCuda.cpp:85-93; force_trace (3*nsteps) may be null
void ref_run_cli(void* hv, int nsteps, float* force_trace)
{
    RefHandle* h = (RefHandle*)hv;
    for (int s = 0; s < nsteps; ++s) {
        if (h->reflection) {
            ReflectionData* r = h->lattice->retrieve_reflection_data();
            if (force_trace) { force_trace[3 * s] = r->x; force_trace[3 * s + 1] = r->y; force_trace[3 * s + 2] = r->z; }
        }
        h->lattice->simulateStreaming();
        h->lattice->simulateReflections();
        h->lattice->simulateCollision();
        h->lattice->simulateFlow();
    }
    cudaDeviceSynchronize();
}

const char* ref_last_cuda_error() { return cudaGetErrorString(cudaGetLastError()); }

}  // extern "C"
