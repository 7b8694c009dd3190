// Lattice.h -- the reference's orchestrator interface (simulation/include/Lattice.h:15-106) on top of
// liblbm_b200.  Public signatures are the reference's; the private part is a handle.  Differences a caller can
// observe: the launch shape (blocks, threads) is accepted and ignored; any nx, ny, nz works (the reference needs
// X == Z and lattice == blocks*threads); consecutive simulateStreaming / Reflections / Collision / Flow calls are
// executed as one fused kernel step; errors are printed like the reference does and additionally kept in
// lastError().
#pragma once

#include <cuda_runtime.h>
#include <array>
#include <string>
#include <utility>
#include <vector>

#include "Geometry.h"
#include "FlowSurface.h"

class Model;
struct lbm_lattice_s;

struct LatticeData
{
    cudaPitchedPtr latticePtr;     // .ptr names the lattice (an opaque handle), not raw AoS memory
    dim3 latticeDimensions;

    LatticeData(cudaPitchedPtr ptr, dim3 dims) : latticePtr(ptr), latticeDimensions(dims) {}
    LatticeData() : latticePtr(cudaPitchedPtr()), latticeDimensions(dim3()) {}
};

struct FluidData
{
    float m_fluidDensity;
    float m_characteristicTimescale;

    FluidData(float fluidDensity, float characteristicTimescale) : m_fluidDensity(fluidDensity), m_characteristicTimescale(characteristicTimescale) {}
    FluidData() : m_fluidDensity(1.0), m_characteristicTimescale(1.0) {}
};

struct ReflectionData
{
    float x = 0.0;
    float y = 0.0;
    float z = 0.0;
};

class Lattice
{
public:
    Lattice(int x, int y, int z, dim3 blocks, dim3 threads, FluidData fluid, float spacing);
    ~Lattice();
    Lattice(const Lattice&) = delete;
    Lattice& operator=(const Lattice&) = delete;

    void load_data(LatticePoint* lattice_array);
    LatticePoint* retrieve_data();                       // new[]-allocated, owned by the caller
    ReflectionData* retrieve_reflection_data();

    dim3 getDimensions() { return dim3(m_xResolution, m_yResolution, m_zResolution); }
    cudaPitchedPtr getCudaDataPointer();

    void simulateStreaming();
    void simulateCollision();
    void simulateReflections();
    void simulateFlow();

    void simulateLattice();

    void insertModel(std::string filename);

    std::array<std::pair<float, int>, 27> distributeVector(Point_3 vector);

    void preProcessModel();

    void setFlowData(FlowData* flowData);

    void setReflectionData(ReflectionData* reflection);

    void setReflectionDirection(LatticePoint* point, const int index, const unsigned int value);
    int getReflectionDirection(LatticePoint* point, const int index);

    // ---- additions -----------------------------------------------------------------------------------
    void simulateSteps(int n);                           // n fused steps: stream, reflect, collide, source
    void synchronize();
    lbm_lattice_s* handle() { flush(); return m_handle; }
    const std::string& lastError() const { return m_lastError; }

private:
    void queue(unsigned phase);
    void flush();
    bool check(int rc, const char* what);

    lbm_lattice_s* m_handle = nullptr;
    unsigned m_pending = 0;                              // phases queued in CLI order, not yet executed

    FluidData m_fluid;
    FlowData* m_flowData = nullptr;
    unsigned short m_xResolution, m_yResolution, m_zResolution;
    float m_latticeSpacing;
    Model* m_simModel = nullptr;
    dim3 m_threads, m_blocks;
    ReflectionData* m_reflectionData = nullptr;
    std::vector<ReflectionValues> m_reflections;         // host copy of the installed tables
    std::vector<long long> m_reflectionCells;
    std::string m_lastError;
};
