// lbm_reference_api.hpp -- the reference's host-side interface, restated in ONE header on top of liblbm_b200.
//
// A program written against CharlieNicholls/LatticeBoltzmannCuda includes `Lattice.h`, `LatticePoint.h`,
// `FlowSurface.h`, `FlowCriterion.cuh`, `SimCalcFuncs.cuh`, `Geometry.h`, `Constants.h`, `Utils.h`, `Model.h`; in
// this directory each of those names is a two-line forwarder to this file, so such a program builds unchanged
// (INTEGRATION.md, section 1).  What must match the reference is the *interface*: type and function names, argument
// lists, the byte layout of the two POD records that cross the boundary.  Everything behind it is ours: the classes
// hold a handle of the C ABI (include/lbm_b200.h) and lbm_compat.cpp forwards to it.
//
// Citations (`inc/` = simulation/include/ of the reference):
//   records       inc/LatticePoint.h:3-17      launch bundle / fluid / force   inc/Lattice.h:15-38
//   orchestrator  inc/Lattice.h:40-106         launch wrappers                 inc/SimCalcFuncs.cuh:13-26
//   source        inc/FlowSurface.h:9-47, simulation/source/FlowCriterion.cu:5-16
//   geometry      inc/Geometry.h:8-31, inc/Model.h:23-47, inc/Constants.h:4-9, inc/Utils.h
#pragma once

#include <cuda_runtime.h>      // dim3, cudaPitchedPtr: part of the reference's signatures

#include <array>
#include <cmath>
#include <string>
#include <utility>
#include <vector>

struct lbm_lattice_s;          // include/lbm_b200.h
struct lbm_mesh_s;

// ===================================================================================================
//  1. Records that cross the host boundary.  Field names, order and sizes are the interface: liblbm_b200
//     reads and writes host arrays of exactly these records (lbm_upload_aos / lbm_download_aos).
// ===================================================================================================

struct ReflectionValues {                        // one per boundary-solid cell
    int   reflection_directions[14];             // 81 slots of 5 bits, six to an int
    float reflection_weight[81];
};
static_assert(sizeof(ReflectionValues) == 380, "ReflectionValues must stay 380 bytes");

struct LatticePoint {
    int   x, y, z;
    float particle_distribution[27];
    bool  isReflected;
    bool  isInternal;
    ReflectionValues* reflections   = nullptr;   // host table of a reflecting cell, filled in by retrieve_data
    ReflectionValues* d_reflections = nullptr;   // unused: the tables live inside the library
};
static_assert(sizeof(LatticePoint) == 144, "LatticePoint must stay 144 bytes");

struct Point_3 {
    float x, y, z;
    Point_3(float x_i, float y_i, float z_i) : x(x_i), y(y_i), z(z_i) {}
    bool operator==(const Point_3& o) const { return o.x == x && o.y == y && o.z == z; }
    void normaliseVector() { const float n = std::sqrt(x * x + y * y + z * z); x /= n; y /= n; z /= n; }
};

struct FluidData {
    float m_fluidDensity            = 1.0f;
    float m_characteristicTimescale = 1.0f;      // tau of the BGK relaxation
    FluidData() = default;
    FluidData(float fluidDensity, float characteristicTimescale)
        : m_fluidDensity(fluidDensity), m_characteristicTimescale(characteristicTimescale) {}
};

struct ReflectionData { float x = 0.0f, y = 0.0f, z = 0.0f; };      // force accumulated on the solid

struct LatticeData {                             // argument bundle of the launch wrappers
    cudaPitchedPtr latticePtr;                   // .ptr NAMES a lattice (opaque handle); it is not raw AoS memory
    dim3 latticeDimensions;
    LatticeData() : latticePtr(cudaPitchedPtr()), latticeDimensions(dim3()) {}
    LatticeData(cudaPitchedPtr ptr, dim3 dims) : latticePtr(ptr), latticeDimensions(dims) {}
};

// ===================================================================================================
//  2. Additive source.  The reference reaches its predicates through a device function pointer on every
//     thread; here they are host functions that Lattice::setFlowData turns into an index box for the kernel.
// ===================================================================================================

namespace FlowCriterion {
    bool testCriteria(int x, int y, int z);          // the single cell (1, 1, 1)
    bool testPlaneCriteria(int x, int y, int z);     // the plane x == 0
    extern bool (*ptestCriteria)(int, int, int);
    extern bool (*ptestPlaneCriteria)(int, int, int);
}

struct FlowData {
    LatticePoint inducedFlow;
    bool (*pointCriterion)(int x, int y, int z);
    int val = 1;
    FlowData(bool (**criterion)(int, int, int), LatticePoint flow) : inducedFlow(flow), pointCriterion(*criterion) {}
    void setCriterion(bool (**criterion)(int, int, int)) { pointCriterion = *criterion; }
    void setInducedFlow(LatticePoint flow) { inducedFlow = flow; }
};

// ===================================================================================================
//  3. Geometry helpers and the mesh wrapper (CGAL-free: points are Point_3, the box is a small struct with
//     the accessors Lattice::insertModel uses).
// ===================================================================================================

namespace CONSTANTS {
    extern const float GEOMETRIC_TOLERANCE;
    extern const int   DIRECTIONS[27][3];
    extern const int   INVERSE_DIRECTIONS[27];
}

namespace Utils {
    unsigned int compressInts(int* array, const int num, const int bitLength);
}

Point_3 rotateVectorAroundAxis(Point_3 axis, Point_3 vector, float angle);

struct BoundingBox {
    float lo[3], hi[3];
    float xmin() const { return lo[0]; }
    float xmax() const { return hi[0]; }
    float ymin() const { return lo[1]; }
    float ymax() const { return hi[1]; }
    float zmin() const { return lo[2]; }
    float zmax() const { return hi[2]; }
};

// What `Model::getMesh()` returns.  The reference hands out its CGAL::Surface_mesh (`inc/Model.h:14,32`); callers use it
// for counts (`getMesh().number_of_vertices()`, test/ModelTest.cpp:11).  CGAL is not a dependency here, so this is a small
// stand-in with the Surface_mesh member names for exactly that.
struct Meshdata {
    std::size_t number_of_vertices() const { return vertices_; }
    std::size_t number_of_faces() const { return faces_; }
    bool is_empty() const { return vertices_ == 0; }
    std::size_t vertices_ = 0, faces_ = 0;
};

class Model {
public:
    Model() = default;
    ~Model();
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    void    importModel(std::string filename);
    bool    checkError();
    bool    isModelClosed();
    const Meshdata& getMesh() const { return modelMesh; }     // counts of the imported mesh (Model.h:32)
    int     number_of_vertices();                    // == getMesh().number_of_vertices()
    BoundingBox bounding_box();
    bool    isPointInsideModel(Point_3 point);
    Point_3 reflectionVector(Point_3 point_1, Point_3 point_2);

    lbm_mesh_s* handle() { return mesh_; }

private:
    lbm_mesh_s* mesh_ = nullptr;
    Meshdata modelMesh;
};

// ===================================================================================================
//  4. The orchestrator.  Differences a caller can observe: the launch shape (blocks, threads) is accepted and
//     ignored; any nx, ny, nz works (the reference needs X == Z and lattice == blocks * threads); consecutive
//     simulateStreaming / Reflections / Collision / Flow calls run as ONE fused kernel step; errors are printed
//     like the reference does and additionally kept in lastError().
// ===================================================================================================

class Lattice {
public:
    // What the reference has no switch for (it has one velocity set, one boundary model, one edge rule, one GPU).  The
    // defaults are the reference's behaviour on one GPU; `fromEnvironment()` reads LBM_GPUS (x slabs across that many
    // devices of this process, joined by peer-memory halos), LBM_Q (19 | 27), LBM_BOUNDARY (reflect | bounce),
    // LBM_BC ("zpw": one letter per axis -- z zero inflow, p periodic, w wall), LBM_MOMENTS (reference | standard),
    // LBM_EQ (fp64 | fp32), LBM_ENGINE (pull2 | inplace).  The constructor with the reference's signature uses it.
    struct Options {
        int gpus = 1;
        int q = 27;
        int boundary_mode = 0;           // lbm_boundary_mode
        int bc[3] = {0, 0, 0};           // lbm_bc per axis
        int moment_mode = 0;             // lbm_moment_mode
        int eq_precision = 1;            // lbm_eq_precision (1 = fp64, the reference's)
        int engine = 0;                  // lbm_engine
        static Options fromEnvironment();
    };

    Lattice(int x, int y, int z, dim3 blocks, dim3 threads, FluidData fluid, float spacing);
    Lattice(int x, int y, int z, dim3 blocks, dim3 threads, FluidData fluid, float spacing, const Options& options);
    ~Lattice();
    Lattice(const Lattice&) = delete;
    Lattice& operator=(const Lattice&) = delete;

    // whole-lattice transfers
    void            load_data(LatticePoint* lattice_array);
    LatticePoint*   retrieve_data();                 // new[]-allocated, owned by the caller
    dim3            getDimensions() { return dim3(nx_, ny_, nz_); }
    cudaPitchedPtr  getCudaDataPointer();

    // geometry side inputs
    void insertModel(std::string filename);
    void preProcessModel();
    std::array<std::pair<float, int>, 27> distributeVector(Point_3 vector);
    void setReflectionDirection(LatticePoint* point, const int index, const unsigned int value);
    int  getReflectionDirection(LatticePoint* point, const int index);

    // source and force
    void            setFlowData(FlowData* flowData);
    void            setReflectionData(ReflectionData* reflection);
    ReflectionData* retrieve_reflection_data();

    // the per-timestep path
    void simulateStreaming();
    void simulateReflections();
    void simulateCollision();
    void simulateFlow();
    void simulateLattice();                          // stream + collide

    // additions
    void simulateSteps(int n);                       // n fused steps: stream, reflect, collide, source
    void synchronize();
    int  deviceCount() const { return (int)slabs_.size(); }
    lbm_lattice_s* handle() { flush(); return handle_; }      // the first slab
    const std::string& lastError() const { return last_error_; }

    // what RunCudaFunctions::* do once they have found the Lattice a LatticeData names
    static Lattice* owner_of(lbm_lattice_s* h);
    void runPhase(unsigned phase, const char* what);
    void runCollision(float timescale);
    void runReflections(ReflectionData* into);

private:
    void create(const Options& options);
    void queue(unsigned phase);
    void flush();
    bool check(int rc, const char* what);
    bool stepPhases(int n, unsigned phases, const char* what);

    lbm_lattice_s* handle_  = nullptr;               // == slabs_[0]
    std::vector<lbm_lattice_s*> slabs_;              // one per device, cut along x
    std::vector<int> slab_x0_;                       // first plane of every slab, and nx_ at the end
    bool           periodic_x_ = false;
    unsigned       pending_ = 0;                     // phases queued in CLI order, not yet executed
    unsigned short nx_, ny_, nz_;
    float          spacing_;
    dim3           blocks_, threads_;
    FluidData      fluid_;
    FlowData*       flow_  = nullptr;
    ReflectionData* force_ = nullptr;
    Model*          model_ = nullptr;
    std::vector<ReflectionValues> tables_;           // host copy of the installed tables
    std::vector<long long>        table_cells_;
    std::string last_error_;
};

// ===================================================================================================
//  5. The launch wrappers.  `lattice.latticePtr.ptr` names a lattice created by `Lattice` (see
//     Lattice::getCudaDataPointer); the launch shape is ignored and `templattice` is not needed.
// ===================================================================================================

namespace RunCudaFunctions {
    void run_prime_points(dim3 blocks, dim3 threads, LatticeData lattice);
    void run_calculate_streaming(dim3 blocks, dim3 threads, LatticeData lattice, LatticeData templattice);
    void run_calculate_reflections(dim3 blocks, dim3 threads, LatticeData lattice, LatticeData templattice);
    void run_calculate_reflections_data(dim3 blocks, dim3 threads, LatticeData lattice, LatticeData templattice, ReflectionData* reflections);
    void run_calculate_collision(dim3 blocks, dim3 threads, LatticeData lattice, float timescale);
    void run_generate_flow(dim3 blocks, dim3 threads, LatticeData lattice, FlowData* flowData);
}
