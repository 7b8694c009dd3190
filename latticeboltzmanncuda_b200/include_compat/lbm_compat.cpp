// lbm_compat.cpp -- the reference's C++ interface (Lattice, RunCudaFunctions, Model, rotateVectorAroundAxis,
// CONSTANTS, Utils, FlowCriterion) implemented on the C ABI of liblbm_b200.so (include/lbm_b200.h).
// Host code only: every line of device work sits behind the C ABI.
#include "lbm_reference_api.hpp"

#include <mutex>
#include <unordered_map>

#include "../../include/lbm_b200.h"

#include <cstdio>
#include <cstring>

// ---- Constants / Utils / Geometry / FlowCriterion ---------------------------------------------------------

namespace CONSTANTS
{
    const float GEOMETRIC_TOLERANCE = 1e-6;
    const int DIRECTIONS[27][3] = {
        { 0, 0, 0},
        {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1}, { 0, 0, 1}, { 0, 1, 0}, { 1, 0, 0},
        {-1,-1, 0}, {-1, 0,-1}, {-1, 0, 1}, {-1, 1, 0}, { 0,-1,-1}, { 0,-1, 1},
        { 0, 1,-1}, { 0, 1, 1}, { 1,-1, 0}, { 1, 0,-1}, { 1, 0, 1}, { 1, 1, 0},
        {-1,-1,-1}, {-1,-1, 1}, {-1, 1,-1}, {-1, 1, 1}, { 1,-1,-1}, { 1,-1, 1}, { 1, 1,-1}, { 1, 1, 1}};
    const int INVERSE_DIRECTIONS[27] = {0, 6, 5, 4, 3, 2, 1, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7,
                                        26, 25, 24, 23, 22, 21, 20, 19};
}

namespace Utils
{
    unsigned int compressInts(int* array, const int num, const int bitLength) { return lbm_compress_ints(array, num, bitLength); }
}

Point_3 rotateVectorAroundAxis(Point_3 axis, Point_3 vector, float angle)
{
    const float a[3] = {axis.x, axis.y, axis.z}, v[3] = {vector.x, vector.y, vector.z};
    float out[3] = {0, 0, 0};
    lbm_rotate_vector_around_axis(a, v, angle, out);
    return Point_3(out[0], out[1], out[2]);
}

namespace FlowCriterion
{
    bool testCriteria(int x, int y, int z) { return x == 1 && y == 1 && z == 1; }
    bool testPlaneCriteria(int x, int, int) { return x == 0; }
    bool (*ptestCriteria)(int, int, int) = &FlowCriterion::testCriteria;
    bool (*ptestPlaneCriteria)(int, int, int) = &FlowCriterion::testPlaneCriteria;
}

// ---- Model ---------------------------------------------------------------------------------------------------

Model::~Model() { lbm_mesh_destroy(mesh_); }

void Model::importModel(std::string filename)
{
    lbm_mesh_destroy(mesh_);
    mesh_ = nullptr;
    lbm_mesh_load(filename.c_str(), &mesh_);
}

bool Model::isModelClosed()
{
    int closed = 0;
    if (mesh_) lbm_mesh_info(mesh_, nullptr, nullptr, &closed, nullptr);
    return closed != 0;
}

int Model::number_of_vertices()
{
    int n = 0;
    if (mesh_) lbm_mesh_info(mesh_, &n, nullptr, nullptr, nullptr);
    return n;
}

bool Model::checkError()
{
    int valid = 1;
    if (mesh_) lbm_mesh_info(mesh_, nullptr, nullptr, nullptr, &valid);
    return !valid;
}

bool Model::isPointInsideModel(Point_3 point)
{
    int inside = 0;
    const float p[3] = {point.x, point.y, point.z};
    if (mesh_) lbm_mesh_point_inside(mesh_, p, &inside);
    return inside != 0;
}

BoundingBox Model::bounding_box()
{
    BoundingBox b;
    std::memset(&b, 0, sizeof b);
    if (mesh_) lbm_mesh_bounding_box(mesh_, b.lo, b.hi);
    return b;
}

Point_3 Model::reflectionVector(Point_3 point_1, Point_3 point_2)
{
    const float a[3] = {point_1.x, point_1.y, point_1.z}, b[3] = {point_2.x, point_2.y, point_2.z};
    float out[3] = {0, 0, 0};
    if (mesh_) lbm_mesh_reflection_vector(mesh_, a, b, out);
    return Point_3(out[0], out[1], out[2]);
}

// ---- Lattice -------------------------------------------------------------------------------------------------

// RunCudaFunctions::* receive a LatticeData by value (the handle travels in its pointer field), not the Lattice that
// owns it: phases that Lattice queued for fusion must run before a direct kernel-wrapper call, so every Lattice is
// findable from its handle.
namespace {
std::mutex& owners_mutex() { static std::mutex m; return m; }
std::unordered_map<lbm_lattice_s*, Lattice*>& owners() { static std::unordered_map<lbm_lattice_s*, Lattice*> m; return m; }
}  // namespace

void Lattice::flush_owner_of(lbm_lattice_s* h)
{
    Lattice* owner = nullptr;
    {
        std::lock_guard<std::mutex> lock(owners_mutex());
        auto it = owners().find(h);
        if (it != owners().end()) owner = it->second;
    }
    if (owner) owner->flush();
}

bool Lattice::check(int rc, const char* what)
{
    if (rc == LBM_OK) return true;
    last_error_ = std::string(what) + ": " + lbm_last_error();
    printf("Lattice::%s failed: %s\n", what, lbm_last_error());     // the reference prints and carries on
    return false;
}

Lattice::Lattice(int x, int y, int z, dim3 blocks, dim3 threads, FluidData fluid, float spacing)
{
    nx_ = x; ny_ = y; nz_ = z;
    blocks_ = blocks; threads_ = threads;
    fluid_ = fluid;
    spacing_ = spacing;
    lbm_config cfg;
    lbm_config_default(&cfg);          // the reference's behaviour: D3Q27, SPLIT moments, fp64 equilibrium, zero inflow, tables
    cfg.nx = x; cfg.ny = y; cfg.nz = z;
    cfg.tau = fluid.m_characteristicTimescale;
    check(lbm_create(&cfg, &handle_), "Lattice");
    if (handle_) {
        std::lock_guard<std::mutex> lock(owners_mutex());
        owners()[handle_] = this;
    }
}

Lattice::~Lattice()
{
    if (handle_) {
        std::lock_guard<std::mutex> lock(owners_mutex());
        owners().erase(handle_);
    }
    lbm_destroy(handle_);
    delete model_;
}

cudaPitchedPtr Lattice::getCudaDataPointer()
{
    flush();
    return make_cudaPitchedPtr((void*)handle_, sizeof(LatticePoint) * nx_, nx_, ny_);
}

// Phases arrive one call at a time in the CLI's order stream, reflect, collide, source
// (LatticeBoltzmannCuda.cpp:89-92).  They are queued while they extend that order and executed as ONE fused
// kernel step when the sequence is complete; anything else executes what is queued phase by phase first.
void Lattice::queue(unsigned phase)
{
    if (!handle_) return;
    // `phase` extends the queued prefix iff it is the next phase in CLI order: S -> S|R -> S|R|C -> S|R|C|F
    const bool extends = pending_ != 0 && (pending_ & LBM_PHASE_STREAM) && phase == ((pending_ + 1) & ~pending_) &&
                         (pending_ | phase) == ((phase << 1) - 1);
    if (!extends) flush();
    pending_ |= phase;
    if (pending_ == LBM_PHASE_ALL) flush();
}

void Lattice::flush()
{
    if (!handle_ || pending_ == 0) return;
    const unsigned p = pending_;
    pending_ = 0;
    check(lbm_step_phases(handle_, 1, p), "simulate");      // one launch for whatever prefix was queued
}

void Lattice::simulateStreaming() { queue(LBM_PHASE_STREAM); }
void Lattice::simulateReflections() { queue(LBM_PHASE_REFLECT); }
void Lattice::simulateCollision() { queue(LBM_PHASE_COLLIDE); }
void Lattice::simulateFlow() { if (flow_ != nullptr) queue(LBM_PHASE_SOURCE); else flush(); }

void Lattice::simulateLattice()
{
    flush();
    if (handle_) check(lbm_step_phases(handle_, 1, LBM_PHASE_STREAM | LBM_PHASE_COLLIDE), "simulateLattice");
}

void Lattice::simulateSteps(int n)
{
    flush();
    if (handle_) check(lbm_step(handle_, n), "simulateSteps");
}

void Lattice::synchronize()
{
    flush();
    if (handle_) check(lbm_sync(handle_), "synchronize");
}

void Lattice::load_data(LatticePoint* lattice_array)
{
    flush();
    if (!handle_) return;
    check(lbm_upload_aos(handle_, lattice_array, (size_t)nx_ * ny_ * nz_), "load_data");
}

LatticePoint* Lattice::retrieve_data()
{
    flush();
    if (!handle_) return nullptr;
    const size_t n = (size_t)nx_ * ny_ * nz_;
    LatticePoint* out = new LatticePoint[n];
    if (!check(lbm_download_aos(handle_, out, n), "retrieve_data")) {
        delete[] out;
        return nullptr;
    }
    // reflecting cells point at their host table, as in the reference (tests read point->reflections)
    for (size_t i = 0; i < table_cells_.size(); ++i) out[table_cells_[i]].reflections = &tables_[i];
    return out;
}

void Lattice::setReflectionData(ReflectionData* reflection)
{
    flush();
    force_ = reflection;
    if (!handle_) return;
    if (!reflection) { check(lbm_set_force_accumulator(handle_, 0, nullptr), "setReflectionData"); return; }
    const float init[3] = {reflection->x, reflection->y, reflection->z};
    check(lbm_set_force_accumulator(handle_, 1, init), "setReflectionData");
}

ReflectionData* Lattice::retrieve_reflection_data()
{
    flush();
    if (!handle_ || !force_) return force_;
    float f[3] = {0, 0, 0};
    if (check(lbm_get_force(handle_, f), "retrieve_reflection_data")) {
        force_->x = f[0]; force_->y = f[1]; force_->z = f[2];
    }
    return force_;
}

void Lattice::setFlowData(FlowData* flowData)
{
    flush();
    flow_ = flowData;
    if (!handle_) return;
    if (!flowData) { check(lbm_set_source(handle_, LBM_SOURCE_NONE, nullptr, nullptr, nullptr), "setFlowData"); return; }
    const int dims[3] = {nx_, ny_, nz_};
    int32_t lo[3] = {0, 0, 0}, hi[3] = {dims[0] - 1, dims[1] - 1, dims[2] - 1};
    if (flowData->pointCriterion == &FlowCriterion::testCriteria) {
        lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 1;
    } else if (flowData->pointCriterion == &FlowCriterion::testPlaneCriteria) {
        hi[0] = 0;
    } else {
        // a user predicate: evaluate it once on the host; the kernel takes an index box
        int32_t l[3] = {dims[0], dims[1], dims[2]}, h[3] = {-1, -1, -1};
        long long count = 0;
        for (int x = 0; x < dims[0]; ++x)
            for (int y = 0; y < dims[1]; ++y)
                for (int z = 0; z < dims[2]; ++z)
                    if (flowData->pointCriterion(x, y, z)) {
                        const int c[3] = {x, y, z};
                        for (int a = 0; a < 3; ++a) { if (c[a] < l[a]) l[a] = c[a]; if (c[a] > h[a]) h[a] = c[a]; }
                        ++count;
                    }
        if (count == 0) { check(lbm_set_source(handle_, LBM_SOURCE_NONE, nullptr, nullptr, nullptr), "setFlowData"); return; }
        if (count != (long long)(h[0] - l[0] + 1) * (h[1] - l[1] + 1) * (h[2] - l[2] + 1)) {
            last_error_ = "setFlowData: the criterion does not select an index box";
            printf("Lattice::setFlowData failed: the criterion does not select an index box\n");
            return;
        }
        for (int a = 0; a < 3; ++a) { lo[a] = l[a]; hi[a] = h[a]; }
    }
    check(lbm_set_source(handle_, LBM_SOURCE_BOX, lo, hi, flowData->inducedFlow.particle_distribution), "setFlowData");
}

void Lattice::insertModel(std::string filename)
{
    flush();
    delete model_;
    model_ = new Model();
    model_->importModel(filename);
    if (!handle_ || !model_->handle()) return;
    int accepted = 0;
    check(lbm_insert_model(handle_, model_->handle(), spacing_, &accepted), "insertModel");
}

void Lattice::preProcessModel()
{
    flush();
    if (model_ == nullptr || !handle_ || !model_->handle()) return;
    if (!check(lbm_preprocess_model(handle_, model_->handle(), spacing_), "preProcessModel")) return;
    int n = 0;
    lbm_get_reflection_tables(handle_, nullptr, nullptr, 0, &n);
    tables_.assign((size_t)n, ReflectionValues());
    std::vector<int64_t> cells((size_t)n);
    if (n) lbm_get_reflection_tables(handle_, tables_.data(), cells.data(), n, &n);
    table_cells_.assign(cells.begin(), cells.end());
}

std::array<std::pair<float, int>, 27> Lattice::distributeVector(Point_3 vector)
{
    const float v[3] = {vector.x, vector.y, vector.z};
    float scores[27];
    int32_t dirs[27];
    lbm_distribute_vector(v, scores, dirs);
    std::array<std::pair<float, int>, 27> out;
    for (int i = 0; i < 27; ++i) out[i] = {scores[i], dirs[i]};
    return out;
}

void Lattice::setReflectionDirection(LatticePoint* point, const int index, const unsigned int value)
{
    const unsigned shift = 5u * (unsigned)(index % 6);
    unsigned word = (unsigned)point->reflections->reflection_directions[index / 6];
    word = (word & ~(0x1Fu << shift)) | (value << shift);
    point->reflections->reflection_directions[index / 6] = (int)word;
}

int Lattice::getReflectionDirection(LatticePoint* point, const int index)
{
    const unsigned word = (unsigned)point->reflections->reflection_directions[index / 6];
    return (int)((word >> (5u * (unsigned)(index % 6))) & 0x1Fu);
}

// ---- RunCudaFunctions ------------------------------------------------------------------------------------------

namespace RunCudaFunctions
{
    static lbm_handle of(const LatticeData& d) { return (lbm_handle)d.latticePtr.ptr; }
    static void report(int rc, const char* what)
    {
        if (rc != LBM_OK) printf("RunCudaFunctions::%s failed: %s\n", what, lbm_last_error());
    }

    // The reference's wrappers take a second lattice (the copy the host made before the launch, Lattice.cpp:115-210):
    // this library keeps both buffers inside the handle, so `templattice` carries nothing and is not looked at.
    // Phases the owning Lattice has queued for fusion run first, so that calls cannot overtake each other.
    static lbm_handle ready(const LatticeData& d)
    {
        lbm_handle h = of(d);
        Lattice::flush_owner_of(h);
        return h;
    }

    void run_calculate_streaming(dim3, dim3, LatticeData lattice, LatticeData) { report(lbm_stream(ready(lattice)), "run_calculate_streaming"); }

    // the timescale of THIS call (the reference kernel divides by its argument, SimCalcFuncs.cu:98-156), not the FluidData's
    void run_calculate_collision(dim3, dim3, LatticeData lattice, float timescale)
    {
        report(lbm_collide_tau(ready(lattice), timescale), "run_calculate_collision");
    }

    void run_calculate_reflections(dim3, dim3, LatticeData lattice, LatticeData) { report(lbm_reflect(ready(lattice)), "run_calculate_reflections"); }

    // the force of this call is added to the caller's ReflectionData (device memory in the reference, Lattice.cpp:190-193;
    // host memory works too), whether or not Lattice::setReflectionData was called
    void run_calculate_reflections_data(dim3, dim3, LatticeData lattice, LatticeData, ReflectionData* reflectionData)
    {
        report(lbm_reflect_accumulate(ready(lattice), reflectionData ? &reflectionData->x : nullptr), "run_calculate_reflections_data");
    }

    void run_prime_points(dim3, dim3, LatticeData lattice) { report(lbm_prime_points(ready(lattice)), "run_prime_points"); }

    void run_generate_flow(dim3, dim3, LatticeData lattice, FlowData*) { report(lbm_source(ready(lattice)), "run_generate_flow"); }
}
