// lbm_compat.cpp -- the reference's C++ interface (Lattice, RunCudaFunctions, Model, rotateVectorAroundAxis,
// CONSTANTS, Utils, FlowCriterion) implemented on the C ABI of liblbm_b200.so (include/lbm_b200.h).
// Host code only: every line of device work sits behind the C ABI.
#include "lbm_reference_api.hpp"

#include <cstdlib>
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
    int nv = 0, nf = 0;
    modelMesh = Meshdata();
    if (mesh_ && lbm_mesh_info(mesh_, &nv, &nf, nullptr, nullptr) == LBM_OK) {
        modelMesh.vertices_ = (std::size_t)nv;
        modelMesh.faces_ = (std::size_t)nf;
    }
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
// owns it: phases that Lattice queued for fusion must run before a direct kernel-wrapper call, and a Lattice that spans
// several GPUs must step all of its slabs, so every Lattice is findable from its handles.
namespace {
std::mutex& owners_mutex() { static std::mutex m; return m; }
std::unordered_map<lbm_lattice_s*, Lattice*>& owners() { static std::unordered_map<lbm_lattice_s*, Lattice*> m; return m; }

int env_choice(const char* name, const char* a, const char* b, int fallback)
{
    const char* e = getenv(name);
    if (!e) return fallback;
    if (strcmp(e, a) == 0) return 0;
    if (strcmp(e, b) == 0) return 1;
    return fallback;
}
}  // namespace

Lattice* Lattice::owner_of(lbm_lattice_s* h)
{
    std::lock_guard<std::mutex> lock(owners_mutex());
    auto it = owners().find(h);
    return it == owners().end() ? nullptr : it->second;
}

Lattice::Options Lattice::Options::fromEnvironment()
{
    Options o;
    if (const char* e = getenv("LBM_GPUS")) o.gpus = atoi(e) > 0 ? atoi(e) : 1;
    if (const char* e = getenv("LBM_Q")) o.q = atoi(e) == 19 ? 19 : 27;
    o.boundary_mode = env_choice("LBM_BOUNDARY", "reflect", "bounce", o.boundary_mode);
    o.moment_mode = env_choice("LBM_MOMENTS", "reference", "standard", o.moment_mode);
    o.eq_precision = env_choice("LBM_EQ", "fp32", "fp64", o.eq_precision);
    o.engine = 1 + env_choice("LBM_ENGINE", "pull2", "inplace", 0);
    if (const char* e = getenv("LBM_BC"))
        for (int a = 0; a < 3 && e[a]; ++a) o.bc[a] = e[a] == 'p' ? LBM_BC_PERIODIC : (e[a] == 'w' ? LBM_BC_WALL : LBM_BC_ZERO_INFLOW);
    return o;
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
    create(Options::fromEnvironment());
}

Lattice::Lattice(int x, int y, int z, dim3 blocks, dim3 threads, FluidData fluid, float spacing, const Options& options)
{
    nx_ = x; ny_ = y; nz_ = z;
    blocks_ = blocks; threads_ = threads;
    fluid_ = fluid;
    spacing_ = spacing;
    create(options);
}

// One handle on one GPU, or `gpus` x slabs (as even as possible), one per device, joined by peer-memory halos: the edge
// cells of a slab address the neighbouring slab's memory over NVLink, lbm_group_step keeps the slabs in lockstep.
void Lattice::create(const Options& o)
{
    lbm_config cfg;
    lbm_config_default(&cfg);          // the reference's behaviour: D3Q27, SPLIT moments, fp64 equilibrium, zero inflow, tables
    cfg.ny = ny_; cfg.nz = nz_;
    cfg.tau = fluid_.m_characteristicTimescale;
    cfg.q = o.q;
    cfg.boundary_mode = o.boundary_mode;
    cfg.moment_mode = o.moment_mode;
    cfg.eq_precision = o.eq_precision;
    cfg.engine = o.engine;
    for (int a = 0; a < 3; ++a) cfg.bc[a] = o.bc[a];
    int gpus = o.gpus < 1 ? 1 : o.gpus;
    if (gpus > nx_ / 2) gpus = nx_ / 2 > 0 ? nx_ / 2 : 1;      // every slab needs two planes
    if (gpus > 1 && cfg.bc[0] != LBM_BC_PERIODIC && cfg.bc[0] != LBM_BC_ZERO_INFLOW) {
        printf("Lattice: a wall on the x axis cannot be cut into slabs; running on one GPU\n");
        gpus = 1;
    }
    int devices = 1;
    lbm_device_count(&devices);
    if (devices < 1) devices = 1;
    const bool ring = cfg.bc[0] == LBM_BC_PERIODIC;
    periodic_x_ = ring;
    slab_x0_.assign((size_t)gpus + 1, 0);
    for (int k = 0; k < gpus; ++k) {
        const int base = nx_ / gpus, extra = nx_ % gpus;
        slab_x0_[k] = k * base + (k < extra ? k : extra);
    }
    slab_x0_[gpus] = nx_;
    if (gpus > 1) cfg.bc[0] = LBM_BC_HALO;
    bool ok = true;
    for (int k = 0; k < gpus && ok; ++k) {
        lbm_lattice_s* h = nullptr;
        cfg.nx = slab_x0_[k + 1] - slab_x0_[k];
        cfg.device = gpus > 1 ? k % devices : -1;
        ok = check(lbm_create(&cfg, &h), "Lattice");
        if (ok) slabs_.push_back(h);
    }
    if (ok && gpus > 1) {
        std::vector<std::vector<unsigned char>> blobs((size_t)gpus, std::vector<unsigned char>(LBM_HALO_BLOB_BYTES));
        for (int k = 0; k < gpus && ok; ++k) {
            ok = check(lbm_set_slab_origin(slabs_[k], slab_x0_[k], nx_), "Lattice") &&
                 check(lbm_halo_export(slabs_[k], blobs[k].data(), blobs[k].size()), "Lattice");
        }
        for (int k = 0; k < gpus && ok; ++k) {
            const void* lo = (ring || k > 0) ? blobs[(k + gpus - 1) % gpus].data() : nullptr;
            const void* hi = (ring || k + 1 < gpus) ? blobs[(k + 1) % gpus].data() : nullptr;
            ok = check(lbm_halo_connect(slabs_[k], lo, hi), "Lattice");
        }
    }
    if (!ok) {
        for (lbm_lattice_s* h : slabs_) lbm_destroy(h);
        slabs_.clear();
        return;
    }
    handle_ = slabs_[0];
    std::lock_guard<std::mutex> lock(owners_mutex());
    for (lbm_lattice_s* h : slabs_) owners()[h] = this;
}

Lattice::~Lattice()
{
    {
        std::lock_guard<std::mutex> lock(owners_mutex());
        for (lbm_lattice_s* h : slabs_) owners().erase(h);
    }
    if (slabs_.size() > 1)
        for (lbm_lattice_s* h : slabs_) lbm_halo_disconnect(h);
    for (lbm_lattice_s* h : slabs_) lbm_destroy(h);
    delete model_;
}

cudaPitchedPtr Lattice::getCudaDataPointer()
{
    flush();
    return make_cudaPitchedPtr((void*)handle_, sizeof(LatticePoint) * nx_, nx_, ny_);
}

bool Lattice::stepPhases(int n, unsigned phases, const char* what)
{
    if (!handle_) return false;
    if (slabs_.size() == 1) return check(lbm_step_phases(handle_, n, phases), what);
    return check(lbm_group_step(slabs_.data(), (int)slabs_.size(), n, phases), what);
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
    stepPhases(1, p, "simulate");      // one launch for whatever prefix was queued
}

void Lattice::simulateStreaming() { queue(LBM_PHASE_STREAM); }
void Lattice::simulateReflections() { queue(LBM_PHASE_REFLECT); }
void Lattice::simulateCollision() { queue(LBM_PHASE_COLLIDE); }
void Lattice::simulateFlow() { if (flow_ != nullptr) queue(LBM_PHASE_SOURCE); else flush(); }

void Lattice::simulateLattice()
{
    flush();
    stepPhases(1, LBM_PHASE_STREAM | LBM_PHASE_COLLIDE, "simulateLattice");
}

void Lattice::simulateSteps(int n)
{
    flush();
    stepPhases(n, LBM_PHASE_ALL, "simulateSteps");
}

void Lattice::synchronize()
{
    flush();
    for (lbm_lattice_s* h : slabs_) check(lbm_sync(h), "synchronize");
}

void Lattice::runPhase(unsigned phase, const char* what)
{
    flush();
    stepPhases(1, phase, what);
}

void Lattice::runCollision(float timescale)
{
    flush();
    if (handle_) check(lbm_group_collide_tau(slabs_.data(), (int)slabs_.size(), timescale), "run_calculate_collision");
}

void Lattice::runReflections(ReflectionData* into)
{
    flush();
    if (handle_) check(lbm_group_reflect_accumulate(slabs_.data(), (int)slabs_.size(), into ? &into->x : nullptr), "run_calculate_reflections_data");
}

void Lattice::load_data(LatticePoint* lattice_array)
{
    flush();
    if (!handle_) return;
    const size_t per = (size_t)ny_ * nz_;
    for (size_t k = 0; k < slabs_.size(); ++k) {
        const int x0 = slab_x0_[k], x1 = slab_x0_[k + 1];
        if (!check(lbm_upload_aos(slabs_[k], lattice_array + (size_t)x0 * per, (size_t)(x1 - x0) * per), "load_data")) return;
        if (slabs_.size() == 1) continue;
        // a slab also needs the solid flags of its neighbours' edge planes (fluid beyond a zero-inflow end, the other
        // end of the lattice beyond a periodic one): isInternal of the records on either side
        std::vector<uint8_t> solid((size_t)(x1 - x0 + 2) * per, 0);
        bool any = false;
        for (int xs = x0 - 1; xs <= x1; ++xs) {
            int xg = xs;
            if (xg < 0 || xg >= nx_) {
                if (!periodic_x_) continue;
                xg = (xg + nx_) % nx_;
            }
            const LatticePoint* src = lattice_array + (size_t)xg * per;
            uint8_t* dst = solid.data() + (size_t)(xs - x0 + 1) * per;
            for (size_t c = 0; c < per; ++c)
                if (src[c].isInternal) { dst[c] = 1; any = true; }
        }
        if (any && !check(lbm_set_solid_mask(slabs_[k], solid.data(), 1), "load_data")) return;
    }
}

LatticePoint* Lattice::retrieve_data()
{
    flush();
    if (!handle_) return nullptr;
    synchronize();
    const size_t per = (size_t)ny_ * nz_, n = (size_t)nx_ * per;
    LatticePoint* out = new LatticePoint[n];
    for (size_t k = 0; k < slabs_.size(); ++k) {
        const int x0 = slab_x0_[k], x1 = slab_x0_[k + 1];
        if (!check(lbm_download_aos(slabs_[k], out + (size_t)x0 * per, (size_t)(x1 - x0) * per), "retrieve_data")) {
            delete[] out;
            return nullptr;
        }
        if (x0 != 0)
            for (size_t c = (size_t)x0 * per; c < (size_t)x1 * per; ++c) out[c].x += x0;      // the slab counts its own planes
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
    for (size_t k = 0; k < slabs_.size(); ++k) {
        if (!reflection) { check(lbm_set_force_accumulator(slabs_[k], 0, nullptr), "setReflectionData"); continue; }
        // the starting value lives on the first slab; the slabs' sums are added when the force is read
        const float init[3] = {k == 0 ? reflection->x : 0.0f, k == 0 ? reflection->y : 0.0f, k == 0 ? reflection->z : 0.0f};
        check(lbm_set_force_accumulator(slabs_[k], 1, init), "setReflectionData");
    }
}

ReflectionData* Lattice::retrieve_reflection_data()
{
    flush();
    if (!handle_ || !force_) return force_;
    double total[3] = {0.0, 0.0, 0.0};
    for (lbm_lattice_s* h : slabs_) {
        double f[3] = {0, 0, 0};
        if (!check(lbm_sync(h), "retrieve_reflection_data") || !check(lbm_get_force_f64(h, f), "retrieve_reflection_data")) return force_;
        for (int a = 0; a < 3; ++a) total[a] += f[a];
    }
    force_->x = (float)total[0]; force_->y = (float)total[1]; force_->z = (float)total[2];
    return force_;
}

void Lattice::setFlowData(FlowData* flowData)
{
    flush();
    flow_ = flowData;
    if (!handle_) return;
    for (size_t k = 0; k < slabs_.size(); ++k) {
        lbm_lattice_s* h = slabs_[k];
        const int x0 = slab_x0_[k], nxl = slab_x0_[k + 1] - slab_x0_[k];
        if (!flowData) { check(lbm_set_source(h, LBM_SOURCE_NONE, nullptr, nullptr, nullptr), "setFlowData"); continue; }
        // the criterion is evaluated once, here, on this slab's cells in GLOBAL coordinates (the reference calls it per
        // cell per step on the device, SimCalcFuncs.cu:206): an index box stays a box, anything else becomes one bit per
        // cell (lbm_set_source_mask)
        const int dims[3] = {nxl, ny_, nz_};
        int32_t l[3] = {dims[0], dims[1], dims[2]}, u[3] = {-1, -1, -1};
        long long count = 0;
        std::vector<uint8_t> sel((size_t)dims[0] * dims[1] * dims[2], 0);
        for (int x = 0; x < dims[0]; ++x)
            for (int y = 0; y < dims[1]; ++y)
                for (int z = 0; z < dims[2]; ++z)
                    if (flowData->pointCriterion(x + x0, y, z)) {
                        const int c[3] = {x, y, z};
                        for (int a = 0; a < 3; ++a) { if (c[a] < l[a]) l[a] = c[a]; if (c[a] > u[a]) u[a] = c[a]; }
                        sel[((size_t)x * dims[1] + y) * dims[2] + z] = 1;
                        ++count;
                    }
        if (count == 0) { check(lbm_set_source(h, LBM_SOURCE_NONE, nullptr, nullptr, nullptr), "setFlowData"); continue; }
        if (count == (long long)(u[0] - l[0] + 1) * (u[1] - l[1] + 1) * (u[2] - l[2] + 1))
            check(lbm_set_source(h, LBM_SOURCE_BOX, l, u, flowData->inducedFlow.particle_distribution), "setFlowData");
        else
            check(lbm_set_source_mask(h, sel.data(), flowData->inducedFlow.particle_distribution), "setFlowData");
    }
}

void Lattice::insertModel(std::string filename)
{
    flush();
    delete model_;
    model_ = new Model();
    model_->importModel(filename);
    if (!handle_ || !model_->handle()) return;
    for (lbm_lattice_s* h : slabs_) {       // (slabs voxelise in global coordinates, lbm_set_slab_origin)
        int accepted = 0;
        check(lbm_insert_model(h, model_->handle(), spacing_, &accepted), "insertModel");
    }
}

void Lattice::preProcessModel()
{
    flush();
    if (model_ == nullptr || !handle_ || !model_->handle()) return;
    tables_.clear();
    table_cells_.clear();
    const size_t per = (size_t)ny_ * nz_;
    for (size_t k = 0; k < slabs_.size(); ++k) {
        lbm_lattice_s* h = slabs_[k];
        if (!check(lbm_preprocess_model(h, model_->handle(), spacing_), "preProcessModel")) return;
        int n = 0;
        lbm_get_reflection_tables(h, nullptr, nullptr, 0, &n);
        std::vector<ReflectionValues> t((size_t)n);
        std::vector<int64_t> cells((size_t)n);
        if (n) lbm_get_reflection_tables(h, t.data(), cells.data(), n, &n);
        for (int i = 0; i < n; ++i) {
            tables_.push_back(t[i]);
            table_cells_.push_back((long long)cells[i] + (long long)slab_x0_[k] * (long long)per);
        }
    }
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
    // this library keeps both buffers inside the handle, so `templattice` carries nothing and is not looked at.  The
    // calls go through the Lattice that owns the handle: phases it has queued for fusion run first (calls cannot
    // overtake each other), and a Lattice that spans several GPUs steps all of its slabs.
    static void phase(const LatticeData& d, unsigned ph, int (*direct)(lbm_handle), const char* what)
    {
        lbm_handle h = of(d);
        if (Lattice* owner = Lattice::owner_of(h)) owner->runPhase(ph, what);
        else report(direct(h), what);
    }

    void run_calculate_streaming(dim3, dim3, LatticeData lattice, LatticeData) { phase(lattice, LBM_PHASE_STREAM, lbm_stream, "run_calculate_streaming"); }

    // the timescale of THIS call (the reference kernel divides by its argument, SimCalcFuncs.cu:98-156), not the FluidData's
    void run_calculate_collision(dim3, dim3, LatticeData lattice, float timescale)
    {
        lbm_handle h = of(lattice);
        if (Lattice* owner = Lattice::owner_of(h)) owner->runCollision(timescale);
        else report(lbm_collide_tau(h, timescale), "run_calculate_collision");
    }

    void run_calculate_reflections(dim3, dim3, LatticeData lattice, LatticeData) { phase(lattice, LBM_PHASE_REFLECT, lbm_reflect, "run_calculate_reflections"); }

    // the force of this call is added to the caller's ReflectionData (device memory in the reference, Lattice.cpp:190-193;
    // host memory works too), whether or not Lattice::setReflectionData was called
    void run_calculate_reflections_data(dim3, dim3, LatticeData lattice, LatticeData, ReflectionData* reflectionData)
    {
        lbm_handle h = of(lattice);
        if (Lattice* owner = Lattice::owner_of(h)) owner->runReflections(reflectionData);
        else report(lbm_reflect_accumulate(h, reflectionData ? &reflectionData->x : nullptr), "run_calculate_reflections_data");
    }

    void run_prime_points(dim3, dim3, LatticeData lattice) { report(lbm_prime_points(of(lattice)), "run_prime_points"); }

    // (the source a Lattice was given with setFlowData; the FlowData argument of a direct call is the device copy of it)
    void run_generate_flow(dim3, dim3, LatticeData lattice, FlowData*) { phase(lattice, LBM_PHASE_SOURCE, lbm_source, "run_generate_flow"); }
}
