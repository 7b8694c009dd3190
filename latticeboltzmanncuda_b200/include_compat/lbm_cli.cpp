// lbm_cli.cpp -- the reference's command line contract (LatticeBoltzmannCuda.cpp:11-96) on the compatibility layer:
//
//   lbm_cli <model.obj> <threads x y z> <blocks x y z> <fluid density> <tau> <lattice spacing> <steps> [print_every]
//           [--gpus N] [--q 19|27] [--boundary reflect|bounce] [--bc zpw] [--moments reference|standard]
//           [--eq fp64|fp32] [--engine pull2|inplace] [--time 1]
//
// The eleven positional arguments are the reference's; the named options select what the reference has no switch for
// (SURVEY.md section 5): x slabs across N GPUs of this process (peer-memory halos), the velocity set, the boundary model,
// the edge rule per axis (one letter each: z zero inflow, p periodic, w wall), the velocity moment, the arithmetic of the
// equilibrium and the streaming engine.  Defaults = the reference's behaviour on one GPU (or the LBM_* environment
// variables, Lattice::Options::fromEnvironment).  --time 1 writes the seconds the step loop took (device work finished) and
// the lattice-cell updates per second to stderr; stdout stays the reference's force trace.
//
// The lattice is threads * blocks cells per axis; every population starts at 0.001 and the +x population at 0.002;
// the same values are injected on the plane x == 0 after every step; the force accumulated on the body is written
// as one "Fx Fy Fz" line before each step.  The optional 12th argument prints that line every K steps only (the
// reference blocks on a device-to-host copy every step; 0 = never).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "lbm_reference_api.hpp"

namespace {

struct Options {
    const char* model = nullptr;
    dim3 threads, blocks;
    float density = 1.0f, tau = 1.0f, spacing = 1.0f;
    int steps = 0, print_every = 1;
    bool time_loop = false;
    Lattice::Options lattice = Lattice::Options::fromEnvironment();
    int nx() const { return (int)(threads.x * blocks.x); }
    int ny() const { return (int)(threads.y * blocks.y); }
    int nz() const { return (int)(threads.z * blocks.z); }
};

bool parse(int argc_all, char* argv_all[], Options& o)
{
    // named options may stand anywhere; what is left are the positional arguments
    std::vector<char*> pos;
    for (int k = 0; k < argc_all; ++k) {
        const std::string a = argv_all[k];
        if (a.rfind("--", 0) != 0) { pos.push_back(argv_all[k]); continue; }
        if (k + 1 >= argc_all) return false;
        const std::string v = argv_all[++k];
        if (a == "--time") o.time_loop = std::atoi(v.c_str()) != 0;
        else if (a == "--gpus") o.lattice.gpus = std::atoi(v.c_str()) > 0 ? std::atoi(v.c_str()) : 1;
        else if (a == "--q") { if (v != "19" && v != "27") return false; o.lattice.q = std::atoi(v.c_str()); }
        else if (a == "--boundary") { if (v != "reflect" && v != "bounce") return false; o.lattice.boundary_mode = v == "bounce"; }
        else if (a == "--moments") { if (v != "reference" && v != "standard") return false; o.lattice.moment_mode = v == "standard"; }
        else if (a == "--eq") { if (v != "fp64" && v != "fp32") return false; o.lattice.eq_precision = v == "fp64"; }
        else if (a == "--engine") { if (v != "pull2" && v != "inplace") return false; o.lattice.engine = v == "inplace" ? 2 : 1; }
        else if (a == "--bc") {
            if (v.size() != 3) return false;
            for (int ax = 0; ax < 3; ++ax) {
                if (v[ax] != 'z' && v[ax] != 'p' && v[ax] != 'w') return false;
                o.lattice.bc[ax] = v[ax] == 'p' ? 1 : (v[ax] == 'w' ? 2 : 0);
            }
        } else return false;
    }
    const int argc = (int)pos.size();
    char** argv = pos.data();
    if (argc < 12) return false;
    auto u = [&](int k) { return (unsigned)std::atoi(argv[k]); };
    o.model = argv[1];
    o.threads = dim3(u(2), u(3), u(4));
    o.blocks = dim3(u(5), u(6), u(7));
    o.density = (float)std::atof(argv[8]);
    o.tau = (float)std::atof(argv[9]);
    o.spacing = (float)std::atof(argv[10]);
    o.steps = std::atoi(argv[11]);
    if (argc > 12) o.print_every = std::atoi(argv[12]);
    return true;
}

// the inflow record: what every cell starts with and what the source plane receives each step
LatticePoint inflow_record()
{
    LatticePoint p{};
    for (float& f : p.particle_distribution) f = 0.001f;
    p.particle_distribution[6] = 0.002f;        // direction 6 is +x
    return p;
}

void fill_uniform(Lattice& lattice, const Options& o, const LatticePoint& value)
{
    const size_t cells = (size_t)o.nx() * o.ny() * o.nz();
    std::unique_ptr<LatticePoint[]> host(new LatticePoint[cells]());
    for (size_t k = 0; k < cells; ++k)
        for (int i = 0; i < 27; ++i) host[k].particle_distribution[i] = value.particle_distribution[i];
    lattice.load_data(host.get());
}

}  // namespace

int main(int argc, char* argv[])
{
    Options o;
    if (!parse(argc, argv, o)) {
        std::fprintf(stderr, "usage: %s model.obj tx ty tz bx by bz density tau spacing steps [print_every] [--gpus N] [--q 19|27] "
                             "[--boundary reflect|bounce] [--bc zpw] [--moments reference|standard] [--eq fp64|fp32] [--engine pull2|inplace]\n", argv[0]);
        return -1;
    }
    Lattice lattice(o.nx(), o.ny(), o.nz(), o.blocks, o.threads, FluidData(o.density, o.tau), o.spacing, o.lattice);
    const LatticePoint inflow = inflow_record();
    fill_uniform(lattice, o, inflow);

    // geometry: voxelise the mesh, build the reflection tables
    lattice.insertModel(o.model);
    lattice.preProcessModel();

    // source on the plane x == 0, force accumulator starting at zero
    FlowData source(&FlowCriterion::ptestPlaneCriteria, inflow);
    lattice.setFlowData(&source);
    ReflectionData force;
    lattice.setReflectionData(&force);

    RunCudaFunctions::run_prime_points(o.blocks, o.threads, LatticeData(lattice.getCudaDataPointer(), lattice.getDimensions()));

    lattice.synchronize();
    const auto t0 = std::chrono::steady_clock::now();
    for (int step = 0; step < o.steps; ++step) {
        if (o.print_every > 0 && step % o.print_every == 0) {
            const ReflectionData* f = lattice.retrieve_reflection_data();
            std::cout << f->x << " " << f->y << " " << f->z << std::endl;
        }
        lattice.simulateStreaming();
        lattice.simulateReflections();
        lattice.simulateCollision();
        lattice.simulateFlow();
    }
    lattice.synchronize();
    if (o.time_loop) {
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "lbm_cli: %d steps of %dx%dx%d in %.6f s = %.1f MLUPS on %d GPU(s)\n", o.steps, o.nx(), o.ny(), o.nz(), sec,
                     (double)o.nx() * o.ny() * o.nz() * o.steps / sec / 1e6, lattice.deviceCount());
    }
    return 0;
}
