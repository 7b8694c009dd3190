// compat_selftest.cpp -- the reference's gtest cases (test/CalculationsTest.cpp, CoordinatesTest.cpp,
// ModelTest.cpp, GeometryTest.cpp) re-expressed against the compatibility headers, as one plain executable
// (GoogleTest is fetched from the network by the reference's CMake and is not available here).
// Each block names the reference test it follows.  Exit code = number of failed checks.
// usage: compat_selftest <dir with the OBJ fixtures>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "Lattice.h"
#include "LatticePoint.h"
#include "SimCalcFuncs.cuh"
#include "Model.h"
#include "Constants.h"
#include "FlowSurface.h"
#include "FlowCriterion.cuh"
#include "Utils.h"

static int g_failed = 0, g_checks = 0;
#define EXPECT_NEAR(a, b, tol) do { ++g_checks; const double a_ = (a), b_ = (b); if (!(std::fabs(a_ - b_) <= (tol))) { ++g_failed; \
    printf("  FAIL %s:%d  %s = %.9g, expected %.9g\n", __FILE__, __LINE__, #a, a_, b_); } } while (0)
#define EXPECT_TRUE(c) do { ++g_checks; if (!(c)) { ++g_failed; printf("  FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))

static LatticePoint* zeroed(size_t n)
{
    LatticePoint* a = new LatticePoint[n];
    std::memset((void*)a, 0, sizeof(LatticePoint) * n);
    return a;
}

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : "../../tests/golden/obj";
    const FluidData fluid(1.0, 1.0);

    printf("CalculationsTest.TestStreaming\n");       // test/CalculationsTest.cpp:19-45
    {
        dim3 threads(1, 2, 1), blocks(1, 1, 1);
        Lattice lat(1, 2, 1, blocks, threads, fluid, 0.1);
        LatticeData data(lat.getCudaDataPointer(), lat.getDimensions());
        LatticePoint* a = zeroed(2);
        a[0].particle_distribution[5] = 2.0;
        lat.load_data(a);
        RunCudaFunctions::run_prime_points(blocks, threads, data);
        lat.simulateStreaming();
        LatticePoint* r = lat.retrieve_data();
        EXPECT_NEAR(r[1].particle_distribution[5], 2.0, 1e-6);
        delete[] a; delete[] r;
    }

    printf("CalculationsTest.TestCollision\n");       // :47-75
    {
        dim3 threads(1, 1, 1), blocks(1, 1, 1);
        Lattice lat(1, 1, 1, blocks, threads, fluid, 0.1);
        LatticePoint* a = zeroed(1);
        a[0].particle_distribution[0] = 2.0;
        lat.load_data(a);
        lat.simulateCollision();
        LatticePoint* r = lat.retrieve_data();
        EXPECT_NEAR(r[0].particle_distribution[0], 0.592592593, 1e-6);
        EXPECT_NEAR(r[0].particle_distribution[1], 0.148148148, 1e-6);
        EXPECT_NEAR(r[0].particle_distribution[2], 0.148148148, 1e-6);
        delete[] a; delete[] r;
    }

    printf("CalculationsTest.TestFull\n");            // :77-105
    {
        dim3 threads(1, 2, 1), blocks(1, 1, 1);
        Lattice lat(1, 2, 1, blocks, threads, fluid, 0.1);
        LatticePoint* a = zeroed(2);
        a[0].particle_distribution[5] = 2.0;
        lat.load_data(a);
        lat.simulateLattice();
        LatticePoint* r = lat.retrieve_data();
        EXPECT_NEAR(r[1].particle_distribution[0], -0.296296296, 1e-6);
        EXPECT_NEAR(r[1].particle_distribution[1], -0.074074074, 1e-6);
        EXPECT_NEAR(r[1].particle_distribution[2], 0.148148148, 1e-6);
        delete[] a; delete[] r;
    }

    for (int with_data = 0; with_data < 2; ++with_data) {
        printf("CalculationsTest.TestReflections%s\n", with_data ? "Data" : "");     // :107-197
        dim3 threads(3, 3, 3), blocks(1, 1, 1);
        Lattice lat(3, 3, 3, blocks, threads, fluid, 0.5);
        lat.insertModel(dir + "/cube_angled.obj");
        lat.preProcessModel();
        auto at = [&](int x, int y, int z) { dim3 d = lat.getDimensions(); return z + (y * d.z) + (x * d.z * d.y); };
        LatticePoint* a = lat.retrieve_data();
        a[at(1, 1, 1)].particle_distribution[1] = 1.0;
        lat.load_data(a);
        lat.preProcessModel();
        ReflectionData* refl = new ReflectionData;
        if (with_data) lat.setReflectionData(refl);
        lat.simulateReflections();
        LatticePoint* r = lat.retrieve_data();
        EXPECT_NEAR(r[at(2, 1, 2)].particle_distribution[17], 0.367402, 1e-6);
        EXPECT_NEAR(r[at(1, 1, 2)].particle_distribution[4], 0.341121, 1e-6);
        EXPECT_NEAR(r[at(2, 2, 2)].particle_distribution[26], 0.291477, 1e-6);
        EXPECT_NEAR(r[at(1, 1, 1)].particle_distribution[1], 0.0, 1e-6);
        if (with_data) {
            refl = lat.retrieve_reflection_data();
            EXPECT_NEAR(refl->x, 0.28086, 1e-6);
            EXPECT_NEAR(refl->y, -0.194318, 1e-6);
            EXPECT_NEAR(refl->z, 0.648261, 1e-6);
        }
        delete[] a; delete[] r; delete refl;
    }

    printf("CalculationsTest.TestFlowGeneration / TestPlaneFlowGeneration\n");       // :199-277
    for (int plane = 0; plane < 2; ++plane) {
        dim3 threads(3, 3, 3), blocks(1, 1, 1);
        Lattice lat(3, 3, 3, blocks, threads, fluid, 0.5);
        auto at = [&](int x, int y, int z) { return z + y * 3 + x * 9; };
        LatticePoint flowRef;
        std::memset((void*)&flowRef, 0, sizeof flowRef);
        flowRef.particle_distribution[0] = 1.0;
        FlowData* flow = new FlowData(plane ? &FlowCriterion::ptestPlaneCriteria : &FlowCriterion::ptestCriteria, flowRef);
        lat.setFlowData(flow);
        lat.simulateFlow();
        LatticePoint* r = lat.retrieve_data();
        if (!plane) {
            EXPECT_NEAR(r[at(1, 1, 1)].particle_distribution[0], 1.0, 1e-6);
            EXPECT_NEAR(r[at(0, 0, 0)].particle_distribution[0], 0.0, 1e-6);
        } else {
            for (int y = 0; y < 3; ++y)
                for (int z = 0; z < 3; ++z) {
                    EXPECT_NEAR(r[at(0, y, z)].particle_distribution[0], 1.0, 1e-6);
                    EXPECT_NEAR(r[at(1, y, z)].particle_distribution[0], 0.0, 1e-6);
                }
        }
        delete[] r; delete flow;
    }

    printf("CalculationsTest.CompressionTest\n");     // :279-284
    {
        int a[6] = {1, 2, 3, 4, 5, 6};
        EXPECT_EQ(Utils::compressInts(a, 6, 5), 35754150u);
    }

    printf("CoordinatesTest.TestLoadAndRetrieve / TestCoordsMatch\n");      // test/CoordinatesTest.cpp:14-86
    {
        dim3 threads(10, 10, 10), blocks(10, 10, 10);
        Lattice lat(100, 100, 100, blocks, threads, fluid, 0.1);
        LatticePoint* a = zeroed(1000000);
        for (int i = 0; i < 1000000; ++i) a[i].particle_distribution[i % 27] = (float)i;
        lat.load_data(a);
        LatticePoint* r = lat.retrieve_data();
        int bad = 0;
        for (int x = 0; x < 100; ++x)
            for (int y = 0; y < 100; ++y)
                for (int z = 0; z < 100; ++z) {
                    const int k = z + (y * 100) + (x * 10000);
                    if (r[k].x != x || r[k].y != y || r[k].z != z || r[k].particle_distribution[k % 27] != (float)k) ++bad;
                }
        EXPECT_EQ(bad, 0);
        delete[] a; delete[] r;
    }

    printf("CoordinatesTest.TestCoordinatesWithModel\n");      // :137-169
    {
        dim3 threads(1, 1, 1), blocks(10, 10, 10);
        Lattice lat(10, 10, 10, blocks, threads, fluid, 0.1);
        lat.insertModel(dir + "/compat_cube.obj");
        LatticePoint* r = lat.retrieve_data();
        for (int y = 0; y < 10; ++y)
            for (int x = 0; x < 10; ++x) EXPECT_TRUE(!r[(10 * y) + x].isInternal);
        for (int y = 1; y < 9; ++y)
            for (int x = 1; x < 9; ++x) EXPECT_TRUE(r[500 + (10 * y) + x].isInternal);
        delete[] r;
    }

    printf("CoordinatesTest.TestDistributeVector\n");          // :171-198
    {
        dim3 one(1, 1, 1);
        Lattice lat(1, 1, 1, one, one, fluid, 0.1);
        auto res = lat.distributeVector(Point_3(1.0, 0.1, 0.1));
        const float norm = res[0].first + res[1].first + res[2].first;
        EXPECT_NEAR(res[0].first / norm, 0.4134132615337746, CONSTANTS::GEOMETRIC_TOLERANCE);
        EXPECT_NEAR(res[1].first / norm, 0.2932933692331128, CONSTANTS::GEOMETRIC_TOLERANCE);
        EXPECT_NEAR(res[2].first / norm, 0.2932933692331128, CONSTANTS::GEOMETRIC_TOLERANCE);
        EXPECT_EQ(res[0].second, 6); EXPECT_EQ(res[1].second, 18); EXPECT_EQ(res[2].second, 17);
    }

    printf("CoordinatesTest.TestReflection\n");                // :200-229
    {
        dim3 threads(3, 3, 3), blocks(1, 1, 1);
        Lattice lat(3, 3, 3, blocks, threads, fluid, 0.5);
        lat.insertModel(dir + "/cube_angled.obj");
        lat.simulateCollision();
        lat.preProcessModel();
        LatticePoint* r = lat.retrieve_data();
        EXPECT_TRUE(r[13].reflections != nullptr);
        if (r[13].reflections) {
            EXPECT_NEAR(r[13].reflections->reflection_weight[1 * 3 + 0], 0.367402, 1e-6);
            EXPECT_NEAR(r[13].reflections->reflection_weight[1 * 3 + 1], 0.341121, 1e-6);
            EXPECT_NEAR(r[13].reflections->reflection_weight[1 * 3 + 2], 0.291477, 1e-6);
            EXPECT_EQ(lat.getReflectionDirection(&r[13], 3), 17);
            EXPECT_EQ(lat.getReflectionDirection(&r[13], 4), 4);
            EXPECT_EQ(lat.getReflectionDirection(&r[13], 5), 26);
        }
        delete[] r;
    }

    printf("ModelTest.*\n");                                   // test/ModelTest.cpp:5-74
    {
        Model m;
        m.importModel(dir + "/triangle.obj");
        EXPECT_EQ(m.number_of_vertices(), 3);
        EXPECT_EQ(m.getMesh().number_of_vertices(), 3u);      // as the reference's test spells it (ModelTest.cpp:11)
        EXPECT_EQ(m.getMesh().number_of_faces(), 1u);
        m.importModel(dir + "/cube.obj");
        EXPECT_TRUE(m.isModelClosed());
        EXPECT_TRUE(m.isPointInsideModel(Point_3(0.0, 0.0, 0.0)));
        EXPECT_TRUE(!m.isPointInsideModel(Point_3(2.0, 0.0, 0.0)));
        m.importModel(dir + "/triangle.obj");
        EXPECT_TRUE(!m.isModelClosed());
        m.importModel(dir + "/quad.obj");
        EXPECT_TRUE(m.checkError());
        m.importModel(dir + "/compat_cube.obj");
        Point_3 r = m.reflectionVector(Point_3(0.5, 0.5, -1.0), Point_3(0.5, 0.5, 1.0));
        EXPECT_NEAR(r.x, 0.0, 1e-6); EXPECT_NEAR(r.y, 0.0, 1e-6); EXPECT_NEAR(r.z, -1.0, 1e-6);
        r = m.reflectionVector(Point_3(0.5, 0.5, -1.0), Point_3(0.5, 0.5, -2.0));
        EXPECT_TRUE(r == Point_3(0.0, 0.0, 0.0));
    }

    printf("GeometryTest.TestQuaternionRoatation\n");          // test/GeometryTest.cpp:11-32
    {
        Point_3 axis(cos(M_PI * 0.25), sin(M_PI * 0.25), 0.0), v(1.0, 0.0, 0.0);
        Point_3 r = rotateVectorAroundAxis(axis, v, 180.0 * (M_PI / 180.0));
        EXPECT_NEAR(r.x, 0.0, CONSTANTS::GEOMETRIC_TOLERANCE); EXPECT_NEAR(r.y, 1.0, CONSTANTS::GEOMETRIC_TOLERANCE);
        EXPECT_NEAR(r.z, 0.0, CONSTANTS::GEOMETRIC_TOLERANCE);
        r = rotateVectorAroundAxis(axis, v, 90.0 * (M_PI / 180.0));
        EXPECT_NEAR(r.x, 0.5, CONSTANTS::GEOMETRIC_TOLERANCE); EXPECT_NEAR(r.y, 0.5, CONSTANTS::GEOMETRIC_TOLERANCE);
        EXPECT_NEAR(r.z, -cos(M_PI * 0.25), CONSTANTS::GEOMETRIC_TOLERANCE);
    }

    printf("CLI loop: four phase calls per step are fused and equal simulateSteps\n");     // LatticeBoltzmannCuda.cpp:85-93
    {
        dim3 threads(8, 8, 8), blocks(2, 2, 2);
        LatticePoint flowRef;
        std::memset((void*)&flowRef, 0, sizeof flowRef);
        for (int i = 0; i < 27; ++i) flowRef.particle_distribution[i] = 0.001;
        flowRef.particle_distribution[6] = 0.002;
        float force[2][3];
        LatticePoint* out[2];
        for (int mode = 0; mode < 2; ++mode) {
            Lattice lat(16, 16, 16, blocks, threads, FluidData(1.0, 0.8), 1.0f / 16);
            LatticePoint* a = zeroed(4096);
            for (int k = 0; k < 4096; ++k)
                for (int i = 0; i < 27; ++i) a[k].particle_distribution[i] = flowRef.particle_distribution[i];
            lat.load_data(a);
            lat.insertModel(dir + "/cube_angled.obj");
            lat.preProcessModel();
            FlowData* flow = new FlowData(&FlowCriterion::ptestPlaneCriteria, flowRef);
            lat.setFlowData(flow);
            ReflectionData* refl = new ReflectionData;
            lat.setReflectionData(refl);
            if (mode == 0) {
                for (int s = 0; s < 12; ++s) { lat.simulateStreaming(); lat.simulateReflections(); lat.simulateCollision(); lat.simulateFlow(); }
            } else {
                lat.simulateSteps(12);
            }
            refl = lat.retrieve_reflection_data();
            force[mode][0] = refl->x; force[mode][1] = refl->y; force[mode][2] = refl->z;
            out[mode] = lat.retrieve_data();
            delete[] a; delete flow; delete refl;
        }
        int diff = 0;
        for (int k = 0; k < 4096; ++k)
            for (int i = 0; i < 27; ++i)
                if (out[0][k].particle_distribution[i] != out[1][k].particle_distribution[i]) ++diff;
        EXPECT_EQ(diff, 0);
        for (int a = 0; a < 3; ++a) EXPECT_NEAR(force[0][a], force[1][a], 1e-6 * (1 + std::fabs(force[1][a])));
        EXPECT_TRUE(std::fabs(force[1][0]) > 0.0);
        delete[] out[0]; delete[] out[1];
    }

    printf("RunCudaFunctions honour their arguments: timescale of the call, ReflectionData of the call, a criterion that is no box\n");
    {
        // SimCalcFuncs.cu:225-228: the kernel divides by the timescale it is GIVEN (the FluidData's is 1.0 here)
        dim3 threads(1, 1, 1), blocks(1, 1, 1);
        Lattice lat(1, 1, 1, blocks, threads, fluid, 0.1);
        LatticeData data(lat.getCudaDataPointer(), lat.getDimensions());
        LatticePoint* a = zeroed(1);
        a[0].particle_distribution[0] = 2.0;
        lat.load_data(a);
        RunCudaFunctions::run_calculate_collision(blocks, threads, data, 2.0f);
        LatticePoint* r = lat.retrieve_data();
        EXPECT_NEAR(r[0].particle_distribution[0], 2.0 + (0.592592593 - 2.0) / 2.0, 1e-6);       // half way to TestCollision's answer
        EXPECT_NEAR(r[0].particle_distribution[1], 0.148148148 / 2.0, 1e-6);
        delete[] a; delete[] r;
    }
    {
        // TestReflectionsData's lattice, but the accumulator is the CALLER's (Lattice::setReflectionData never called),
        // and a phase queued by the Lattice before the direct call must run first
        dim3 threads(3, 3, 3), blocks(1, 1, 1);
        Lattice lat(3, 3, 3, blocks, threads, fluid, 0.5);
        LatticeData data(lat.getCudaDataPointer(), lat.getDimensions());
        LatticePoint* a = zeroed(27);
        lat.load_data(a);
        lat.insertModel(dir + "/cube_angled.obj");
        lat.preProcessModel();
        LatticePoint* b = lat.retrieve_data();
        b[13].particle_distribution[1] = 1.0;
        lat.load_data(b);
        lat.insertModel(dir + "/cube_angled.obj");
        lat.preProcessModel();
        ReflectionData mine;
        mine.x = 10.0f;
        RunCudaFunctions::run_calculate_reflections_data(blocks, threads, data, data, &mine);
        EXPECT_NEAR(mine.x, 10.0 + 0.28086, 1e-5);
        EXPECT_NEAR(mine.y, -0.194318, 1e-5);
        EXPECT_NEAR(mine.z, 0.648261, 1e-5);
        LatticePoint* r = lat.retrieve_data();
        EXPECT_NEAR(r[2 * 9 + 1 * 3 + 2].particle_distribution[17], 0.367402, 1e-5);
        delete[] a; delete[] b; delete[] r;
    }
    {
        struct Pattern { static bool hit(int x, int y, int z) { return (x + 2 * y + 3 * z) % 5 == 0 && x >= 1; } };
        dim3 threads(4, 4, 4), blocks(1, 1, 1);
        Lattice lat(4, 4, 4, blocks, threads, fluid, 0.1);
        LatticePoint* a = zeroed(64);
        lat.load_data(a);
        LatticePoint flowRef;
        std::memset((void*)&flowRef, 0, sizeof flowRef);
        flowRef.particle_distribution[0] = 1.0;
        static bool (*ppattern)(int, int, int) = &Pattern::hit;      // (FlowData takes the address of a function-pointer symbol)
        FlowData* flow = new FlowData(&ppattern, flowRef);
        lat.setFlowData(flow);
        lat.simulateFlow();
        LatticePoint* r = lat.retrieve_data();
        int wrong = 0;
        for (int x = 0; x < 4; ++x) for (int y = 0; y < 4; ++y) for (int z = 0; z < 4; ++z)
            if (r[z + 4 * y + 16 * x].particle_distribution[0] != (Pattern::hit(x, y, z) ? 1.0f : 0.0f)) ++wrong;
        EXPECT_EQ(wrong, 0);
        EXPECT_TRUE(lat.lastError().empty());
        delete[] a; delete[] r; delete flow;
    }

    printf("A Lattice cut into x slabs (one per device, peer-memory halos) equals the Lattice on one GPU\n");
    {
        dim3 threads(8, 8, 8), blocks(2, 2, 2);
        LatticePoint flowRef;
        std::memset((void*)&flowRef, 0, sizeof flowRef);
        for (int i = 0; i < 27; ++i) flowRef.particle_distribution[i] = 0.001;
        flowRef.particle_distribution[6] = 0.002;
        float force[3][3];
        LatticePoint* out[3];
        const int gpus[3] = {1, 2, 3};
        for (int mode = 0; mode < 3; ++mode) {
            Lattice::Options o;                       // the reference's behaviour ...
            o.gpus = gpus[mode];                      // ... on that many slabs (devices are reused when the box has fewer)
            Lattice lat(16, 16, 16, blocks, threads, FluidData(1.0, 0.8), 1.0f / 16, o);
            EXPECT_EQ(lat.deviceCount(), gpus[mode]);
            LatticeData data(lat.getCudaDataPointer(), lat.getDimensions());
            LatticePoint* a = zeroed(4096);
            for (int k = 0; k < 4096; ++k)
                for (int i = 0; i < 27; ++i) a[k].particle_distribution[i] = flowRef.particle_distribution[i];
            lat.load_data(a);
            lat.insertModel(dir + "/cube_angled.obj");
            lat.preProcessModel();
            FlowData* flow = new FlowData(&FlowCriterion::ptestPlaneCriteria, flowRef);
            lat.setFlowData(flow);
            ReflectionData* refl = new ReflectionData;
            lat.setReflectionData(refl);
            for (int s = 0; s < 6; ++s) { lat.simulateStreaming(); lat.simulateReflections(); lat.simulateCollision(); lat.simulateFlow(); }
            // and the kernel wrappers, which see only the LatticeData
            RunCudaFunctions::run_calculate_streaming(blocks, threads, data, data);
            RunCudaFunctions::run_calculate_reflections(blocks, threads, data, data);
            RunCudaFunctions::run_calculate_collision(blocks, threads, data, 0.8f);
            lat.simulateSteps(5);
            refl = lat.retrieve_reflection_data();
            force[mode][0] = refl->x; force[mode][1] = refl->y; force[mode][2] = refl->z;
            out[mode] = lat.retrieve_data();
            EXPECT_TRUE(lat.lastError().empty());
            delete[] a; delete flow; delete refl;
        }
        for (int mode = 1; mode < 3; ++mode) {
            int diff = 0, coords = 0, flags = 0;
            for (int k = 0; k < 4096; ++k) {
                for (int i = 0; i < 27; ++i)
                    if (out[0][k].particle_distribution[i] != out[mode][k].particle_distribution[i]) ++diff;
                if (out[0][k].x != out[mode][k].x || out[0][k].y != out[mode][k].y || out[0][k].z != out[mode][k].z) ++coords;
                if (out[0][k].isInternal != out[mode][k].isInternal || out[0][k].isReflected != out[mode][k].isReflected) ++flags;
            }
            EXPECT_EQ(diff, 0);                        // bit for bit
            EXPECT_EQ(coords, 0);
            EXPECT_EQ(flags, 0);
            for (int a = 0; a < 3; ++a) EXPECT_NEAR(force[mode][a], force[0][a], 1e-5 * (1 + std::fabs(force[0][a])));
        }
        EXPECT_TRUE(std::fabs(force[0][0]) > 0.0);
        for (int mode = 0; mode < 3; ++mode) delete[] out[mode];
    }

    printf("%d checks, %d failed\n", g_checks, g_failed);
    return g_failed;
}
