/*
 * lbm_oracle.h -- CPU restatement of the reference's per-timestep path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (latticeboltzmanncuda_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle_golden.py
 * against the known-answer vectors of the reference's own gtest suite
 * (test/CalculationsTest.cpp, test/CoordinatesTest.cpp, test/ModelTest.cpp,
 * test/GeometryTest.cpp), against the reference's host sources compiled unmodified
 * into oracle/_ref/libref_host.so (Geometry.cpp, Constants.cpp, Utils.cpp), and -- on
 * the GPU box -- against the reference's own CUDA kernels compiled unmodified into
 * oracle/_ref/libref_gpu.so (SimCalcFuncs.cu, FlowCriterion.cu, Lattice.cpp).
 * The one part that cannot be pinned beyond the reference's fixtures is the CGAL
 * point-in-mesh / first-hit query (CGAL is an absent third-party dependency,
 * libcgal-dev unpinned, Dockerfile:5-9): see geom_oracle.c.
 *
 * All citations are relative to /root/reference.
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* simulation/include/LatticePoint.h:3-7 -- 380 bytes */
typedef struct {
    int32_t reflection_directions[14];
    float   reflection_weight[81];
} lbo_reflection_values;

/* simulation/include/LatticePoint.h:9-17 -- 144 bytes, same field offsets as the
 * reference struct on x86-64 / nvcc (x,y,z @0/4/8, f @12, flags @120/121, ptrs @128/136). */
typedef struct {
    int32_t x, y, z;
    float   f[27];
    uint8_t is_reflected;
    uint8_t is_internal;
    uint8_t pad_[6];
    uint64_t reflections;    /* host pointer in the reference; opaque here */
    uint64_t d_reflections;  /* device pointer in the reference; opaque here */
} lbo_lattice_point;

/* velocity set, simulation/source/SimCalcFuncs.cu:12-14 */
extern const int   LBO_DIRECTIONS[27][3];
extern const int   LBO_INVERSE[27];
extern const float LBO_SPLIT[27];

enum { LBO_SRC_NONE = 0, LBO_SRC_CELL_111 = 1, LBO_SRC_PLANE_X0 = 2,
       LBO_SRC_LATTICE_PATTERN = 3 /* (x + 2y + 3z) % 5 == 0 && x >= 1: a criterion that is not an index box (test only) */ };
enum { LBO_BC_ZERO_INFLOW = 0, LBO_BC_PERIODIC = 1 };
enum { LBO_MOMENT_REFERENCE = 0, LBO_MOMENT_STANDARD = 1 };

/* host index of cell (x,y,z): z + y*Z + x*Z*Y  (simulation/source/Lattice.cpp:254) */
static inline size_t lbo_index(int x, int y, int z, int Y, int Z)
{
    return (size_t)z + (size_t)y * (size_t)Z + (size_t)x * (size_t)Z * (size_t)Y;
}

/* CudaFunctions::prime_points, SimCalcFuncs.cu:70-77 */
void lbo_prime_points(lbo_lattice_point *cells, int X, int Y, int Z);

/* CudaFunctions::calculate_streaming, SimCalcFuncs.cu:79-96 with the host sequence of
 * Lattice::simulateStreaming (Lattice.cpp:115-154): dst starts as a copy of src.
 * bc = LBO_BC_ZERO_INFLOW is the reference; LBO_BC_PERIODIC is the build extension. */
void lbo_stream(const lbo_lattice_point *src, lbo_lattice_point *dst, int X, int Y, int Z, int bc);

/* CudaFunctions::calculate_collision, SimCalcFuncs.cu:98-156 (fp32 moments, fp64 equilibrium).
 * q = 27 is the reference.  q = 19 / moment = STANDARD are build extensions
 * (D3Q19 weights 1/3, 1/18, 1/36 over directions 0..18). */
void lbo_collide(lbo_lattice_point *cells, int X, int Y, int Z, float tau, int q, int moment);

/* CudaFunctions::calculate_reflections, SimCalcFuncs.cu:158-198 with the host sequence of
 * Lattice::simulateReflections (Lattice.cpp:163-210): dst starts as a copy of src.
 * tables[table_index[cell]] replaces the d_reflections pointer (table_index < 0: none).
 * force: NULL or float[3] accumulated in place (ReflectionData, Lattice.h:33-38).
 * Force terms whose slot index is >= 27 read past the reference's tables (undefined
 * there); they are skipped here. */
void lbo_reflect(const lbo_lattice_point *src, lbo_lattice_point *dst, int X, int Y, int Z,
                 const lbo_reflection_values *tables, const int32_t *table_index, float *force);

/* CudaFunctions::generate_flow, SimCalcFuncs.cu:200-215 with FlowCriterion.cu:5-13 */
void lbo_source(lbo_lattice_point *cells, int X, int Y, int Z, int kind, const float *flow27);

/* One step in CLI order (LatticeBoltzmannCuda.cpp:85-93): stream, reflect, collide, source.
 * cells is updated in place; scratch must hold X*Y*Z records. */
void lbo_step_cli(lbo_lattice_point *cells, lbo_lattice_point *scratch, int X, int Y, int Z,
                  float tau, const lbo_reflection_values *tables, const int32_t *table_index,
                  float *force, int src_kind, const float *flow27);

/* Lattice::simulateLattice, Lattice.cpp:222-226: stream then collide. */
void lbo_step_lattice(lbo_lattice_point *cells, lbo_lattice_point *scratch, int X, int Y, int Z, float tau);

/* ---- build extensions, used to check the throughput configurations -------------- */

/* Plain SoA stepper: f[q][x][y][z] (dense, no padding), pull-stream + (bounce-back on
 * solid[] bytes) + BGK collide.  bc[3] per axis: 0 zero-inflow, 1 periodic, 2 wall
 * (half-way bounce-back at the domain face; wall_u[axis*2+side][3] is the wall velocity).
 * moment: REFERENCE/STANDARD.  eq64: evaluate equilibrium in double like the reference.
 * force: NULL or double[3] += momentum handed to obstacle cells by bounce-back this step. */
void lbo_step_soa(const float *fin, float *fout, int X, int Y, int Z, int q, float tau,
                  const int bc[3], const float wall_u[6][3], const uint8_t *solid,
                  int moment, int eq64, double *force);

/* 5-bit packing helpers, Lattice.cpp:450-477 / SimCalcFuncs.cu:18-30 */
void lbo_set_reflection_direction(lbo_reflection_values *t, int slot, unsigned value);
int  lbo_get_reflection_direction(const lbo_reflection_values *t, int slot);

/* Utils::compressInts, Utils.cpp:4-29 */
unsigned lbo_compress_ints(const int *array, int num, int bit_length);

int lbo_num_threads(void);
void lbo_set_num_threads(int n);

/* ---- geometry side input (geom_oracle.c) ---------------------------------------- */

typedef struct lbo_mesh lbo_mesh;

/* Model::importModel, Model.cpp:22-44.  Returns NULL only on allocation failure; a
 * missing file gives an empty, "valid" mesh exactly like the reference (it returns early). */
lbo_mesh *lbo_mesh_load(const char *path);
void  lbo_mesh_free(lbo_mesh *m);
int   lbo_mesh_num_vertices(const lbo_mesh *m);
int   lbo_mesh_num_faces(const lbo_mesh *m);
int   lbo_mesh_is_closed(const lbo_mesh *m);     /* Model::isModelClosed, Model.cpp:46-49 */
int   lbo_mesh_check_error(const lbo_mesh *m);   /* Model::checkError, Model.h:34 */
void  lbo_mesh_bbox(const lbo_mesh *m, float lo[3], float hi[3]);   /* Model.cpp:70-73 */
int   lbo_mesh_point_inside(const lbo_mesh *m, float px, float py, float pz); /* Model.cpp:51-63 */
/* Model::reflectionVector, Model.cpp:75-91 */
void  lbo_mesh_reflection_vector(const lbo_mesh *m, const float p1[3], const float p2[3], float out[3]);

/* rotateVectorAroundAxis, Geometry.cpp:22-42 */
void lbo_rotate_vector_around_axis(const float axis[3], const float vec[3], float angle, float out[3]);

/* Lattice::distributeVector, Lattice.cpp:262-321: scores[27], dirs[27] ranked descending */
void lbo_distribute_vector(const float v[3], float scores[27], int dirs[27]);

/* Lattice::insertModel, Lattice.cpp:228-260: fills is_internal.  Returns 0 when the
 * bounding-box guard rejects the mesh (nothing written), 1 otherwise. */
int lbo_insert_model(const lbo_mesh *m, lbo_lattice_point *cells, int X, int Y, int Z, float spacing);

/* Lattice::preProcessModel, Lattice.cpp:323-426.  tables must hold up to X*Y*Z entries
 * (or pass max_tables); table_index gets the table number of each reflecting cell or -1.
 * Returns the number of tables written. */
int lbo_preprocess_model(const lbo_mesh *m, lbo_lattice_point *cells, int X, int Y, int Z, float spacing,
                         lbo_reflection_values *tables, int max_tables, int32_t *table_index);

#ifdef __cplusplus
}
#endif
#endif
