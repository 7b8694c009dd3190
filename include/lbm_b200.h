/*
 * lbm_b200.h -- C ABI of liblbm_b200.so: the B200-native (sm_100a) replacement for the
 * per-timestep path of CharlieNicholls/LatticeBoltzmannCuda.
 *
 * The reference has no FFI layer; its narrowest waist is `namespace RunCudaFunctions`
 * (simulation/include/SimCalcFuncs.cuh:13-26) driven by `class Lattice`
 * (simulation/include/Lattice.h:40-71).  Every entry point below names the reference
 * interface it replaces.  The C++ shim in latticeboltzmanncuda_b200/include_compat/ keeps
 * the reference's class and function signatures on top of these calls.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or an
 * lbm_status code, and lbm_last_error() describes the last failure on the calling thread.
 * The handle owns all device memory; the caller owns every host buffer.  Calls on one
 * handle must come from one thread at a time.  Work is queued on the handle's CUDA stream;
 * functions that return data to the host synchronise that stream first.
 *
 * Host cell order is the reference's: index = z + y*nz + x*nz*ny
 * (simulation/source/Lattice.cpp:254).
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LBM_API
#else
#define LBM_API __attribute__((visibility("default")))
#endif

typedef struct lbm_lattice_s *lbm_handle;

enum lbm_status {
    LBM_OK = 0,
    LBM_ERR_INVALID = 1,      /* bad argument */
    LBM_ERR_CUDA = 2,         /* a CUDA runtime call or kernel failed */
    LBM_ERR_NOMEM = 3,
    LBM_ERR_STATE = 4,        /* call not valid in the handle's current state */
    LBM_ERR_UNSUPPORTED = 5
};

/* domain-edge rule per axis.  ZERO_INFLOW is the reference's only rule: a pull from
 * outside the box yields 0.0 (simulation/source/SimCalcFuncs.cu:37-40,87-94). */
enum lbm_bc {
    LBM_BC_ZERO_INFLOW = 0,
    LBM_BC_PERIODIC = 1,
    LBM_BC_WALL = 2,          /* half-way bounce-back at the face, optional wall velocity */
    LBM_BC_HALO = 3           /* x axis only: ghost planes are filled by the caller (slab decomposition) */
};

/* velocity moment.  REFERENCE scales c_i by SPLIT[i] = 0,1,1/2,1/3
 * (simulation/source/SimCalcFuncs.cu:14,107-113); STANDARD is sum f_i c_i. */
enum lbm_moment_mode { LBM_MOMENT_REFERENCE = 0, LBM_MOMENT_STANDARD = 1 };

/* arithmetic of the collision.  FP64 evaluates the equilibrium exactly like the reference
 * (sequential fp32 moments, double equilibrium, divide by tau; SimCalcFuncs.cu:98-156);
 * FP32 is the throughput mode (fp32 FMA, multiply by 1/tau), within 1e-5 relative. */
enum lbm_eq_precision { LBM_EQ_FP32 = 0, LBM_EQ_FP64 = 1 };

/* treatment of solid cells.  REFLECT is the reference: solid cells receive populations,
 * cells carrying a ReflectionValues table scatter them into fluid neighbours, then solid
 * cells are zeroed (SimCalcFuncs.cu:158-198).  BOUNCE_BACK is plain half-way bounce-back
 * from the packed solid bitmask (the throughput configurations of BASELINE.json). */
enum lbm_boundary_mode { LBM_BOUNDARY_REFLECT = 0, LBM_BOUNDARY_BOUNCE_BACK = 1 };

/* force accumulator.  REFERENCE_SLOT_INDEX reproduces SimCalcFuncs.cu:181-183 (incoming
 * direction taken from the slot index; slots >= 27 read past the reference's tables and are
 * skipped); PHYSICAL uses slot/3 for every slot. */
enum lbm_force_mode { LBM_FORCE_REFERENCE_SLOT_INDEX = 0, LBM_FORCE_PHYSICAL = 1 };

/* how the force is summed.  F64_PARTIALS: per-warp sums into 128 double slots (order-independent to ~1e-15, the
 * default).  REFERENCE_FP32: the reference's accumulator -- ONE float triple that every reflected slot adds to
 * (SimCalcFuncs.cu:181-183) -- in a fixed order: solid cells in host order, slots ascending, one fused multiply-add per
 * term.  The reference's atomics arrive in no particular order, so its own sum is one of many; this mode fixes the order
 * the CPU oracle uses (oracle/lbm_oracle.c lbo_reflect) and reproduces that float sum, including the increments a large
 * fp32 accumulator absorbs.  Table reflection on the two-buffer engine, one handle (not slabs); it adds a short serial
 * kernel per step. */
enum lbm_force_accum { LBM_FORCE_ACCUM_F64_PARTIALS = 0, LBM_FORCE_ACCUM_REFERENCE_FP32 = 1 };

/* phases of one reference step, in the CLI's order (LatticeBoltzmannCuda.cpp:89-92) */
enum lbm_phase {
    LBM_PHASE_STREAM = 1,     /* Lattice::simulateStreaming   -> run_calculate_streaming   */
    LBM_PHASE_REFLECT = 2,    /* Lattice::simulateReflections -> run_calculate_reflections */
    LBM_PHASE_COLLIDE = 4,    /* Lattice::simulateCollision   -> run_calculate_collision   */
    LBM_PHASE_SOURCE = 8,     /* Lattice::simulateFlow        -> run_generate_flow         */
    LBM_PHASE_ALL = 15
};

enum lbm_source_kind { LBM_SOURCE_NONE = 0, LBM_SOURCE_BOX = 1 };

/* streaming engine.  PULL2 keeps two copies of the populations (A -> B) and runs everything.
 * INPLACE is the single-buffer AA pattern (half the footprint, same traffic): fused
 * stream + collide steps (lbm_step / lbm_step_phases with STREAM and COLLIDE set),
 * periodic or wall edges, LBM_BC_HALO on x for slabs (lbm_halo_connect: the edge cells work on the
 * neighbouring slab's memory; or, without a connection, the caller moves the edge planes around
 * every odd step through lbm_plane_ptr / lbm_inplace_merge_edges), all-fluid lattices or a solid
 * mask in LBM_BOUNDARY_BOUNCE_BACK mode, source (box or mask), force accumulator, SoA transfers;
 * everything else (table reflection, unfused phases, zero-inflow edges, AoS transfers,
 * lbm_step_range) returns LBM_ERR_UNSUPPORTED.  On this engine lbm_set_solid_mask is accepted
 * before stepping or after an even number of steps (LBM_ERR_STATE otherwise).  AUTO = PULL2. */
enum lbm_engine { LBM_ENGINE_AUTO = 0, LBM_ENGINE_PULL2 = 1, LBM_ENGINE_INPLACE = 2 };

typedef struct lbm_config {
    int32_t nx, ny, nz;           /* Lattice::Lattice(x, y, z, ...), Lattice.cpp:22-39 */
    int32_t q;                    /* 27 (reference) or 19 (its first 19 directions) */
    float   tau;                  /* FluidData::m_characteristicTimescale, Lattice.h:24-31 */
    int32_t bc[3];                /* enum lbm_bc per axis x, y, z */
    int32_t moment_mode;          /* enum lbm_moment_mode */
    int32_t eq_precision;         /* enum lbm_eq_precision */
    int32_t boundary_mode;        /* enum lbm_boundary_mode */
    int32_t force_mode;           /* enum lbm_force_mode */
    int32_t engine;               /* enum lbm_engine */
    int32_t device;               /* CUDA device ordinal, -1: current device */
    float   wall_u[6][3];         /* wall velocity for LBM_BC_WALL faces, index axis*2 + side(0 low,1 high) */
    int32_t force_accum;          /* enum lbm_force_accum */
    int32_t reserved[7];
} lbm_config;

/* fills *cfg with the reference's behaviour: q=27, tau=1, zero inflow, SPLIT moments,
 * fp64 equilibrium, table reflection */
LBM_API int lbm_config_default(lbm_config *cfg);

/* Lattice::Lattice / ~Lattice (Lattice.cpp:22-44).  All populations start at zero (the
 * reference's tests assume that of cudaMalloc3D memory). */
LBM_API int lbm_create(const lbm_config *cfg, lbm_handle *out);
LBM_API int lbm_destroy(lbm_handle h);

/* Lattice::load_data / retrieve_data (Lattice.cpp:73-113): `cells` is an array of
 * nx*ny*nz 144-byte LatticePoint records (simulation/include/LatticePoint.h:9-17) in host
 * order.  Upload reads particle_distribution and isInternal.  Download writes x, y, z,
 * particle_distribution, isReflected, isInternal and leaves the pointer fields alone. */
LBM_API int lbm_upload_aos(lbm_handle h, const void *cells, size_t count);
LBM_API int lbm_download_aos(lbm_handle h, void *cells, size_t count);

/* populations only, dense float[q][nx][ny][nz]; f may point to host (pageable or pinned) or device memory */
LBM_API int lbm_upload_f_soa(lbm_handle h, const float *f);
LBM_API int lbm_download_f_soa(lbm_handle h, float *f);

/* LatticePoint::isInternal for every cell (what Lattice::insertModel writes,
 * Lattice.cpp:248-257): one byte per cell in host order, non-zero = solid.  With
 * x_ghost_planes != 0 the array carries nx+2 planes (the neighbours' edge planes first and
 * last), used by slab decomposition. */
LBM_API int lbm_set_solid_mask(lbm_handle h, const uint8_t *solid, int x_ghost_planes);
LBM_API int lbm_get_solid_mask(lbm_handle h, uint8_t *solid);

/* the table Lattice::preProcessModel builds (Lattice.cpp:323-426): n 380-byte
 * ReflectionValues records (LatticePoint.h:3-7) and, for each, the host index of the solid
 * cell that owns it.  Replaces LatticePoint::d_reflections / isReflected. */
LBM_API int lbm_set_reflection_table(lbm_handle h, const void *tables, const int64_t *cell_index, int n);

/* the same for a slab (x axis LBM_BC_HALO): the cell indices count nx + 2 planes, the neighbouring slabs' edge planes first
 * and last (index = z + y*nz + (x+1)*nz*ny), like lbm_set_solid_mask with x_ghost_planes.  A boundary-solid cell in a
 * neighbour's edge plane reflects into this slab too, so its table is needed here as well. */
LBM_API int lbm_set_reflection_table_slab(lbm_handle h, const void *tables, const int64_t *cell_index, int n);
/* where this slab sits in the whole lattice: its first plane is global plane x0 of gx.  lbm_insert_model and
 * lbm_preprocess_model then work in global coordinates (lattice point = global index * spacing) and produce the
 * neighbours' edge planes as well; the ends of the x axis are the reference's zero-inflow edges. */
LBM_API int lbm_set_slab_origin(lbm_handle h, int x0, int gx);

/* Lattice::setFlowData (Lattice.cpp:428-436) with the criterion as data instead of a device
 * function pointer: every cell with lo[a] <= coord[a] <= hi[a] gets f_i += f[i] each step.
 * FlowCriterion::testCriteria is the box (1,1,1)-(1,1,1); testPlaneCriteria is x in [0,0]. */
LBM_API int lbm_set_source(lbm_handle h, int kind, const int32_t lo[3], const int32_t hi[3], const float f[27]);

/* the same for ANY criterion (FlowSurface.h:12 takes an arbitrary device predicate, SimCalcFuncs.cu:206 calls it per cell
 * per step): the caller evaluates it once and hands over one byte per cell in host order, non-zero = the cell takes the
 * source.  Kept on the device as one bit per cell; the kernels look at it only inside the box around the selected cells. */
LBM_API int lbm_set_source_mask(lbm_handle h, const uint8_t *cells, const float f[27]);

/* Lattice::setReflectionData / retrieve_reflection_data (Lattice.cpp:61-71,438-443) */
LBM_API int lbm_set_force_accumulator(lbm_handle h, int enable, const float init[3]);
LBM_API int lbm_set_force_accumulator_f64(lbm_handle h, int enable, const double init[3]);
LBM_API int lbm_get_force(lbm_handle h, float out[3]);
/* the same sum before it is narrowed to the reference's float triple (the kernels keep per-warp partial sums in double):
 * for checkpoints and for adding up the slabs of a decomposed lattice */
LBM_API int lbm_get_force_f64(lbm_handle h, double out[3]);

/* RunCudaFunctions::run_prime_points (SimCalcFuncs.cu:240-243): coordinates are implicit
 * in this layout, so this only validates the handle. */
LBM_API int lbm_prime_points(lbm_handle h);

/* one phase at a time, as Lattice::simulate* call them (Lattice.cpp:115-220) */
LBM_API int lbm_stream(lbm_handle h);
LBM_API int lbm_reflect(lbm_handle h);
LBM_API int lbm_collide(lbm_handle h);
LBM_API int lbm_source(lbm_handle h);

/* RunCudaFunctions::run_calculate_collision(blocks, threads, lattice, timescale) (SimCalcFuncs.cu:225-228): one collision
 * phase with the timescale of THIS call; the handle's own tau (FluidData) is unchanged. */
LBM_API int lbm_collide_tau(lbm_handle h, float timescale);
/* RunCudaFunctions::run_calculate_reflections_data(blocks, threads, lattice, temp, ReflectionData *) (SimCalcFuncs.cu:
 * 235-238): one reflect phase whose force is ADDED to the three floats at `xyz` -- the caller's ReflectionData, in device
 * memory like the reference's d_reflectionData, or in host memory -- whether or not the handle's own accumulator
 * (lbm_set_force_accumulator) is on.  NULL: a plain lbm_reflect. */
LBM_API int lbm_reflect_accumulate(lbm_handle h, float *xyz);

/* n fused steps; each equals stream, reflect, collide, source in that order
 * (LatticeBoltzmannCuda.cpp:89-92).  lbm_step_phases runs n fused steps of a subset, e.g.
 * STREAM|COLLIDE is Lattice::simulateLattice (Lattice.cpp:222-226). */
LBM_API int lbm_step(lbm_handle h, int n);
LBM_API int lbm_step_phases(lbm_handle h, int n, unsigned phases);

/* n fused steps bracketed by CUDA events on the handle's stream; *ms gets the elapsed
 * device time and *launches the number of kernels launched */
LBM_API int lbm_step_timed(lbm_handle h, int n, unsigned phases, float *ms, int *launches);

/* density and velocity moments of the current state, dense float[nx*ny*nz] each (any may
 * be NULL), computed with the handle's moment_mode; the outputs may be host or device memory */
LBM_API int lbm_moments(lbm_handle h, float *rho, float *ux, float *uy, float *uz);
/* sum over all cells of sum_i f_i, and of sum_i f_i c_i, accumulated in double */
LBM_API int lbm_mass(lbm_handle h, double *mass);
LBM_API int lbm_momentum(lbm_handle h, double out[3]);

LBM_API int lbm_sync(lbm_handle h);
LBM_API const char *lbm_last_error(void);
LBM_API const char *lbm_version(void);

/* ---- slab decomposition (one handle per GPU, x axis LBM_BC_HALO) ----------------------- */

/* device pointer and byte size of plane xp (0 = low ghost, 1..nx interior, nx+1 = high
 * ghost) of population q in the CURRENT state.  Planes are contiguous. */
LBM_API int lbm_plane_ptr(lbm_handle h, int q, int xp, void **dev_ptr, size_t *bytes);
/* populations that cross the +x face (c_x = +1, side = 1) or the -x face (side = 0) */
LBM_API int lbm_face_populations(lbm_handle h, int side, int32_t *qs, int *nq);
/* the handle's stream as a cudaStream_t */
LBM_API int lbm_get_stream(lbm_handle h, void **stream);
/* run all further work of this handle on a caller-owned cudaStream_t (NULL: back to the handle's own
 * stream), e.g. the stream a communication library orders its transfers against */
LBM_API int lbm_set_stream(lbm_handle h, void *stream);
/* step only planes [x0, x1) of the slab (0-based interior planes), without swapping
 * buffers; lbm_step_commit swaps.  Lets the caller compute edge planes first, start the
 * halo exchange, then the interior. */
LBM_API int lbm_step_range(lbm_handle h, unsigned phases, int x0, int x1, int on_side_stream);
LBM_API int lbm_step_commit(lbm_handle h);
/* next-state plane pointer (valid between lbm_step_range and lbm_step_commit) */
LBM_API int lbm_next_plane_ptr(lbm_handle h, int q, int xp, void **dev_ptr, size_t *bytes);
/* fill the y and z ghost layers (periodic axes) of the current state, including the x ghost planes */
LBM_API int lbm_fill_ghosts(lbm_handle h);
/* single-buffer slabs (LBM_ENGINE_INPLACE with LBM_BC_HALO on x), after an odd step: the populations the
 * neighbouring slabs pushed across the faces have been delivered into the unused slots of the ghost planes
 * (c_x = +1 slots of padded plane 0, c_x = -1 slots of padded plane nx+1); this moves them into the edge planes,
 * except where the slot holds this slab's own bounced population (y/z wall, solid neighbour). */
LBM_API int lbm_inplace_merge_edges(lbm_handle h);

/* ---- slabs with peer-memory halos: the exchange below this interface ---------------------------------------------
 * The reference runs on one GPU (no cudaSetDevice anywhere); a `Lattice` that spans the GPUs of a box is
 * a build extension with one requirement: the same answer as one GPU.  One handle per slab (x axis LBM_BC_HALO).  After
 * lbm_halo_connect a slab's edge cells read (two-buffer engine) or read and write (single-buffer engine) the
 * neighbouring slab's edge plane where it lies -- peer memory over NVLink, no ghost-plane copies, no communication
 * library -- and lbm_step / lbm_step_phases order the slabs with one epoch word per face (csrc/lbm_halo.cu).  All slabs
 * of a domain must take the same steps.  Before reading state back (downloads, moments, mass) synchronise every slab
 * (lbm_sync on each; across processes also a barrier): on the single-buffer engine after an odd number of steps part of
 * an edge cell's state sits in the neighbour's memory and is fetched from there.
 *   one process, several GPUs : create the handles (cfg.device), export + connect them, then lbm_group_step;
 *   one process per GPU       : export, exchange the blobs (any transport), connect, then lbm_step on every rank. */
#define LBM_HALO_BLOB_BYTES 320
/* describes this slab to its neighbours (same-process pointers and CUDA IPC handles of its buffers); resets the epoch,
 * so every slab exports before any slab connects, and no slab steps between its export and its connect */
LBM_API int lbm_halo_export(lbm_handle h, void *blob, size_t blob_bytes);
/* lo_blob / hi_blob: the export of the slab below / above (NULL: none, the domain ends there with zero inflow; the
 * single-buffer engine needs both).  A ring closes a periodic x axis; one slab may be its own neighbour. */
LBM_API int lbm_halo_connect(lbm_handle h, const void *lo_blob, const void *hi_blob);
LBM_API int lbm_halo_disconnect(lbm_handle h);
/* number of handshakes that gave up waiting for a neighbour (LBM_HALO_TIMEOUT_MS, default 20000); 0 in a healthy run.
 * lbm_sync returns LBM_ERR_STATE when it is not 0. */
LBM_API int lbm_halo_timeouts(lbm_handle h, unsigned long long *count);
/* n steps of all slabs held by ONE process, in lockstep (every slab's "ready" signal is queued before any slab waits) */
LBM_API int lbm_group_step(lbm_handle *slabs, int count, int n, unsigned phases);

/* lbm_collide_tau / lbm_reflect_accumulate for all slabs of one process (the sum over the slabs is added to xyz) */
LBM_API int lbm_group_collide_tau(lbm_handle *slabs, int count, float timescale);
LBM_API int lbm_group_reflect_accumulate(lbm_handle *slabs, int count, float *xyz);
/* number of CUDA devices this process sees */
LBM_API int lbm_device_count(int *count);

/* layout facts for tools: padded dims and strides in floats */
typedef struct lbm_layout {
    int32_t nx, ny, nz, q;
    int32_t xp, yp, zp;       /* padded extents */
    int32_t pad_z;            /* first interior z index within a padded row */
    int64_t plane;            /* yp*zp */
    int64_t pop_stride;       /* xp*yp*zp */
    int64_t state_bytes;      /* bytes of one copy of the populations */
    int32_t n_buffers;
} lbm_layout;
LBM_API int lbm_get_layout(lbm_handle h, lbm_layout *out);

/* classification of the 32-cell words of the lattice (set when the geometry is set): ordinary fluid words,
 * dead words (every cell solid, never touched by a step) and special words (next to a solid cell or
 * holding a reflection target) */
typedef struct lbm_stats {
    int64_t words_total, words_dead, words_special;
    int64_t links;            /* reflection links (table mode) */
    int64_t kernel_launches;  /* launches issued by this handle so far */
} lbm_stats;
LBM_API int lbm_get_stats(lbm_handle h, lbm_stats *out);

/* ---- geometry side inputs (the callers of the path) ---------------------- */

typedef struct lbm_mesh_s *lbm_mesh;

/* Model::importModel (simulation/source/Model.cpp:22-44): reads an OBJ file.  A missing file gives an empty
 * model, like the reference; a mesh that is not all triangles or not closed is "invalid" (nothing is inside). */
LBM_API int lbm_mesh_load(const char *path, lbm_mesh *out);
LBM_API int lbm_mesh_destroy(lbm_mesh m);
/* number_of_vertices / faces, Model::isModelClosed (Model.cpp:46-49), !Model::checkError (Model.h:34) */
LBM_API int lbm_mesh_info(lbm_mesh m, int *n_vertices, int *n_faces, int *closed, int *valid);
/* Model::bounding_box (Model.cpp:70-73) */
LBM_API int lbm_mesh_bounding_box(lbm_mesh m, float lo[3], float hi[3]);
/* Model::isPointInsideModel (Model.cpp:51-63): inside or on the boundary */
LBM_API int lbm_mesh_point_inside(lbm_mesh m, const float p[3], int *inside);
/* Model::reflectionVector (Model.cpp:75-91): mirror of the direction p1 - p2 about the first facet the ray
 * from p1 towards p2 hits; (0,0,0) when it hits nothing */
LBM_API int lbm_mesh_reflection_vector(lbm_mesh m, const float p1[3], const float p2[3], float out[3]);
/* rotateVectorAroundAxis (simulation/source/Geometry.cpp:22-42) */
LBM_API int lbm_rotate_vector_around_axis(const float axis[3], const float vec[3], float angle, float out[3]);
/* Lattice::distributeVector (Lattice.cpp:262-321): the 27 directions ranked by closeness to v */
LBM_API int lbm_distribute_vector(const float v[3], float scores[27], int32_t dirs[27]);
/* Utils::compressInts (simulation/source/Utils.cpp:4-29) */
LBM_API unsigned lbm_compress_ints(const int *array, int num, int bit_length);

/* Lattice::insertModel (Lattice.cpp:228-260): voxelise the mesh on the lattice points (index * spacing) on
 * the GPU and make the result the solid mask.  *accepted = 0 when the bounding-box guard rejects the mesh
 * (or the file was missing): the lattice is left alone, as in the reference. */
LBM_API int lbm_insert_model(lbm_handle h, lbm_mesh m, float spacing, int *accepted);
/* Lattice::preProcessModel (Lattice.cpp:323-426): build the reflection tables of the boundary-solid cells on
 * the GPU and install them (as lbm_set_reflection_table would). */
LBM_API int lbm_preprocess_model(lbm_handle h, lbm_mesh m, float spacing);
/* the tables currently installed: up to `capacity` 380-byte records and the host index of their cells.
 * Pass NULL buffers to query *n only. */
LBM_API int lbm_get_reflection_tables(lbm_handle h, void *tables, int64_t *cell_index, int capacity, int *n);

#ifdef __cplusplus
}
#endif
#endif
