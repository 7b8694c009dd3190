// lbm_device.h -- layout of the lattice in HBM and the launch interface of the kernels.
//
// Populations live in a padded structure-of-arrays  f[q][xp][yp][zp]  (fp32, z fastest):
//   * one ghost layer on every side (xp = nx+2, yp = ny+2); ghost cells implement the
//     domain-edge rule, so the hot kernel has no bounds tests: zero-inflow ghosts hold 0
//     (the reference's rule, simulation/source/SimCalcFuncs.cu:37-40,87-94), periodic
//     ghosts are refreshed from the opposite face, wall ghosts are refreshed with the bounced-back
//     populations of the cells that pull from them, halo ghosts come from the neighbouring slab
//     (x and y faces; the z edge is resolved by address selection when the populations are pulled);
//   * the first interior z sits at PAD_Z = 32 floats, so every interior row starts on a
//     128-byte line and a warp's 32 lanes map to exactly one solid-mask word;
//   * zp is a multiple of 32, so every row, plane and population starts 128-byte aligned.
// Solid flags are one bit per padded cell: mask[xp][yp][zp/32].
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "d3q.h"

namespace lbm {

constexpr int PAD_Z = 32;
constexpr int FORCE_SLOTS = 128;   // partial force sums, summed on read-back

enum : unsigned { PH_STREAM = 1u, PH_REFLECT = 2u, PH_COLLIDE = 4u, PH_SOURCE = 8u };

struct Geom {
    int nx, ny, nz;
    int xp, yp, zp;       // padded extents
    int wz;               // mask words per row = zp / 32
    long long plane;      // yp * zp
    long long np;         // xp * yp * zp  (floats per population)
    __host__ __device__ long long cell(int x, int y, int z) const
    {
        return ((long long)(x + 1) * yp + (y + 1)) * zp + (z + PAD_Z);
    }
};

// Measurement knobs (profiles/r02_notes.md, the 1024^3 study): LBM_PAD_ROWS extra rows per plane and LBM_PAD_LINES
// extra 128-byte lines per row, so that plane and row strides stop being powers of two.  The padding is never touched.
inline int geom_pad(const char* name)
{
    const char* e = getenv(name);
    const int v = e ? atoi(e) : 0;
    return v > 0 && v <= 64 ? v : 0;
}

inline Geom make_geom(int nx, int ny, int nz)
{
    Geom g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.xp = nx + 2; g.yp = ny + 2 + geom_pad("LBM_PAD_ROWS");
    g.zp = ((PAD_Z + nz + 1 + 31) / 32) * 32 + 32 * geom_pad("LBM_PAD_LINES");
    g.wz = g.zp / 32;
    g.plane = (long long)g.yp * g.zp;
    g.np = (long long)g.xp * g.plane;
    return g;
}

// one reflection link: population j that streamed into solid cell S is handed to fluid
// cell T = S + c_d as population d with weight w (reference: one (cell, slot) pair of
// ReflectionValues, simulation/include/LatticePoint.h:3-7, applied at SimCalcFuncs.cu:177)
struct Link {
    uint32_t src_cell;    // padded cell index of S
    uint8_t  j;           // population read at S (= slot / 3)
    uint8_t  d;           // population written at T
    uint8_t  slot;        // reference slot index (force bookkeeping)
    uint8_t  pad;
    float    w;
};

struct StepParams {
    const float* fin;
    float*       fout;
    Geom  g;
    int   x0, x1;             // interior plane range [x0, x1) this launch covers
    float tau, inv_tau;
    unsigned phases;
    int   boundary_mode;      // 0 reflect (reference), 1 bounce-back
    const uint32_t* solid;    // padded bit mask or nullptr
    const uint8_t*  cls;      // class of every interior 32-cell word, [x][word][y], pitch cls_pitch
    int   cls_pitch;
    const uint32_t* special_list;   // mask-word indices of the special words
    long long n_special;      // entries this launch covers ...
    long long special_k0;     // ... starting at this one (the list is sorted, so a range of planes is a range of entries)
    int   dead_mode;          // all-solid words: 0 process like any other, 1 write zeros, 2 skip
    const uint32_t* cellinfo; // [n_special][32], list order: bits 0..26 solid flags of the cell's 3x3x3 neighbourhood, bit 27 reflection target
    const uint32_t* tmask;    // link-target bits or nullptr
    const uint32_t* tprefix;  // exclusive popcount prefix per mask word
    const uint32_t* toff;     // CSR offsets per target cell
    const Link*     links;
    double* force;            // FORCE_SLOTS x 3 partial sums or nullptr
    int   force_mode;
    // LBM_FORCE_ACCUM_REFERENCE_FP32: instead of summing, every link stores the amount it moved at its rank in the
    // oracle's order (link_rank, 0xffffffff = contributes nothing); force_sequential_kernel adds them up afterwards
    const uint32_t* link_rank;
    float* force_moved;
    int   src_on;
    int   src_lo[3], src_hi[3];
    float src_f[27];
    const uint32_t* src_mask; // optional: one bit per padded cell (the layout of `solid`), cells of the box that take the source
    int   bc[3];
    float wall_u[6][3];       // wall velocities (the z walls are applied at pull time, x and y walls by ghost fills)
    int   wall_moving;
    // element offsets from a cell, precomputed on the host so that the kernels address with one constant-bank add
    // and keep no per-population strides in registers: pop_off[i] = i * np (the cell's own slot of population i),
    // pull_off[i] = i * np - (c_ix * plane + c_iy * zp + c_iz) (the slot population i is pulled from)
    long long pop_off[27];
    long long pull_off[27];
    // x slabs with peer-memory halos (lbm_halo_connect): a pull that leaves the slab through the low / high x face reads
    // the neighbouring slab's current buffer directly (NVLink): xdelta[side] = element distance from this slab's ghost
    // plane to the neighbour's edge plane in the neighbour's allocation.  0 = the own ghost plane (zero inflow at the
    // end of the domain, or planes moved by the caller).
    // Slabs may differ in thickness, hence in population stride: the distance is kept per population.
    long long xdelta[2][27];
    int   xdelta_on;
};

// single-buffer (AA pattern) engine, lbm_inplace.cu
#ifndef LBM_DEAD_BOXES
#define LBM_DEAD_BOXES 4      // boxes of all-solid words the single-buffer kernels skip without a load (find_dead_boxes)
#endif
struct InplaceParams {
    float* f;
    Geom  g;
    float tau, inv_tau;
    int   parity;             // 0: f*_i(x) parked in A[inv(i)][x] (next step pulls and pushes); 1: pushed to A[i][x + c_i]
    int   bc[3];              // 1 periodic, 2 wall
    float wall_u[6][3];
    // element offsets from a cell, precomputed on the host (see StepParams): pop_off[i] = i * np,
    // pull_off[i] = inv(i) * np - (c_i as an offset), push_off[i] = i * np + (c_i as an offset)
    long long pop_off[27];
    long long pull_off[27];
    long long push_off[27];
    long long park_off[27];   // = pop_off[inv(i)]: where the even step parks population i (its own array, so that the row
                              // loop's loads and stores do not share offsets the compiler then keeps in registers)
    // obstacles (half-way bounce-back), shared with the two-buffer engine: padded bit mask, word classes, the
    // sorted list of special words and their neighbourhood words (all nullptr / 0 on an all-fluid lattice)
    const uint32_t* solid;
    const uint8_t*  cls;
    int   cls_pitch;
    const uint32_t* special_list;
    long long n_special;
    const uint32_t* cellinfo;
    double* force;            // FORCE_SLOTS x 3 partial sums of the momentum exchange, or nullptr
    // The momentum a cell hands to an obstacle in an EVEN step is carried by the populations the preceding odd step
    // parked in the cell's own slots: the odd step's special-word kernel sums it ahead of time into force_pending
    // (counted once the even step has run, lbm_get_force), so that the even step needs no special-word pass.
    // nullptr: the even step sums it itself (aa_even_special_kernel; LBM_AA_FORCE_SPLIT=1, kept for comparison).
    double* force_pending;
    int   src_on;             // additive box source, applied after the collision
    int   src_lo[3], src_hi[3];
    float src_f[27];
    const uint32_t* src_mask; // optional: one bit per padded cell, cells of the box that take the source
    int   x0, x1;             // interior planes [x0, x1) this launch covers (the list entries of those planes: special_k0, n_special)
    long long special_k0;
    // shift[a][0 / 1]: element offset added to a neighbour address that leaves the lattice through the low / high edge
    // of axis a under a periodic or halo rule.  Periodic: the far side of the own array (+/- n_a * stride_a).  Halo (x
    // slabs): the distance from this slab's ghost plane to the neighbouring slab's edge plane IN THE NEIGHBOUR'S
    // ALLOCATION (peer memory over NVLink, lbm_halo_connect), so that edge cells pull from and push to the neighbour
    // directly; 0 when the caller moves ghost planes itself (lbm_plane_ptr / lbm_inplace_merge_edges).
    // The x entries are kept per population slot (xshift[side][slot]; shift[0] is unused): neighbouring slabs may differ
    // in thickness and therefore in population stride.
    long long shift[3][2];
    long long xshift[2][27];
    // Lattices with obstacles: two box tests on these parameters keep loads off the critical path of the big kernels
    // (csrc/lbm_capi.cu find_dead_boxes; measured in profiles/r02_notes.md section 4).
    int   dead_box[LBM_DEAD_BOXES][6];   // boxes of all-solid words, origin (x, y, word) and extent; extent 0 = unused
    int   busy_box[6];        // bounding box of the words that are not ordinary, origin and extent; outside it no class byte is read
};

// does cell (x, y, z) take the additive source?  The reference asks a device predicate for every cell
// (SimCalcFuncs.cu:200-215, FlowSurface.h:9-47); here the predicate was evaluated once on the host: an index box (the two
// criteria the reference ships are boxes) and, for anything else, one bit per cell inside the box.
template <class Params>
__device__ __forceinline__ bool source_hit(const Params& p, int x, int y, int z)
{
    if (x < p.src_lo[0] || x > p.src_hi[0] || y < p.src_lo[1] || y > p.src_hi[1] || z < p.src_lo[2] || z > p.src_hi[2]) return false;
    if (!p.src_mask) return true;
    const uint32_t w = p.src_mask[((long long)(x + 1) * p.g.yp + (y + 1)) * p.g.wz + ((z + PAD_Z) >> 5)];
    return (w >> ((unsigned)(z + PAD_Z) & 31u)) & 1u;
}

// offsets of the in-place engine for a lattice of geometry g (everything else is left untouched)
inline void set_inplace_offsets(InplaceParams& ip, const Geom& g)
{
    ip.g = g;
    for (int i = 0; i < 27; ++i) {
        const long long c = (long long)dir_c(i, 0) * g.plane + (long long)(dir_c(i, 1) * g.zp + dir_c(i, 2));
        ip.pop_off[i] = (long long)i * g.np;
        ip.pull_off[i] = (long long)dir_inv(i) * g.np - c;
        ip.push_off[i] = (long long)i * g.np + c;
        ip.park_off[i] = (long long)dir_inv(i) * g.np;
    }
    const long long span[3] = {(long long)g.nx * g.plane, (long long)g.ny * g.zp, (long long)g.nz};
    for (int a = 0; a < 3; ++a) { ip.shift[a][0] = span[a]; ip.shift[a][1] = -span[a]; }
    for (int i = 0; i < 27; ++i) { ip.xshift[0][i] = span[0]; ip.xshift[1][i] = -span[0]; }
    ip.x0 = 0; ip.x1 = g.nx;
    ip.special_k0 = 0;
}

cudaError_t launch_inplace_zero_solid(const InplaceParams& p, int q, cudaStream_t stream);
cudaError_t launch_inplace_merge_edges(const InplaceParams& p, int q, cudaStream_t stream);
// `side` (optional): the face and special-word kernels of an odd step run there, next to the interior launch on `stream`
// (they touch disjoint slots); the caller orders the two streams before and after
cudaError_t launch_inplace_step(const InplaceParams& p, int q, int moment, int eq64, cudaStream_t stream, cudaStream_t side, int* launches);
cudaError_t launch_inplace_fold_pending(double* force, double* pending, cudaStream_t stream);
cudaError_t launch_inplace_scatter(const float* dense, float* f, const Geom& g, int q_index, int x0, int nxc, cudaStream_t stream);
cudaError_t launch_inplace_gather(const InplaceParams& p, int q, float* dense, int q_index, int x0, int nxc, cudaStream_t stream);
cudaError_t launch_inplace_moments(const InplaceParams& p, int q, int moment_mode, float* rho, float* ux, float* uy, float* uz, cudaStream_t stream);
cudaError_t launch_inplace_mass(const InplaceParams& p, int q, double* out4, cudaStream_t stream);

struct LaunchCfg {
    int q;
    int moment_mode;
    int eq64;
    int general;              // 0: pure stream+collide fluid fast path
};

// launches (all asynchronous on `stream`)
// stream_special (optional): the special-word kernel of the general path runs there, concurrently with the rest
// *launches (optional) is incremented by the number of kernels launched
cudaError_t launch_step(const LaunchCfg& cfg, const StepParams& p, cudaStream_t stream, cudaStream_t stream_special, int* launches);

// one thread: acc[a] = fma(moved[r], coef[3 r + a], acc[a]) for r = 0 .. n-1, in that order (the reference's single fp32 triple)
cudaError_t launch_force_sequential(const float* moved, const float* coef, long long n, float* acc3, cudaStream_t stream);

cudaError_t launch_fill_ghosts(float* f, const Geom& g, int q, const int bc[3], const float wall_u[6][3], bool include_x,
                               cudaStream_t stream, int* launches);

// AoS (144-byte LatticePoint records, host order) <-> padded SoA; x range [x0, x0+nxc)
cudaError_t launch_unpack_aos(const void* aos, float* f, uint8_t* solid_bytes, const Geom& g, int q, int x0, int nxc, cudaStream_t stream);
cudaError_t launch_pack_aos(void* aos, const float* f, const uint32_t* solid, const uint32_t* reflected, const Geom& g, int q, int x0, int nxc, cudaStream_t stream);
// dense [nxc][ny][nz] chunk of one population <-> padded
cudaError_t launch_scatter_dense(const float* dense, float* fq, const Geom& g, int x0, int nxc, cudaStream_t stream);
cudaError_t launch_gather_dense(float* dense, const float* fq, const Geom& g, int x0, int nxc, cudaStream_t stream);

// solid bytes (host order, optionally with x ghost planes) -> padded bit mask incl. ghost bits
cudaError_t launch_build_mask(const uint8_t* solid_bytes, int has_x_ghost, uint32_t* mask, const Geom& g, const int bc[3], cudaStream_t stream);
// classify interior mask words and list the special ones; then (list sorted) write their neighbourhood words
cudaError_t launch_build_cellinfo(const uint32_t* solid, const uint32_t* tmask, uint8_t* cls, int cls_pitch,
                                  uint32_t* special_list, unsigned* counter, const Geom& g, cudaStream_t stream);
cudaError_t launch_fill_cellinfo(const uint32_t* solid, const uint32_t* tmask, const uint32_t* special_list, long long n_special,
                                 uint32_t* cellinfo, const Geom& g, cudaStream_t stream);
cudaError_t launch_mask_to_bytes(const uint32_t* mask, uint8_t* solid_bytes, const Geom& g, cudaStream_t stream);

// moments / reductions
cudaError_t launch_moments(const float* f, const Geom& g, int q, int moment_mode, float* rho, float* ux, float* uy, float* uz, cudaStream_t stream);
cudaError_t launch_mass_momentum(const float* f, const Geom& g, int q, double* out4, cudaStream_t stream);

}  // namespace lbm
