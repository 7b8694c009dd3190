// lbm_kernels.cu -- sm_100a kernels of the per-timestep path.
//
// step_kernel fuses what the reference runs as four kernels with two whole-lattice copies
// in between (simulation/source/Lattice.cpp:115-220 driving SimCalcFuncs.cu:79-215):
//   pull-stream  ->  solid-boundary treatment  ->  BGK collide  ->  source
// One thread owns one cell; a warp owns 32 consecutive z of one (x, y) row, so every one of
// the 2*Q population accesses of a warp is a single contiguous 128-byte request.
// Roofline: HBM.  Algorithmic traffic is 2*Q*4 bytes per cell update (152 B for D3Q19,
// 216 B for D3Q27); the solid mask adds 1 bit per cell.
#include "lbm_device.h"

#include <cstdio>
#include <cstdlib>

namespace lbm {

namespace {

constexpr unsigned FULL = 0xffffffffu;

#include "lbm_collide.inl"

// velocity set for run-time indices (link lists); compile-time indices use dir_c()
__constant__ int8_t c_dir[27][3] = {
    { 0, 0, 0},
    {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1}, { 0, 0, 1}, { 0, 1, 0}, { 1, 0, 0},
    {-1,-1, 0}, {-1, 0,-1}, {-1, 0, 1}, {-1, 1, 0}, { 0,-1,-1}, { 0,-1, 1},
    { 0, 1,-1}, { 0, 1, 1}, { 1,-1, 0}, { 1, 0,-1}, { 1, 0, 1}, { 1, 1, 0},
    {-1,-1,-1}, {-1,-1, 1}, {-1, 1,-1}, {-1, 1, 1}, { 1,-1,-1}, { 1,-1, 1}, { 1, 1,-1}, { 1, 1, 1}};

// offset (in floats) from a cell to the cell population i is pulled from: x - c_i
__device__ __forceinline__ long long pull_offset_dyn(const Geom& g, int i)
{
    return (long long)c_dir[i][0] * g.plane + (long long)(c_dir[i][1] * g.zp + c_dir[i][2]);
}

// 27-bit field: bit nb_bit(dx,dy,dz) = solid flag of the neighbour at (dx,dy,dz).  The warp's
// 32 cells share one mask word per row; the words before and after it supply z-1 of lane 0 and
// z+1 of lane 31 through a funnel shift, so all loads are warp-uniform and nothing branches.
__device__ __forceinline__ uint32_t neighbour_bits(const uint32_t* __restrict__ solid, long long row, int yp, int wz,
                                                   int w, unsigned lane)
{
    uint32_t nb = 0;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const uint32_t* r = solid + (row + (long long)dx * yp + dy) * wz + w;
            // when nz is not a multiple of 32 the last interior word is also the last word of the row
            const uint32_t prev = __ldg(r - 1), c = __ldg(r), next = (w + 1 < wz) ? __ldg(r + 1) : 0u;
            const uint32_t below = __funnelshift_l(prev, c, 1);   // bit k = solid(z_k - 1)
            const uint32_t above = __funnelshift_r(c, next, 1);   // bit k = solid(z_k + 1)
            nb |= (((below >> lane) & 1u) << nb_bit(dx, dy, -1)) | (((c >> lane) & 1u) << nb_bit(dx, dy, 0)) |
                  (((above >> lane) & 1u) << nb_bit(dx, dy, 1));
        }
    }
    return nb;
}

// Boundary work of one warp that touches a solid cell, a wall ghost or a reflection target.
// nb is the cell's precomputed neighbourhood word (see build_cellinfo_kernel): bit nb_bit(dx,dy,dz)
// = solid flag of the neighbour at (dx,dy,dz), bit 13 = the cell itself.
// Returns true when the cell is dead (bounce-back solid): no collision, no source.
template <int Q>
__device__ __forceinline__ bool boundary_path(const StepParams& p, float (&f)[Q], const float* __restrict__ fin, long long widx,
                                              int x, int y, int z, unsigned lane, bool active, bool do_stream, uint32_t nb)
{
    const Geom& g = p.g;
    const bool solid_here = (nb >> nb_bit(0, 0, 0)) & 1u;
    bool dead = false;
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;     // this thread's share of the force on the solid
    bool has_force = false;
    if (p.boundary_mode == 1) {
        // ---- half-way bounce-back off solid neighbours and wall ghosts
        if (do_stream && (nb & 0x7ffffffu)) {
#pragma unroll
            for (int i = 1; i < Q; ++i) {
                if ((nb >> nb_bit(-dir_c(i, 0), -dir_c(i, 1), -dir_c(i, 2))) & 1u) {
                    const float back = fin[p.pop_off[dir_inv(i)]];
                    f[i] = back;
                    // momentum exchange with an obstacle cell (domain walls are ghost fills and never get
                    // here; periodic or halo images of obstacle cells count): population inv(i) goes in
                    // with -c_i and comes back as i with +c_i
                    if (p.force) {
                        if (active && !solid_here) {
                            const float m = back + back;
                            fx -= (float)dir_c(i, 0) * m;
                            fy -= (float)dir_c(i, 1) * m;
                            fz -= (float)dir_c(i, 2) * m;
                            has_force = true;
                        }
                    }
                }
            }
        }
        if (solid_here) {
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = 0.0f;
            dead = true;
        }
    }
    if ((p.phases & PH_REFLECT) && p.boundary_mode == 0) {
        // ---- reference reflection as a gather (scatter form: SimCalcFuncs.cu:162-187).  The links of a
        // target cell are sorted by the population they feed, so the walk below only ever touches f[i]
        // with a compile-time i and f stays in registers.
        if (p.tmask && ((nb >> 27) & 1u)) {
            const uint32_t tw = __ldg(p.tmask + widx);
            if (active) {
                const uint32_t rank = __ldg(p.tprefix + widx) + __popc(tw & ((1u << lane) - 1u));
                uint32_t e = __ldg(p.toff + rank);
                const uint32_t e1 = __ldg(p.toff + rank + 1);
#pragma unroll
                for (int i = 1; i < Q; ++i) {
                    while (e < e1 && p.links[e].d == i) {
                        const Link L = p.links[e];
                        long long src = (long long)L.j * g.np + (long long)L.src_cell;
                        if (do_stream) {
                            src -= pull_offset_dyn(g, L.j);
                            if (p.bc[2] == 1 && c_dir[L.j][2] != 0) {       // periodic z: wrap a source beyond the edge
                                const int sz = (int)(L.src_cell % (uint32_t)g.zp) - PAD_Z - c_dir[L.j][2];
                                if (sz < 0) src += g.nz; else if (sz >= g.nz) src -= g.nz;
                            }
                        }
                        if (p.xdelta_on) {
                            // the value may sit in the neighbouring slab (source plane below 1 or above nx): read it there
                            const int sp = (int)(L.src_cell / (uint32_t)g.plane) - (do_stream ? c_dir[L.j][0] : 0);
                            if (sp < 1) src += p.xdelta[0][L.j]; else if (sp > g.nx) src += p.xdelta[1][L.j];
                        }
                        const float moved = p.fin[src] * L.w;
                        f[i] += moved;
                        if (p.force_moved) {
                            const uint32_t r = __ldg(p.link_rank + e);
                            if (r != 0xffffffffu) p.force_moved[r] = moved;
                        } else if (p.force) {
                            // SimCalcFuncs.cu:181-183; incoming direction from the slot index (reference) or slot/3
                            const int in_dir = (p.force_mode == 0) ? (int)L.slot : (int)L.j;
                            if (in_dir < 27) {
                                const float so = dir_split(i), si = dir_split(in_dir);
                                fx += moved * (((float)dir_c(i, 0) * so) - ((float)c_dir[in_dir][0] * si));
                                fy += moved * (((float)dir_c(i, 1) * so) - ((float)c_dir[in_dir][1] * si));
                                fz += moved * (((float)dir_c(i, 2) * so) - ((float)c_dir[in_dir][2] * si));
                                has_force = true;
                            }
                        }
                        ++e;
                    }
                }
            }
        }
        // solid cells are emptied after the scatter (SimCalcFuncs.cu:189-197)
        if (solid_here) {
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = 0.0f;
        }
    }
    if (p.force && __any_sync(FULL, has_force)) {
        // one atomic triple per warp (the reference issues three atomics per link on one address)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(FULL, fx, o);
            fy += __shfl_xor_sync(FULL, fy, o);
            fz += __shfl_xor_sync(FULL, fz, o);
        }
        if (lane == 0) {
            // spread over FORCE_SLOTS partial sums so that neighbouring blocks do not serialise on one address
            double* slot = p.force + 3 * ((unsigned)(widx >> 2) % FORCE_SLOTS);
            atomicAdd(slot + 0, (double)fx);
            atomicAdd(slot + 1, (double)fy);
            atomicAdd(slot + 2, (double)fz);
        }
    }
    return dead;
}

// moving-wall term 6 w_i (c_i . u_wall), rho_wall = 1, for a population that bounced off the z edge; where walls
// meet, the one named first in x, y, z order supplies the velocity.  Kept out of line: it runs on the edge
// lanes of moving-wall set-ups only and must not cost the pull loop registers.
template <int Q>
__device__ __noinline__ float moving_wall_term(const StepParams& p, int i, int x, int y)
{
    const Geom& g = p.g;
    const int sx = x - c_dir[i][0], sy = y - c_dir[i][1];
    const float* u = p.wall_u[4 + (c_dir[i][2] > 0 ? 0 : 1)];
    if (p.bc[0] == 2 && (sx < 0 || sx >= g.nx)) u = p.wall_u[sx < 0 ? 0 : 1];
    else if (p.bc[1] == 2 && (sy < 0 || sy >= g.ny)) u = p.wall_u[2 + (sy < 0 ? 0 : 1)];
    const float w = (Q == 27) ? (i < 7 ? (float)(2.0 / 27.0) : (i < 19 ? (float)(1.0 / 54.0) : (float)(1.0 / 216.0)))
                              : (i < 7 ? (float)(1.0 / 18.0) : (float)(1.0 / 36.0));
    return __fmul_rn(6.0f * w, (float)c_dir[i][0] * u[0] + (float)c_dir[i][1] * u[1] + (float)c_dir[i][2] * u[2]);
}

// Launch shape of the general and special kernels; the defaults are the values measured best on B200 at 512^3
// (profiles/r01_notes.md).  Variants for measurement are built with -D (latticeboltzmanncuda_b200/build.py).
#ifndef LBM_GEN_WARPS
#define LBM_GEN_WARPS 4        // z-words (warps) per block
#endif
#ifndef LBM_GEN_MINB
#define LBM_GEN_MINB 8         // resident blocks per SM the register allocation aims at (64 registers; 10 and 12 measured slower)
#endif
#ifndef LBM_GEN_ROWS
#define LBM_GEN_ROWS 8         // rows a warp walks (4, 8 or 16)
#endif
#ifndef LBM_GEN_PREFETCH
#define LBM_GEN_PREFETCH 0     // experiment: per-lane prefetch.global.L2 of the next row's lines. Measured on B200: 2.08 -> 3.08 ms
#endif                         // (19 divergent requests per instruction load the L1 request path; profiles/r01_notes.md)
#ifndef LBM_GEN_PAIR
#define LBM_GEN_PAIR 0         // experiment: two ordinary rows at once, both rows' pulls issued before the first collision
#endif
#ifndef LBM_SPC_PREFETCH
#define LBM_SPC_PREFETCH 0     // the same for the next list entry of the special kernel: 0.35 -> 0.46 ms
#endif
#ifndef LBM_STREAM_HINTS
#define LBM_STREAM_HINTS 0     // experiment: 1 = st.global.cs for population stores, 2 = also ld.global.cs for the pulls
#endif
#ifndef LBM_SPC_ENTRIES
#define LBM_SPC_ENTRIES 1      // consecutive list entries a warp of the special kernel takes (1 = no look-ahead)
#endif
#ifndef LBM_SPC_MINB
#define LBM_SPC_MINB 8          // 63 registers, no spills (10 blocks/SM = 48 registers spills in the pull loop)
#endif

// population loads / stores of the step kernels (LBM_STREAM_HINTS selects cache-streaming variants for measurement)
__device__ __forceinline__ float ld_pop(const float* a)
{
#if LBM_STREAM_HINTS >= 2
    return __ldcs(a);
#else
    return *a;
#endif
}
__device__ __forceinline__ void st_pop(float* a, float v)
{
#if LBM_STREAM_HINTS >= 1
    __stcs(a, v);
#else
    *a = v;
#endif
}

// Pull f_i(x) <- f_i(x - c_i)  (SimCalcFuncs.cu:83-94, j = INVERSE[i] relabelled).
// x and y edges are ghost planes / rows (filled by launch_fill_ghosts or left at zero); the z edge is
// resolved here, by choosing the address the edge lane reads from, because a z ghost shares its row with
// the interior and any separate pass over it would move one 32-byte sector per 4-byte value:
//   zero inflow: the ghost column holds 0 (the reference's rule);
//   periodic   : the lane reads the cell on the far side of the row (+/- nz);
//   wall       : half-way bounce-back, the lane reads its own cell's opposite population and adds
//                6 w_i (c_i . u_wall) for a moving wall (rho_wall = 1).
// edge_word is warp-uniform: only the first and last word of a row take the second branch.
// XE: the launch covers x edge planes of a slab with peer-memory halos; on_lo / on_hi = this cell's plane is the first /
// last of the slab, so a pull through that face reads the neighbouring slab's memory (xdelta, 0 without a neighbour).
// One exception: beyond a y WALL the value would be the wall ghost of the neighbour's plane, which only this slab can
// produce (it is this very cell's bounced-back population, launch_fill_ghosts / fill_axis_kernel): the lane takes it
// from its own opposite slot, like a z wall.
template <int Q>
__device__ __forceinline__ bool through_face(int i, bool on_lo, bool on_hi)
{
    return dir_c(i, 0) != 0 && (dir_c(i, 0) > 0 ? on_lo : on_hi);
}
template <int Q>
__device__ __forceinline__ bool beyond_y_wall(const StepParams& p, int i, int y)
{
    return p.bc[1] == 2 && dir_c(i, 1) != 0 && (dir_c(i, 1) > 0 ? y == 0 : y == p.g.ny - 1);
}

template <int Q, bool XE = false>
__device__ __forceinline__ void pull_populations(const StepParams& p, const float* __restrict__ fin, int x, int y, int z,
                                                 bool edge_word, float (&f)[Q], bool on_lo = false, bool on_hi = false)
{
    const Geom& g = p.g;
    uint32_t ywalls = 0;                         // (XE) populations that bounced off a y wall next to a slab face
    if (!edge_word || p.bc[2] == 0) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            long long a = p.pull_off[i];
            if (XE && through_face<Q>(i, on_lo, on_hi)) {
                if (beyond_y_wall<Q>(p, i, y)) { a = p.pop_off[dir_inv(i)]; ywalls |= 1u << i; }
                else a += p.xdelta[dir_c(i, 0) > 0 ? 0 : 1][i];
            }
            f[i] = ld_pop(fin + a);
        }
    } else {
        const bool lo = (z == 0), hi = (z == g.nz - 1);
        const bool wall = p.bc[2] == 2;
        // all Q loads first, as one batch (only the address differs on the edge lane) ...
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            long long a = p.pull_off[i];
            bool own = false;
            if (dir_c(i, 2) != 0) {
                const bool outside = dir_c(i, 2) > 0 ? lo : hi;          // the cell this population comes from is beyond the z edge
                if (outside) {
                    if (wall) { a = p.pop_off[dir_inv(i)]; own = true; }
                    else a += (dir_c(i, 2) > 0 ? g.nz : -g.nz);
                }
            }
            if (XE && through_face<Q>(i, on_lo, on_hi)) {
                if (beyond_y_wall<Q>(p, i, y)) { a = p.pop_off[dir_inv(i)]; ywalls |= 1u << i; }
                else if (!own) a += p.xdelta[dir_c(i, 0) > 0 ? 0 : 1][i];
            }
            f[i] = ld_pop(fin + a);
        }
        // ... then the moving-wall terms of the populations that bounced (edge lanes of moving-wall set-ups only)
        if (wall && p.wall_moving && (lo || hi)) {
#pragma unroll
            for (int i = 1; i < Q; ++i)
                if (dir_c(i, 2) != 0 && (dir_c(i, 2) > 0 ? lo : hi) && !((ywalls >> i) & 1u)) f[i] += moving_wall_term<Q>(p, i, x, y);
        }
    }
    if (XE && p.wall_moving && ywalls) {
#pragma unroll
        for (int i = 1; i < Q; ++i)
            if ((ywalls >> i) & 1u) f[i] += moving_wall_term<Q>(p, i, x, y);
    }
}

// Dense kernel for an all-fluid lattice: one warp per 32 consecutive z of an (x, y) row, pull +
// collide, nothing else.  Used when there is no geometry, no source, the phases are stream + collide and
// the z edge is open (zero inflow), so that no z ghost values have to be produced.
template <int Q, int MOMENT, bool EQ64>
__global__ void __launch_bounds__(256) step_kernel(const __grid_constant__ StepParams p)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = p.x0 + blockIdx.z;
    if (z >= g.nz || y >= g.ny) return;

    const long long idx = ((long long)(x + 1) * g.yp + (y + 1)) * g.zp + (z + PAD_Z);
    const float* __restrict__ fin = p.fin + idx;
    float f[Q];
    pull_populations<Q>(p, fin, x, y, z, false, f);      // open z edge: the ghost columns hold zeros
    // ---- BGK relaxation (SimCalcFuncs.cu:98-156)
    if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
    else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    float* __restrict__ fout = p.fout + idx;
#pragma unroll
    for (int i = 0; i < Q; ++i) st_pop(fout + p.pop_off[i], f[i]);
}

// General kernel.  Every 32-cell word of every interior row was classified when the geometry was set
// (build_cellinfo_kernel): ordinary fluid, dead (all cells solid) or special (touches a solid cell, a
// wall ghost or a reflection target).  A warp owns one z-word and walks ROWS_PER_WARP consecutive rows;
// the classes of those rows are adjacent bytes, fetched with one 4-byte load from an array that is kept
// resident in L2 (persisting access window), so dead words cost no population traffic at all and
// ordinary words run the same code as the dense kernel:
//   ordinary  pull, collide, (source), store;
//   dead      nothing to read; zeros are written while the buffers may still hold stale values there
//             (dead_mode 1) and nothing at all afterwards (dead_mode 2);
//   special   skipped here: step_special_kernel walks the compact list of special words.
enum : unsigned { CLS_ORDINARY = 0, CLS_DEAD = 1, CLS_SPECIAL = 2 };

template <int Q, int MOMENT, bool EQ64, int ROWS_PER_WARP, bool SRC>
__global__ void __launch_bounds__(32 * LBM_GEN_WARPS, LBM_GEN_MINB) step_general_kernel(const __grid_constant__ StepParams p)
{
    const Geom& g = p.g;
    const unsigned lane = threadIdx.x & 31u;
    const int wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // z-word of this warp
    const int wcount = (g.nz + 31) >> 5;
    if (wi >= wcount) return;
    const int x = p.x0 + blockIdx.z;
    const int y0 = blockIdx.y * ROWS_PER_WARP;
    const int z = wi * 32 + (int)lane;
    const bool active = z < g.nz;
    const bool do_stream = p.phases & PH_STREAM;
    const uint8_t* cp = p.cls + ((long long)x * wcount + wi) * p.cls_pitch + y0;
    // the class bytes of this warp's rows, in two 64-bit registers (a small array indexed by the row would live
    // in local memory)
    unsigned long long cw0 = 0, cw1 = 0;
    if (ROWS_PER_WARP == 4) {
        cw0 = __ldg(reinterpret_cast<const uint32_t*>(cp));
    } else if (ROWS_PER_WARP == 8) {
        cw0 = __ldg(reinterpret_cast<const unsigned long long*>(cp));
    } else {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(cp));
        cw0 = v.x; cw1 = v.y;
    }

#if LBM_GEN_PREFETCH
    // ncu: 63 % of this kernel's stall samples sit on the first use of a pulled value (profiles/r01d_*): with one
    // row of loads in flight per warp the kernel is bound by DRAM latency, not by issue or bandwidth.  Lane i < Q
    // therefore owns population i's pull offset and, while the warp works on row r, asks L2 for the two 128-byte
    // lines that population's pull of row r + 1 will touch (first and last lane's address): two instructions per
    // row, no registers held across the collision besides the offset.
    const long long my_off = lane < (unsigned)Q ? (do_stream ? p.pull_off[lane] : p.pop_off[lane]) : 0;
#endif

#pragma unroll 1
    for (int r = 0; r < ROWS_PER_WARP; ++r) {
        const int y = y0 + r;
        if (y >= g.ny) break;
        const unsigned cls = (unsigned)(((ROWS_PER_WARP > 8 && r >= 8) ? cw1 : cw0) >> (8 * (r & 7))) & 0xffu;
#if LBM_GEN_PREFETCH
        if (r + 1 < ROWS_PER_WARP && y + 1 < g.ny && lane < (unsigned)Q) {
            const unsigned cn = (unsigned)(((ROWS_PER_WARP > 8 && r + 1 >= 8) ? cw1 : cw0) >> (8 * ((r + 1) & 7))) & 0xffu;
            if (cn == CLS_ORDINARY || (cn == CLS_DEAD && !p.dead_mode)) {
                const float* a = p.fin + (((long long)(x + 1) * g.yp + (y + 2)) * g.zp + (wi * 32 + PAD_Z)) + my_off;
                asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(a + 31));
            }
        }
#endif
        if (cls == CLS_SPECIAL) continue;              // step_special_kernel owns these words
        const long long idx = ((long long)(x + 1) * g.yp + (y + 1)) * g.zp + (z + PAD_Z);
#if LBM_GEN_PAIR
        // software pipelining by hand: when this row and the next are both ordinary, the 2 x Q pulls of both rows are in
        // flight before the first collision starts (the kernel waits for memory: 9.8 warps per issue on the long scoreboard)
        if (!SRC && (r & 1) == 0 && r + 1 < ROWS_PER_WARP && y + 1 < g.ny && do_stream && (p.phases & PH_COLLIDE) && cls == CLS_ORDINARY &&
            ((unsigned)(((ROWS_PER_WARP > 8 && r + 1 >= 8) ? cw1 : cw0) >> (8 * ((r + 1) & 7))) & 0xffu) == CLS_ORDINARY) {
            const bool edge = wi == 0 || wi == wcount - 1;
            float f0[Q], f1[Q];
            pull_populations<Q>(p, p.fin + idx, x, y, z, edge, f0);
            pull_populations<Q>(p, p.fin + idx + g.zp, x, y + 1, z, edge, f1);
            if (EQ64) { collide_fp64<Q, MOMENT>(f0, p.tau); collide_fp64<Q, MOMENT>(f1, p.tau); }
            else      { collide_fp32<Q, MOMENT>(f0, p.inv_tau); collide_fp32<Q, MOMENT>(f1, p.inv_tau); }
            if (active) {
                float* __restrict__ fo = p.fout + idx;
#pragma unroll
                for (int i = 0; i < Q; ++i) st_pop(fo + p.pop_off[i], f0[i]);
#pragma unroll
                for (int i = 0; i < Q; ++i) st_pop(fo + p.pop_off[i] + g.zp, f1[i]);
            }
            ++r;
            continue;
        }
#endif
        float f[Q];
        bool dead = false;
        if (cls == CLS_DEAD && p.dead_mode) {
            if (p.dead_mode == 2) continue;
            dead = true;
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = 0.0f;
        } else {
            const float* __restrict__ fin = p.fin + idx;
            if (do_stream) {
                if (SRC) {          // (the variant for the planes that need more: source box, slab faces next to peer memory)
                    pull_populations<Q, true>(p, fin, x, y, z, wi == 0 || wi == wcount - 1, f, p.xdelta_on && x == 0, p.xdelta_on && x == g.nx - 1);
                } else {
                    pull_populations<Q>(p, fin, x, y, z, wi == 0 || wi == wcount - 1, f);
                }
            } else {
#pragma unroll
                for (int i = 0; i < Q; ++i) f[i] = fin[p.pop_off[i]];
            }
            // ---- BGK relaxation (SimCalcFuncs.cu:98-156); runs on every live cell like the reference
            if (p.phases & PH_COLLIDE) {
                if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
                else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
            }
        }
        // ---- additive source on the selected cells (SimCalcFuncs.cu:200-215); bounce-back solids take none
        // (SRC is set for the launches that cover the planes the source box intersects, see launch_step_general_r)
        if (SRC && (p.phases & PH_SOURCE) && p.src_on && !(dead && p.boundary_mode == 1)) {
            if (source_hit(p, x, y, z)) {
#pragma unroll
                for (int i = 0; i < Q; ++i) f[i] += p.src_f[i];
            }
        }
        if (active) {
            float* __restrict__ fout = p.fout + idx;
#pragma unroll
            for (int i = 0; i < Q; ++i) st_pop(fout + p.pop_off[i], f[i]);
        }
    }
}

// Bounce-back pull of one cell of a special word: exactly one load per population.  A population whose
// upstream neighbour is solid (bit of nb) or lies beyond a z wall is read from the cell's own opposite slot
// (half-way bounce-back) instead of from the neighbour, so solid cells are never read; solid and inactive
// lanes load nothing at all.  Momentum exchange with obstacle cells is summed into (fx, fy, fz).
// ZEDGE (warp-uniform) = the word holds a z edge that is periodic or a wall; kept out of the common path so
// that its Q loads stay one uninterrupted batch.
template <int Q, bool ZEDGE, bool XE>
__device__ __forceinline__ void pull_bounce_back(const StepParams& p, const float* __restrict__ fin, int x, int y, int z,
                                                 uint32_t nb, bool live, float (&f)[Q],
                                                 float& fx, float& fy, float& fz, bool& has_force, bool on_lo, bool on_hi)
{
    const Geom& g = p.g;
    const bool lo = ZEDGE && (z == 0), hi = ZEDGE && (z == g.nz - 1);
    const bool wall = p.bc[2] == 2;
    uint32_t zwalls = 0;                       // populations that bounced off a z wall (moving-wall term below)
    f[0] = live ? fin[0] : 0.0f;
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        long long a = p.pull_off[i];
        bool own = (nb >> nb_bit(-dir_c(i, 0), -dir_c(i, 1), -dir_c(i, 2))) & 1u;
        if (ZEDGE && dir_c(i, 2) != 0) {
            const bool outside = dir_c(i, 2) > 0 ? lo : hi;
            if (outside && !own) {
                if (wall) { own = true; zwalls |= 1u << i; }
                else a += (dir_c(i, 2) > 0 ? g.nz : -g.nz);
            }
        }
        if (XE && !own && through_face<Q>(i, on_lo, on_hi)) {
            // through a slab face: the neighbour's memory, or (beyond a y wall) the cell's own bounced population
            if (beyond_y_wall<Q>(p, i, y)) { own = true; zwalls |= 1u << i; }
            else a += p.xdelta[dir_c(i, 0) > 0 ? 0 : 1][i];
        }
        if (own) a = p.pop_off[dir_inv(i)];
        f[i] = live ? ld_pop(fin + a) : 0.0f;
    }
    if (p.force) {
#pragma unroll
        for (int i = 1; i < Q; ++i) {
            if (live && ((nb >> nb_bit(-dir_c(i, 0), -dir_c(i, 1), -dir_c(i, 2))) & 1u)) {
                const float m = f[i] + f[i];
                fx -= (float)dir_c(i, 0) * m;
                fy -= (float)dir_c(i, 1) * m;
                fz -= (float)dir_c(i, 2) * m;
                has_force = true;
            }
        }
    }
    if ((ZEDGE || XE) && p.wall_moving && live && zwalls) {
#pragma unroll
        for (int i = 1; i < Q; ++i)
            if ((zwalls >> i) & 1u) f[i] += moving_wall_term<Q>(p, i, x, y);
    }
}

// Special words (touching a solid cell, a wall ghost or a reflection target): one warp per entry of the
// compact list built with the classification; cellinfo holds one neighbourhood word per cell of each entry,
// in list order.  Same step as everywhere else plus the boundary treatment.
// XE: the launch covers the first and / or last plane of a slab with peer-memory halos (lbm_halo_connect)
template <int Q, int MOMENT, bool EQ64, bool XE>
__global__ void __launch_bounds__(128, LBM_SPC_MINB) step_special_kernel(const __grid_constant__ StepParams p)
{
    const Geom& g = p.g;
    const unsigned lane = threadIdx.x & 31u;
    // the launch covers the entries [special_k0, special_k0 + n_special) of the list: those of the planes [x0, x1);
    // a warp takes LBM_SPC_ENTRIES consecutive entries.  The list entry and the neighbourhood word of the next
    // entry are fetched, and its population lines requested from L2, while the current entry is in flight:
    // the kernel is latency-bound (two dependent loads before the first population load, ncu: long scoreboard).
    const long long kfirst = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * LBM_SPC_ENTRIES;
    if (kfirst >= p.n_special) return;
    const bool do_stream = p.phases & PH_STREAM;
#if LBM_SPC_ENTRIES > 1
#if LBM_SPC_PREFETCH
    const long long my_off = lane < (unsigned)Q ? (do_stream ? p.pull_off[lane] : p.pop_off[lane]) : 0;
#endif
    uint32_t widx_next = __ldg(p.special_list + p.special_k0 + kfirst);
    uint32_t nb_next = __ldg(p.cellinfo + (p.special_k0 + kfirst) * 32 + lane);
#pragma unroll 1
    for (int e = 0; e < LBM_SPC_ENTRIES; ++e) {
    const long long kl = kfirst + e;
    if (kl >= p.n_special) break;
    const long long k = p.special_k0 + kl;
    const uint32_t widx = widx_next, nb = nb_next;
    if (e + 1 < LBM_SPC_ENTRIES && kl + 1 < p.n_special) {
        widx_next = __ldg(p.special_list + k + 1);
        nb_next = __ldg(p.cellinfo + (k + 1) * 32 + lane);
#if LBM_SPC_PREFETCH
        if (lane < (unsigned)Q) {
            const float* a = p.fin + (long long)widx_next * 32 + my_off;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(a + 31));
        }
#endif
    }
#else
    const long long kl = kfirst;
    const long long k = p.special_k0 + kl;
    const uint32_t widx = __ldg(p.special_list + k);          // mask-word index = padded cell index / 32
    const uint32_t nb = __ldg(p.cellinfo + k * 32 + lane);
    {
#endif
    const int w = (int)(widx % (uint32_t)g.wz);
    const uint32_t row = widx / (uint32_t)g.wz;
    const int x = (int)(row / (uint32_t)g.yp) - 1, y = (int)(row % (uint32_t)g.yp) - 1;
    const int z = w * 32 - PAD_Z + (int)lane;
    const bool active = z < g.nz;
    const bool edge_word = w == PAD_Z / 32 || w == PAD_Z / 32 + ((g.nz + 31) >> 5) - 1;
    const long long idx = (long long)widx * 32 + lane;
    const float* __restrict__ fin = p.fin + idx;
    const bool on_lo = XE && x == 0, on_hi = XE && x == g.nx - 1;
    float f[Q];
    bool dead, skip_store = false;
    if (p.boundary_mode == 1 && do_stream) {
        // ---- half-way bounce-back, fused with the pull
        const bool solid_here = (nb >> nb_bit(0, 0, 0)) & 1u;
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;
        bool has_force = false;
        if (edge_word && p.bc[2] != 0) pull_bounce_back<Q, true, XE>(p, fin, x, y, z, nb, active && !solid_here, f, fx, fy, fz, has_force, on_lo, on_hi);
        else                           pull_bounce_back<Q, false, XE>(p, fin, x, y, z, nb, active && !solid_here, f, fx, fy, fz, has_force, on_lo, on_hi);
        dead = solid_here;
        // both buffers already hold zeros in the solid cells: leave out the 32-byte sectors (8 lanes) that
        // are solid throughout; a sector with a live cell is written whole, a partial one would cost a read
        const uint32_t solid_lanes = __ballot_sync(FULL, solid_here || !active);
        skip_store = p.dead_mode == 2 && ((solid_lanes >> (lane & 24u)) & 0xffu) == 0xffu;
        if (p.force && __any_sync(FULL, has_force)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                fx += __shfl_xor_sync(FULL, fx, o);
                fy += __shfl_xor_sync(FULL, fy, o);
                fz += __shfl_xor_sync(FULL, fz, o);
            }
            if (lane == 0) {
                double* slot = p.force + 3 * ((unsigned)(widx >> 2) % FORCE_SLOTS);
                atomicAdd(slot + 0, (double)fx);
                atomicAdd(slot + 1, (double)fy);
                atomicAdd(slot + 2, (double)fz);
            }
        }
    } else {
        if (do_stream) {
            pull_populations<Q, XE>(p, fin, x, y, z, edge_word, f, on_lo, on_hi);
        } else {
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = fin[p.pop_off[i]];
        }
        dead = boundary_path<Q>(p, f, fin, (long long)widx, x, y, z, lane, active, do_stream, nb);
    }
    if ((p.phases & PH_COLLIDE) && !dead) {
        if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
        else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    }
    if ((p.phases & PH_SOURCE) && p.src_on && !(dead && p.boundary_mode == 1)) {
        if (source_hit(p, x, y, z)) {
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] += p.src_f[i];
        }
    }
    if (active && !skip_store) {
        float* __restrict__ fout = p.fout + idx;
#pragma unroll
        for (int i = 0; i < Q; ++i) st_pop(fout + p.pop_off[i], f[i]);
    }
    }   // entries of this warp
}

// threads per block of the step kernel (tuning knob, LBM_BLOCK_THREADS = 64 / 128 / 256)
static int step_block_threads()
{
    static int v = [] {
        const char* e = getenv("LBM_BLOCK_THREADS");
        int t = e ? atoi(e) : 128;
        return (t == 64 || t == 128 || t == 256) ? t : 128;
    }();
    return v;
}

template <int Q, int MOMENT, bool EQ64>
cudaError_t launch_step_dense(const StepParams& p, cudaStream_t stream)
{
    const Geom& g = p.g;
    const int threads = step_block_threads();
    int bz = ((g.nz + 31) / 32) * 32;
    if (bz > 128) bz = 128;
    if (bz > threads) bz = threads;
    int by = threads / bz;
    if (by > g.ny) by = g.ny;
    dim3 block(bz, by, 1);
    dim3 grid((g.nz + bz - 1) / bz, (g.ny + by - 1) / by, p.x1 - p.x0);
    if (p.x1 <= p.x0) return cudaSuccess;
    static int smem = [] { const char* e = getenv("LBM_DENSE_SMEM"); return e ? atoi(e) : 0; }();   // occupancy experiment
    if (smem > 48 * 1024) cudaFuncSetAttribute(step_kernel<Q, MOMENT, EQ64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    step_kernel<Q, MOMENT, EQ64><<<grid, block, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int Q, int MOMENT, bool EQ64, int ROWS, bool SRC>
cudaError_t launch_step_general_x(const StepParams& p, int xa, int xb, cudaStream_t stream)
{
    if (xb <= xa) return cudaSuccess;
    const Geom& g = p.g;
    StepParams q = p;
    q.x0 = xa; q.x1 = xb;
    const int wcount = (g.nz + 31) / 32;
    const int warps = wcount < LBM_GEN_WARPS ? wcount : LBM_GEN_WARPS;
    dim3 block(32 * warps, 1, 1);
    dim3 grid((wcount + warps - 1) / warps, (g.ny + ROWS - 1) / ROWS, xb - xa);
    step_general_kernel<Q, MOMENT, EQ64, ROWS, SRC><<<grid, block, 0, stream>>>(q);
    return cudaGetLastError();
}

// The additive source touches the planes src_lo[0]..src_hi[0] only (the reference's plane x == 0 or one cell), and
// only the first and last plane of a slab pull from the neighbouring slab's memory: runs of such planes get their own
// launch of the variant that carries that code, every other plane runs the lean variant.
template <int Q, int MOMENT, bool EQ64, int ROWS>
cudaError_t launch_step_general_r(const StepParams& p, cudaStream_t stream, int* launches)
{
    const Geom& g = p.g;
    int sa = p.x0, sb = p.x0;      // [sa, sb) = planes of this launch the source box intersects
    if ((p.phases & PH_SOURCE) && p.src_on) {
        sa = p.src_lo[0] > p.x0 ? p.src_lo[0] : p.x0;
        sb = (p.src_hi[0] + 1) < p.x1 ? (p.src_hi[0] + 1) : p.x1;
        if (sb <= sa) sa = sb = p.x0;
    }
    const bool peer = p.xdelta_on && (p.phases & PH_STREAM);
    auto extra = [&](int x) { return (x >= sa && x < sb) || (peer && (x == 0 || x == g.nx - 1)); };
    cudaError_t e = cudaSuccess;
    int x = p.x0;
    while (x < p.x1 && e == cudaSuccess) {        // maximal runs of planes of one kind (at most five)
        const bool ex = extra(x);
        int xe = x + 1;
        if (!ex) {
            // the next plane that needs the extra variant: sa, nx - 1, or the end of the range
            xe = p.x1;
            if (sb > sa && sa > x && sa < xe) xe = sa;
            if (peer && g.nx - 1 > x && g.nx - 1 < xe) xe = g.nx - 1;
        } else {
            while (xe < p.x1 && extra(xe)) ++xe;
        }
        e = ex ? launch_step_general_x<Q, MOMENT, EQ64, ROWS, true>(p, x, xe, stream)
               : launch_step_general_x<Q, MOMENT, EQ64, ROWS, false>(p, x, xe, stream);
        if (launches) *launches += 1;
        x = xe;
    }
    return e;
}

template <int Q, int MOMENT, bool EQ64>
cudaError_t launch_step_tasks(const StepParams& p, cudaStream_t stream, cudaStream_t stream_special, int* launches)
{
    if (p.x1 <= p.x0) return cudaSuccess;
    static_assert(LBM_GEN_ROWS == 4 || LBM_GEN_ROWS == 8 || LBM_GEN_ROWS == 16, "rows per warp");
    cudaError_t e = launch_step_general_r<Q, MOMENT, EQ64, LBM_GEN_ROWS>(p, stream, launches);   // cls_pitch is a multiple of 16
    if (e != cudaSuccess) return e;
    if (p.n_special > 0) {
        const long long blocks = (p.n_special + 4 * LBM_SPC_ENTRIES - 1) / (4 * LBM_SPC_ENTRIES);
        cudaStream_t ss = stream_special ? stream_special : stream;
        const bool xe = p.xdelta_on && (p.phases & PH_STREAM) && (p.x0 == 0 || p.x1 == p.g.nx);
        if (xe) step_special_kernel<Q, MOMENT, EQ64, true><<<(unsigned)blocks, 128, 0, ss>>>(p);
        else    step_special_kernel<Q, MOMENT, EQ64, false><<<(unsigned)blocks, 128, 0, ss>>>(p);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

template <int Q>
cudaError_t launch_step_q(const LaunchCfg& c, const StepParams& p, cudaStream_t s, cudaStream_t s2, int* n)
{
    const int key = (c.moment_mode ? 4 : 0) | (c.eq64 ? 2 : 0) | (c.general ? 1 : 0);
    if (!c.general && n) *n += 1;
    switch (key) {
        case 0: return launch_step_dense<Q, 0, false>(p, s);
        case 1: return launch_step_tasks<Q, 0, false>(p, s, s2, n);
        case 2: return launch_step_dense<Q, 0, true>(p, s);
        case 3: return launch_step_tasks<Q, 0, true>(p, s, s2, n);
        case 4: return launch_step_dense<Q, 1, false>(p, s);
        case 5: return launch_step_tasks<Q, 1, false>(p, s, s2, n);
        case 6: return launch_step_dense<Q, 1, true>(p, s);
        default: return launch_step_tasks<Q, 1, true>(p, s, s2, n);
    }
}

// ---- ghost layers -----------------------------------------------------------------------

// One launch per axis fills both ghost faces of that axis for the populations that enter through them
// (c_i[axis] = +1 through the low face, -1 through the high face).
//   periodic: G_i(g) = F_i(g shifted by n along the axis)
//   wall    : half-way bounce-back written into the ghost cell, G_i(g) = F_inv(i)(g + c_i) + 6 w_i (c_i.u_wall)
//             with rho_wall = 1, only where g + c_i is an interior cell (the only cell that pulls G_i(g)).
// Faces span the full padded range of the other two axes, so running the axes in order fills edges and
// corners; walls are filled after all periodic axes and in z, y, x order so that the wall named first
// in x, y, z order wins where two walls meet.
struct FillParams {
    float* f;
    Geom g;
    int axis, mode;          // mode 1 periodic, 2 wall
    int bc[3];
    float wall_u[2][3];      // low / high face velocity of this axis
};

template <int Q, int AX>
__global__ void fill_axis_kernel(const __grid_constant__ FillParams p)
{
    const Geom& g = p.g;
    const int n[3] = {g.nx, g.ny, g.nz};
    // face coordinates (a, b): the two axes other than AX, b is the faster one in memory
    constexpr int ax = AX;
    constexpr int a_ax = ax == 0 ? 1 : 0, b_ax = ax == 2 ? 1 : 2;
    const int nb = n[b_ax] + 2, na = n[a_ax] + 2;
    const long long cell_count = (long long)na * nb;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cell_count) return;
    const int side = blockIdx.y;                     // 0 low face, 1 high face
    int c[3];
    c[a_ax] = (int)(k / nb) - 1;
    c[b_ax] = (int)(k % nb) - 1;
    c[ax] = side ? n[ax] : -1;
    const int inward = side ? -1 : 1;
    // ghost cells that also lie outside a wall axis belong to that wall (bounce-back wins over the wrap)
    if (p.mode == 1 && ((p.bc[a_ax] == 2 && (c[a_ax] < 0 || c[a_ax] >= n[a_ax])) ||
                        (p.bc[b_ax] == 2 && (c[b_ax] < 0 || c[b_ax] >= n[b_ax])))) return;
    const long long gidx = g.cell(c[0], c[1], c[2]);
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        if (dir_c(i, ax) != inward) continue;
        if (p.mode == 1) {
            int s[3] = {c[0], c[1], c[2]};
            s[ax] += side ? -n[ax] : n[ax];
            p.f[(long long)i * g.np + gidx] = p.f[(long long)i * g.np + g.cell(s[0], s[1], s[2])];
        } else {
            int t[3] = {c[0] + dir_c(i, 0), c[1] + dir_c(i, 1), c[2] + dir_c(i, 2)};
            if (p.bc[2] == 1 && ax != 2) t[2] = (t[2] + n[2]) % n[2];
            if (t[0] < 0 || t[0] >= n[0] || t[1] < 0 || t[1] >= n[1] || t[2] < 0 || t[2] >= n[2]) continue;
            const float* u = p.wall_u[side];
            const float cu = (float)dir_c(i, 0) * u[0] + (float)dir_c(i, 1) * u[1] + (float)dir_c(i, 2) * u[2];
            // (explicitly rounded: on a slab face the same value is formed by moving_wall_term() at pull time, and a product
            // contracted into the add here would differ from it in the last bit)
            p.f[(long long)i * g.np + gidx] =
                __fadd_rn(p.f[(long long)dir_inv(i) * g.np + g.cell(t[0], t[1], t[2])], __fmul_rn(6.0f * (float)dir_weight<Q>(i), cu));
        }
    }
}

// ---- AoS <-> SoA (Lattice::load_data / retrieve_data, Lattice.cpp:73-113) ------------------

constexpr int AOS_BYTES = 144;   // sizeof(LatticePoint), simulation/include/LatticePoint.h:9-17
constexpr int AOS_F = 12;        // offsetof particle_distribution
constexpr int AOS_REFLECTED = 120;
constexpr int AOS_INTERNAL = 121;

__global__ void unpack_aos_kernel(const unsigned char* __restrict__ aos, float* __restrict__ f, uint8_t* __restrict__ solid_bytes,
                                  Geom g, int q, int x0, int nxc)
{
    const long long per = (long long)g.ny * g.nz;
    const long long n = per * nxc;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int xl = (int)(k / per);
    const int rem = (int)(k - (long long)xl * per);
    const int y = rem / g.nz, z = rem - y * g.nz;
    const unsigned char* rec = aos + k * AOS_BYTES;
    const long long idx = g.cell(x0 + xl, y, z);
    const float* pf = reinterpret_cast<const float*>(rec + AOS_F);
    for (int i = 0; i < q; ++i) f[(long long)i * g.np + idx] = pf[i];
    if (solid_bytes) solid_bytes[(long long)(x0 + xl) * per + rem] = rec[AOS_INTERNAL] ? 1 : 0;
}

__global__ void pack_aos_kernel(unsigned char* __restrict__ aos, const float* __restrict__ f, const uint32_t* __restrict__ solid,
                                const uint32_t* __restrict__ reflected, Geom g, int q, int x0, int nxc)
{
    const long long per = (long long)g.ny * g.nz;
    const long long n = per * nxc;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int xl = (int)(k / per);
    const int rem = (int)(k - (long long)xl * per);
    const int y = rem / g.nz, z = rem - y * g.nz;
    const int x = x0 + xl;
    unsigned char* rec = aos + k * AOS_BYTES;
    const long long idx = g.cell(x, y, z);
    int* pc = reinterpret_cast<int*>(rec);
    pc[0] = x; pc[1] = y; pc[2] = z;
    float* pf = reinterpret_cast<float*>(rec + AOS_F);
    for (int i = 0; i < q; ++i) pf[i] = f[(long long)i * g.np + idx];
    for (int i = q; i < 27; ++i) pf[i] = 0.0f;
    const long long word = ((long long)(x + 1) * g.yp + (y + 1)) * g.wz + ((z + PAD_Z) >> 5);
    const unsigned bit = (unsigned)(z + PAD_Z) & 31u;
    rec[AOS_INTERNAL] = solid ? ((solid[word] >> bit) & 1u) : 0;
    rec[AOS_REFLECTED] = reflected ? ((reflected[word] >> bit) & 1u) : 0;
}

__global__ void scatter_dense_kernel(const float* __restrict__ dense, float* __restrict__ fq, Geom g, int x0, int nxc)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int xl = blockIdx.z;
    if (z >= g.nz) return;
    fq[g.cell(x0 + xl, y, z)] = dense[((long long)xl * g.ny + y) * g.nz + z];
}

__global__ void gather_dense_kernel(float* __restrict__ dense, const float* __restrict__ fq, Geom g, int x0, int nxc)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int xl = blockIdx.z;
    if (z >= g.nz) return;
    dense[((long long)xl * g.ny + y) * g.nz + z] = fq[g.cell(x0 + xl, y, z)];
}

// ---- solid mask ---------------------------------------------------------------------------

// One warp builds one mask word.  Ghost cells follow the domain-edge rule of each axis they
// lie outside of: wall and zero-inflow -> fluid (their populations are ghost values); periodic ->
// the wrapped interior cell; halo (x only) -> the plane handed in by the caller.
__global__ void build_mask_kernel(const uint8_t* __restrict__ sb, int has_x_ghost, uint32_t* __restrict__ mask, Geom g,
                                  int bcx, int bcy, int bcz)
{
    const long long nwords = (long long)g.xp * g.yp * g.wz;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (wid >= nwords) return;
    const int w = (int)(wid % g.wz);
    const long long r = wid / g.wz;
    const int ypad = (int)(r % g.yp), xpad = (int)(r / g.yp);
    int c[3] = {xpad - 1, ypad - 1, w * 32 + (int)lane - PAD_Z};
    const int n[3] = {g.nx, g.ny, g.nz};
    const int bc[3] = {bcx, bcy, bcz};
    // padding beyond the single ghost layer in z is never read: leave it clear
    const bool padding = (c[2] < -1 || c[2] > g.nz);
    bool wall = false, empty = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (c[a] >= 0 && c[a] < n[a]) continue;
        if (bc[a] == 2 || bc[a] == 0) empty = true;     // wall and zero-inflow ghosts carry values, not solid bits
        else if (bc[a] == 1) c[a] = (c[a] + n[a]) % n[a];
        else if (!(a == 0 && has_x_ghost)) empty = true;   // halo: keep x = -1 / nx, read the caller's plane
    }
    bool bit = false;
    if (!padding && !wall && !empty && sb) {
        const long long per = (long long)g.ny * g.nz;
        const long long xi = has_x_ghost ? (long long)c[0] + 1 : (long long)c[0];
        bit = sb[xi * per + (long long)c[1] * g.nz + c[2]] != 0;
    }
    const uint32_t word = __ballot_sync(FULL, bit);
    if (lane == 0) mask[wid] = word;
}

// Classifies every interior mask word and precomputes the neighbourhood words of the special ones.
// One warp per interior (row, word).  A word is special if any of the 27 mask words around it has a
// solid bit or it holds a reflection target; it is dead if every one of its interior cells is solid.
// cls is laid out [x][word][y] (y fastest, pitch cls_pitch) so that the classes of consecutive rows of
// one z-word are adjacent bytes.  The neighbourhood words of the special words are written afterwards, in
// the order of the sorted list (fill_cellinfo_kernel).
__global__ void build_cellinfo_kernel(const uint32_t* __restrict__ solid, const uint32_t* __restrict__ tmask,
                                      uint8_t* __restrict__ cls, int cls_pitch,
                                      uint32_t* __restrict__ special_list, unsigned* counter, Geom g)
{
    const int wlo = PAD_Z / 32, wcount = (g.nz + 31) / 32;            // words that hold interior cells
    const long long nwarps = (long long)g.nx * g.ny * wcount;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (wid >= nwarps) return;
    const int wi = (int)(wid % wcount), w = wlo + wi;
    const long long r = wid / wcount;
    const int y = (int)(r % g.ny), x = (int)(r / g.ny);
    const long long row = (long long)(x + 1) * g.yp + (y + 1);
    const long long widx = row * g.wz + w;
    uint32_t near = 0, nb = 0, own = 0;
    if (solid) {
        // lanes 0..26 each look at one of the 27 surrounding words
        if (lane < 27) {
            const int dx = (int)lane / 9 - 1, dy = ((int)lane / 3) % 3 - 1, dw = (int)lane % 3 - 1;
            if (w + dw < g.wz) near = solid[(row + (long long)dx * g.yp + dy) * g.wz + w + dw];
        }
        own = __shfl_sync(FULL, near, 13);
        near = __any_sync(FULL, near != 0u);
        if (near) nb = neighbour_bits(solid, row, g.yp, g.wz, w, lane);
    }
    const uint32_t tw = tmask ? tmask[widx] : 0u;
    const int cells = min(32, g.nz - wi * 32);
    const uint32_t live_lanes = cells == 32 ? 0xffffffffu : ((1u << cells) - 1u);
    unsigned c = CLS_ORDINARY;
    const bool lane_live = (live_lanes >> lane) & 1u;
    const bool touches = __any_sync(FULL, lane_live && (nb & 0x7ffffffu) != 0u);   // exact: some cell sees a solid cell
    if ((own & live_lanes) == live_lanes) c = CLS_DEAD;
    else if (touches || (tw & live_lanes)) {
        c = CLS_SPECIAL;
        if (lane == 0) special_list[atomicAdd(counter, 1u)] = (uint32_t)widx;
    }
    if (lane == 0) cls[((long long)x * wcount + wi) * cls_pitch + y] = (uint8_t)c;
}

// cellinfo[k * 32 + lane] for the k-th entry of the (sorted) special list: bits 0..26 = solid flags of the
// 3x3x3 neighbourhood of the cell (bit 13 the cell itself), bit 27 = reflection target.  One warp per entry;
// 128 bytes per special word, contiguous in the order the step kernel walks them.
__global__ void fill_cellinfo_kernel(const uint32_t* __restrict__ solid, const uint32_t* __restrict__ tmask,
                                     const uint32_t* __restrict__ special_list, long long n_special,
                                     uint32_t* __restrict__ cellinfo, Geom g)
{
    const long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (k >= n_special) return;
    const uint32_t widx = special_list[k];
    const int w = (int)(widx % (uint32_t)g.wz);
    const long long row = widx / (uint32_t)g.wz;
    const uint32_t nb = solid ? neighbour_bits(solid, row, g.yp, g.wz, w, lane) : 0u;
    const uint32_t tw = tmask ? tmask[widx] : 0u;
    cellinfo[k * 32 + lane] = nb | (((tw >> lane) & 1u) << 27);
}

__global__ void mask_to_bytes_kernel(const uint32_t* __restrict__ mask, uint8_t* __restrict__ sb, Geom g)
{
    const long long per = (long long)g.ny * g.nz;
    const long long n = per * g.nx;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int x = (int)(k / per);
    const int rem = (int)(k - (long long)x * per);
    const int y = rem / g.nz, z = rem - y * g.nz;
    const long long word = ((long long)(x + 1) * g.yp + (y + 1)) * g.wz + ((z + PAD_Z) >> 5);
    sb[k] = (mask[word] >> ((unsigned)(z + PAD_Z) & 31u)) & 1u;
}

// ---- moments and conservation diagnostics ----------------------------------------------------

template <int Q>
__global__ void moments_kernel(const float* __restrict__ f, Geom g, int moment_mode, float* rho, float* ux, float* uy, float* uz)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, x = blockIdx.z;
    if (z >= g.nz) return;
    const long long idx = g.cell(x, y, z);
    float r = 0.0f, mx = 0.0f, my = 0.0f, mz = 0.0f;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const float v = f[(long long)i * g.np + idx];
        const float s = (moment_mode == 0) ? dir_split(i) : 1.0f;
        r += v;
        mx += v * (float)dir_c(i, 0) * s;
        my += v * (float)dir_c(i, 1) * s;
        mz += v * (float)dir_c(i, 2) * s;
    }
    if (r != 0.0f) { mx /= r; my /= r; mz /= r; }
    const long long k = ((long long)x * g.ny + y) * g.nz + z;
    if (rho) rho[k] = r;
    if (ux) ux[k] = mx;
    if (uy) uy[k] = my;
    if (uz) uz[k] = mz;
}

// mass and momentum in double: warps stride over the interior rows, lanes over z; shuffle reduction
// per warp, shared-memory reduction per block, one atomic quadruple per block
template <int Q>
__global__ void __launch_bounds__(256) mass_momentum_kernel(const float* __restrict__ f, Geom g, double* out4)
{
    const long long rows = (long long)g.nx * g.ny;
    const int warps_per_block = blockDim.x >> 5;
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    double m = 0.0, px = 0.0, py = 0.0, pz = 0.0;
    for (long long r = (long long)blockIdx.x * warps_per_block + wib; r < rows; r += (long long)gridDim.x * warps_per_block) {
        const int x = (int)(r / g.ny), y = (int)(r % g.ny);
        const long long base = g.cell(x, y, 0);
        for (int z = (int)lane; z < g.nz; z += 32) {
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                const double v = (double)f[(long long)i * g.np + base + z];
                m += v;
                px += v * dir_c(i, 0);
                py += v * dir_c(i, 1);
                pz += v * dir_c(i, 2);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m += __shfl_xor_sync(FULL, m, o);
        px += __shfl_xor_sync(FULL, px, o);
        py += __shfl_xor_sync(FULL, py, o);
        pz += __shfl_xor_sync(FULL, pz, o);
    }
    __shared__ double part[8][4];
    if (lane == 0) { part[wib][0] = m; part[wib][1] = px; part[wib][2] = py; part[wib][3] = pz; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int k = 0; k < warps_per_block; ++k) t += part[k][threadIdx.x];
        atomicAdd(out4 + threadIdx.x, t);
    }
}

}  // namespace

// ---- launch wrappers ---------------------------------------------------------------------------

namespace {
// The reference keeps the force in one float triple that every reflected slot adds to with an atomic
// (SimCalcFuncs.cu:181-183).  This restates that accumulator with the order fixed: one thread, terms in the oracle's order,
// one fused multiply-add each -- slow by design (a serial chain), it exists for parity (LBM_FORCE_ACCUM_REFERENCE_FP32).
__global__ void force_sequential_kernel(const float* __restrict__ moved, const float* __restrict__ coef, long long n, float* acc3)
{
    float fx = acc3[0], fy = acc3[1], fz = acc3[2];
    for (long long r = 0; r < n; ++r) {
        const float m = moved[r];
        fx = fmaf(m, coef[3 * r + 0], fx);
        fy = fmaf(m, coef[3 * r + 1], fy);
        fz = fmaf(m, coef[3 * r + 2], fz);
    }
    acc3[0] = fx; acc3[1] = fy; acc3[2] = fz;
}
}  // namespace

cudaError_t launch_force_sequential(const float* moved, const float* coef, long long n, float* acc3, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    force_sequential_kernel<<<1, 1, 0, stream>>>(moved, coef, n, acc3);
    return cudaGetLastError();
}

cudaError_t launch_step(const LaunchCfg& cfg, const StepParams& p, cudaStream_t stream, cudaStream_t stream_special, int* launches)
{
    if (cfg.q == 27) return launch_step_q<27>(cfg, p, stream, stream_special, launches);
    if (cfg.q == 19) return launch_step_q<19>(cfg, p, stream, stream_special, launches);
    return cudaErrorInvalidValue;
}

static cudaError_t launch_fill_axis(float* f, const Geom& g, int q, int axis, int mode, const int bc[3], const float wall_u[6][3], cudaStream_t stream)
{
    FillParams fp;
    fp.f = f; fp.g = g; fp.axis = axis; fp.mode = mode;
    for (int a = 0; a < 3; ++a) fp.bc[a] = bc[a];
    for (int s = 0; s < 2; ++s)
        for (int a = 0; a < 3; ++a) fp.wall_u[s][a] = wall_u ? wall_u[axis * 2 + s][a] : 0.0f;
    const int n[3] = {g.nx, g.ny, g.nz};
    const int a_ax = axis == 0 ? 1 : 0, b_ax = axis == 2 ? 1 : 2;
    const long long cells = (long long)(n[a_ax] + 2) * (n[b_ax] + 2);
    dim3 grid((unsigned)((cells + 127) / 128), 2);
    if (axis == 0) {
        if (q == 27) fill_axis_kernel<27, 0><<<grid, 128, 0, stream>>>(fp);
        else         fill_axis_kernel<19, 0><<<grid, 128, 0, stream>>>(fp);
    } else {
        if (q == 27) fill_axis_kernel<27, 1><<<grid, 128, 0, stream>>>(fp);
        else         fill_axis_kernel<19, 1><<<grid, 128, 0, stream>>>(fp);
    }
    return cudaGetLastError();
}

// x and y faces only: the z edge is resolved by the step kernels at pull time (pull_populations)
cudaError_t launch_fill_ghosts(float* f, const Geom& g, int q, const int bc[3], const float wall_u[6][3], bool include_x,
                               cudaStream_t stream, int* launches)
{
    int n = 0;
    cudaError_t e = cudaSuccess;
    const int first = include_x ? 0 : 1;
    // periodic axes first, walls afterwards in y, x order so that the wall named first wins where two meet
    for (int axis = first; axis <= 1 && e == cudaSuccess; ++axis)
        if (bc[axis] == 1) { e = launch_fill_axis(f, g, q, axis, 1, bc, wall_u, stream); ++n; }
    for (int axis = 1; axis >= first && e == cudaSuccess; --axis)
        if (bc[axis] == 2) { e = launch_fill_axis(f, g, q, axis, 2, bc, wall_u, stream); ++n; }
    if (launches) *launches = n;
    return e;
}

cudaError_t launch_unpack_aos(const void* aos, float* f, uint8_t* solid_bytes, const Geom& g, int q, int x0, int nxc, cudaStream_t stream)
{
    const long long n = (long long)g.ny * g.nz * nxc;
    if (n <= 0) return cudaSuccess;
    unpack_aos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const unsigned char*)aos, f, solid_bytes, g, q, x0, nxc);
    return cudaGetLastError();
}

cudaError_t launch_pack_aos(void* aos, const float* f, const uint32_t* solid, const uint32_t* reflected, const Geom& g, int q, int x0, int nxc, cudaStream_t stream)
{
    const long long n = (long long)g.ny * g.nz * nxc;
    if (n <= 0) return cudaSuccess;
    pack_aos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((unsigned char*)aos, f, solid, reflected, g, q, x0, nxc);
    return cudaGetLastError();
}

cudaError_t launch_scatter_dense(const float* dense, float* fq, const Geom& g, int x0, int nxc, cudaStream_t stream)
{
    if (nxc <= 0) return cudaSuccess;
    dim3 grid((g.nz + 127) / 128, g.ny, nxc);
    scatter_dense_kernel<<<grid, 128, 0, stream>>>(dense, fq, g, x0, nxc);
    return cudaGetLastError();
}

cudaError_t launch_gather_dense(float* dense, const float* fq, const Geom& g, int x0, int nxc, cudaStream_t stream)
{
    if (nxc <= 0) return cudaSuccess;
    dim3 grid((g.nz + 127) / 128, g.ny, nxc);
    gather_dense_kernel<<<grid, 128, 0, stream>>>(dense, fq, g, x0, nxc);
    return cudaGetLastError();
}

cudaError_t launch_build_mask(const uint8_t* solid_bytes, int has_x_ghost, uint32_t* mask, const Geom& g, const int bc[3], cudaStream_t stream)
{
    const long long nwords = (long long)g.xp * g.yp * g.wz;
    const long long threads = nwords * 32;
    build_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(solid_bytes, has_x_ghost, mask, g, bc[0], bc[1], bc[2]);
    return cudaGetLastError();
}

cudaError_t launch_build_cellinfo(const uint32_t* solid, const uint32_t* tmask, uint8_t* cls, int cls_pitch,
                                  uint32_t* special_list, unsigned* counter, const Geom& g, cudaStream_t stream)
{
    const long long nwarps = (long long)g.nx * g.ny * ((g.nz + 31) / 32);
    const long long threads = nwarps * 32;
    build_cellinfo_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(solid, tmask, cls, cls_pitch, special_list, counter, g);
    return cudaGetLastError();
}

cudaError_t launch_fill_cellinfo(const uint32_t* solid, const uint32_t* tmask, const uint32_t* special_list, long long n_special,
                                 uint32_t* cellinfo, const Geom& g, cudaStream_t stream)
{
    if (n_special <= 0) return cudaSuccess;
    const long long threads = n_special * 32;
    fill_cellinfo_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(solid, tmask, special_list, n_special, cellinfo, g);
    return cudaGetLastError();
}

cudaError_t launch_mask_to_bytes(const uint32_t* mask, uint8_t* solid_bytes, const Geom& g, cudaStream_t stream)
{
    const long long n = (long long)g.nx * g.ny * g.nz;
    mask_to_bytes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(mask, solid_bytes, g);
    return cudaGetLastError();
}

cudaError_t launch_moments(const float* f, const Geom& g, int q, int moment_mode, float* rho, float* ux, float* uy, float* uz, cudaStream_t stream)
{
    dim3 grid((g.nz + 127) / 128, g.ny, g.nx);
    if (q == 27) moments_kernel<27><<<grid, 128, 0, stream>>>(f, g, moment_mode, rho, ux, uy, uz);
    else         moments_kernel<19><<<grid, 128, 0, stream>>>(f, g, moment_mode, rho, ux, uy, uz);
    return cudaGetLastError();
}

cudaError_t launch_mass_momentum(const float* f, const Geom& g, int q, double* out4, cudaStream_t stream)
{
    const int grid = 148 * 8;      // one resident wave: 8 blocks of 256 threads per SM
    if (q == 27) mass_momentum_kernel<27><<<grid, 256, 0, stream>>>(f, g, out4);
    else         mass_momentum_kernel<19><<<grid, 256, 0, stream>>>(f, g, out4);
    return cudaGetLastError();
}

}  // namespace lbm
