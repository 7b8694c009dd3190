// lbm_inplace.cu -- single-buffer streaming (the "AA" pattern): one copy of the populations instead of two,
// same HBM traffic per step (read Q + write Q), half the footprint (1024^3 D3Q19 = 83 GB fits one B200).
//
// Two kernels alternate; each thread reads and writes exactly the same set of slots, so the update is in
// place without any ordering between threads:
//   even step  reads  A[i][x]              (populations that have arrived at x)
//              writes A[inv(i)][x]         (post-collision f*_i(x), parked in the opposite slot of the same cell)
//   odd step   reads  A[inv(i)][x - c_i]   (= f*_i(x - c_i): the pull, straight from where the even step parked it)
//              writes A[i][x + c_i]        (post-collision f*_i(x), pushed to where the next even step reads it)
// The state a caller sees is always f*_i(x) (what the two-buffer engine holds): `parity` says which of the
// two storage forms the buffer is in, and aa_fetch() translates (lbm_download_f_soa, moments, mass).
//
// Domain edges are resolved by address selection on the edge cells, no ghost traffic at all:
//   periodic: the neighbour on the far side;
//   wall    : half-way bounce-back -- the pull reads the cell's own slot i (= f*_inv(i)(x)), the push parks
//             f*_i(x) in the cell's own slot inv(i); a moving wall adds 6 w (c . u_wall), rho_wall = 1.
// Obstacles (packed solid bitmask, half-way bounce-back) follow the same rule as a wall, cell by cell: a
// population whose upstream neighbour x - c_i is solid is pulled from the cell's own slot i, and a population
// whose downstream neighbour x + c_i is solid is parked in the cell's own slot inv(i) -- again the very slots
// the cell read, so the update stays in place.  Solid cells are never read or written (they hold zeros).  The
// word classes of the two-buffer engine are reused: ordinary words run the lean code, all-solid words cost
// nothing, special words (a cell is solid or sees a solid cell) come from the compact list with their
// neighbourhood words.  In either storage form the population that bounces into direction i at cell x sits in
// A[i][x] and is among the values the cell reads anyway, so the momentum exchange with the obstacle is summed by
// the special-word kernels of both steps at no extra traffic.
// x slabs (LBM_BC_HALO on x): the ghost planes are ordinary neighbours.  Before an odd step they receive the
// neighbouring slabs' edge planes of the populations the edge cells pull (c_x = -1 slots below, c_x = +1 slots
// above); the edge cells push into the same slots, and after the step those planes go back to their owners.  Even
// steps are cell-local and need no exchange at all (latticeboltzmanncuda_b200/slab.py).
// Scope: bounce-back obstacles, additive box source, stream + collide fused, periodic / wall / halo edges.
// Reflection tables, unfused phases and zero-inflow edges run on the two-buffer engine.
#include "lbm_device.h"

// measurement knob (python -m latticeboltzmanncuda_b200.build --variant NAME -DMACRO=VALUE)
#ifndef LBM_AA_ODD_MINB
#define LBM_AA_ODD_MINB 12      // blocks/SM of the D3Q19 fp32 odd interior kernel with obstacles (40 registers, no spills)
#endif
#ifndef LBM_AA_ODD_WARPS
#define LBM_AA_ODD_WARPS 4      // z-words (warps) per block of the odd interior kernel
#endif
#ifndef LBM_AA_SPC_MINB
#define LBM_AA_SPC_MINB 8       // blocks/SM of the D3Q19 odd special-word kernel (64 registers)
#endif
#ifndef LBM_AA_EVEN_ROWS
#define LBM_AA_EVEN_ROWS 1      // rows a thread of the even kernel takes at once (2: both rows' loads in flight before the first collision)
#endif
#ifndef LBM_AA_EVEN_ROWS_MINB
#define LBM_AA_EVEN_ROWS_MINB 8
#endif
#ifndef LBM_AA_EVEN_THREADS
#define LBM_AA_EVEN_THREADS 128 // threads per block of the even kernel along z
#endif

namespace lbm {

namespace {

// the collision operators are shared with lbm_kernels.cu through this header-like include
#include "lbm_collide.inl"

// Moving-wall terms are formed and added with explicitly rounded operations: the same cell goes through different code
// (interior kernel, face kernel, special-word kernel) depending on where a slab boundary falls, and a product that one
// of them contracted into the following add would differ in the last bit from the others.
__device__ __forceinline__ float wall_term(float w6, float sgn, float cu) { return __fmul_rn(__fmul_rn(w6, sgn), cu); }
__device__ __forceinline__ float plus_term(float v, float t) { return __fadd_rn(v, t); }

struct Edge {          // where population i of cell (x,y,z) comes from / goes to
    long long off;     // complete element offset from the cell: population slot + neighbour cell (valid when kind == 0)
    int kind;          // 0 neighbour inside (possibly wrapped, possibly in the neighbouring slab), 1 wall
    float term;        // moving-wall term for the bounced population
};

// neighbour of (x,y,z) in direction c_i under the per-axis edge rules; `sign` = +1 for the push target
// x + c_i (slot i), -1 for the pull source x - c_i (slot inv(i)).  The plain address is the precomputed
// pull_off / push_off; an axis on which the neighbour leaves the lattice adds shift[a][side] (periodic: the far side
// of the array, halo: the neighbouring slab's edge plane) or turns the access into the cell's own slot (wall).
// For a wall the returned term is 6 w_i (c_b . u) where c_b is the direction of the population after the bounce.
// AXES: bit a set = the cell may sit on an edge of axis a (the others are known to be interior).
template <int Q, int AXES = 7>
__device__ __forceinline__ Edge edge_lookup(const InplaceParams& p, int i, int x, int y, int z, int sign)
{
    const Geom& g = p.g;
    const int c[3] = {sign * dir_c(i, 0), sign * dir_c(i, 1), sign * dir_c(i, 2)};
    const int n[3] = {g.nx, g.ny, g.nz};
    const int pos[3] = {x, y, z};
    Edge e;
    e.off = sign < 0 ? p.pull_off[i] : p.push_off[i];
    e.kind = 0; e.term = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!((AXES >> a) & 1) || c[a] == 0) continue;
        const int t = pos[a] + c[a];
        if (t >= 0 && t < n[a]) continue;
        if (p.bc[a] == 2) {
            if (e.kind == 0) {       // the wall named first in x, y, z order supplies the velocity
                const float* u = p.wall_u[a * 2 + (t < 0 ? 0 : 1)];
                // pull (sign -1): the population arriving is i itself; push (sign +1): it comes back as inv(i) = -c_i
                const float s = (sign < 0) ? 1.0f : -1.0f;
                e.term = wall_term(6.0f * (float)dir_weight<Q>(i), s,
                                   (float)dir_c(i, 0) * u[0] + (float)dir_c(i, 1) * u[1] + (float)dir_c(i, 2) * u[2]);
            }
            e.kind = 1;
        } else {
            // periodic wrap / neighbouring slab (x: per slot, the slot being inv(i) for a pull and i for a push)
            if (a == 0) e.off += p.xshift[t < 0 ? 0 : 1][sign < 0 ? dir_inv(i) : i];
            else        e.off += p.shift[a][t < 0 ? 0 : 1];
        }
    }
    return e;
}

// Word (x, y, w) lies in one of the boxes of all-solid words the host found (csrc/lbm_capi.cu find_dead_boxes): the warp
// leaves on a compare of kernel parameters, before any load.  An unused box has extent 0 and matches nothing.
__device__ __forceinline__ bool in_dead_box(const InplaceParams& p, int x, int y, int w)
{
#pragma unroll
    for (int b = 0; b < LBM_DEAD_BOXES; ++b)
        if ((unsigned)(x - p.dead_box[b][0]) < (unsigned)p.dead_box[b][3] && (unsigned)(y - p.dead_box[b][1]) < (unsigned)p.dead_box[b][4] &&
            (unsigned)(w - p.dead_box[b][2]) < (unsigned)p.dead_box[b][5]) return true;
    return false;
}

// Word (x, y, w) lies in the bounding box of the words that are not ordinary (all-solid, next to a solid cell or holding a
// reflection target).  Outside it every cell is plain fluid: the warp asks for its populations at once instead of waiting
// for its class byte or mask word first -- one dependent load less in front of every block upstream, downstream and
// beside an obstacle (measured: profiles/r02_notes.md section 4).
__device__ __forceinline__ bool in_busy_box(const InplaceParams& p, int x, int y, int w)
{
    return (unsigned)(x - p.busy_box[0]) < (unsigned)p.busy_box[3] && (unsigned)(y - p.busy_box[1]) < (unsigned)p.busy_box[4] &&
           (unsigned)(w - p.busy_box[2]) < (unsigned)p.busy_box[5];
}

// solid flag of the padded cell (xpad, ypad, zpad); ghost cells hold the periodic images, wall ghosts are clear
__device__ __forceinline__ bool solid_bit(const InplaceParams& p, int xpad, int ypad, int zpad)
{
    const uint32_t w = __ldg(p.solid + ((long long)xpad * p.g.yp + ypad) * p.g.wz + (zpad >> 5));
    return (w >> (zpad & 31)) & 1u;
}

// additive source on the cells of the box (SimCalcFuncs.cu:200-215), after the collision like the two-buffer engine
template <int Q>
__device__ __forceinline__ void add_source(const InplaceParams& p, float (&f)[Q], int x, int y, int z)
{
    if (source_hit(p, x, y, z)) {
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] += p.src_f[i];
    }
}

enum : unsigned { AA_ORDINARY = 0, AA_DEAD = 1, AA_SPECIAL = 2 };      // word classes, as in lbm_kernels.cu

// EXTRA: the lattice has a solid mask and / or a source (kept out of the all-fluid instantiation)
template <int Q, int MOMENT, bool EQ64, bool EXTRA>
__global__ void __launch_bounds__(LBM_AA_EVEN_THREADS > 256 ? LBM_AA_EVEN_THREADS : 256, ((Q == 19 && !EQ64) ? 6 : 4) * 256 / (LBM_AA_EVEN_THREADS > 256 ? LBM_AA_EVEN_THREADS : 256)) aa_even_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = p.x0 + blockIdx.z;                 // the launch covers the planes [x0, x1)
    if (z >= g.nz || y >= g.ny) return;
    if (EXTRA && p.solid && in_busy_box(p, x, y, z >> 5)) {
        if (in_dead_box(p, x, y, z >> 5)) return;
        // a warp is 32 consecutive z of one row = one 32-cell word (blockDim.x is a multiple of 32)
        if (p.force && !p.force_pending) {
            // the momentum exchange is summed by aa_even_special_kernel, which takes the special words: ordinary words only
            if (__ldg(p.cls + ((long long)x * ((g.nz + 31) >> 5) + (z >> 5)) * p.cls_pitch + y) != AA_ORDINARY) return;
        } else {
            // everything in one launch (measured: 2.35 ms against 2.11 + 0.34 ms for the split): one broadcast load of
            // the mask word, solid lanes leave; going through the L2-resident class byte first measured the same
            if (solid_bit(p, x + 1, y + 1, z + PAD_Z)) return;
        }
    }
    float* a = p.f + g.cell(x, y, z);
    float f[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) f[i] = a[p.pop_off[i]];
    if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
    else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    if (EXTRA && p.src_on) {
        // the coordinates are read again from the special registers (volatile: no common subexpression with the ones above),
        // so that x, y, z do not have to stay in registers across the collision: the kernel sits at its 40-register cap
        unsigned bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z, tx = threadIdx.x, ty = threadIdx.y;
#ifdef __CUDACC__
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
        asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
        asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tx));
        asm volatile("mov.u32 %0, %%tid.y;" : "=r"(ty));
#endif
        add_source<Q>(p, f, p.x0 + (int)bz, (int)(by * blockDim.y + ty), (int)(bx * blockDim.x + tx));
    }
    asm volatile("" : "+l"(a));          // rebuild the Q addresses next to the stores instead of keeping them alive
#pragma unroll
    for (int i = 0; i < Q; ++i) a[p.pop_off[dir_inv(i)]] = f[i];
}

// Experiment (LBM_AA_EVEN_ROWS = 2): a thread takes the same z of ROWS consecutive rows and issues the loads of all of
// them before the first collision -- ROWS x Q requests in flight per thread at fewer resident warps.
template <int Q, int MOMENT, bool EQ64, bool EXTRA, int ROWS>
__global__ void __launch_bounds__(128, LBM_AA_EVEN_ROWS_MINB) aa_even_rows_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * ROWS;
    const int x = p.x0 + blockIdx.z;
    if (z >= g.nz) return;
    float* a = p.f + g.cell(x, y0, z);
    float f[ROWS][Q];
    bool live[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        live[r] = y0 + r < g.ny;
        if (EXTRA && p.solid && live[r]) live[r] = !solid_bit(p, x + 1, y0 + r + 1, z + PAD_Z);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
        if (live[r]) {
#pragma unroll
            for (int i = 0; i < Q; ++i) f[r][i] = a[p.pop_off[i] + r * g.zp];
        }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
        if (live[r]) {
            if (EQ64) collide_fp64<Q, MOMENT>(f[r], p.tau);
            else      collide_fp32<Q, MOMENT>(f[r], p.inv_tau);
            if (EXTRA && p.src_on) add_source<Q>(p, f[r], x, y0 + r, z);
            asm volatile("" : "+l"(a));
#pragma unroll
            for (int i = 0; i < Q; ++i) a[p.park_off[i] + r * g.zp] = f[r][i];
        }
}

// pull + collide + push of one cell that may sit on a domain edge of the axes in AXES
template <int Q, int MOMENT, bool EQ64, int AXES, bool EXTRA>
__device__ __forceinline__ void aa_odd_edge_cell(const InplaceParams& p, float* a, int x, int y, int z)
{
    float f[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const Edge e = edge_lookup<Q, AXES>(p, i, x, y, z, -1);
        f[i] = (e.kind == 0) ? a[e.off] : plus_term(a[p.pop_off[i]], e.term);
    }
    if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
    else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    if (EXTRA && p.src_on) add_source<Q>(p, f, x, y, z);
    asm volatile("" : "+l"(a), "+r"(x), "+r"(y), "+r"(z));      // rebuild the addresses instead of keeping Q of them alive
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const Edge e = edge_lookup<Q, AXES>(p, i, x, y, z, +1);
        if (e.kind == 0) a[e.off] = f[i];
        else             a[p.park_off[i]] = plus_term(f[i], e.term);
    }
}

// pull + collide + push of a cell of the first or last z-word of an interior row: only the z rule applies, and only
// to the populations with c_z != 0 of the one lane that sits on the edge.  Everything else is the plain code
// (constant-bank offsets): the generic edge_lookup costs ~4x the instructions of a plain word, and with one edge
// word in every other block the kernel waited for those warps (ncu: 570 warp instructions per word on average).
template <int Q, int MOMENT, bool EQ64, bool EXTRA>
__device__ __forceinline__ void aa_odd_zedge_cell(const InplaceParams& p, float* a, int x, int y, int z)
{
    const Geom& g = p.g;
    const bool lo = z == 0, hi = z == g.nz - 1;
    const bool wall = p.bc[2] == 2;
    float f[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        long long off = p.pull_off[i];
        if (dir_c(i, 2) != 0) {
            const bool outside = dir_c(i, 2) > 0 ? lo : hi;           // x - c_i lies beyond the z edge
            if (outside) off = wall ? p.pop_off[i] : off + (dir_c(i, 2) > 0 ? g.nz : -g.nz);
        }
        f[i] = a[off];
    }
    if (wall && (lo || hi)) {       // moving-wall terms of the populations that came back from the wall
#pragma unroll
        for (int i = 1; i < Q; ++i) {
            if (dir_c(i, 2) != 0 && (dir_c(i, 2) > 0 ? lo : hi)) {
                const float* u = p.wall_u[dir_c(i, 2) > 0 ? 4 : 5];
                f[i] = plus_term(f[i], wall_term(6.0f * (float)dir_weight<Q>(i), 1.0f,
                                                 (float)dir_c(i, 0) * u[0] + (float)dir_c(i, 1) * u[1] + (float)dir_c(i, 2) * u[2]));
            }
        }
    }
    if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
    else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    if (EXTRA && p.src_on) add_source<Q>(p, f, x, y, z);
    asm volatile("" : "+l"(a), "+r"(z));
    const bool lo2 = z == 0, hi2 = z == g.nz - 1;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        long long off = p.push_off[i];
        float v = f[i];
        if (dir_c(i, 2) != 0) {
            const bool outside = dir_c(i, 2) < 0 ? lo2 : hi2;         // x + c_i lies beyond the z edge
            if (outside) {
                off = wall ? p.park_off[i] : off + (dir_c(i, 2) < 0 ? g.nz : -g.nz);
                if (wall) {
                    const float* u = p.wall_u[dir_c(i, 2) < 0 ? 4 : 5];
                    v = plus_term(v, wall_term(6.0f * (float)dir_weight<Q>(i), -1.0f,
                                               (float)dir_c(i, 0) * u[0] + (float)dir_c(i, 1) * u[1] + (float)dir_c(i, 2) * u[2]));
                }
            }
        }
        a[off] = v;
    }
}

// Three launches cover the planes [x0, x1) of the lattice, each thread touching only its own slots (no ordering
// between them):
//   MODE 0  the interior box, rows 1..ny-2 of the planes 1..nx-2: a block owns 4 consecutive z-words of a row.  Words
//           away from the z edge run plain offsets (lean code); the first and last word of the row take the z rule
//           (periodic wrap / wall) for the populations with c_z != 0 of their first / last lane -- in the same
//           launch, next to the rest of their row, so that the row is still one contiguous run in DRAM;
//   MODE 2  the x faces (planes 0 and nx-1, every row), with the per-direction rules of all three axes: on x slabs
//           these are the cells that pull from and push to the neighbouring slab's memory;
//   MODE 3  the y faces (rows 0 and ny-1) of the planes 1..nx-2, with the y and z rules.
// With geometry (EXTRA and a class array) all launches take the ordinary words only: all-solid words are
// skipped and the special words are left to aa_odd_special_kernel.
template <int Q, int MOMENT, bool EQ64, int MODE, bool EXTRA>
__global__ void __launch_bounds__(MODE == 0 ? 32 * LBM_AA_ODD_WARPS : 128,
                                  (MODE == 0 && Q == 19 && !EQ64) ? (EXTRA ? LBM_AA_ODD_MINB : 12) * 4 / LBM_AA_ODD_WARPS : (MODE == 0 ? 8 * 4 / LBM_AA_ODD_WARPS : 8)) aa_odd_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const int wcount = (g.nz + 31) >> 5;
    // (MODE 0 and 3 are launched with [x0, x1) already clipped to the interior planes 1..nx-2, see launch_t)
    int x, y, w;
    if constexpr (MODE == 0) {
        w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (w >= wcount) return;
        y = 1 + blockIdx.y;
        x = p.x0 + blockIdx.z;
    } else if constexpr (MODE == 2) {
        // face planes of the range: plane 0 if x0 == 0, plane nx-1 if x1 == nx (one plane when nx == 1)
        const long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const long long per = (long long)g.ny * wcount;
        const bool lo = p.x0 == 0, hi = p.x1 == g.nx && g.nx > 1;
        if (k >= per * ((lo ? 1 : 0) + (hi ? 1 : 0))) return;
        const int plane = (int)(k / per);
        const int rem = (int)(k - (long long)plane * per);
        x = (plane == 0 && lo) ? 0 : g.nx - 1;
        y = rem / wcount;
        w = rem - y * wcount;
    } else {
        const long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int y2 = min(g.ny, 2);
        if (k >= (long long)(p.x1 - p.x0) * y2 * wcount) return;
        const int xs = (int)(k / (y2 * wcount));
        const int rem = (int)(k - (long long)xs * (y2 * wcount));
        x = p.x0 + xs;
        y = (rem / wcount) == 0 ? 0 : g.ny - 1;
        w = rem % wcount;
    }
    const int z = w * 32 + (int)(threadIdx.x & 31u);
    if (z >= g.nz) return;
    if (EXTRA && p.cls && (MODE != 0 || in_busy_box(p, x, y, w))) {      // warp-uniform: one class byte per 32-cell word, resident in L2
        if (MODE == 0 && in_dead_box(p, x, y, w)) return;
        if (__ldg(p.cls + ((long long)x * wcount + w) * p.cls_pitch + y) != AA_ORDINARY) return;
    }
    float* a = p.f + g.cell(x, y, z);
    if constexpr (MODE == 0) {
        if (w == 0 || w == wcount - 1) {           // warp-uniform
            aa_odd_zedge_cell<Q, MOMENT, EQ64, EXTRA>(p, a, x, y, z);
            return;
        }
        float f[Q];
        // pull_off[i] = inv(i) * np - (c_i as an element offset), push_off[i] = i * np + (c_i as an element offset),
        // precomputed on the host: every address is the cell plus one constant-bank operand, and no stride or
        // address has to stay in a register across the collision (the push goes to the very slots the pull read;
        // with the strides in registers the compiler kept all Q addresses alive: 78 registers for D3Q19)
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = a[p.pull_off[i]];
        if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
        else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
        if (EXTRA && p.src_on) add_source<Q>(p, f, x, y, z);
        asm volatile("" : "+l"(a));
#pragma unroll
        for (int i = 0; i < Q; ++i) a[p.push_off[i]] = f[i];
    } else if constexpr (MODE == 2) {
        aa_odd_edge_cell<Q, MOMENT, EQ64, 7, EXTRA>(p, a, x, y, z);
    } else {
        aa_odd_edge_cell<Q, MOMENT, EQ64, 6, EXTRA>(p, a, x, y, z);
    }
}

// Momentum handed to the obstacle cells by this step: population i of a live cell whose upstream neighbour is
// solid is f*_inv(i)(x) coming back, the obstacle takes -c_i * 2 * that value (the two-buffer engine's and the
// oracle's sum, oracle/lbm_oracle.c lbo_step_soa).  f[i] holds exactly that value in both steps: the odd step
// pulled it from the cell's own slot i, the even step finds it arrived in slot i.
template <int Q>
__device__ __forceinline__ void bounce_momentum(const float (&f)[Q], uint32_t nb, float& fx, float& fy, float& fz)
{
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        if ((nb >> nb_bit(-dir_c(i, 0), -dir_c(i, 1), -dir_c(i, 2))) & 1u) {
            const float m = f[i] + f[i];
            fx -= (float)dir_c(i, 0) * m;
            fy -= (float)dir_c(i, 1) * m;
            fz -= (float)dir_c(i, 2) * m;
        }
    }
}

// one atomic triple per warp into one of FORCE_SLOTS partial sums (all 32 lanes of the warp arrive here)
__device__ __forceinline__ void warp_force_add(double* sums, uint32_t widx, float fx, float fy, float fz)
{
    double* slot = sums + 3 * ((unsigned)(widx >> 2) % FORCE_SLOTS);
    if (!__any_sync(0xffffffffu, fx != 0.0f || fy != 0.0f || fz != 0.0f)) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, o);
        fy += __shfl_xor_sync(0xffffffffu, fy, o);
        fz += __shfl_xor_sync(0xffffffffu, fz, o);
    }
    if ((threadIdx.x & 31u) == 0) {
        atomicAdd(slot + 0, (double)fx);
        atomicAdd(slot + 1, (double)fy);
        atomicAdd(slot + 2, (double)fz);
    }
}

// pull + collide + push of one live cell of a special word.  nb = the cell's neighbourhood word (bit
// nb_bit(dx,dy,dz) = solid flag of that neighbour, ghost bits = periodic images).  A solid neighbour wins over
// everything else (own slot, no term); then the domain-edge rule of the axes in AXES (wall: own slot + term,
// periodic: wrapped neighbour); AXES = 0 is the word away from every edge, with constant-bank offsets only.
template <int Q, int MOMENT, bool EQ64, int AXES, bool FORCE>
__device__ __forceinline__ void aa_odd_geo_cell(const InplaceParams& p, float* a, int x, int y, int z, uint32_t nb,
                                                float& fx, float& fy, float& fz, float& gx, float& gy, float& gz)
{
    float f[Q];
    f[0] = a[p.pop_off[0]];
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        const bool own = (nb >> nb_bit(-dir_c(i, 0), -dir_c(i, 1), -dir_c(i, 2))) & 1u;      // x - c_i is solid
        if constexpr (AXES == 0) {
            f[i] = a[own ? p.pop_off[i] : p.pull_off[i]];
        } else {
            const Edge e = edge_lookup<Q, AXES>(p, i, x, y, z, -1);
            const bool self = own || e.kind != 0;
            const float v = a[self ? p.pop_off[i] : e.off];
            f[i] = (!own && e.kind != 0) ? plus_term(v, e.term) : v;
        }
    }
    if (FORCE) bounce_momentum<Q>(f, nb, fx, fy, fz);
    if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
    else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    if (p.src_on) add_source<Q>(p, f, x, y, z);
    asm volatile("" : "+l"(a), "+r"(x), "+r"(y), "+r"(z), "+r"(nb));
    a[p.pop_off[0]] = f[0];
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        const bool own = (nb >> nb_bit(dir_c(i, 0), dir_c(i, 1), dir_c(i, 2))) & 1u;         // x + c_i is solid
        if (FORCE && own && p.force_pending) {
            // parked for the next (even) step, where it arrives as population inv(i) with the obstacle upstream: the
            // momentum that step hands over, -c_inv(i) * 2 f = +c_i * 2 f (bounce_momentum), summed here ahead of time
            const float m = f[i] + f[i];
            gx += (float)dir_c(i, 0) * m;
            gy += (float)dir_c(i, 1) * m;
            gz += (float)dir_c(i, 2) * m;
        }
        if constexpr (AXES == 0) {
            a[own ? p.pop_off[dir_inv(i)] : p.push_off[i]] = f[i];
        } else {
            const Edge e = edge_lookup<Q, AXES>(p, i, x, y, z, +1);
            const bool self = own || e.kind != 0;
            a[self ? p.pop_off[dir_inv(i)] : e.off] = (!own && e.kind != 0) ? plus_term(f[i], e.term) : f[i];
        }
    }
}

// odd step of the special words: one warp per entry of the compact list (sorted, so neighbouring warps touch
// neighbouring rows).  Which edge rules a word needs is warp-uniform.  FORCE: the momentum exchange is summed here
// (solid and out-of-range lanes then stay for the warp sum; without it they leave at once -- keeping them cost the
// kernel 0.09 ms of 0.41 in a measurement, hence the two instantiations).
template <int Q, int MOMENT, bool EQ64, bool FORCE>
__global__ void __launch_bounds__(128, Q == 19 ? LBM_AA_SPC_MINB : 6) aa_odd_special_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const unsigned lane = threadIdx.x & 31u;
    const long long kl = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (kl >= p.n_special) return;
    const long long k = p.special_k0 + kl;                    // the launch covers the entries [special_k0, special_k0 + n_special)
    const uint32_t widx = __ldg(p.special_list + k);          // mask-word index = padded cell index / 32
    const uint32_t nb = __ldg(p.cellinfo + k * 32 + lane);
    const int wpad = (int)(widx % (uint32_t)g.wz);
    const uint32_t row = widx / (uint32_t)g.wz;
    const int x = (int)(row / (uint32_t)g.yp) - 1, y = (int)(row % (uint32_t)g.yp) - 1;
    const int w = wpad - PAD_Z / 32;
    const int z = w * 32 + (int)lane;
    const bool live = z < g.nz && !((nb >> nb_bit(0, 0, 0)) & 1u);
    if (!FORCE && !live) return;
    float fx = 0.0f, fy = 0.0f, fz = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
    if (live) {
        float* a = p.f + (long long)widx * 32 + lane;
        const int wcount = (g.nz + 31) >> 5;
        if (x == 0 || x == g.nx - 1 || y == 0 || y == g.ny - 1) aa_odd_geo_cell<Q, MOMENT, EQ64, 7, FORCE>(p, a, x, y, z, nb, fx, fy, fz, gx, gy, gz);
        else if (w == 0 || w == wcount - 1)                     aa_odd_geo_cell<Q, MOMENT, EQ64, 4, FORCE>(p, a, x, y, z, nb, fx, fy, fz, gx, gy, gz);
        else                                                    aa_odd_geo_cell<Q, MOMENT, EQ64, 0, FORCE>(p, a, x, y, z, nb, fx, fy, fz, gx, gy, gz);
    }
    if (FORCE) {
        warp_force_add(p.force, widx, fx, fy, fz);
        if (p.force_pending) warp_force_add(p.force_pending, widx, gx, gy, gz);
    }
}

// even step of the special words when the momentum exchange is summed (otherwise aa_even_kernel takes them too):
// cell-local like aa_even_kernel, solid lanes skipped; the populations that arrived by bounce-back are among the Q
// values read, so the sum costs no extra traffic
template <int Q, int MOMENT, bool EQ64>
__global__ void __launch_bounds__(128, 8) aa_even_special_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const unsigned lane = threadIdx.x & 31u;
    const long long kl = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (kl >= p.n_special) return;
    const long long k = p.special_k0 + kl;
    const uint32_t widx = __ldg(p.special_list + k);
    const uint32_t nb = __ldg(p.cellinfo + k * 32 + lane);
    const int wpad = (int)(widx % (uint32_t)g.wz);
    const uint32_t row = widx / (uint32_t)g.wz;
    const int x = (int)(row / (uint32_t)g.yp) - 1, y = (int)(row % (uint32_t)g.yp) - 1;
    const int z = (wpad - PAD_Z / 32) * 32 + (int)lane;
    const bool live = z < g.nz && !((nb >> nb_bit(0, 0, 0)) & 1u);
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    if (live) {
        float* a = p.f + (long long)widx * 32 + lane;
        float f[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = a[p.pop_off[i]];
        bounce_momentum<Q>(f, nb, fx, fy, fz);
        if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
        else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
        if (p.src_on) add_source<Q>(p, f, x, y, z);
        asm volatile("" : "+l"(a));
#pragma unroll
        for (int i = 0; i < Q; ++i) a[p.park_off[i]] = f[i];
    }
    warp_force_add(p.force, widx, fx, fy, fz);
}

// x slabs, after an odd step.  The neighbouring slab pushed populations across the shared face into its ghost copy
// of my edge plane; slab.py delivers those planes into the UNUSED slots of my ghost planes (the low ghost plane
// only ever holds the c_x = -1 slots, so its c_x = +1 slots are free, and vice versa).  This kernel moves them
// into the edge plane -- but only the entries the neighbour really wrote: where the pushing cell would lie beyond
// a y/z wall or is solid, the slot holds my own cell's bounced population (written by my own edge cell in the same
// step) and the neighbour's copy of it is stale.
template <int Q>
__global__ void aa_merge_edge_kernel(const __grid_constant__ InplaceParams p, int side)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (z >= g.nz) return;
    const int x = side ? g.nx - 1 : 0;                 // my edge plane
    const int xg = side ? g.nx : -1;                   // the ghost plane next to it
    if (p.solid && solid_bit(p, x + 1, y + 1, z + PAD_Z)) return;
#pragma unroll
    for (int k = 1; k < Q; ++k) {
        if (dir_c(k, 0) != (side ? -1 : 1)) continue;  // populations that ENTER my slab through this face
        const int ys = y - dir_c(k, 1), zs = z - dir_c(k, 2);      // the cell that pushed it sits at (xg, ys, zs)
        if (p.bc[1] == 2 && (ys < 0 || ys >= g.ny)) continue;
        if (p.bc[2] == 2 && (zs < 0 || zs >= g.nz)) continue;
        if (p.solid && solid_bit(p, xg + 1, ys + 1, zs + PAD_Z)) continue;      // (ghost rows / columns hold the periodic images)
        const long long slot = (long long)k * g.np;
        p.f[slot + g.cell(x, y, z)] = p.f[slot + g.cell(xg, y, z)];
    }
}

// zero the populations of the solid cells (once after the mask or the state was set: from then on nothing
// reads or writes them, and the state reads back as the two-buffer engine's, which writes zeros there)
template <int Q>
__global__ void aa_zero_solid_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, x = blockIdx.z;
    if (z >= g.nz) return;
    if (!solid_bit(p, x + 1, y + 1, z + PAD_Z)) return;
    float* a = p.f + g.cell(x, y, z);
#pragma unroll
    for (int i = 0; i < Q; ++i) a[p.pop_off[i]] = 0.0f;
}

// f*_i(x) from either storage form
template <int Q>
__device__ __forceinline__ float aa_fetch(const InplaceParams& p, int i, int x, int y, int z)
{
    const Geom& g = p.g;
    const float* a = p.f + g.cell(x, y, z);
    if (p.parity == 0) return a[(long long)dir_inv(i) * g.np];
    if (p.solid) {
        // solid cells are never touched (parked form, zeros); a population that bounced off a solid neighbour
        // is parked in the cell's own opposite slot
        if (solid_bit(p, x + 1, y + 1, z + PAD_Z)) return a[(long long)dir_inv(i) * g.np];
        if (solid_bit(p, x + 1 + dir_c(i, 0), y + 1 + dir_c(i, 1), z + PAD_Z + dir_c(i, 2))) return a[(long long)dir_inv(i) * g.np];
    }
    const Edge e = edge_lookup<Q>(p, i, x, y, z, +1);
    // a population that bounced off a wall is parked in the opposite slot with the wall term already added
    return (e.kind == 0) ? a[e.off] : plus_term(a[(long long)dir_inv(i) * g.np], -e.term);
}

// The Q populations of a cell for the diagnostics.  Warp-uniform short cuts: in the parked form every population is in the
// cell's own slots; in the arrived form a word of ordinary cells away from every edge finds population i at the plain push
// offset.  Everything else (cells next to obstacles or edges) goes through aa_fetch.
template <int Q>
__device__ __forceinline__ void aa_fetch_cell(const InplaceParams& p, int x, int y, int z, float (&f)[Q])
{
    const Geom& g = p.g;
    const int w = z >> 5, wcount = (g.nz + 31) >> 5;
    const float* a = p.f + g.cell(x, y, z);
    if (p.parity == 0) {
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = a[p.park_off[i]];
    } else if (x > 0 && x < g.nx - 1 && y > 0 && y < g.ny - 1 && w > 0 && w < wcount - 1 &&
               (!p.cls || __ldg(p.cls + ((long long)x * wcount + w) * p.cls_pitch + y) == AA_ORDINARY)) {
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = a[p.push_off[i]];
    } else {
#pragma unroll
        for (int i = 0; i < Q; ++i) f[i] = aa_fetch<Q>(p, i, x, y, z);
    }
}

template <int Q>
__global__ void aa_gather_kernel(const __grid_constant__ InplaceParams p, float* __restrict__ dense, int q, int x0, int nxc)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, xl = blockIdx.z;
    if (z >= g.nz) return;
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < Q; ++i)
        if (i == q) v = aa_fetch<Q>(p, i, x0 + xl, y, z);
    dense[((long long)xl * g.ny + y) * g.nz + z] = v;
}

__global__ void aa_scatter_kernel(const float* __restrict__ dense, float* __restrict__ fq_inv, Geom g, int x0, int nxc)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, xl = blockIdx.z;
    if (z >= g.nz) return;
    fq_inv[g.cell(x0 + xl, y, z)] = dense[((long long)xl * g.ny + y) * g.nz + z];
}

template <int Q>
__global__ void aa_moments_kernel(const __grid_constant__ InplaceParams p, int moment_mode, float* rho, float* ux, float* uy, float* uz)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, x = blockIdx.z;
    if (z >= g.nz) return;
    float r = 0.0f, mx = 0.0f, my = 0.0f, mz = 0.0f;
    float f[Q];
    aa_fetch_cell<Q>(p, x, y, z, f);
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const float v = f[i];
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

template <int Q>
__global__ void __launch_bounds__(256) aa_mass_kernel(const __grid_constant__ InplaceParams p, double* out4)
{
    const Geom& g = p.g;
    const long long rows = (long long)g.nx * g.ny;
    const int warps_per_block = blockDim.x >> 5;
    const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    double m = 0.0, px = 0.0, py = 0.0, pz = 0.0;
    for (long long r = (long long)blockIdx.x * warps_per_block + wib; r < rows; r += (long long)gridDim.x * warps_per_block) {
        const int x = (int)(r / g.ny), y = (int)(r % g.ny);
        for (int z = (int)lane; z < g.nz; z += 32) {
            float f[Q];
            aa_fetch_cell<Q>(p, x, y, z, f);
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                const double v = (double)f[i];
                m += v; px += v * dir_c(i, 0); py += v * dir_c(i, 1); pz += v * dir_c(i, 2);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m += __shfl_xor_sync(0xffffffffu, m, o);
        px += __shfl_xor_sync(0xffffffffu, px, o);
        py += __shfl_xor_sync(0xffffffffu, py, o);
        pz += __shfl_xor_sync(0xffffffffu, pz, o);
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

// one step of the planes [p.x0, p.x1) (and of the special-list entries [p.special_k0, p.special_k0 + p.n_special),
// which the caller sets to the entries of those planes)
template <int Q, int MOMENT, bool EQ64, bool EXTRA>
cudaError_t launch_t(const InplaceParams& p, cudaStream_t stream, cudaStream_t side, int* launches)
{
    const Geom& g = p.g;
    int n = 0;
    if (p.x1 <= p.x0) return cudaSuccess;
    if (!side) side = stream;
    const unsigned special_blocks = (unsigned)((p.n_special + 3) / 4);
    if (p.parity == 0) {        // parked form -> pull / push
        const int wcount = (g.nz + 31) / 32;
        const int xi0 = p.x0 > 1 ? p.x0 : 1, xi1 = p.x1 < g.nx - 1 ? p.x1 : g.nx - 1;
        if (xi1 > xi0) {
            InplaceParams pi = p;          // the interior planes of the range
            pi.x0 = xi0; pi.x1 = xi1;
            if (g.ny > 2) {
                dim3 grid((wcount + LBM_AA_ODD_WARPS - 1) / LBM_AA_ODD_WARPS, g.ny - 2, xi1 - xi0);
                aa_odd_kernel<Q, MOMENT, EQ64, 0, EXTRA><<<grid, 32 * LBM_AA_ODD_WARPS, 0, stream>>>(pi);
                ++n;
            }
            const long long warps = (long long)(xi1 - xi0) * (g.ny < 2 ? g.ny : 2) * wcount;
            aa_odd_kernel<Q, MOMENT, EQ64, 3, EXTRA><<<(unsigned)((warps + 3) / 4), 128, 0, side>>>(pi);
            ++n;
        }
        const int faces = (p.x0 == 0 ? 1 : 0) + ((p.x1 == g.nx && g.nx > 1) ? 1 : 0);
        if (faces) {
            const long long warps = (long long)faces * g.ny * wcount;
            aa_odd_kernel<Q, MOMENT, EQ64, 2, EXTRA><<<(unsigned)((warps + 3) / 4), 128, 0, side>>>(p);
            ++n;
        }
        if (EXTRA && p.cls && p.n_special > 0) {
            if (p.force) aa_odd_special_kernel<Q, MOMENT, EQ64, true><<<special_blocks, 128, 0, side>>>(p);
            else         aa_odd_special_kernel<Q, MOMENT, EQ64, false><<<special_blocks, 128, 0, side>>>(p);
            ++n;
        }
    } else {                    // arrived form -> local
        int bz = ((g.nz + 31) / 32) * 32;
        if (bz > LBM_AA_EVEN_THREADS) bz = LBM_AA_EVEN_THREADS;
        int by = bz >= 128 ? 1 : 128 / bz;
        if (by > g.ny) by = g.ny;
        dim3 block(bz, by, 1), grid((g.nz + bz - 1) / bz, (g.ny + by - 1) / by, p.x1 - p.x0);
        if (LBM_AA_EVEN_ROWS > 1 && bz == 128 && !(p.force && !p.force_pending)) {
            dim3 grid2((g.nz + 127) / 128, (g.ny + LBM_AA_EVEN_ROWS - 1) / LBM_AA_EVEN_ROWS, p.x1 - p.x0);
            aa_even_rows_kernel<Q, MOMENT, EQ64, EXTRA, LBM_AA_EVEN_ROWS><<<grid2, 128, 0, stream>>>(p);
            ++n;
        } else {
            aa_even_kernel<Q, MOMENT, EQ64, EXTRA><<<grid, block, 0, stream>>>(p);
            ++n;
        }
        if (EXTRA && p.cls && p.force && !p.force_pending && p.n_special > 0) {
            aa_even_special_kernel<Q, MOMENT, EQ64><<<special_blocks, 128, 0, side>>>(p);
            ++n;
        }
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

template <int Q>
cudaError_t launch_q(const InplaceParams& p, int moment, int eq64, cudaStream_t s, cudaStream_t s2, int* n)
{
    // the all-fluid, source-free lattice keeps its own instantiation (no class or mask look-ups at all)
    const bool extra = p.solid != nullptr || p.src_on;
    if (extra) {
        if (moment == 0) return eq64 ? launch_t<Q, 0, true, true>(p, s, s2, n) : launch_t<Q, 0, false, true>(p, s, s2, n);
        return eq64 ? launch_t<Q, 1, true, true>(p, s, s2, n) : launch_t<Q, 1, false, true>(p, s, s2, n);
    }
    if (moment == 0) return eq64 ? launch_t<Q, 0, true, false>(p, s, s2, n) : launch_t<Q, 0, false, false>(p, s, s2, n);
    return eq64 ? launch_t<Q, 1, true, false>(p, s, s2, n) : launch_t<Q, 1, false, false>(p, s, s2, n);
}

}  // namespace

cudaError_t launch_inplace_step(const InplaceParams& p, int q, int moment, int eq64, cudaStream_t stream, cudaStream_t side, int* launches)
{
    return q == 27 ? launch_q<27>(p, moment, eq64, stream, side, launches) : launch_q<19>(p, moment, eq64, stream, side, launches);
}

namespace {
// before an odd step: what the previous odd step summed ahead of time has been earned by the even step in between
__global__ void fold_pending_kernel(double* force, double* pending)
{
    const int k = threadIdx.x;
    if (k < 3 * FORCE_SLOTS) {
        force[k] += pending[k];
        pending[k] = 0.0;
    }
}
}  // namespace

cudaError_t launch_inplace_fold_pending(double* force, double* pending, cudaStream_t stream)
{
    fold_pending_kernel<<<1, 3 * FORCE_SLOTS, 0, stream>>>(force, pending);
    return cudaGetLastError();
}

cudaError_t launch_inplace_zero_solid(const InplaceParams& p, int q, cudaStream_t stream)
{
    if (!p.solid) return cudaSuccess;
    dim3 grid((p.g.nz + 127) / 128, p.g.ny, p.g.nx);
    if (q == 27) aa_zero_solid_kernel<27><<<grid, 128, 0, stream>>>(p);
    else         aa_zero_solid_kernel<19><<<grid, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_inplace_merge_edges(const InplaceParams& p, int q, cudaStream_t stream)
{
    dim3 grid((p.g.nz + 127) / 128, p.g.ny, 1);
    for (int side = 0; side < 2; ++side) {
        if (q == 27) aa_merge_edge_kernel<27><<<grid, 128, 0, stream>>>(p, side);
        else         aa_merge_edge_kernel<19><<<grid, 128, 0, stream>>>(p, side);
    }
    return cudaGetLastError();
}

cudaError_t launch_inplace_scatter(const float* dense, float* f, const Geom& g, int q_index, int x0, int nxc, cudaStream_t stream)
{
    if (nxc <= 0) return cudaSuccess;
    // the loaded state is f*_i(x): park it like an even step would, in the opposite slot of the same cell
    dim3 grid((g.nz + 127) / 128, g.ny, nxc);
    aa_scatter_kernel<<<grid, 128, 0, stream>>>(dense, f + (long long)dir_inv(q_index) * g.np, g, x0, nxc);
    return cudaGetLastError();
}

cudaError_t launch_inplace_gather(const InplaceParams& p, int q, float* dense, int q_index, int x0, int nxc, cudaStream_t stream)
{
    if (nxc <= 0) return cudaSuccess;
    dim3 grid((p.g.nz + 127) / 128, p.g.ny, nxc);
    if (q == 27) aa_gather_kernel<27><<<grid, 128, 0, stream>>>(p, dense, q_index, x0, nxc);
    else         aa_gather_kernel<19><<<grid, 128, 0, stream>>>(p, dense, q_index, x0, nxc);
    return cudaGetLastError();
}

cudaError_t launch_inplace_moments(const InplaceParams& p, int q, int moment_mode, float* rho, float* ux, float* uy, float* uz, cudaStream_t stream)
{
    dim3 grid((p.g.nz + 127) / 128, p.g.ny, p.g.nx);
    if (q == 27) aa_moments_kernel<27><<<grid, 128, 0, stream>>>(p, moment_mode, rho, ux, uy, uz);
    else         aa_moments_kernel<19><<<grid, 128, 0, stream>>>(p, moment_mode, rho, ux, uy, uz);
    return cudaGetLastError();
}

cudaError_t launch_inplace_mass(const InplaceParams& p, int q, double* out4, cudaStream_t stream)
{
    if (q == 27) aa_mass_kernel<27><<<148 * 8, 256, 0, stream>>>(p, out4);
    else         aa_mass_kernel<19><<<148 * 8, 256, 0, stream>>>(p, out4);
    return cudaGetLastError();
}

}  // namespace lbm
