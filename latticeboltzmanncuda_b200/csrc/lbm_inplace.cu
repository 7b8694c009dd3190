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
// Scope: all-fluid lattices (no solid mask, no reflection tables, no source), stream + collide fused, periodic
// or wall edges.  Everything else runs on the two-buffer engine.
#include "lbm_device.h"

namespace lbm {

namespace {

// the collision operators are shared with lbm_kernels.cu through this header-like include
#include "lbm_collide.inl"

struct Edge {          // where population i of cell (x,y,z) comes from / goes to
    long long off;     // element offset of the neighbour cell relative to the cell (valid when kind == 0)
    int kind;          // 0 neighbour inside (possibly wrapped), 1 wall
    float term;        // moving-wall term for the bounced population
};

// neighbour of (x,y,z) in direction (dx,dy,dz) under the per-axis edge rules; `sign` = +1 for the push target
// x + c_i, -1 for the pull source x - c_i.  For a wall the returned term is 6 w_i (c_b . u) where c_b is the
// direction of the population after the bounce.
// AXES: bit a set = the cell may sit on an edge of axis a (the others are known to be interior).
template <int Q, int AXES = 7>
__device__ __forceinline__ Edge edge_lookup(const InplaceParams& p, int i, int x, int y, int z, int sign)
{
    const Geom& g = p.g;
    const int c[3] = {sign * dir_c(i, 0), sign * dir_c(i, 1), sign * dir_c(i, 2)};
    const int n[3] = {g.nx, g.ny, g.nz};
    const int pos[3] = {x, y, z};
    const long long stride[3] = {g.plane, (long long)g.zp, 1};
    Edge e;
    e.off = 0; e.kind = 0; e.term = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int t = pos[a] + c[a];
        if (!((AXES >> a) & 1) || (t >= 0 && t < n[a])) { e.off += (long long)c[a] * stride[a]; continue; }
        if (p.bc[a] == 2) {
            if (e.kind == 0) {       // the wall named first in x, y, z order supplies the velocity
                const float* u = p.wall_u[a * 2 + (t < 0 ? 0 : 1)];
                // pull (sign -1): the population arriving is i itself; push (sign +1): it comes back as inv(i) = -c_i
                const float s = (sign < 0) ? 1.0f : -1.0f;
                e.term = 6.0f * (float)dir_weight<Q>(i) * s *
                         ((float)dir_c(i, 0) * u[0] + (float)dir_c(i, 1) * u[1] + (float)dir_c(i, 2) * u[2]);
            }
            e.kind = 1;
        } else {
            e.off += (long long)(c[a] - (t < 0 ? -n[a] : n[a])) * stride[a];     // periodic wrap
        }
    }
    return e;
}

template <int Q, int MOMENT, bool EQ64>
__global__ void __launch_bounds__(256, (Q == 19 && !EQ64) ? 6 : 4) aa_even_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (z >= g.nz || y >= g.ny) return;
    float* a = p.f + g.cell(x, y, z);
    float f[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) f[i] = a[p.pop_off[i]];
    if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
    else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    asm volatile("" : "+l"(a));          // rebuild the Q addresses next to the stores instead of keeping them alive
#pragma unroll
    for (int i = 0; i < Q; ++i) a[p.pop_off[dir_inv(i)]] = f[i];
}

// pull + collide + push of one cell that may sit on a domain edge of the axes in AXES
template <int Q, int MOMENT, bool EQ64, int AXES>
__device__ __forceinline__ void aa_odd_edge_cell(const InplaceParams& p, float* a, int x, int y, int z)
{
    const Geom& g = p.g;
    float f[Q];
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const Edge e = edge_lookup<Q, AXES>(p, i, x, y, z, -1);
        f[i] = (e.kind == 0) ? a[(long long)dir_inv(i) * g.np + e.off] : a[(long long)i * g.np] + e.term;
    }
    if (EQ64) collide_fp64<Q, MOMENT>(f, p.tau);
    else      collide_fp32<Q, MOMENT>(f, p.inv_tau);
    asm volatile("" : "+l"(a), "+r"(x), "+r"(y), "+r"(z));      // rebuild the addresses instead of keeping Q of them alive
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const Edge e = edge_lookup<Q, AXES>(p, i, x, y, z, +1);
        if (e.kind == 0) a[(long long)i * g.np + e.off] = f[i];
        else             a[(long long)dir_inv(i) * g.np] = f[i] + e.term;
    }
}

// Two launches cover the lattice, each thread touching only its own slots (no ordering between them):
//   MODE 0  the interior box, rows 1..ny-2 of planes 1..nx-2: a block owns 4 consecutive z-words of a row.  Words
//           away from the z edge run plain offsets (lean code); the first and last word of the row take the z rule
//           (periodic wrap / wall) for the populations with c_z != 0 of their first / last lane -- in the same
//           launch, next to the rest of their row, so that the row is still one contiguous run in DRAM;
//   MODE 2  the x and y faces, with the per-direction rules of all three axes.
template <int Q, int MOMENT, bool EQ64, int MODE>
__global__ void __launch_bounds__(128, (MODE == 0 && Q == 19 && !EQ64) ? 12 : 8) aa_odd_kernel(const __grid_constant__ InplaceParams p)
{
    const Geom& g = p.g;
    const int wcount = (g.nz + 31) >> 5;
    int x, y, w;
    if constexpr (MODE == 0) {
        w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (w >= wcount) return;
        y = 1 + blockIdx.y;
        x = 1 + blockIdx.z;
    } else {
        long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int x2 = min(g.nx, 2), y2 = min(g.ny, 2);
        const int xi = g.nx - x2;
        const long long na = (long long)x2 * g.ny * wcount, nb = (long long)xi * y2 * wcount;
        if (k < na) {
            const int plane = (int)(k / ((long long)g.ny * wcount));
            const int rem = (int)(k - (long long)plane * g.ny * wcount);
            x = plane == 0 ? 0 : g.nx - 1;
            y = rem / wcount;
            w = rem - y * wcount;
        } else if (k < na + nb) {
            k -= na;
            const int xs = (int)(k / (y2 * wcount));
            const int rem = (int)(k - (long long)xs * (y2 * wcount));
            x = 1 + xs;
            y = (rem / wcount) == 0 ? 0 : g.ny - 1;
            w = rem % wcount;
        } else {
            return;
        }
    }
    const int z = w * 32 + (int)(threadIdx.x & 31u);
    if (z >= g.nz) return;
    float* a = p.f + g.cell(x, y, z);
    if constexpr (MODE == 0) {
        if (w == 0 || w == wcount - 1) {           // warp-uniform
            aa_odd_edge_cell<Q, MOMENT, EQ64, 4>(p, a, x, y, z);
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
        asm volatile("" : "+l"(a));
#pragma unroll
        for (int i = 0; i < Q; ++i) a[p.push_off[i]] = f[i];
    } else {
        aa_odd_edge_cell<Q, MOMENT, EQ64, 7>(p, a, x, y, z);
    }
}

// f*_i(x) from either storage form
template <int Q>
__device__ __forceinline__ float aa_fetch(const InplaceParams& p, int i, int x, int y, int z)
{
    const Geom& g = p.g;
    const float* a = p.f + g.cell(x, y, z);
    if (p.parity == 0) return a[(long long)dir_inv(i) * g.np];
    const Edge e = edge_lookup<Q>(p, i, x, y, z, +1);
    // a population that bounced off a wall is parked in the opposite slot with the wall term already added
    return (e.kind == 0) ? a[(long long)i * g.np + e.off] : a[(long long)dir_inv(i) * g.np] - e.term;
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
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const float v = aa_fetch<Q>(p, i, x, y, z);
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
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                const double v = (double)aa_fetch<Q>(p, i, x, y, z);
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

template <int Q, int MOMENT, bool EQ64>
cudaError_t launch_t(const InplaceParams& p, cudaStream_t stream, int* launches)
{
    const Geom& g = p.g;
    int n = 0;
    if (p.parity == 0) {        // parked form -> pull / push
        const int wcount = (g.nz + 31) / 32;
        if (g.nx > 2 && g.ny > 2) {
            dim3 grid((wcount + 3) / 4, g.ny - 2, g.nx - 2);
            aa_odd_kernel<Q, MOMENT, EQ64, 0><<<grid, 128, 0, stream>>>(p);
            ++n;
        }
        const long long x2 = g.nx < 2 ? g.nx : 2, y2 = g.ny < 2 ? g.ny : 2;
        const long long xi = g.nx - x2;
        // when there is no interior box (nx or ny <= 2) the faces are the whole lattice
        const long long warps = x2 * g.ny * wcount + xi * y2 * wcount;
        aa_odd_kernel<Q, MOMENT, EQ64, 2><<<(unsigned)((warps + 3) / 4), 128, 0, stream>>>(p);
        ++n;
    } else {                    // arrived form -> local
        int bz = ((g.nz + 31) / 32) * 32;
        if (bz > 128) bz = 128;
        int by = 128 / bz;
        if (by > g.ny) by = g.ny;
        dim3 block(bz, by, 1), grid((g.nz + bz - 1) / bz, (g.ny + by - 1) / by, g.nx);
        aa_even_kernel<Q, MOMENT, EQ64><<<grid, block, 0, stream>>>(p);
        ++n;
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

template <int Q>
cudaError_t launch_q(const InplaceParams& p, int moment, int eq64, cudaStream_t s, int* n)
{
    if (moment == 0) return eq64 ? launch_t<Q, 0, true>(p, s, n) : launch_t<Q, 0, false>(p, s, n);
    return eq64 ? launch_t<Q, 1, true>(p, s, n) : launch_t<Q, 1, false>(p, s, n);
}

}  // namespace

cudaError_t launch_inplace_step(const InplaceParams& p, int q, int moment, int eq64, cudaStream_t stream, int* launches)
{
    return q == 27 ? launch_q<27>(p, moment, eq64, stream, launches) : launch_q<19>(p, moment, eq64, stream, launches);
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
