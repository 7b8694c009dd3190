// cuda_shim.h -- TEST INFRASTRUCTURE ONLY (tests/emu/): lets g++ compile the single-buffer engine's kernel source
// (latticeboltzmanncuda_b200/csrc/lbm_inplace.cu, launches rewritten by tests/_emu.py) so that its indexing
// logic can be unit-tested in the CPU test suite, one thread after the other.  Nothing in the package loads
// this; the product path is the CUDA library and fails without it.  Only kernels whose threads are independent
// (no shuffles, no shared memory) are meaningful here.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>

#define LBM_HOST_EMULATION 1

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }     // (only in kernels the harness never runs)
inline int __any_sync(unsigned, int pred) { return pred; }
inline void __syncthreads() {}
inline double atomicAdd(double* p, double v) { const double o = *p; *p += v; return o; }

template <class F>
inline void emu_launch(dim3 grid, dim3 block, F&& body)
{
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz)
        for (unsigned ty = 0; ty < block.y; ++ty)
        for (unsigned tx = 0; tx < block.x; ++tx) {
            blockIdx = uint3{bx, by, bz};
            threadIdx = uint3{tx, ty, tz};
            body();
        }
}
#define EMU_LAUNCH(grid, block, ...) emu_launch(dim3(grid), dim3(block), [&] { __VA_ARGS__; })
#define cudaGetLastError() cudaSuccess       // no runtime behind the harness
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
