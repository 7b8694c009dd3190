// simt.h -- TEST INFRASTRUCTURE ONLY (tests/emu/).  A small SIMT executor that lets g++ compile the CUDA sources of
// liblbm_b200 (latticeboltzmanncuda_b200/csrc/*.cu, launches rewritten by tests/_emu.py) and run them on the host:
// every CUDA thread of a block is a fiber, a warp's fibers meet at the warp intrinsics (__shfl_*_sync, __ballot_sync,
// __any_sync) and a block's at __syncthreads(), device memory is host memory.  It exists so that the kernels' logic
// and the host logic of the C ABI can be unit-tested against the oracle in the CPU test suite.  Nothing in the
// package loads it; the product is the CUDA library and fails without it (capi.load: "there is no CPU fallback").
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>

#define __DEVICE_LAUNCH_PARAMETERS_H__      // (its `extern const uint3 threadIdx` would clash with the executor's variables)
#undef __shared__
#define __shared__ static          // blocks run one after the other, so one static copy per kernel is the block's copy

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

namespace simt {
enum Wait { NONE = 0, SHFL_XOR, SHFL_IDX, BALLOT, ANY, SYNCTHREADS };
uint64_t warp_exchange(int kind, uint64_t value, int arg);       // suspends the calling thread until its warp has arrived
void block_barrier();
void run_grid(dim3 grid, dim3 block, const std::function<void()>& thread_body);
}  // namespace simt

#define EMU_LAUNCH(grid, block, ...) simt::run_grid(dim3(grid), dim3(block), [&] { __VA_ARGS__; })

template <class T> inline T simt_xchg(int kind, T v, int arg)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof v);
    bits = simt::warp_exchange(kind, bits, arg);
    T out;
    std::memcpy(&out, &bits, sizeof out);
    return out;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return simt_xchg(simt::SHFL_XOR, v, lane_mask); }
template <class T> inline T __shfl_sync(unsigned, T v, int src_lane) { return simt_xchg(simt::SHFL_IDX, v, src_lane); }
inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)simt::warp_exchange(simt::BALLOT, pred ? 1u : 0u, 0); }
inline int __any_sync(unsigned, int pred) { return (int)simt::warp_exchange(simt::ANY, pred ? 1u : 0u, 0); }
inline void __syncthreads() { simt::block_barrier(); }

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline void __threadfence_system() {}
inline void __threadfence() {}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
    return (unsigned)(((((uint64_t)hi << 32) | lo) << (shift & 31u)) >> 32);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
    return (unsigned)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

// one host thread runs every fiber: plain read-modify-write is atomic here
template <class T> inline T simt_atomic_add(T* p, T v) { const T old = *p; *p = old + v; return old; }
inline double atomicAdd(double* p, double v) { return simt_atomic_add(p, v); }
inline float atomicAdd(float* p, float v) { return simt_atomic_add(p, v); }
inline int atomicAdd(int* p, int v) { return simt_atomic_add(p, v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return simt_atomic_add(p, v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return simt_atomic_add(p, v); }
inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned old = *p; *p = old | v; return old; }

template <class T> inline cudaError_t cudaFuncSetAttribute(T*, enum cudaFuncAttribute, int) { return cudaSuccess; }

