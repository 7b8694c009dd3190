// simt_runtime.cpp -- TEST INFRASTRUCTURE ONLY: the fiber scheduler behind simt.h and a host-memory stand-in for the
// few CUDA runtime calls liblbm_b200 makes (see simt.h for what this is and is not).
#include "simt.h"

#include <setjmp.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace simt {
namespace {

constexpr size_t STACK_BYTES = 256 * 1024;

// A fiber is created once (makecontext) and reused for every CUDA thread that gets its slot: it runs the current
// kernel body, reports done, and parks.  Switches are _setjmp / _longjmp pairs (no signal-mask system calls).
struct Fiber {
    ucontext_t boot;
    jmp_buf jb;
    char* stack = nullptr;
    bool booted = false;
};

struct Lane {
    uint3 tid;
    bool done = false;
    int wait = NONE;
    uint64_t value = 0, result = 0;
    int arg = 0;
};

jmp_buf g_scheduler;
Lane* g_current = nullptr;
Fiber* g_fiber = nullptr;
const std::function<void()>* g_body = nullptr;
std::vector<Fiber*> g_fibers;

void yield_to_scheduler()
{
    Fiber* me = g_fiber;
    if (_setjmp(me->jb) == 0) _longjmp(g_scheduler, 1);
}

void fiber_main()
{
    for (;;) {
        (*g_body)();
        g_current->done = true;
        yield_to_scheduler();
    }
}

void resume(Lane& lane, size_t index)
{
    while (g_fibers.size() <= index) {
        Fiber* f = new Fiber();
        f->stack = (char*)std::malloc(STACK_BYTES);
        getcontext(&f->boot);
        f->boot.uc_stack.ss_sp = f->stack;
        f->boot.uc_stack.ss_size = STACK_BYTES;
        f->boot.uc_link = nullptr;
        makecontext(&f->boot, fiber_main, 0);
        g_fibers.push_back(f);
    }
    g_current = &lane;
    g_fiber = g_fibers[index];
    threadIdx = lane.tid;
    if (_setjmp(g_scheduler) == 0) {
        if (!g_fiber->booted) {
            g_fiber->booted = true;
            setcontext(&g_fiber->boot);      // first entry: onto the fiber's own stack
            std::abort();
        }
        _longjmp(g_fiber->jb, 1);
    }
}

// every live lane of the warp has arrived at a warp intrinsic: hand out the results
void resolve_warp(std::vector<Lane>& lanes, size_t first, size_t last)
{
    int kind = NONE;
    for (size_t t = first; t < last; ++t) {
        if (lanes[t].done) continue;
        if (kind == NONE) kind = lanes[t].wait;
        if (lanes[t].wait != kind) {
            std::fprintf(stderr, "simt: lanes of one warp wait at different intrinsics (%d / %d)\n", kind, lanes[t].wait);
            std::abort();
        }
    }
    uint64_t ballot = 0;
    for (size_t t = first; t < last; ++t)
        if (!lanes[t].done && lanes[t].value) ballot |= 1ull << (t - first);
    for (size_t t = first; t < last; ++t) {
        Lane& l = lanes[t];
        if (l.done) continue;
        const size_t me = t - first;
        size_t src = me;
        if (kind == SHFL_XOR) src = me ^ (size_t)(l.arg & 31);
        if (kind == SHFL_IDX) src = (size_t)(l.arg & 31);
        switch (kind) {
            case SHFL_XOR:
            case SHFL_IDX:
                l.result = (first + src < last && !lanes[first + src].done) ? lanes[first + src].value : l.value;
                break;
            case BALLOT: l.result = ballot; break;
            case ANY: l.result = ballot != 0; break;
            default: std::abort();
        }
    }
    for (size_t t = first; t < last; ++t)
        if (!lanes[t].done) lanes[t].wait = NONE;
}

void run_block(std::vector<Lane>& lanes)
{
    const size_t n = lanes.size();
    for (;;) {
        bool any_live = false;
        for (size_t first = 0; first < n; first += 32) {
            const size_t last = first + 32 < n ? first + 32 : n;
            for (;;) {
                for (size_t t = first; t < last; ++t)
                    while (!lanes[t].done && lanes[t].wait == NONE) resume(lanes[t], t);
                bool live = false, at_barrier = false;
                for (size_t t = first; t < last; ++t)
                    if (!lanes[t].done) { live = true; at_barrier |= lanes[t].wait == SYNCTHREADS; }
                if (!live || at_barrier) { any_live |= live; break; }
                resolve_warp(lanes, first, last);
            }
        }
        if (!any_live) return;
        for (Lane& l : lanes) {          // every live thread of the block stands at __syncthreads()
            if (l.done) continue;
            if (l.wait != SYNCTHREADS) { std::fprintf(stderr, "simt: barrier reached by part of a warp only\n"); std::abort(); }
            l.wait = NONE;
        }
    }
}

}  // namespace

uint64_t warp_exchange(int kind, uint64_t value, int arg)
{
    Lane* me = g_current;
    me->value = value;
    me->arg = arg;
    me->wait = kind;
    yield_to_scheduler();
    return me->result;
}

void block_barrier()
{
    Lane* me = g_current;
    me->wait = SYNCTHREADS;
    yield_to_scheduler();
}

void run_grid(dim3 grid, dim3 block, const std::function<void()>& thread_body)
{
    const std::function<void()>* outer = g_body;
    g_body = &thread_body;
    gridDim = grid;
    blockDim = block;
    const size_t n = (size_t)block.x * block.y * block.z;
    std::vector<Lane> lanes(n);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                for (size_t t = 0; t < n; ++t) {
                    Lane& l = lanes[t];
                    l.done = false;
                    l.wait = NONE;
                    l.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((size_t)block.x * block.y))};
                }
                run_block(lanes);
            }
    g_body = outer;
}

}  // namespace simt

// ---- the CUDA runtime calls of liblbm_b200, on host memory ------------------------------------------------------

namespace {
struct EmuEvent { std::chrono::steady_clock::time_point t; };
int g_dummy_stream;
std::map<void*, std::pair<void*, size_t>> g_allocations;      // user pointer -> mapping
}

extern "C" {

cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "simt emulation error"; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, enum cudaDeviceAttr, int) { *v = 0; return cudaSuccess; }
cudaError_t cudaDeviceSetLimit(enum cudaLimit, size_t) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, enum cudaFuncAttribute, int) { return cudaSuccess; }

// "Device" allocations sit flush against an inaccessible page (electric-fence style): a kernel that reads or writes
// past the end of a buffer faults at once instead of passing silently.  SIMT_GUARD=start puts the page in front of
// the buffer instead (underruns); SIMT_GUARD=off uses plain malloc.
cudaError_t cudaMalloc(void** p, size_t n)
{
    static const int mode = [] {
        const char* e = std::getenv("SIMT_GUARD");
        return !e ? 1 : (std::strcmp(e, "start") == 0 ? 2 : (std::strcmp(e, "off") == 0 ? 0 : 1));
    }();
    *p = nullptr;
    if (n == 0) n = 1;
    if (mode == 0) return posix_memalign(p, 256, n) == 0 ? cudaSuccess : cudaErrorMemoryAllocation;
    const size_t page = 4096, body = (n + page - 1) / page * page, total = body + page;
    char* base = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == (char*)MAP_FAILED) return cudaErrorMemoryAllocation;
    char* user;
    if (mode == 1) {
        mprotect(base + body, page, PROT_NONE);
        user = base + body - (n + 7) / 8 * 8;
    } else {
        mprotect(base, page, PROT_NONE);
        user = base + page;
    }
    g_allocations[user] = std::make_pair((void*)base, total);
    *p = user;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p)
{
    if (!p) return cudaSuccess;
    auto it = g_allocations.find(p);
    if (it == g_allocations.end()) { std::free(p); return cudaSuccess; }
    munmap(it->second.first, it->second.second);
    g_allocations.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, enum cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, enum cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy3DAsync(const struct cudaMemcpy3DParms* p, cudaStream_t)
{
    const cudaPitchedPtr &s = p->srcPtr, &d = p->dstPtr;
    for (size_t z = 0; z < p->extent.depth; ++z)
        for (size_t y = 0; y < p->extent.height; ++y) {
            const char* src = (const char*)s.ptr + ((p->srcPos.z + z) * s.ysize + (p->srcPos.y + y)) * s.pitch + p->srcPos.x;
            char* dst = (char*)d.ptr + ((p->dstPos.z + z) * d.ysize + (p->dstPos.y + y)) * d.pitch + p->dstPos.x;
            std::memcpy(dst, src, p->extent.width);
        }
    return cudaSuccess;
}

cudaError_t cudaMemcpy3D(const struct cudaMemcpy3DParms* p) { return cudaMemcpy3DAsync(p, nullptr); }
cudaError_t cudaMalloc3D(struct cudaPitchedPtr* out, struct cudaExtent e)
{
    void* p = nullptr;
    const cudaError_t rc = cudaMalloc(&p, e.width * e.height * e.depth);
    *out = make_cudaPitchedPtr(p, e.width, e.width, e.height);
    return rc;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
// slabs in one process share the host's address space; nothing here ever crosses a process
cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes* a, const void*) { std::memset(a, 0, sizeof *a); a->type = cudaMemoryTypeDevice; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void*) { std::memset(h, 0, sizeof *h); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
// (the reference fetches a device function pointer through the address of a __device__ variable; on this executor
// that variable is an ordinary global, initialised with the function's address)
cudaError_t cudaMemcpyFromSymbol(void* dst, const void* symbol, size_t count, size_t offset, enum cudaMemcpyKind)
{
    std::memcpy(dst, (const char*)symbol + offset, count);
    return cudaSuccess;
}
cudaError_t cudaMemcpyToSymbol(const void* symbol, const void* src, size_t count, size_t offset, enum cudaMemcpyKind)
{
    std::memcpy((char*)symbol + offset, src, count);
    return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)&g_dummy_stream; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)&g_dummy_stream; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamSetAttribute(cudaStream_t, cudaStreamAttrID, const cudaStreamAttrValue*) { return cudaSuccess; }

cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new EmuEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (EmuEvent*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ((EmuEvent*)e)->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(((EmuEvent*)b)->t - ((EmuEvent*)a)->t).count();
    return cudaSuccess;
}

}  // extern "C"
