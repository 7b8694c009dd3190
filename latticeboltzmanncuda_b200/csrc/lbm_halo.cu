// lbm_halo.cu -- x slabs across the GPUs of one box with the halo exchange BELOW the C ABI: no ghost-plane copies and
// no communication library in the step.  A slab's edge cells read (two-buffer engine) or read and write (single-buffer
// engine) the neighbouring slab's edge plane where it lies, in the neighbour's own allocation, through peer memory over
// NVLink (same process: cudaDeviceEnablePeerAccess; one process per GPU: CUDA IPC handles exchanged by the caller).
// The reference is single-GPU; the requirement is "the same answer as one GPU".
//
// Ordering between slabs is one 64-bit epoch word per face, written with a peer store and polled locally:
//   signal(t): "everything I did in step t - 1 is complete; my edge planes are in the state step t reads"
//   wait(t)  : both neighbours have signalled t
// Each step of a slab is  [main] ghost fills -> [edge] signal(t), wait(t), edge planes  ||  [main] interior planes ->
// join.  The edge planes (1 or 2 of the nx planes) run on a high-priority stream next to the interior launch, so the
// NVLink latency of their pulls and the handshake hide behind the interior; a neighbour's signal normally arrived a
// whole step earlier.  A wait that sees no signal for LBM_HALO_TIMEOUT_MS (default 20000) gives up and counts the
// time-out in the flag block -- lbm_sync then reports LBM_ERR_STATE -- so a lost neighbour cannot hang the GPU.
#include "../../include/lbm_b200.h"
#include "lbm_device.h"
#include "lbm_internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

using namespace lbm;

namespace {

#define HCK(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char buf_[384];                                                                            \
            snprintf(buf_, sizeof buf_, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return lbm_internal_fail(e_ == cudaErrorMemoryAllocation ? LBM_ERR_NOMEM : LBM_ERR_CUDA, buf_);   \
        }                                                                                              \
    } while (0)

constexpr uint32_t BLOB_MAGIC = 0x4c424d48u;      // "LBMH"
constexpr size_t FLAG_BYTES = size_t(2) << 20;    // a whole 2 MiB block, so that its IPC handle maps exactly this allocation

struct HaloBlob {
    uint32_t magic, version;
    int32_t pid, device;
    int32_t nx, ny, nz, q;
    int32_t inplace, cur, aa_parity, zp;
    uint64_t handle;                  // the exporting handle (same-process shortcut)
    uint64_t buf[2], flags;           // device pointers, valid in the exporting process
    cudaIpcMemHandle_t ipc_buf[2], ipc_flags;
};
static_assert(sizeof(HaloBlob) <= LBM_HALO_BLOB_BYTES, "blob size");

__device__ __forceinline__ unsigned long long now_ns()
{
#if defined(__CUDACC__)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#else
    return 0;
#endif
}

// one thread: my epoch into the flag words the neighbours poll
__global__ void halo_signal_kernel(unsigned long long* lo_word, unsigned long long* hi_word, unsigned long long epoch)
{
    __threadfence_system();           // whatever earlier kernels of this GPU wrote (also into peer memory) is visible first
    if (lo_word) *(volatile unsigned long long*)lo_word = epoch;
    if (hi_word) *(volatile unsigned long long*)hi_word = epoch;
    __threadfence_system();
}

// one thread: spin until both neighbours have reached `epoch`
__global__ void halo_wait_kernel(unsigned long long* flags, int wait_lo, int wait_hi, unsigned long long epoch,
                                 unsigned long long timeout_ns)
{
    volatile unsigned long long* f = flags;
    const unsigned long long t0 = now_ns();
    while ((wait_lo && f[0] < epoch) || (wait_hi && f[1] < epoch)) {
#if defined(__CUDACC__)
        __nanosleep(200);
        if (now_ns() - t0 > timeout_ns) {
            atomicAdd(flags + 2, 1ull);
            break;
        }
#else
        flags[2] += 1;                // the host executor runs kernels one after the other: a missing signal never arrives
        break;
#endif
    }
    __threadfence_system();
}

unsigned long long timeout_ns()
{
    static const unsigned long long v = [] {
        const char* e = getenv("LBM_HALO_TIMEOUT_MS");
        const long long ms = e ? atoll(e) : 20000;
        return (unsigned long long)(ms > 0 ? ms : 20000) * 1000000ull;
    }();
    return v;
}

int ensure_streams(lbm_handle h)
{
    lbm_halo_state& s = h->halo;
    if (!s.edge) {
        int lo_prio = 0, hi_prio = 0;
        HCK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        HCK(cudaStreamCreateWithPriority(&s.edge, cudaStreamNonBlocking, hi_prio));
        HCK(cudaEventCreateWithFlags(&s.ev_pre, cudaEventDisableTiming));
        HCK(cudaEventCreateWithFlags(&s.ev_edge, cudaEventDisableTiming));
    }
    return LBM_OK;
}

void close_side(lbm_handle h, int side, bool shared_with_other_side)
{
    lbm_halo_state& s = h->halo;
    if (s.has[side] && s.ipc[side] && !shared_with_other_side) {
        for (int b = 0; b < 2; ++b)
            if (s.peer_buf[side][b]) cudaIpcCloseMemHandle(s.peer_buf[side][b]);
        if (s.peer_flag[side]) cudaIpcCloseMemHandle((void*)(s.peer_flag[side] - (side == 0 ? 1 : 0)));
    }
    s.has[side] = s.ipc[side] = false;
    s.peer_buf[side][0] = s.peer_buf[side][1] = nullptr;
    s.peer_flag[side] = nullptr;
    s.delta_inplace[side] = 0;
    s.np_diff[side] = 0;
}

}  // namespace

int lbm_internal_halo_signal(lbm_handle h)
{
    lbm_halo_state& s = h->halo;
    halo_signal_kernel<<<1, 1, 0, s.edge>>>(s.peer_flag[0], s.peer_flag[1], s.epoch + 1);
    HCK(cudaGetLastError());
    h->launches += 1;
    return LBM_OK;
}

int lbm_internal_halo_wait(lbm_handle h)
{
    lbm_halo_state& s = h->halo;
    s.epoch += 1;
    halo_wait_kernel<<<1, 1, 0, s.edge>>>(s.flags, s.has[0] ? 1 : 0, s.has[1] ? 1 : 0, s.epoch, timeout_ns());
    HCK(cudaGetLastError());
    h->launches += 1;
    return LBM_OK;
}

void lbm_internal_halo_release(lbm_handle h)
{
    lbm_halo_state& s = h->halo;
    const bool same = s.has[0] && s.has[1] && s.ipc[0] && s.ipc[1] && s.peer_buf[0][0] == s.peer_buf[1][0];
    close_side(h, 0, same);
    close_side(h, 1, false);
    s.connected = false;
    if (s.edge) { cudaStreamSynchronize(s.edge); cudaStreamDestroy(s.edge); s.edge = nullptr; }
    if (s.ev_pre) { cudaEventDestroy(s.ev_pre); s.ev_pre = nullptr; }
    if (s.ev_edge) { cudaEventDestroy(s.ev_edge); s.ev_edge = nullptr; }
    cudaFree(s.flags);
    s.flags = nullptr;
    cudaGetLastError();
}

extern "C" {

LBM_API int lbm_halo_export(lbm_handle h, void* blob, size_t blob_bytes)
{
    int rc = lbm_internal_use(h);
    if (rc) return rc;
    if (!blob || blob_bytes < LBM_HALO_BLOB_BYTES) return lbm_internal_fail(LBM_ERR_INVALID, "lbm_halo_export: the blob needs LBM_HALO_BLOB_BYTES bytes");
    if (h->cfg.bc[0] != LBM_BC_HALO) return lbm_internal_fail(LBM_ERR_STATE, "lbm_halo_export: the handle's x axis is not LBM_BC_HALO");
    lbm_halo_state& s = h->halo;
    if (s.connected) return lbm_internal_fail(LBM_ERR_STATE, "lbm_halo_export: already connected (lbm_halo_disconnect first)");
    if ((rc = ensure_streams(h))) return rc;
    if (!s.flags) HCK(cudaMalloc(&s.flags, FLAG_BYTES));
    HCK(cudaStreamSynchronize(h->stream));
    HCK(cudaMemset(s.flags, 0, 64));
    HCK(cudaDeviceSynchronize());
    s.epoch = 0;
    HaloBlob b;
    std::memset(&b, 0, sizeof b);
    b.magic = BLOB_MAGIC; b.version = 1;
    b.pid = (int32_t)getpid(); b.device = h->device;
    b.nx = h->g.nx; b.ny = h->g.ny; b.nz = h->g.nz; b.q = h->cfg.q; b.zp = h->g.zp;
    b.inplace = h->inplace ? 1 : 0; b.cur = h->cur; b.aa_parity = h->aa_parity;
    b.handle = (uint64_t)(uintptr_t)h;
    b.buf[0] = (uint64_t)(uintptr_t)h->buf[0]; b.buf[1] = (uint64_t)(uintptr_t)h->buf[1];
    b.flags = (uint64_t)(uintptr_t)s.flags;
    // the IPC handles only matter to another process; a failure here (e.g. on a platform without IPC) leaves them zero
    // and same-process connections still work
    if (cudaIpcGetMemHandle(&b.ipc_buf[0], h->buf[0]) != cudaSuccess) cudaGetLastError();
    if (h->buf[1] && cudaIpcGetMemHandle(&b.ipc_buf[1], h->buf[1]) != cudaSuccess) cudaGetLastError();
    if (cudaIpcGetMemHandle(&b.ipc_flags, s.flags) != cudaSuccess) cudaGetLastError();
    std::memset(blob, 0, LBM_HALO_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    return LBM_OK;
}

LBM_API int lbm_halo_connect(lbm_handle h, const void* lo_blob, const void* hi_blob)
{
    int rc = lbm_internal_use(h);
    if (rc) return rc;
    lbm_halo_state& s = h->halo;
    if (h->cfg.bc[0] != LBM_BC_HALO) return lbm_internal_fail(LBM_ERR_STATE, "lbm_halo_connect: the handle's x axis is not LBM_BC_HALO");
    if (s.connected) return lbm_internal_fail(LBM_ERR_STATE, "lbm_halo_connect: already connected");
    if (!s.flags) return lbm_internal_fail(LBM_ERR_STATE, "lbm_halo_connect: call lbm_halo_export first");
    if (h->inplace && (!lo_blob || !hi_blob))
        return lbm_internal_fail(LBM_ERR_UNSUPPORTED, "the in-place engine needs a neighbour on both sides (periodic global x axis)");
    const void* blobs[2] = {lo_blob, hi_blob};
    HaloBlob bb[2];
    for (int side = 0; side < 2; ++side) {
        if (!blobs[side]) continue;
        std::memcpy(&bb[side], blobs[side], sizeof(HaloBlob));
        const HaloBlob& b = bb[side];
        if (b.magic != BLOB_MAGIC || b.version != 1) return lbm_internal_fail(LBM_ERR_INVALID, "lbm_halo_connect: not a blob of lbm_halo_export");
        if (b.ny != h->g.ny || b.nz != h->g.nz || b.zp != h->g.zp || b.q != h->cfg.q || b.inplace != (h->inplace ? 1 : 0))
            return lbm_internal_fail(LBM_ERR_INVALID, "lbm_halo_connect: the neighbouring slab has another cross-section, velocity set, layout or engine");
        if (h->inplace && b.aa_parity != h->aa_parity)
            return lbm_internal_fail(LBM_ERR_STATE, "lbm_halo_connect: the neighbouring slab is in the other storage form (step counts differ)");
    }
    for (int side = 0; side < 2; ++side) {
        if (!blobs[side]) continue;
        const HaloBlob& b = bb[side];
        float* pb[2] = {nullptr, nullptr};
        unsigned long long* pf = nullptr;
        bool ipc = false;
        if (b.pid == (int32_t)getpid()) {
            pb[0] = (float*)(uintptr_t)b.buf[0]; pb[1] = (float*)(uintptr_t)b.buf[1];
            pf = (unsigned long long*)(uintptr_t)b.flags;
            if (b.device != h->device) {
                int can = 0;
                HCK(cudaDeviceCanAccessPeer(&can, h->device, b.device));
                if (!can) return lbm_internal_fail(LBM_ERR_UNSUPPORTED, "lbm_halo_connect: no peer access between the two devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) HCK(e);
                cudaGetLastError();
            }
        } else if (side == 1 && blobs[0] && bb[0].pid == b.pid && bb[0].handle == b.handle && s.ipc[0]) {
            // a ring of two: the slab above is the slab below, already opened
            pb[0] = s.peer_buf[0][0]; pb[1] = s.peer_buf[0][1];
            pf = s.peer_flag[0] - 1;
            ipc = true;
        } else {
            ipc = true;
            HCK(cudaIpcOpenMemHandle((void**)&pb[0], b.ipc_buf[0], cudaIpcMemLazyEnablePeerAccess));
            if (!h->inplace) HCK(cudaIpcOpenMemHandle((void**)&pb[1], b.ipc_buf[1], cudaIpcMemLazyEnablePeerAccess));
            HCK(cudaIpcOpenMemHandle((void**)&pf, b.ipc_flags, cudaIpcMemLazyEnablePeerAccess));
        }
        s.has[side] = true;
        s.ipc[side] = ipc;
        s.peer_buf[side][0] = pb[0]; s.peer_buf[side][1] = pb[1];
        // I am the HIGH neighbour of the slab below me (its word 1) and the LOW neighbour of the slab above me (its word 0)
        s.peer_flag[side] = pf + (side == 0 ? 1 : 0);
        s.peer_nx[side] = b.nx;
        s.peer_cur_xor[side] = (b.cur ^ h->cur) & 1;
        const long long apart = (long long)(((intptr_t)pb[0] - (intptr_t)h->buf[0]) / (intptr_t)sizeof(float));
        s.delta_inplace[side] = apart + (side == 0 ? (long long)b.nx : -(long long)h->g.nx) * h->g.plane;
        s.np_diff[side] = (long long)(b.nx - h->g.nx) * h->g.plane;
    }
    s.connected = true;
    return LBM_OK;
}

LBM_API int lbm_halo_disconnect(lbm_handle h)
{
    int rc = lbm_internal_use(h);
    if (rc) return rc;
    lbm_halo_state& s = h->halo;
    if (s.edge) HCK(cudaStreamSynchronize(s.edge));
    HCK(cudaStreamSynchronize(h->stream));
    const bool same = s.has[0] && s.has[1] && s.ipc[0] && s.ipc[1] && s.peer_buf[0][0] == s.peer_buf[1][0];
    close_side(h, 0, same);
    close_side(h, 1, false);
    s.connected = false;
    cudaGetLastError();
    return LBM_OK;
}

LBM_API int lbm_halo_timeouts(lbm_handle h, unsigned long long* count)
{
    int rc = lbm_internal_use(h);
    if (rc) return rc;
    if (!count) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    *count = 0;
    if (!h->halo.flags) return LBM_OK;
    HCK(cudaMemcpy(count, h->halo.flags + 2, sizeof *count, cudaMemcpyDeviceToHost));
    return LBM_OK;
}

LBM_API int lbm_group_step(lbm_handle* slabs, int count, int n, unsigned phases)
{
    if (!slabs || count < 1 || n < 0 || (phases & ~15u) || phases == 0) return lbm_internal_fail(LBM_ERR_INVALID, "lbm_group_step: bad arguments");
    for (int k = 0; k < count; ++k)
        if (!slabs[k]) return lbm_internal_fail(LBM_ERR_INVALID, "lbm_group_step: null handle");
    for (int step = 0; step < n; ++step) {
        // every slab's signal is queued before any slab's wait: one host thread can drive all GPUs without a deadlock
        for (int k = 0; k < count; ++k) {
            int rc = lbm_internal_use(slabs[k]);
            if (rc) return rc;
            if ((rc = lbm_internal_step_a(slabs[k], phases))) return rc;
        }
        for (int k = 0; k < count; ++k) {
            int rc = lbm_internal_use(slabs[k]);
            if (rc) return rc;
            if ((rc = lbm_internal_step_b(slabs[k], phases))) return rc;
        }
    }
    return LBM_OK;
}

}  // extern "C"
