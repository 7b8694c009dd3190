// lbm_capi.cu -- the C ABI of include/lbm_b200.h: handle, device memory, host<->device
// transfers, and the step driver that replaces Lattice::simulate* (the reference allocates a
// second lattice, copies the whole lattice and frees the old one twice per step,
// simulation/source/Lattice.cpp:115-210; here the two population buffers are allocated once
// and only their roles swap).
#include "../../include/lbm_b200.h"
#include "lbm_device.h"
#include "lbm_internal.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace lbm;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? LBM_ERR_NOMEM : LBM_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                               \
    } while (0)

constexpr size_t STAGING_BYTES = size_t(256) << 20;
constexpr int AOS_BYTES = 144;     // sizeof(LatticePoint), simulation/include/LatticePoint.h:9-17
constexpr int TABLE_BYTES = 380;   // sizeof(ReflectionValues), LatticePoint.h:3-7

}  // namespace

namespace {

int use(lbm_handle h)
{
    if (!h) return fail(LBM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    return LBM_OK;
}

int ensure_staging(lbm_handle h)
{
    if (!h->staging) CK(cudaMalloc(&h->staging, STAGING_BYTES));
    return LBM_OK;
}

__global__ void any_nonzero_kernel(const uint32_t* w, size_t n, unsigned* flag)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = (i < n) && w[i] != 0u;
    if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

// Keep the class array (main stream, read by every warp of the general kernel before anything else) and the
// neighbourhood words of the special list (side stream, read by every warp of the special kernel before it can
// issue a single population load) resident in L2: a few MB and a few tens of MB that the population stream
// would otherwise evict every step.
void pin_tables_in_l2(lbm_handle h)
{
    // measured on B200 (profiles/r01_notes.md): pinning the 45 MB of neighbourhood words of the 512^3 obstacle costs
    // the general kernel 20 % (it loses that much L2 as a write buffer), so only the class bytes are pinned
    static const bool pin_info = [] { const char* e = getenv("LBM_PIN_CELLINFO"); return e ? atoi(e) != 0 : false; }();
    static const bool pin_cls = [] { const char* e = getenv("LBM_PIN_CLS"); return e ? atoi(e) != 0 : true; }();
    if (!pin_cls) return;
    int dev = h->device, max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persist <= 0 || max_window <= 0) { cudaGetLastError(); return; }
    const size_t info_bytes = (pin_info && h->cellinfo) ? h->n_special * 32 * sizeof(uint32_t) : 0;
    size_t want = h->cls_bytes + info_bytes;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = h->cls;
    attr.accessPolicyWindow.num_bytes = h->cls_bytes < (size_t)max_window ? h->cls_bytes : (size_t)max_window;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    if (info_bytes) {
        const size_t room = want > h->cls_bytes ? want - h->cls_bytes : 0;
        attr.accessPolicyWindow.base_ptr = h->cellinfo;
        attr.accessPolicyWindow.num_bytes = info_bytes < (size_t)max_window ? info_bytes : (size_t)max_window;
        attr.accessPolicyWindow.hitRatio = room >= info_bytes ? 1.0f : (float)((double)room / (double)info_bytes);
    }
    cudaStreamSetAttribute(h->side, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();   // best effort: an unsupported attribute must not poison later calls
}

// Up to LBM_DEAD_BOXES disjoint boxes (in x, y and 32-cell z-words) that hold all-solid words only, as large as a greedy search finds:
// the warps of the single-buffer kernels that land inside leave on a compare of kernel parameters instead of waiting for
// their class byte or mask word (a third of the warps of the 512^3 obstacle).  Each box is grown from a seed word, one
// layer at a time over its six faces in turn, while the new layer is all solid (3-D prefix sums answer that in O(1));
// the best of a few hundred seeds wins.  Conservative by construction: a word outside every box is treated as before.
void find_dead_boxes(const std::vector<uint8_t>& cls, int nx, int ny, int wc, int pitch, int n_boxes, int (*box)[6], double* covered)
{
    const size_t sy = (size_t)wc + 1, sx = ((size_t)ny + 1) * sy;
    std::vector<uint8_t> dead((size_t)nx * ny * wc);
    std::vector<int32_t> S(((size_t)nx + 1) * sx);
    size_t n_dead = 0, taken = 0;
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int w = 0; w < wc; ++w) {
                const bool d = cls[((size_t)x * wc + w) * pitch + y] == 1;      // 1 = every cell of the word is solid
                dead[((size_t)x * ny + y) * wc + w] = d;
                n_dead += d;
            }
    for (int b = 0; b < LBM_DEAD_BOXES; ++b)
        for (int k = 0; k < 6; ++k) box[b][k] = 0;
    *covered = 0.0;
    if (n_dead == 0) return;
    auto count = [&](int x0, int x1, int y0, int y1, int w0, int w1) {
        auto at = [&](int x, int y, int w) { return S[(size_t)x * sx + (size_t)y * sy + w]; };
        return at(x1, y1, w1) - at(x0, y1, w1) - at(x1, y0, w1) - at(x1, y1, w0) + at(x0, y0, w1) + at(x0, y1, w0) + at(x1, y0, w0) - at(x0, y0, w0);
    };
    for (int b = 0; b < n_boxes && b < LBM_DEAD_BOXES; ++b) {
        std::fill(S.begin(), S.end(), 0);
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y) {
                int32_t row = 0;
                for (int w = 0; w < wc; ++w) {
                    row += dead[((size_t)x * ny + y) * wc + w];
                    S[(size_t)(x + 1) * sx + (size_t)(y + 1) * sy + (w + 1)] =
                        row + S[(size_t)x * sx + (size_t)(y + 1) * sy + (w + 1)] + S[(size_t)(x + 1) * sx + (size_t)y * sy + (w + 1)] -
                        S[(size_t)x * sx + (size_t)y * sy + (w + 1)];
                }
            }
        const int32_t left = S[(size_t)nx * sx + (size_t)ny * sy + wc];
        if (left == 0) break;
        // seeds: every (left / 384)-th remaining all-solid word in storage order
        const int32_t stride = left / 384 + 1;
        long long best = 0;
        int bb[6] = {0, 0, 0, 0, 0, 0};
        int32_t seen = 0;
        for (size_t k = 0; k < dead.size(); ++k) {
            if (!dead[k]) continue;
            if (seen++ % stride != stride / 2) continue;
            int lo[3] = {(int)(k / ((size_t)ny * wc)), (int)(k / wc % ny), (int)(k % wc)};
            int hi[3] = {lo[0] + 1, lo[1] + 1, lo[2] + 1};
            const int lim[3] = {nx, ny, wc};
            for (bool grew = true; grew;) {
                grew = false;
                for (int a = 0; a < 3; ++a)
                    for (int side = 0; side < 2; ++side) {
                        int l2[3] = {lo[0], lo[1], lo[2]}, h2[3] = {hi[0], hi[1], hi[2]};
                        if (side == 0) { if (lo[a] == 0) continue; l2[a] = lo[a] - 1; h2[a] = lo[a]; }
                        else           { if (hi[a] == lim[a]) continue; l2[a] = hi[a]; h2[a] = hi[a] + 1; }
                        const long long layer = (long long)(h2[0] - l2[0]) * (h2[1] - l2[1]) * (h2[2] - l2[2]);
                        if (count(l2[0], h2[0], l2[1], h2[1], l2[2], h2[2]) != layer) continue;
                        if (side == 0) lo[a] -= 1; else hi[a] += 1;
                        grew = true;
                    }
            }
            const long long vol = (long long)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
            if (vol > best) { best = vol; for (int a = 0; a < 3; ++a) { bb[a] = lo[a]; bb[3 + a] = hi[a] - lo[a]; } }
        }
        if (best == 0) break;
        for (int k = 0; k < 6; ++k) box[b][k] = bb[k];
        taken += (size_t)best;
        for (int x = bb[0]; x < bb[0] + bb[3]; ++x)          // the next box is searched in what is left
            for (int y = bb[1]; y < bb[1] + bb[4]; ++y)
                for (int w = bb[2]; w < bb[2] + bb[5]; ++w) dead[((size_t)x * ny + y) * wc + w] = 0;
    }
    *covered = (double)taken / (double)n_dead;
}

// rebuild the classification of every interior 32-cell word and the neighbourhood words of the special
// ones; call after the solid mask or the link lists change
int rebuild_special(lbm_handle h)
{
    const Geom& g = h->g;
    const int wcount = (g.nz + 31) / 32;
    if (!h->cls) {
        h->cls_pitch = ((g.ny + 15) / 16) * 16;
        h->cls_bytes = (size_t)g.nx * wcount * h->cls_pitch;
        CK(cudaMalloc(&h->cls, h->cls_bytes));
    }
    const uint32_t* solid = h->any_solid ? h->solid : nullptr;
    const uint32_t* tmask = h->nlinks ? h->tmask : nullptr;
    if ((solid || tmask) && !h->special_list)
        CK(cudaMalloc(&h->special_list, (size_t)g.nx * g.ny * wcount * sizeof(uint32_t)));
    if (!h->flag) CK(cudaMalloc(&h->flag, sizeof(unsigned)));
    CK(cudaMemsetAsync(h->flag, 0, sizeof(unsigned), h->stream));
    CK(cudaMemsetAsync(h->cls, 0, h->cls_bytes, h->stream));
    CK(launch_build_cellinfo(solid, tmask, h->cls, h->cls_pitch, h->special_list, h->flag, g, h->stream));
    unsigned n = 0;
    CK(cudaMemcpyAsync(&n, h->flag, sizeof n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_special = n;
    h->special_start.assign((size_t)g.nx + 1, 0u);
    if (n > 0) {
        // the list is filled in atomic order; sorting it puts the z-words of a row and the rows of a plane next
        // to each other again, so neighbouring warps share sectors and DRAM pages like in the dense kernels,
        // and the entries of a range of planes become a range of the list
        std::vector<uint32_t> list(n);
        // (all transfers on the handle's stream: it is a non-blocking stream, so a cudaMemcpy on the legacy default stream
        // would not be ordered against the kernels that read the list next)
        CK(cudaMemcpyAsync(list.data(), h->special_list, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        std::sort(list.begin(), list.end());
        CK(cudaMemcpyAsync(h->special_list, list.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));        // `list` is pageable and goes out of scope
        for (int x = 0; x <= g.nx; ++x) {
            const uint64_t first_word = (uint64_t)(x + 1) * g.yp * g.wz;      // mask words of padded plane x + 1 start here
            h->special_start[x] = (uint32_t)(std::lower_bound(list.begin(), list.end(), first_word,
                                                              [](uint32_t a, uint64_t b) { return (uint64_t)a < b; }) - list.begin());
        }
    }
    if (n > h->cellinfo_cap) {
        cudaFree(h->cellinfo);
        h->cellinfo = nullptr;
        h->cellinfo_cap = 0;
        CK(cudaMalloc(&h->cellinfo, (size_t)n * 32 * sizeof(uint32_t)));
        h->cellinfo_cap = n;
    }
    CK(launch_fill_cellinfo(solid, tmask, h->special_list, (long long)n, h->cellinfo, g, h->stream));
    std::memset(h->dead_box, 0, sizeof h->dead_box);
    const int whole[6] = {0, 0, 0, g.nx, g.ny, wcount};
    std::memcpy(h->busy_box, whole, sizeof whole);
    static const int n_boxes = [] { const char* e = getenv("LBM_AA_DEAD_BOXES"); return e ? atoi(e) : LBM_DEAD_BOXES; }();
    if (n_boxes > 0 && h->inplace && solid) {
        std::vector<uint8_t> cls(h->cls_bytes);
        CK(cudaMemcpyAsync(cls.data(), h->cls, h->cls_bytes, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        double covered = 0.0;
        find_dead_boxes(cls, g.nx, g.ny, wcount, h->cls_pitch, n_boxes, h->dead_box, &covered);
        static const bool use_busy = [] { const char* e = getenv("LBM_AA_BUSY_BOX"); return e ? atoi(e) != 0 : true; }();
        if (use_busy) {
            int lo[3] = {g.nx, g.ny, wcount}, hi[3] = {0, 0, 0};
            for (int x = 0; x < g.nx; ++x)
                for (int w = 0; w < wcount; ++w) {
                    const uint8_t* row = cls.data() + ((size_t)x * wcount + w) * h->cls_pitch;
                    for (int y = 0; y < g.ny; ++y)
                        if (row[y] != 0) {           // 0 = ordinary
                            lo[0] = std::min(lo[0], x); hi[0] = std::max(hi[0], x + 1);
                            lo[1] = std::min(lo[1], y); hi[1] = std::max(hi[1], y + 1);
                            lo[2] = std::min(lo[2], w); hi[2] = std::max(hi[2], w + 1);
                        }
                }
            for (int a = 0; a < 3; ++a) { h->busy_box[a] = hi[a] > lo[a] ? lo[a] : 0; h->busy_box[3 + a] = hi[a] > lo[a] ? hi[a] - lo[a] : 0; }
        }
        if (getenv("LBM_VERBOSE")) {
            for (int b = 0; b < LBM_DEAD_BOXES; ++b)
                fprintf(stderr, "lbm: all-solid box %d: origin %d %d %d extent %d %d %d (x, y, z-words)\n", b, h->dead_box[b][0], h->dead_box[b][1],
                        h->dead_box[b][2], h->dead_box[b][3], h->dead_box[b][4], h->dead_box[b][5]);
            fprintf(stderr, "lbm: the boxes hold %.1f %% of the all-solid words; words that are not ordinary lie in origin %d %d %d extent %d %d %d of %d %d %d\n",
                    100.0 * covered, h->busy_box[0], h->busy_box[1], h->busy_box[2], h->busy_box[3], h->busy_box[4], h->busy_box[5], g.nx, g.ny, wcount);
        }
    }
    pin_tables_in_l2(h);
    h->dead_dirty = 2;
    return LBM_OK;
}

int rebuild_mask(lbm_handle h, const uint8_t* dev_bytes, int has_x_ghost)
{
    if (!h->solid) CK(cudaMalloc(&h->solid, h->mask_words() * sizeof(uint32_t)));
    CK(launch_build_mask(dev_bytes, has_x_ghost, h->solid, h->g, h->cfg.bc, h->stream));
    if (!h->flag) CK(cudaMalloc(&h->flag, sizeof(unsigned)));
    CK(cudaMemsetAsync(h->flag, 0, sizeof(unsigned), h->stream));
    const size_t n = h->mask_words();
    any_nonzero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->solid, n, h->flag);
    CK(cudaGetLastError());
    unsigned f = 0;
    CK(cudaMemcpyAsync(&f, h->flag, sizeof f, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->any_solid = f != 0;
    return rebuild_special(h);
}

void fill_params(lbm_handle h, StepParams& p, unsigned phases, int x0, int x1)
{
    std::memset(&p, 0, sizeof p);
    p.fin = h->buf[h->cur];
    p.fout = h->buf[h->cur ^ 1];
    p.g = h->g;
    p.x0 = x0; p.x1 = x1;
    p.tau = h->cfg.tau;
    p.inv_tau = 1.0f / h->cfg.tau;
    p.phases = phases;
    p.boundary_mode = h->cfg.boundary_mode;
    p.solid = h->any_solid ? h->solid : nullptr;
    p.cls = h->cls;
    p.cls_pitch = h->cls_pitch;
    p.cellinfo = h->cellinfo;
    p.special_list = h->special_list;
    if (h->n_special && (int)h->special_start.size() == h->g.nx + 1 && x0 >= 0 && x1 <= h->g.nx && x0 < x1) {
        p.special_k0 = h->special_start[x0];
        p.n_special = (long long)h->special_start[x1] - (long long)h->special_start[x0];
    }
    // all-solid words hold zeros and are never read: bounce-back solids are dead in every phase, the
    // reference's solids are emptied by the reflect phase.  They are written while stale values may remain.
    const bool dead_ok = h->cfg.boundary_mode == LBM_BOUNDARY_BOUNCE_BACK || (phases & PH_REFLECT);
    const bool src_hits_dead = h->cfg.boundary_mode == LBM_BOUNDARY_REFLECT && (phases & PH_SOURCE) && h->src_on;
    p.dead_mode = !dead_ok ? 0 : ((h->dead_dirty > 0 || src_hits_dead) ? 1 : 2);
    if (h->nlinks) {
        p.tmask = h->tmask; p.tprefix = h->tprefix; p.toff = h->toff; p.links = h->links;
    }
    p.force = (h->force_on && (h->nlinks || h->cfg.boundary_mode == LBM_BOUNDARY_BOUNCE_BACK)) ? h->force : nullptr;
    p.force_mode = h->cfg.force_mode;
    if (p.force && h->force32 && h->n_force_terms) {
        p.link_rank = h->link_rank;
        p.force_moved = h->force_moved;
    }
    p.src_on = h->src_on;
    for (int a = 0; a < 3; ++a) { p.src_lo[a] = h->src_lo[a]; p.src_hi[a] = h->src_hi[a]; p.bc[a] = h->cfg.bc[a]; }
    std::memcpy(p.wall_u, h->cfg.wall_u, sizeof p.wall_u);
    p.wall_moving = 0;
    for (int k = 0; k < 6; ++k)
        for (int a = 0; a < 3; ++a)
            if (h->cfg.wall_u[k][a] != 0.0f) p.wall_moving = 1;
    std::memcpy(p.src_f, h->src_f, sizeof p.src_f);
    p.src_mask = h->src_masked ? h->src_mask : nullptr;
    for (int i = 0; i < 27; ++i) {
        p.pop_off[i] = (long long)i * h->g.np;
        p.pull_off[i] = p.pop_off[i] - ((long long)dir_c(i, 0) * h->g.plane + (long long)(dir_c(i, 1) * h->g.zp + dir_c(i, 2)));
    }
    if (h->halo.connected) {
        // a pull through a slab face reads the neighbour's current buffer where it lies: ghost plane 0 is the low
        // neighbour's last plane, ghost plane nx + 1 the high neighbour's first plane
        p.xdelta_on = 1;
        for (int side = 0; side < 2; ++side) {
            if (!h->halo.has[side]) continue;
            const float* peer = h->halo.peer_buf[side][h->cur ^ h->halo.peer_cur_xor[side]];
            const long long planes = side == 0 ? (long long)h->halo.peer_nx[0] : -(long long)h->g.nx;
            const long long apart = (long long)(((intptr_t)peer - (intptr_t)p.fin) / (intptr_t)sizeof(float)) + planes * h->g.plane;
            for (int i = 0; i < 27; ++i) p.xdelta[side][i] = apart + (long long)i * h->halo.np_diff[side];
        }
    }
}

// is the dense kernel (pull + collide on an all-fluid lattice) enough for these phases?
bool fast_path_ok(lbm_handle h, unsigned phases)
{
    if (!(phases & PH_STREAM) || !(phases & PH_COLLIDE)) return false;
    if (h->cfg.bc[2] != LBM_BC_ZERO_INFLOW) return false;     // periodic / wall z edges are resolved by the general kernels
    if (h->any_solid) return false;
    if (h->halo.connected) return false;                      // the dense kernel knows nothing of neighbouring slabs
    if ((phases & PH_REFLECT) && h->nlinks) return false;
    if ((phases & PH_SOURCE) && h->src_on) return false;
    return true;
}

int launch_range(lbm_handle h, unsigned phases, int x0, int x1, cudaStream_t s)
{
    StepParams p;
    fill_params(h, p, phases, x0, x1);
    LaunchCfg c;
    c.q = h->cfg.q;
    c.moment_mode = h->cfg.moment_mode;
    c.eq64 = h->cfg.eq_precision == LBM_EQ_FP64;
    c.general = fast_path_ok(h, phases) ? 0 : 1;
    // On the handle's main stream the special-word kernel runs next to the general kernel on the side stream:
    // the two touch disjoint cells of the output and only read the input, and the latency-bound special
    // kernel fills bandwidth the general kernel leaves.
    static const bool concurrent = [] { const char* e = getenv("LBM_SPECIAL_CONCURRENT"); return e ? atoi(e) != 0 : true; }();
    const bool split = concurrent && c.general && p.n_special > 0 && s == h->stream;
    if (split) {
        CK(cudaEventRecord(h->ev_main, h->stream));
        CK(cudaStreamWaitEvent(h->side, h->ev_main, 0));
    }
    int launched = 0;
    CK(launch_step(c, p, s, split ? h->side : nullptr, &launched));
    if (split) {
        CK(cudaEventRecord(h->ev_side, h->side));
        CK(cudaStreamWaitEvent(h->stream, h->ev_side, 0));
    }
    h->launches += launched;
    return LBM_OK;
}

// after a step: the buffer just written has clean (zero) all-solid words if the step wrote or skipped them
void note_step_done(lbm_handle h, unsigned phases)
{
    const bool dead_ok = h->cfg.boundary_mode == LBM_BOUNDARY_BOUNCE_BACK || (phases & PH_REFLECT);
    if (!dead_ok) h->dead_dirty = 2;
    else if (h->dead_dirty > 0) h->dead_dirty -= 1;
}

int fill_ghosts(lbm_handle h, bool include_x)
{
    int n = 0;
    CK(launch_fill_ghosts(h->buf[h->cur], h->g, h->cfg.q, h->cfg.bc, h->cfg.wall_u, include_x, h->stream, &n));
    h->launches += n;
    return LBM_OK;
}

void fill_inplace(lbm_handle h, InplaceParams& ip)
{
    std::memset(&ip, 0, sizeof ip);
    set_inplace_offsets(ip, h->g);
    ip.f = h->buf[0];
    ip.tau = h->cfg.tau;
    ip.inv_tau = 1.0f / h->cfg.tau;
    ip.parity = h->aa_parity;
    for (int a = 0; a < 3; ++a) ip.bc[a] = h->cfg.bc[a];
    if (h->cfg.bc[0] == LBM_BC_HALO) {
        // x slabs: the neighbouring slab's edge plane in its own allocation (lbm_halo_connect), or this slab's own
        // ghost plane when the caller moves the planes (no connection)
        for (int side = 0; side < 2; ++side)
            for (int i = 0; i < 27; ++i)
                ip.xshift[side][i] = h->halo.delta_inplace[side] + (long long)i * h->halo.np_diff[side];
    }
    std::memcpy(ip.wall_u, h->cfg.wall_u, sizeof ip.wall_u);
    if (h->any_solid) {       // obstacles: the mask, the word classes and the special list of rebuild_special
        ip.solid = h->solid;
        ip.cls = h->cls;
        ip.cls_pitch = h->cls_pitch;
        ip.special_list = h->special_list;
        ip.n_special = (long long)h->n_special;
        ip.cellinfo = h->cellinfo;
        ip.force = h->force_on ? h->force : nullptr;
        static const bool split_even = [] { const char* e = getenv("LBM_AA_FORCE_SPLIT"); return e ? atoi(e) != 0 : false; }();
        ip.force_pending = (ip.force && !split_even) ? h->force + 3 * FORCE_SLOTS : nullptr;
        static_assert(sizeof ip.dead_box == sizeof h->dead_box, "box tables differ");
        std::memcpy(ip.dead_box, h->dead_box, sizeof ip.dead_box);
        std::memcpy(ip.busy_box, h->busy_box, sizeof ip.busy_box);
    }
    ip.src_on = h->src_on;
    for (int a = 0; a < 3; ++a) { ip.src_lo[a] = h->src_lo[a]; ip.src_hi[a] = h->src_hi[a]; }
    std::memcpy(ip.src_f, h->src_f, sizeof ip.src_f);
    ip.src_mask = h->src_masked ? h->src_mask : nullptr;
}

int inplace_only(lbm_handle h, const char* what)
{
    if (h->inplace) return fail(LBM_ERR_UNSUPPORTED, "%s is not available on the in-place engine (bounce-back obstacles, box source, fused steps, SoA transfers)", what);
    return LBM_OK;
}

// LBM_FORCE_ACCUM_REFERENCE_FP32: the links of this step have stored what they moved; add the terms up in the oracle's order
int force_reference_sum(lbm_handle h, unsigned phases)
{
    if (!(phases & PH_REFLECT) || !h->force_on || !h->force32 || !h->n_force_terms) return LBM_OK;
    CK(launch_force_sequential(h->force_moved, h->force_coef, (long long)h->n_force_terms, h->force32, h->stream));
    h->launches += 1;
    return LBM_OK;
}

// planes [xa, xb) of one in-place step on stream s
int inplace_range(lbm_handle h, unsigned phases, int xa, int xb, cudaStream_t s)
{
    if (xb <= xa) return LBM_OK;
    InplaceParams ip;
    fill_inplace(h, ip);
    if (!(phases & PH_SOURCE)) ip.src_on = 0;
    ip.x0 = xa; ip.x1 = xb;
    if (ip.n_special && (int)h->special_start.size() == h->g.nx + 1) {
        ip.special_k0 = h->special_start[xa];
        ip.n_special = (long long)h->special_start[xb] - (long long)h->special_start[xa];
    }
    // on the handle's main stream the face and special-word kernels of an odd step run next to the interior launch (on
    // the high-priority side stream): they touch disjoint slots, and the tail of one launch overlaps the other
    static const bool concurrent = [] { const char* e = getenv("LBM_AA_CONCURRENT"); return e ? atoi(e) != 0 : true; }();
    const bool fork = concurrent && s == h->stream && h->aa_parity == 0;
    if (fork) {
        CK(cudaEventRecord(h->ev_main, h->stream));
        CK(cudaStreamWaitEvent(h->side, h->ev_main, 0));
    }
    int launched = 0;
    CK(launch_inplace_step(ip, h->cfg.q, h->cfg.moment_mode, h->cfg.eq_precision == LBM_EQ_FP64, s, fork ? h->side : nullptr, &launched));
    if (fork) {
        CK(cudaEventRecord(h->ev_side, h->side));
        CK(cudaStreamWaitEvent(h->stream, h->ev_side, 0));
    }
    h->launches += launched;
    return LBM_OK;
}

// First half of a step: everything that needs no neighbour.  On a slab with peer-memory halos it ends with the signal
// "my previous step is complete, my edge planes are yours to read" to both neighbours.
int step_a(lbm_handle h, unsigned phases)
{
    if (h->inplace) {
        if (!(phases & PH_STREAM) || !(phases & PH_COLLIDE))
            return fail(LBM_ERR_UNSUPPORTED, "the in-place engine runs fused stream + collide steps only");
        if (h->aa_dirty) {
            // first step after the mask or the state was set: solid cells hold zeros from here on
            if (h->any_solid) {
                if (h->aa_parity != 0) return fail(LBM_ERR_STATE, "in-place engine: solid cells can only be cleared in the parked form");
                InplaceParams ip;
                fill_inplace(h, ip);
                CK(launch_inplace_zero_solid(ip, h->cfg.q, h->stream));
                h->launches += 1;
            }
            h->aa_dirty = false;
        }
        if (h->aa_parity == 0 && h->force_on && h->any_solid) {
            // the momentum the previous odd step summed ahead of time belongs to the even step that has run since
            CK(launch_inplace_fold_pending(h->force, h->force + 3 * FORCE_SLOTS, h->stream));
            h->launches += 1;
        }
    } else {
        if (h->pending_commit) return fail(LBM_ERR_STATE, "lbm_step_range in flight: call lbm_step_commit first");
        if (phases & PH_STREAM) {
            int rc = fill_ghosts(h, true);
            if (rc) return rc;
        }
    }
    if (h->halo.connected) {
        CK(cudaEventRecord(h->halo.ev_pre, h->stream));
        CK(cudaStreamWaitEvent(h->halo.edge, h->halo.ev_pre, 0));
        int rc = lbm_internal_halo_signal(h);
        if (rc) return rc;
    }
    return LBM_OK;
}

// Second half: (slabs: wait for both neighbours' signal, then) the edge planes on the edge stream next to the interior
// planes on the main stream, joined at the end.
int step_b(lbm_handle h, unsigned phases)
{
    const int nx = h->g.nx;
    int rc;
    if (!h->halo.connected) {
        if (h->inplace) {
            if ((rc = inplace_range(h, phases, 0, nx, h->stream))) return rc;
            h->aa_parity ^= 1;
            return LBM_OK;
        }
        if ((rc = launch_range(h, phases, 0, nx, h->stream))) return rc;
        if ((rc = force_reference_sum(h, phases))) return rc;
        h->cur ^= 1;
        note_step_done(h, phases);
        return LBM_OK;
    }
    if ((rc = lbm_internal_halo_wait(h))) return rc;
    // table reflection gathers through a solid cell of the neighbouring plane: two planes per face see the neighbour
    int e = (!h->inplace && h->nlinks) ? 2 : 1;
    cudaStream_t es = h->halo.edge;
    if (nx < 2 * e + 1) {
        // no interior to overlap with: the whole slab runs behind the handshake
        rc = h->inplace ? inplace_range(h, phases, 0, nx, es) : launch_range(h, phases, 0, nx, es);
        if (rc) return rc;
    } else {
        if (h->inplace) {
            if ((rc = inplace_range(h, phases, 0, e, es))) return rc;
            if ((rc = inplace_range(h, phases, nx - e, nx, es))) return rc;
            if ((rc = inplace_range(h, phases, e, nx - e, h->stream))) return rc;
        } else {
            if ((rc = launch_range(h, phases, 0, e, es))) return rc;
            if ((rc = launch_range(h, phases, nx - e, nx, es))) return rc;
            if ((rc = launch_range(h, phases, e, nx - e, h->stream))) return rc;
        }
    }
    CK(cudaEventRecord(h->halo.ev_edge, es));
    CK(cudaStreamWaitEvent(h->stream, h->halo.ev_edge, 0));
    if (h->inplace) {
        h->aa_parity ^= 1;
    } else {
        h->cur ^= 1;
        note_step_done(h, phases);
    }
    return LBM_OK;
}

int one_step(lbm_handle h, unsigned phases)
{
    int rc = step_a(h, phases);
    if (rc) return rc;
    return step_b(h, phases);
}

}  // namespace

// doors for the other translation units of the library (lbm_internal.h)
int lbm_internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }
int lbm_internal_use(lbm_handle h) { return use(h); }
int lbm_internal_rebuild_mask(lbm_handle h, const uint8_t* dev_bytes, int has_x_ghost) { return rebuild_mask(h, dev_bytes, has_x_ghost); }
int lbm_internal_step_a(lbm_handle h, unsigned phases) { return step_a(h, phases); }
int lbm_internal_step_b(lbm_handle h, unsigned phases) { return step_b(h, phases); }

// ---- C ABI -------------------------------------------------------------------------------------

extern "C" {

LBM_API const char* lbm_last_error(void) { return g_last_error.c_str(); }
LBM_API const char* lbm_version(void) { return "lbm_b200 0.1 (sm_100a)"; }

LBM_API int lbm_config_default(lbm_config* cfg)
{
    if (!cfg) return fail(LBM_ERR_INVALID, "null config");
    std::memset(cfg, 0, sizeof *cfg);
    cfg->nx = cfg->ny = cfg->nz = 1;
    cfg->q = 27;
    cfg->tau = 1.0f;
    cfg->moment_mode = LBM_MOMENT_REFERENCE;
    cfg->eq_precision = LBM_EQ_FP64;
    cfg->boundary_mode = LBM_BOUNDARY_REFLECT;
    cfg->force_mode = LBM_FORCE_REFERENCE_SLOT_INDEX;
    cfg->engine = LBM_ENGINE_AUTO;
    cfg->device = -1;
    return LBM_OK;
}

LBM_API int lbm_create(const lbm_config* cfg, lbm_handle* out)
{
    if (!cfg || !out) return fail(LBM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->nx < 1 || cfg->ny < 1 || cfg->nz < 1) return fail(LBM_ERR_INVALID, "lattice extents must be >= 1");
    if (cfg->nx > 65535 || cfg->ny > 65535) return fail(LBM_ERR_INVALID, "lattice extent too large for one launch (nx, ny <= 65535)");
    if (cfg->q != 19 && cfg->q != 27) return fail(LBM_ERR_INVALID, "q must be 19 or 27");
    if (!(cfg->tau > 0.0f)) return fail(LBM_ERR_INVALID, "tau must be positive");
    for (int a = 0; a < 3; ++a) {
        if (cfg->bc[a] < 0 || cfg->bc[a] > 3) return fail(LBM_ERR_INVALID, "bad bc");
        if (cfg->bc[a] == LBM_BC_HALO && a != 0) return fail(LBM_ERR_INVALID, "LBM_BC_HALO is for the x axis only");
    }
    if (cfg->engine != LBM_ENGINE_AUTO && cfg->engine != LBM_ENGINE_PULL2 && cfg->engine != LBM_ENGINE_INPLACE)
        return fail(LBM_ERR_UNSUPPORTED, "engine %d is not built in this version", cfg->engine);
    if (cfg->force_accum != LBM_FORCE_ACCUM_F64_PARTIALS && cfg->force_accum != LBM_FORCE_ACCUM_REFERENCE_FP32)
        return fail(LBM_ERR_INVALID, "bad force_accum");
    if (cfg->force_accum == LBM_FORCE_ACCUM_REFERENCE_FP32 &&
        (cfg->engine == LBM_ENGINE_INPLACE || cfg->boundary_mode != LBM_BOUNDARY_REFLECT || cfg->bc[0] == LBM_BC_HALO))
        return fail(LBM_ERR_UNSUPPORTED, "LBM_FORCE_ACCUM_REFERENCE_FP32 restates the reference's accumulator: table reflection on the two-buffer engine, one handle");
    if (cfg->engine == LBM_ENGINE_INPLACE)
        for (int a = 0; a < 3; ++a)
            if (cfg->bc[a] != LBM_BC_PERIODIC && cfg->bc[a] != LBM_BC_WALL && cfg->bc[a] != LBM_BC_HALO)
                return fail(LBM_ERR_UNSUPPORTED, "the in-place engine takes periodic, wall or halo (x slab) edges only");
    const Geom g = make_geom(cfg->nx, cfg->ny, cfg->nz);
    if (g.np >= (1ll << 32)) return fail(LBM_ERR_INVALID, "more than 2^32 padded cells per slab");

    int dev = cfg->device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    CK(cudaSetDevice(dev));

    lbm_lattice_s* h = new (std::nothrow) lbm_lattice_s();
    if (!h) return fail(LBM_ERR_NOMEM, "host allocation failed");
    h->cfg = *cfg;
    h->cfg.device = dev;
    h->g = g;
    h->device = dev;
    h->inplace = cfg->engine == LBM_ENGINE_INPLACE;
    auto bail = [&](int rc) { lbm_destroy(h); return rc; };
#define CKB(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return bail(fail(e_ == cudaErrorMemoryAllocation ? LBM_ERR_NOMEM : LBM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_))); \
    } while (0)
    CKB(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    {
        // the side stream outranks the main one: its (smaller) kernels get SM slots as soon as blocks of the
        // main kernel retire, instead of waiting for the main grid to drain
        int lo_prio = 0, hi_prio = 0;
        CKB(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CKB(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi_prio));
        // the bulk stream carries the interior planes of a slab at the lowest priority, so that the halo
        // transfer posted meanwhile (NCCL or copies) is not starved of SM slots by that large grid
        CKB(cudaStreamCreateWithPriority(&h->bulk, cudaStreamNonBlocking, lo_prio));
    }
    CKB(cudaEventCreate(&h->ev0));
    CKB(cudaEventCreate(&h->ev1));
    CKB(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
    CKB(cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
    CKB(cudaEventCreateWithFlags(&h->ev_bulk, cudaEventDisableTiming));
    for (int b = 0; b < (h->inplace ? 1 : 2); ++b) {
        CKB(cudaMalloc(&h->buf[b], h->state_bytes()));
        CKB(cudaMemsetAsync(h->buf[b], 0, h->state_bytes(), h->stream));
    }
    // (two sets of partial sums: the second holds what the single-buffer engine sums ahead of time, see InplaceParams)
    CKB(cudaMalloc(&h->force, 2 * 3 * FORCE_SLOTS * sizeof(double)));
    CKB(cudaMemsetAsync(h->force, 0, 2 * 3 * FORCE_SLOTS * sizeof(double), h->stream));
    CKB(cudaMalloc(&h->diag, 4 * sizeof(double)));
    if (cfg->force_accum == LBM_FORCE_ACCUM_REFERENCE_FP32) {
        CKB(cudaMalloc(&h->force32, 3 * sizeof(float)));
        CKB(cudaMemsetAsync(h->force32, 0, 3 * sizeof(float), h->stream));
    }
#undef CKB
    {
        int rc = rebuild_special(h);      // the word classes exist even without geometry
        if (rc) return bail(rc);
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return bail(fail(LBM_ERR_CUDA, "create: %s", cudaGetErrorString(e)));
    *out = h;
    return LBM_OK;
}

LBM_API int lbm_destroy(lbm_handle h)
{
    if (!h) return LBM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->side) cudaStreamSynchronize(h->side);
    lbm_internal_halo_release(h);
    cudaFree(h->moments_scratch);
    for (int b = 0; b < 2; ++b) cudaFree(h->buf[b]);
    cudaFree(h->solid); cudaFree(h->cls); cudaFree(h->cellinfo); cudaFree(h->special_list); cudaFree(h->solid_bytes); cudaFree(h->tmask); cudaFree(h->tprefix); cudaFree(h->toff);
    cudaFree(h->reflected); cudaFree(h->links); cudaFree(h->force); cudaFree(h->diag); cudaFree(h->flag);
    cudaFree(h->link_rank); cudaFree(h->force_moved); cudaFree(h->force_coef); cudaFree(h->force32);
    cudaFree(h->src_mask);
    cudaFree(h->staging);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_side) cudaEventDestroy(h->ev_side);
    if (h->ev_main) cudaEventDestroy(h->ev_main);
    if (h->ev_bulk) cudaEventDestroy(h->ev_bulk);
    if (h->bulk) { cudaStreamSynchronize(h->bulk); cudaStreamDestroy(h->bulk); }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side) cudaStreamDestroy(h->side);
    delete h;
    return LBM_OK;
}

LBM_API int lbm_sync(lbm_handle h)
{
    int rc = use(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->side));
    CK(cudaStreamSynchronize(h->bulk));
    if (h->halo.edge) CK(cudaStreamSynchronize(h->halo.edge));
    CK(cudaStreamSynchronize(h->stream));
    if (h->halo.connected) {
        unsigned long long lost = 0;
        CK(cudaMemcpy(&lost, h->halo.flags + 2, sizeof lost, cudaMemcpyDeviceToHost));
        if (lost) return fail(LBM_ERR_STATE, "%llu halo handshake(s) timed out: a neighbouring slab did not take the same steps", lost);
    }
    return LBM_OK;
}

LBM_API int lbm_get_stats(lbm_handle h, lbm_stats* o)
{
    int rc = use(h);
    if (rc) return rc;
    if (!o) return fail(LBM_ERR_INVALID, "null argument");
    const Geom& g = h->g;
    const int wcount = (g.nz + 31) / 32;
    std::vector<uint8_t> cls(h->cls_bytes);
    CK(cudaMemcpy(cls.data(), h->cls, h->cls_bytes, cudaMemcpyDeviceToHost));
    int64_t dead = 0, special = 0;
    for (long long r = 0; r < (long long)g.nx * wcount; ++r)
        for (int y = 0; y < g.ny; ++y) {
            const uint8_t c = cls[(size_t)r * h->cls_pitch + y];
            dead += c == 1;
            special += c == 2;
        }
    o->words_total = (int64_t)g.nx * g.ny * wcount;
    o->words_dead = dead;
    o->words_special = special;
    o->links = (int64_t)h->nlinks;
    o->kernel_launches = h->launches;
    return LBM_OK;
}

LBM_API int lbm_get_layout(lbm_handle h, lbm_layout* o)
{
    if (!h || !o) return fail(LBM_ERR_INVALID, "null argument");
    o->nx = h->g.nx; o->ny = h->g.ny; o->nz = h->g.nz; o->q = h->cfg.q;
    o->xp = h->g.xp; o->yp = h->g.yp; o->zp = h->g.zp; o->pad_z = PAD_Z;
    o->plane = h->g.plane; o->pop_stride = h->g.np;
    o->state_bytes = (int64_t)h->state_bytes();
    o->n_buffers = h->inplace ? 1 : 2;
    return LBM_OK;
}

// ---- transfers --------------------------------------------------------------------------------------

LBM_API int lbm_upload_aos(lbm_handle h, const void* cells, size_t count)
{
    int rc = use(h);
    if (rc) return rc;
    if ((rc = inplace_only(h, "lbm_upload_aos"))) return rc;
    const Geom& g = h->g;
    const size_t n = (size_t)g.nx * g.ny * g.nz;
    if (!cells || count != n) return fail(LBM_ERR_INVALID, "upload_aos: expected %zu records, got %zu", n, count);
    if ((rc = ensure_staging(h))) return rc;
    if (!h->solid_bytes) CK(cudaMalloc(&h->solid_bytes, n + 2 * (size_t)g.ny * g.nz));
    const size_t per_plane = (size_t)g.ny * g.nz * AOS_BYTES;
    if (per_plane > STAGING_BYTES) return fail(LBM_ERR_UNSUPPORTED, "one x-plane of AoS records exceeds the staging buffer");
    const int chunk = (int)std::max<size_t>(1, STAGING_BYTES / per_plane);
    for (int x0 = 0; x0 < g.nx; x0 += chunk) {
        const int nxc = std::min(chunk, g.nx - x0);
        CK(cudaMemcpyAsync(h->staging, (const char*)cells + (size_t)x0 * per_plane, (size_t)nxc * per_plane,
                           cudaMemcpyHostToDevice, h->stream));
        CK(launch_unpack_aos(h->staging, h->buf[h->cur], h->solid_bytes, g, h->cfg.q, x0, nxc, h->stream));
        CK(cudaStreamSynchronize(h->stream));   // staging is reused by the next chunk
    }
    h->dead_dirty = 2;
    return rebuild_mask(h, h->solid_bytes, 0);
}

LBM_API int lbm_download_aos(lbm_handle h, void* cells, size_t count)
{
    int rc = use(h);
    if (rc) return rc;
    if ((rc = inplace_only(h, "lbm_download_aos"))) return rc;
    const Geom& g = h->g;
    const size_t n = (size_t)g.nx * g.ny * g.nz;
    if (!cells || count != n) return fail(LBM_ERR_INVALID, "download_aos: expected %zu records, got %zu", n, count);
    if ((rc = ensure_staging(h))) return rc;
    const size_t per_plane = (size_t)g.ny * g.nz * AOS_BYTES;
    if (per_plane > STAGING_BYTES) return fail(LBM_ERR_UNSUPPORTED, "one x-plane of AoS records exceeds the staging buffer");
    const int chunk = (int)std::max<size_t>(1, STAGING_BYTES / per_plane);
    for (int x0 = 0; x0 < g.nx; x0 += chunk) {
        const int nxc = std::min(chunk, g.nx - x0);
        CK(cudaMemsetAsync(h->staging, 0, (size_t)nxc * per_plane, h->stream));   // pointer fields come back null
        CK(launch_pack_aos(h->staging, h->buf[h->cur], h->solid, h->reflected, g, h->cfg.q, x0, nxc, h->stream));
        CK(cudaMemcpyAsync((char*)cells + (size_t)x0 * per_plane, h->staging, (size_t)nxc * per_plane,
                           cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return LBM_OK;
}

// dense [nx][ny][nz] block of one population <-> its padded device array, as one strided DMA transfer
static cudaError_t copy_population(lbm_handle h, float* dev_pop, float* host_pop, bool to_device, cudaStream_t s)
{
    const Geom& g = h->g;
    cudaMemcpy3DParms p;
    std::memset(&p, 0, sizeof p);
    cudaPitchedPtr host = make_cudaPitchedPtr(host_pop, (size_t)g.nz * sizeof(float), (size_t)g.nz * sizeof(float), (size_t)g.ny);
    cudaPitchedPtr dev = make_cudaPitchedPtr(dev_pop, (size_t)g.zp * sizeof(float), (size_t)g.zp * sizeof(float), (size_t)g.yp);
    const cudaPos interior = make_cudaPos((size_t)PAD_Z * sizeof(float), 1, 1);
    p.srcPtr = to_device ? host : dev;
    p.dstPtr = to_device ? dev : host;
    p.srcPos = to_device ? make_cudaPos(0, 0, 0) : interior;
    p.dstPos = to_device ? interior : make_cudaPos(0, 0, 0);
    p.extent = make_cudaExtent((size_t)g.nz * sizeof(float), (size_t)g.ny, (size_t)g.nx);
    p.kind = cudaMemcpyDefault;      // the caller's buffer may be host (pageable or pinned) or device memory
    return cudaMemcpy3DAsync(&p, s);
}

LBM_API int lbm_upload_f_soa(lbm_handle h, const float* f)
{
    int rc = use(h);
    if (rc) return rc;
    if (!f) return fail(LBM_ERR_INVALID, "null buffer");
    const Geom& g = h->g;
    CK(cudaStreamSynchronize(h->side));
    // one strided host-to-device transfer per population, straight into the padded layout (no staging, no
    // scatter kernel); the in-place engine parks population i in the slot of its opposite (lbm_inplace.cu)
    for (int q = 0; q < h->cfg.q; ++q) {
        float* src = const_cast<float*>(f) + (size_t)q * g.nx * g.ny * g.nz;
        float* dst = h->inplace ? h->buf[0] + (size_t)dir_inv(q) * g.np : h->buf[h->cur] + (size_t)q * g.np;
        CK(copy_population(h, dst, src, true, h->stream));
    }
    if (h->inplace && h->aa_parity == 1)      // momentum summed ahead of time for an even step that the replaced state will never take
        CK(cudaMemsetAsync(h->force + 3 * FORCE_SLOTS, 0, 3 * FORCE_SLOTS * sizeof(double), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->dead_dirty = 2;
    h->aa_parity = 0;
    h->aa_dirty = true;
    return LBM_OK;
}

LBM_API int lbm_download_f_soa(lbm_handle h, float* f)
{
    int rc = use(h);
    if (rc) return rc;
    if (!f) return fail(LBM_ERR_INVALID, "null buffer");
    const Geom& g = h->g;
    CK(cudaStreamSynchronize(h->side));
    if (h->halo.edge) CK(cudaStreamSynchronize(h->halo.edge));
    if (!h->inplace) {
        for (int q = 0; q < h->cfg.q; ++q)
            CK(copy_population(h, h->buf[h->cur] + (size_t)q * g.np, f + (size_t)q * g.nx * g.ny * g.nz, false, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return LBM_OK;
    }
    // single-buffer engine: the state is translated from the storage form through a staging buffer
    if ((rc = ensure_staging(h))) return rc;
    const size_t per_plane = (size_t)g.ny * g.nz * sizeof(float);
    if (per_plane > STAGING_BYTES) return fail(LBM_ERR_UNSUPPORTED, "one x-plane exceeds the staging buffer");
    const int chunk = (int)std::max<size_t>(1, STAGING_BYTES / per_plane);
    for (int q = 0; q < h->cfg.q; ++q) {
        float* dst = f + (size_t)q * g.nx * g.ny * g.nz;
        for (int x0 = 0; x0 < g.nx; x0 += chunk) {
            const int nxc = std::min(chunk, g.nx - x0);
            if (h->inplace) {
                InplaceParams ip;
                fill_inplace(h, ip);
                CK(launch_inplace_gather(ip, h->cfg.q, (float*)h->staging, q, x0, nxc, h->stream));
            } else {
                CK(launch_gather_dense((float*)h->staging, h->buf[h->cur] + (size_t)q * g.np, g, x0, nxc, h->stream));
            }
            CK(cudaMemcpyAsync(dst + (size_t)x0 * g.ny * g.nz, h->staging, (size_t)nxc * per_plane, cudaMemcpyDefault, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
    }
    return LBM_OK;
}

LBM_API int lbm_set_solid_mask(lbm_handle h, const uint8_t* solid, int x_ghost_planes)
{
    int rc = use(h);
    if (rc) return rc;
    if (h->inplace) {
        // obstacles on the single-buffer engine are half-way bounce-back; how a population is stored after an odd
        // number of steps depends on the mask, so the mask may only change in the parked form
        if (solid && h->cfg.boundary_mode != LBM_BOUNDARY_BOUNCE_BACK)
            return fail(LBM_ERR_UNSUPPORTED, "the in-place engine takes a solid mask in bounce-back mode only (table reflection runs on the two-buffer engine)");
        if (!solid && !h->solid) return LBM_OK;
        if (h->aa_parity != 0)
            return fail(LBM_ERR_STATE, "in-place engine: set the solid mask before stepping or after an even number of steps");
        h->aa_dirty = true;
    }
    const Geom& g = h->g;
    const size_t per = (size_t)g.ny * g.nz;
    const size_t n = per * (size_t)g.nx;
    if (!h->solid_bytes) CK(cudaMalloc(&h->solid_bytes, n + 2 * per));
    if (!solid) {
        CK(cudaMemsetAsync(h->solid_bytes, 0, n + 2 * per, h->stream));
        return rebuild_mask(h, h->solid_bytes, 1);
    }
    const size_t bytes = x_ghost_planes ? n + 2 * per : n;
    CK(cudaMemcpyAsync(h->solid_bytes, solid, bytes, cudaMemcpyHostToDevice, h->stream));
    return rebuild_mask(h, h->solid_bytes, x_ghost_planes ? 1 : 0);
}

LBM_API int lbm_get_solid_mask(lbm_handle h, uint8_t* solid)
{
    int rc = use(h);
    if (rc) return rc;
    if (!solid) return fail(LBM_ERR_INVALID, "null buffer");
    const Geom& g = h->g;
    const size_t n = (size_t)g.nx * g.ny * g.nz;
    if (!h->solid) { std::memset(solid, 0, n); return LBM_OK; }
    if (!h->solid_bytes) CK(cudaMalloc(&h->solid_bytes, n + 2 * (size_t)g.ny * g.nz));
    CK(launch_mask_to_bytes(h->solid, h->solid_bytes, g, h->stream));
    CK(cudaMemcpyAsync(solid, h->solid_bytes, n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

// Turns the reference's per-solid-cell scatter tables into per-fluid-cell gather lists.  `ghost` = the cell indices
// count nx + 2 planes (the neighbouring slabs' edge planes first and last, like lbm_set_solid_mask with ghost planes):
// a table of a ghost-plane cell feeds the cells of this slab it reflects into, a table of an edge cell feeds only the
// targets inside this slab -- the neighbour holds the same table for the others.
static int set_reflection_table(lbm_handle h, const void* tables, const int64_t* cell_index, int n, bool ghost)
{
    int rc = use(h);
    if (rc) return rc;
    if ((rc = inplace_only(h, "lbm_set_reflection_table"))) return rc;
    if (n < 0 || (n > 0 && (!tables || !cell_index))) return fail(LBM_ERR_INVALID, "bad reflection table arguments");
    if (h->cfg.bc[0] == LBM_BC_HALO && n > 0 && !ghost)
        return fail(LBM_ERR_UNSUPPORTED, "slab handles take their tables with the neighbours' edge planes: lbm_set_reflection_table_slab");
    const Geom& g = h->g;
    const int xlo = ghost ? -1 : 0;
    const long long ncell = (long long)(g.nx + (ghost ? 2 : 0)) * g.ny * g.nz;
    struct HostLink { uint32_t tgt; Link l; long long order; };
    std::vector<HostLink> hl;
    const size_t words = h->mask_words();
    std::vector<uint32_t> refl(words, 0u);
    std::vector<int> interior;              // tables of cells inside the slab (what lbm_get_reflection_tables reports)
    for (int t = 0; t < n; ++t) {
        const long long k = cell_index[t];
        if (k < 0 || k >= ncell) return fail(LBM_ERR_INVALID, "reflection table %d: cell index out of range", t);
        const int sx = (int)(k / ((long long)g.ny * g.nz)) + xlo;
        const int rem = (int)(k % ((long long)g.ny * g.nz));
        const int sy = rem / g.nz, sz = rem % g.nz;
        const unsigned char* rec = (const unsigned char*)tables + (size_t)t * TABLE_BYTES;
        int32_t dirs[14];
        float w[81];
        std::memcpy(dirs, rec, sizeof dirs);
        std::memcpy(w, rec + sizeof dirs, sizeof w);
        const long long scell = g.cell(sx, sy, sz);
        if (sx >= 0 && sx < g.nx) {
            refl[(size_t)(scell >> 5)] |= 1u << (scell & 31);
            interior.push_back(t);
        }
        for (int s = 1; s < 81; ++s) {      // slot 0 is never read (SimCalcFuncs.cu:164)
            const int d = (int)(((uint32_t)dirs[s / 6] >> (5 * (s % 6))) & 0x1Fu);   // SimCalcFuncs.cu:18-30
            if (d == 0) continue;
            if (d >= 27) return fail(LBM_ERR_INVALID, "reflection table %d slot %d: direction %d", t, s, d);
            const int j = s / 3;
            if (d >= h->cfg.q || j >= h->cfg.q) continue;
            const int tx = sx + dir_c(d, 0), ty = sy + dir_c(d, 1), tz = sz + dir_c(d, 2);
            if (tx < 0 || ty < 0 || tz < 0 || tx >= g.nx || ty >= g.ny || tz >= g.nz) continue;   // SimCalcFuncs.cu:175 (slabs: the neighbour's target)
            HostLink e;
            e.tgt = (uint32_t)g.cell(tx, ty, tz);
            e.l.src_cell = (uint32_t)scell;
            e.l.j = (uint8_t)j; e.l.d = (uint8_t)d; e.l.slot = (uint8_t)s; e.l.pad = 0;
            e.l.w = w[s];
            e.order = k * 81 + s;
            hl.push_back(e);
        }
    }
    h->host_tables.clear();
    h->host_table_cells.clear();
    for (int t : interior) {
        const unsigned char* rec = (const unsigned char*)tables + (size_t)t * TABLE_BYTES;
        h->host_tables.insert(h->host_tables.end(), rec, rec + TABLE_BYTES);
        h->host_table_cells.push_back(cell_index[t] - (ghost ? (int64_t)g.ny * g.nz : 0));      // slab-local host index
    }
    std::sort(hl.begin(), hl.end(), [](const HostLink& a, const HostLink& b) {
        if (a.tgt != b.tgt) return a.tgt < b.tgt;
        if (a.l.d != b.l.d) return a.l.d < b.l.d;      // the kernel walks a target's links population by population
        return a.order < b.order;
    });
    std::vector<uint32_t> tmask(words, 0u), tprefix(words, 0u), toff;
    std::vector<Link> links(hl.size());
    for (size_t i = 0; i < hl.size(); ++i) {
        links[i] = hl[i].l;
        if (i == 0 || hl[i].tgt != hl[i - 1].tgt) {
            toff.push_back((uint32_t)i);
            tmask[hl[i].tgt >> 5] |= 1u << (hl[i].tgt & 31);
        }
    }
    toff.push_back((uint32_t)hl.size());
    // LBM_FORCE_ACCUM_REFERENCE_FP32: the terms of the force sum in the oracle's order (solid cells in host order, slots
    // ascending = ascending `order`), each with its direction coefficients (SimCalcFuncs.cu:181-183, expression shapes kept)
    std::vector<uint32_t> rank(hl.size(), 0xffffffffu);
    std::vector<float> coef;
    if (h->force32) {
        std::vector<size_t> by_order;
        for (size_t i = 0; i < hl.size(); ++i) {
            const int in_dir = (h->cfg.force_mode == LBM_FORCE_REFERENCE_SLOT_INDEX) ? (int)hl[i].l.slot : (int)hl[i].l.j;
            if (in_dir < 27) by_order.push_back(i);
        }
        std::sort(by_order.begin(), by_order.end(), [&](size_t a, size_t b) { return hl[a].order < hl[b].order; });
        coef.resize(3 * by_order.size());
        for (size_t r = 0; r < by_order.size(); ++r) {
            const Link& L = hl[by_order[r]].l;
            const int d = L.d, in_dir = (h->cfg.force_mode == LBM_FORCE_REFERENCE_SLOT_INDEX) ? (int)L.slot : (int)L.j;
            rank[by_order[r]] = (uint32_t)r;
            for (int a = 0; a < 3; ++a)
                coef[3 * r + a] = ((float)dir_c(d, a) * dir_split(d)) - ((float)dir_c(in_dir, a) * dir_split(in_dir));
        }
    }
    uint32_t run = 0;
    for (size_t wi = 0; wi < words; ++wi) { tprefix[wi] = run; run += (uint32_t)__builtin_popcount(tmask[wi]); }

    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->tmask); cudaFree(h->tprefix); cudaFree(h->toff); cudaFree(h->links); cudaFree(h->reflected);
    cudaFree(h->link_rank); cudaFree(h->force_moved); cudaFree(h->force_coef);
    h->tmask = h->tprefix = h->toff = h->reflected = nullptr;
    h->link_rank = nullptr; h->force_moved = h->force_coef = nullptr;
    h->n_force_terms = 0;
    h->links = nullptr;
    h->nlinks = 0;
    CK(cudaMalloc(&h->reflected, words * sizeof(uint32_t)));
    // uploads on the handle's stream, one synchronisation before the host vectors go away (see rebuild_special)
    CK(cudaMemcpyAsync(h->reflected, refl.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    if (!links.empty()) {
        CK(cudaMalloc(&h->tmask, words * sizeof(uint32_t)));
        CK(cudaMalloc(&h->tprefix, words * sizeof(uint32_t)));
        CK(cudaMalloc(&h->toff, toff.size() * sizeof(uint32_t)));
        CK(cudaMalloc(&h->links, links.size() * sizeof(Link)));
        CK(cudaMemcpyAsync(h->tmask, tmask.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->tprefix, tprefix.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->toff, toff.data(), toff.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->links, links.data(), links.size() * sizeof(Link), cudaMemcpyHostToDevice, h->stream));
        h->nlinks = links.size();
        if (h->force32 && !coef.empty()) {
            h->n_force_terms = coef.size() / 3;
            CK(cudaMalloc(&h->link_rank, rank.size() * sizeof(uint32_t)));
            CK(cudaMalloc(&h->force_coef, coef.size() * sizeof(float)));
            CK(cudaMalloc(&h->force_moved, h->n_force_terms * sizeof(float)));
            CK(cudaMemcpyAsync(h->link_rank, rank.data(), rank.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(h->force_coef, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemsetAsync(h->force_moved, 0, h->n_force_terms * sizeof(float), h->stream));
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    return rebuild_special(h);
}

LBM_API int lbm_set_reflection_table(lbm_handle h, const void* tables, const int64_t* cell_index, int n)
{
    return set_reflection_table(h, tables, cell_index, n, false);
}

LBM_API int lbm_set_reflection_table_slab(lbm_handle h, const void* tables, const int64_t* cell_index, int n)
{
    return set_reflection_table(h, tables, cell_index, n, true);
}

LBM_API int lbm_set_slab_origin(lbm_handle h, int x0, int gx)
{
    if (!h) return fail(LBM_ERR_INVALID, "null handle");
    if (x0 < 0 || gx < x0 + h->g.nx) return fail(LBM_ERR_INVALID, "lbm_set_slab_origin: the slab [%d, %d) does not fit a domain of %d planes", x0, x0 + h->g.nx, gx);
    h->slab_x0 = x0;
    h->slab_gx = gx;
    return LBM_OK;
}

LBM_API int lbm_set_source(lbm_handle h, int kind, const int32_t lo[3], const int32_t hi[3], const float f[27])
{
    if (!h) return fail(LBM_ERR_INVALID, "null handle");
    if (kind == LBM_SOURCE_NONE) { h->src_on = 0; h->src_masked = false; return LBM_OK; }
    if (kind != LBM_SOURCE_BOX || !lo || !hi || !f) return fail(LBM_ERR_INVALID, "bad source arguments");
    h->src_on = 1;
    h->src_masked = false;
    for (int a = 0; a < 3; ++a) { h->src_lo[a] = lo[a]; h->src_hi[a] = hi[a]; }
    for (int i = 0; i < 27; ++i) h->src_f[i] = (i < h->cfg.q) ? f[i] : 0.0f;
    return LBM_OK;
}

LBM_API int lbm_set_source_mask(lbm_handle h, const uint8_t* cells, const float f[27])
{
    int rc = use(h);
    if (rc) return rc;
    if (!cells || !f) return fail(LBM_ERR_INVALID, "bad source arguments");
    const Geom& g = h->g;
    const size_t per = (size_t)g.ny * g.nz, n = per * (size_t)g.nx;
    // the box around the selected cells keeps every launch that cannot contain one on the lean kernel variant
    int lo[3] = {g.nx, g.ny, g.nz}, hi[3] = {-1, -1, -1};
    for (int x = 0; x < g.nx; ++x)
        for (int y = 0; y < g.ny; ++y) {
            const uint8_t* row = cells + (size_t)x * per + (size_t)y * g.nz;
            for (int z = 0; z < g.nz; ++z)
                if (row[z]) {
                    const int c[3] = {x, y, z};
                    for (int a = 0; a < 3; ++a) { if (c[a] < lo[a]) lo[a] = c[a]; if (c[a] > hi[a]) hi[a] = c[a]; }
                }
        }
    if (hi[0] < 0) { h->src_on = 0; h->src_masked = false; return LBM_OK; }      // the predicate selects nothing
    if (!h->solid_bytes) CK(cudaMalloc(&h->solid_bytes, n + 2 * per));
    if (!h->src_mask) CK(cudaMalloc(&h->src_mask, h->mask_words() * sizeof(uint32_t)));
    CK(cudaMemcpyAsync(h->solid_bytes, cells, n, cudaMemcpyHostToDevice, h->stream));
    const int open_edges[3] = {LBM_BC_ZERO_INFLOW, LBM_BC_ZERO_INFLOW, LBM_BC_ZERO_INFLOW};      // no images in the ghost cells
    CK(launch_build_mask(h->solid_bytes, 0, h->src_mask, g, open_edges, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->src_on = 1;
    h->src_masked = true;
    for (int a = 0; a < 3; ++a) { h->src_lo[a] = lo[a]; h->src_hi[a] = hi[a]; }
    for (int i = 0; i < 27; ++i) h->src_f[i] = (i < h->cfg.q) ? f[i] : 0.0f;
    return LBM_OK;
}

LBM_API int lbm_set_force_accumulator(lbm_handle h, int enable, const float init[3])
{
    if (!init) return lbm_set_force_accumulator_f64(h, enable, nullptr);
    const double v[3] = {init[0], init[1], init[2]};
    return lbm_set_force_accumulator_f64(h, enable, v);
}

LBM_API int lbm_set_force_accumulator_f64(lbm_handle h, int enable, const double init[3])
{
    int rc = use(h);
    if (rc) return rc;
    h->force_on = enable != 0;
    double v[2 * 3 * FORCE_SLOTS] = {0.0};
    if (init) for (int a = 0; a < 3; ++a) v[a] = init[a];
    CK(cudaMemcpyAsync(h->force, v, sizeof v, cudaMemcpyHostToDevice, h->stream));
    if (h->force32) {
        const float v32[3] = {init ? (float)init[0] : 0.0f, init ? (float)init[1] : 0.0f, init ? (float)init[2] : 0.0f};
        CK(cudaMemcpyAsync(h->force32, v32, sizeof v32, cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

LBM_API int lbm_get_force_f64(lbm_handle h, double out[3])
{
    int rc = use(h);
    if (rc) return rc;
    if (!out) return fail(LBM_ERR_INVALID, "null buffer");
    if (h->force32) {
        float v32[3];
        CK(cudaMemcpyAsync(v32, h->force32, sizeof v32, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int a = 0; a < 3; ++a) out[a] = v32[a];
        return LBM_OK;
    }
    double v[2 * 3 * FORCE_SLOTS];
    if (h->halo.edge) CK(cudaStreamSynchronize(h->halo.edge));
    CK(cudaMemcpyAsync(v, h->force, sizeof v, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // single-buffer engine: the sums made ahead of time count once the even step they belong to has run
    const int sets = (h->inplace && h->aa_parity == 0) ? 2 : 1;
    for (int a = 0; a < 3; ++a) {
        double t = 0.0;
        for (int k = 0; k < sets * FORCE_SLOTS; ++k) t += v[3 * k + a];
        out[a] = t;
    }
    return LBM_OK;
}

LBM_API int lbm_get_force(lbm_handle h, float out[3])
{
    if (!out) return fail(LBM_ERR_INVALID, "null buffer");
    double v[3];
    int rc = lbm_get_force_f64(h, v);
    if (rc) return rc;
    for (int a = 0; a < 3; ++a) out[a] = (float)v[a];
    return LBM_OK;
}

// ---- stepping -----------------------------------------------------------------------------------------

LBM_API int lbm_prime_points(lbm_handle h) { return use(h); }

LBM_API int lbm_step_phases(lbm_handle h, int n, unsigned phases)
{
    int rc = use(h);
    if (rc) return rc;
    if (n < 0 || (phases & ~15u) || phases == 0) return fail(LBM_ERR_INVALID, "bad step arguments");
#if defined(__CUDACC__)
    // LBM_STEP_GRAPH=1: for long runs two consecutive steps (the buffers / the storage form swap with the step parity) are
    // captured once into a CUDA graph and replayed -- tried for BASELINE configs[0] (64^3, three launches per step) and
    // measured on B200 to change nothing (0.100-0.102 ms/step with the graph, 0.090-0.106 without,
    // profiles/r02_notes.md): that lattice is too small to fill the GPU, it is not waiting for launches.  Off by default.
    // Only in the steady state: no slab neighbours (their handshake spans processes), solid words settled, nothing pending.
    static const int graph_mode = [] { const char* e = getenv("LBM_STEP_GRAPH"); return e ? atoi(e) : 0; }();
    const bool small = false;
    if (graph_mode != 0 && (small || graph_mode == 1) && n >= 16 && !h->halo.connected && !h->pending_commit && !h->force32) {
        int done = 0;
        while (done < 4 && done < n) {            // settle dead words / first-step work outside the graph
            if ((rc = one_step(h, phases))) return rc;
            ++done;
        }
        const int pairs = (n - done) / 2;
        if (pairs >= 4 && h->dead_dirty == 0 && !h->aa_dirty) {
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            const long long launches0 = h->launches;
            if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                rc = one_step(h, phases);
                if (!rc) rc = one_step(h, phases);
                const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                if (!rc && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                    const long long per_pair = h->launches - launches0;
                    h->launches = launches0;          // (captured, not run)
                    cudaError_t le = cudaSuccess;
                    for (int k = 0; k < pairs && le == cudaSuccess; ++k) le = cudaGraphLaunch(exec, h->stream);
                    cudaGraphExecDestroy(exec);
                    cudaGraphDestroy(graph);
                    if (le != cudaSuccess) return fail(LBM_ERR_CUDA, "graph launch: %s", cudaGetErrorString(le));
                    h->launches += per_pair * pairs;
                    done += 2 * pairs;
                } else {
                    if (graph) cudaGraphDestroy(graph);
                    cudaGetLastError();
                    if (rc) return rc;
                    // nothing ran: the capture only recorded the two steps, and the host-side state (buffer roles, storage
                    // form) flipped twice, i.e. is unchanged -- take the two steps for real
                    h->launches = launches0;
                    if ((rc = one_step(h, phases))) return rc;
                    if ((rc = one_step(h, phases))) return rc;
                    done += 2;
                }
            }
        }
        for (; done < n; ++done)
            if ((rc = one_step(h, phases))) return rc;
        return LBM_OK;
    }
#endif
    for (int s = 0; s < n; ++s)
        if ((rc = one_step(h, phases))) return rc;
    return LBM_OK;
}

LBM_API int lbm_step(lbm_handle h, int n) { return lbm_step_phases(h, n, LBM_PHASE_ALL); }
LBM_API int lbm_stream(lbm_handle h) { return lbm_step_phases(h, 1, LBM_PHASE_STREAM); }
LBM_API int lbm_reflect(lbm_handle h) { return lbm_step_phases(h, 1, LBM_PHASE_REFLECT); }
LBM_API int lbm_collide(lbm_handle h) { return lbm_step_phases(h, 1, LBM_PHASE_COLLIDE); }
LBM_API int lbm_source(lbm_handle h) { return lbm_step_phases(h, 1, LBM_PHASE_SOURCE); }

LBM_API int lbm_collide_tau(lbm_handle h, float timescale)
{
    return lbm_group_collide_tau(&h, 1, timescale);
}

LBM_API int lbm_reflect_accumulate(lbm_handle h, float* xyz)
{
    return lbm_group_reflect_accumulate(&h, 1, xyz);
}

// one reflect phase on every slab of a group (or on one handle) whose force is added to the caller's three floats
LBM_API int lbm_group_reflect_accumulate(lbm_handle* slabs, int count, float* xyz)
{
    if (!slabs || count < 1) return fail(LBM_ERR_INVALID, "bad arguments");
    for (int k = 0; k < count; ++k) {
        if (!slabs[k]) return fail(LBM_ERR_INVALID, "null handle");
        if (slabs[k]->force32) return fail(LBM_ERR_UNSUPPORTED, "lbm_reflect_accumulate: not with LBM_FORCE_ACCUM_REFERENCE_FP32 (use the handle's accumulator)");
    }
    auto step = [&]() { return count == 1 ? lbm_step_phases(slabs[0], 1, LBM_PHASE_REFLECT) : lbm_group_step(slabs, count, 1, LBM_PHASE_REFLECT); };
    if (!xyz) return step();
    const size_t words = 2 * 3 * FORCE_SLOTS;
    std::vector<std::vector<double>> before(count, std::vector<double>(words)), after(count, std::vector<double>(words));
    std::vector<char> was_on(count);
    int rc = LBM_OK;
    for (int k = 0; k < count && rc == LBM_OK; ++k) {
        lbm_handle h = slabs[k];
        if ((rc = use(h))) break;
        CK(cudaMemcpyAsync(before[k].data(), h->force, words * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        was_on[k] = h->force_on;
        h->force_on = true;
    }
    if (rc == LBM_OK) rc = step();
    double add[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < count; ++k) {
        lbm_handle h = slabs[k];
        h->force_on = was_on[k] != 0;
        if (rc != LBM_OK) continue;
        if ((rc = use(h))) continue;
        if (h->halo.edge) CK(cudaStreamSynchronize(h->halo.edge));
        CK(cudaMemcpyAsync(after[k].data(), h->force, words * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int a = 0; a < 3; ++a)
            for (int s = 0; s < FORCE_SLOTS; ++s) add[a] += after[k][3 * s + a] - before[k][3 * s + a];
        if (!was_on[k])      // the handle's own accumulator is off: it keeps what it held
            CK(cudaMemcpyAsync(h->force, before[k].data(), words * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (rc != LBM_OK) return rc;
    // the caller's triple may live in device memory (the reference's d_reflectionData) or in host memory
    float cur[3];
    CK(cudaMemcpy(cur, xyz, sizeof cur, cudaMemcpyDefault));
    for (int a = 0; a < 3; ++a) cur[a] += (float)add[a];
    CK(cudaMemcpy(xyz, cur, sizeof cur, cudaMemcpyDefault));
    return LBM_OK;
}

LBM_API int lbm_group_collide_tau(lbm_handle* slabs, int count, float timescale)
{
    if (!slabs || count < 1) return fail(LBM_ERR_INVALID, "bad arguments");
    if (!(timescale > 0.0f)) return fail(LBM_ERR_INVALID, "the timescale must be positive");
    std::vector<float> keep(count);
    for (int k = 0; k < count; ++k) {
        if (!slabs[k]) return fail(LBM_ERR_INVALID, "null handle");
        keep[k] = slabs[k]->cfg.tau;
        slabs[k]->cfg.tau = timescale;
    }
    const int rc = count == 1 ? lbm_step_phases(slabs[0], 1, LBM_PHASE_COLLIDE) : lbm_group_step(slabs, count, 1, LBM_PHASE_COLLIDE);
    for (int k = 0; k < count; ++k) slabs[k]->cfg.tau = keep[k];
    return rc;
}

LBM_API int lbm_device_count(int* count)
{
    if (!count) return fail(LBM_ERR_INVALID, "null argument");
    *count = 0;
    CK(cudaGetDeviceCount(count));
    return LBM_OK;
}

LBM_API int lbm_step_timed(lbm_handle h, int n, unsigned phases, float* ms, int* launches)
{
    int rc = use(h);
    if (rc) return rc;
    if (!ms) return fail(LBM_ERR_INVALID, "null ms");
    const long long l0 = h->launches;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    rc = lbm_step_phases(h, n, phases);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    if (launches) *launches = (int)(h->launches - l0);
    return LBM_OK;
}

// ---- diagnostics ------------------------------------------------------------------------------------------

LBM_API int lbm_moments(lbm_handle h, float* rho, float* ux, float* uy, float* uz)
{
    int rc = use(h);
    if (rc) return rc;
    const Geom& g = h->g;
    const size_t n = (size_t)g.nx * g.ny * g.nz;
    if (!h->moments_scratch) CK(cudaMalloc(&h->moments_scratch, 4 * n * sizeof(float)));   // kept for the next call
    float* dev = h->moments_scratch;
    if (h->halo.edge) CK(cudaStreamSynchronize(h->halo.edge));
    cudaError_t e;
    if (h->inplace) {
        InplaceParams ip;
        fill_inplace(h, ip);
        e = launch_inplace_moments(ip, h->cfg.q, h->cfg.moment_mode, dev, dev + n, dev + 2 * n, dev + 3 * n, h->stream);
    } else {
        e = launch_moments(h->buf[h->cur], g, h->cfg.q, h->cfg.moment_mode, dev, dev + n, dev + 2 * n, dev + 3 * n, h->stream);
    }
    float* outs[4] = {rho, ux, uy, uz};
    for (int k = 0; k < 4 && e == cudaSuccess; ++k)
        if (outs[k]) e = cudaMemcpyAsync(outs[k], dev + (size_t)k * n, n * sizeof(float), cudaMemcpyDefault, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "moments: %s", cudaGetErrorString(e));
    return LBM_OK;
}

static int mass_momentum(lbm_handle h, double out4[4])
{
    int rc = use(h);
    if (rc) return rc;
    CK(cudaMemsetAsync(h->diag, 0, 4 * sizeof(double), h->stream));
    if (h->inplace) {
        InplaceParams ip;
        fill_inplace(h, ip);
        CK(launch_inplace_mass(ip, h->cfg.q, h->diag, h->stream));
    } else {
        CK(launch_mass_momentum(h->buf[h->cur], h->g, h->cfg.q, h->diag, h->stream));
    }
    CK(cudaMemcpyAsync(out4, h->diag, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

LBM_API int lbm_mass(lbm_handle h, double* mass)
{
    if (!mass) return fail(LBM_ERR_INVALID, "null buffer");
    double v[4];
    int rc = mass_momentum(h, v);
    if (rc) return rc;
    *mass = v[0];
    return LBM_OK;
}

LBM_API int lbm_momentum(lbm_handle h, double out[3])
{
    if (!out) return fail(LBM_ERR_INVALID, "null buffer");
    double v[4];
    int rc = mass_momentum(h, v);
    if (rc) return rc;
    out[0] = v[1]; out[1] = v[2]; out[2] = v[3];
    return LBM_OK;
}

// ---- slab decomposition ------------------------------------------------------------------------------------

static int plane_ptr(lbm_handle h, float* base, int q, int xp, void** dev_ptr, size_t* bytes)
{
    if (!h || !dev_ptr) return fail(LBM_ERR_INVALID, "null argument");
    if (q < 0 || q >= h->cfg.q || xp < 0 || xp >= h->g.xp) return fail(LBM_ERR_INVALID, "plane (%d, %d) out of range", q, xp);
    *dev_ptr = base + (size_t)q * h->g.np + (size_t)xp * h->g.plane;
    if (bytes) *bytes = (size_t)h->g.plane * sizeof(float);
    return LBM_OK;
}

LBM_API int lbm_plane_ptr(lbm_handle h, int q, int xp, void** dev_ptr, size_t* bytes)
{
    if (!h) return fail(LBM_ERR_INVALID, "null handle");
    if (h->inplace) return plane_ptr(h, h->buf[0], q, xp, dev_ptr, bytes);      // one copy: slot q of padded plane xp
    return plane_ptr(h, h->buf[h->cur], q, xp, dev_ptr, bytes);
}

LBM_API int lbm_next_plane_ptr(lbm_handle h, int q, int xp, void** dev_ptr, size_t* bytes)
{
    if (!h) return fail(LBM_ERR_INVALID, "null handle");
    if (h->inplace) return plane_ptr(h, h->buf[0], q, xp, dev_ptr, bytes);      // the update is in place
    return plane_ptr(h, h->buf[h->cur ^ 1], q, xp, dev_ptr, bytes);
}

LBM_API int lbm_face_populations(lbm_handle h, int side, int32_t* qs, int* nq)
{
    if (!h || !qs || !nq) return fail(LBM_ERR_INVALID, "null argument");
    int n = 0;
    for (int i = 0; i < h->cfg.q; ++i)
        if (dir_c(i, 0) == (side ? 1 : -1)) qs[n++] = i;
    *nq = n;
    return LBM_OK;
}

LBM_API int lbm_get_stream(lbm_handle h, void** stream)
{
    if (!h || !stream) return fail(LBM_ERR_INVALID, "null argument");
    *stream = (void*)h->stream;
    return LBM_OK;
}

LBM_API int lbm_set_stream(lbm_handle h, void* stream)
{
    int rc = use(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    if (h->cls) pin_tables_in_l2(h);
    return LBM_OK;
}

LBM_API int lbm_fill_ghosts(lbm_handle h)
{
    int rc = use(h);
    if (rc) return rc;
    return fill_ghosts(h, h->cfg.bc[0] != LBM_BC_HALO);
}

LBM_API int lbm_step_range(lbm_handle h, unsigned phases, int x0, int x1, int on_side_stream)
{
    int rc = use(h);
    if (rc) return rc;
    if ((rc = inplace_only(h, "lbm_step_range"))) return rc;
    if (x0 < 0 || x1 > h->g.nx || x0 > x1 || (phases & ~15u) || phases == 0) return fail(LBM_ERR_INVALID, "bad range");
    if (on_side_stream) {
        // the bulk stream starts after everything queued so far on the main stream
        CK(cudaEventRecord(h->ev_main, h->stream));
        CK(cudaStreamWaitEvent(h->bulk, h->ev_main, 0));
        rc = launch_range(h, phases, x0, x1, h->bulk);
        h->bulk_used = true;
    } else {
        rc = launch_range(h, phases, x0, x1, h->stream);
    }
    if (rc) return rc;
    h->pending_commit = true;
    h->pending_phases = phases;
    return LBM_OK;
}

LBM_API int lbm_inplace_merge_edges(lbm_handle h)
{
    int rc = use(h);
    if (rc) return rc;
    if (!h->inplace || h->cfg.bc[0] != LBM_BC_HALO) return fail(LBM_ERR_STATE, "lbm_inplace_merge_edges is for single-buffer slab handles");
    if (h->aa_parity != 1) return fail(LBM_ERR_STATE, "lbm_inplace_merge_edges follows an odd (pull / push) step");
    InplaceParams ip;
    fill_inplace(h, ip);
    CK(launch_inplace_merge_edges(ip, h->cfg.q, h->stream));
    h->launches += 2;
    return LBM_OK;
}

LBM_API int lbm_step_commit(lbm_handle h)
{
    int rc = use(h);
    if (rc) return rc;
    if (!h->pending_commit) return fail(LBM_ERR_STATE, "nothing to commit");
    // the main stream continues only after the bulk stream's share of this step
    if (h->bulk_used) {
        CK(cudaEventRecord(h->ev_bulk, h->bulk));
        CK(cudaStreamWaitEvent(h->stream, h->ev_bulk, 0));
        h->bulk_used = false;
    }
    if ((rc = force_reference_sum(h, h->pending_phases))) return rc;
    h->cur ^= 1;
    h->pending_commit = false;
    note_step_done(h, h->pending_phases);
    return LBM_OK;
}

}  // extern "C"
