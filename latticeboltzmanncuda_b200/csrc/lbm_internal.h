// lbm_internal.h -- the handle behind lbm_handle, shared by the translation units of liblbm_b200.so.
#pragma once

#include "../../include/lbm_b200.h"
#include "lbm_device.h"

#include <vector>

// x slabs with peer-memory halos (lbm_halo_export / lbm_halo_connect, lbm_halo.cu): what a slab knows about its two
// neighbours.  Side 0 = the slab below (lower x), side 1 = the slab above.
struct lbm_halo_state {
    bool connected = false;
    bool has[2] = {false, false};            // there is a neighbour on that side (a missing one = end of a zero-inflow domain)
    bool ipc[2] = {false, false};            // its memory was opened through CUDA IPC (another process) and must be closed
    float* peer_buf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side][buffer]: the neighbour's population buffers
    unsigned long long* peer_flag[2] = {nullptr, nullptr};   // the word of the NEIGHBOUR's flag block that this slab writes
    int peer_nx[2] = {0, 0};
    int peer_cur_xor[2] = {0, 0};            // two-buffer engine: (neighbour's current buffer) ^ (mine), constant in lockstep
    long long delta_inplace[2] = {0, 0};     // single-buffer engine: element distance ghost plane -> neighbour's edge plane (slot 0)
    long long np_diff[2] = {0, 0};           // neighbour's population stride - mine (slabs may differ in thickness)
    // own flag block (device memory, written by the neighbours through peer stores): [0] the low neighbour's epoch,
    // [1] the high neighbour's epoch, [2] number of handshakes that timed out, [3] spare
    unsigned long long* flags = nullptr;
    unsigned long long epoch = 0;            // handshakes issued so far; all slabs of a group step in lockstep
    cudaStream_t edge = nullptr;             // highest priority: handshake + the edge planes
    cudaEvent_t ev_pre = nullptr, ev_edge = nullptr;
};

struct lbm_lattice_s {
    // (all members are plain data; the C ABI functions in lbm_capi.cu own the logic)
    lbm_config cfg;
    lbm::Geom g;
    int device = 0;
    cudaStream_t stream = nullptr, side = nullptr, own_stream = nullptr, bulk = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_side = nullptr, ev_main = nullptr, ev_bulk = nullptr;
    float* buf[2] = {nullptr, nullptr};
    int cur = 0;
    bool inplace = false;            // single-buffer AA engine: only buf[0] exists
    int aa_parity = 0;
    bool aa_dirty = false;           // in-place engine: the solid cells have not been cleared since the mask / state was set
    bool pending_commit = false;
    bool bulk_used = false;
    unsigned pending_phases = 0;
    uint32_t* solid = nullptr;       // padded bit mask
    uint8_t* cls = nullptr;          // class of every interior 32-cell word, [x][word][y]
    int cls_pitch = 0;
    size_t cls_bytes = 0;
    int dead_dirty = 2;              // steps for which all-solid words must still be written as zeros
    uint32_t* cellinfo = nullptr;    // neighbourhood words of the special words, [n_special][32] in list order
    size_t cellinfo_cap = 0;         // entries allocated
    uint32_t* special_list = nullptr;   // mask-word indices of the special words
    size_t n_special = 0;
    std::vector<uint32_t> special_start;   // [nx + 1]: first list entry of every interior plane (host side)
    int busy_box[6] = {};                  // bounding box of the words that are not ordinary (whole slab until it is known)
    int dead_box[LBM_DEAD_BOXES][6] = {};               // single-buffer engine: boxes of all-solid words (find_dead_boxes), extent 0 = none
    bool any_solid = false;          // some bit set (geometry or wall ghosts)
    uint8_t* solid_bytes = nullptr;  // nx*ny*nz (+2 planes) scratch
    uint32_t *tmask = nullptr, *tprefix = nullptr, *toff = nullptr, *reflected = nullptr;
    lbm::Link* links = nullptr;
    size_t nlinks = 0;
    // LBM_FORCE_ACCUM_REFERENCE_FP32: rank of every link in the oracle's order, the amounts moved this step, the
    // direction coefficients of every term, and the float triple itself
    uint32_t* link_rank = nullptr;
    float* force_moved = nullptr;
    float* force_coef = nullptr;
    float* force32 = nullptr;
    size_t n_force_terms = 0;
    double* force = nullptr;         // FORCE_SLOTS x 3 doubles
    bool force_on = false;
    int src_on = 0;
    int src_lo[3] = {0, 0, 0}, src_hi[3] = {0, 0, 0};
    float src_f[27] = {0};
    uint32_t* src_mask = nullptr;    // lbm_set_source_mask: the cells of the source box that take the source (padded bit mask)
    bool src_masked = false;
    double* diag = nullptr;          // 4 doubles
    unsigned* flag = nullptr;
    void* staging = nullptr;
    long long launches = 0;
    lbm_halo_state halo;
    float* moments_scratch = nullptr;   // 4 n floats, lbm_moments
    int slab_x0 = 0, slab_gx = 0;     // lbm_set_slab_origin: first global plane of this slab and the global x extent (0 = not a slab)
    // host copy of the reflection tables currently set (380-byte records) and their owning cells
    std::vector<unsigned char> host_tables;
    std::vector<int64_t> host_table_cells;

    size_t pop_bytes() const { return (size_t)g.np * sizeof(float); }
    size_t state_bytes() const { return pop_bytes() * (size_t)cfg.q; }
    size_t mask_words() const { return (size_t)g.xp * g.yp * g.wz; }
};


int lbm_internal_fail(int code, const char* msg);
int lbm_internal_use(lbm_handle h);
int lbm_internal_rebuild_mask(lbm_handle h, const uint8_t* dev_bytes, int has_x_ghost);
// one step of a handle in two halves (lbm_capi.cu): a = everything that needs no neighbour (ghost fills, the
// "my previous step is complete" signal to the neighbouring slabs), b = wait for the neighbours, edge planes,
// interior planes, join.  lbm_step_phases runs a then b; lbm_group_step runs a on every slab of a group first.
int lbm_internal_step_a(lbm_handle h, unsigned phases);
int lbm_internal_step_b(lbm_handle h, unsigned phases);
// lbm_halo.cu
int lbm_internal_halo_signal(lbm_handle h);       // epoch + 1 -> the neighbours' flag blocks (edge stream)
int lbm_internal_halo_wait(lbm_handle h);         // until both neighbours have signalled epoch + 1 (edge stream); epoch += 1
void lbm_internal_halo_release(lbm_handle h);
