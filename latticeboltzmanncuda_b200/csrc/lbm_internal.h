// lbm_internal.h -- the handle behind lbm_handle, shared by the translation units of liblbm_b200.so.
#pragma once

#include "../../include/lbm_b200.h"
#include "lbm_device.h"

#include <vector>

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
    bool any_solid = false;          // some bit set (geometry or wall ghosts)
    uint8_t* solid_bytes = nullptr;  // nx*ny*nz (+2 planes) scratch
    uint32_t *tmask = nullptr, *tprefix = nullptr, *toff = nullptr, *reflected = nullptr;
    lbm::Link* links = nullptr;
    size_t nlinks = 0;
    double* force = nullptr;         // 3 doubles
    bool force_on = false;
    int src_on = 0;
    int src_lo[3] = {0, 0, 0}, src_hi[3] = {0, 0, 0};
    float src_f[27] = {0};
    double* diag = nullptr;          // 4 doubles
    unsigned* flag = nullptr;
    void* staging = nullptr;
    long long launches = 0;
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
