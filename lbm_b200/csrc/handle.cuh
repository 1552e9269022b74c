// handle.cuh — the handle behind blbm_t and the few internals shared by api.cu (one slab on one GPU) and
// group.cu (a lattice split into y-slabs over several GPUs behind ONE handle).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/blbm.h"
#include "blbm_internal.cuh"

namespace blbmh {

// text of the last failure on the calling thread (blbm_last_error)
int fail(int code, const char *fmt, ...);

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return blbmh::fail(BLBM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                               __LINE__);                                                                 \
    } while (0)

constexpr uint32_t PEER_MAGIC = 0xB1B30001u;

// what a neighbour needs to know to store into our halo rows
struct PeerBlob {
    uint32_t magic;
    uint32_t W, P, rows;
    uint64_t row0, row1, Hg;
    uint64_t pool_bytes;
    uint64_t off_f[2][8];
    uint64_t off_mx, off_my;
    uint64_t off_flags;
    int32_t device;
    int32_t pid;
    uint64_t local_ptr;  // pool base in the exporting process (used when pid matches)
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(PeerBlob) <= BLBM_PEER_HANDLE_BYTES, "peer blob too large");

struct Peer {
    bool linked = false;
    bool ipc_opened = false;
    char *base = nullptr;  // neighbour's pool mapped into this process / device
    PeerBlob info{};
};

struct Group;  // group.cu

}  // namespace blbmh

struct blbm_handle {
    // non-null: this handle is a whole lattice split into y-slabs over several GPUs (blbm_create_group); every
    // entry point fans out to the slab handles it owns and none of the members below is used
    blbmh::Group *group = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // asynchronous read-back of the output field: a second stream so the copy overlaps later steps
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_sum = nullptr, ev_copy = nullptr;
    bool copy_pending = false;
    uint32_t W = 0, P = 0, rows = 0;
    uint64_t Hg = 0, row0 = 0, row1 = 0;
    size_t plane = 0;  // elements per population plane, (rows+3)*P
    char *pool = nullptr;
    size_t pool_bytes = 0;
    float *f[2][8] = {};
    float *R = nullptr, *mx = nullptr, *my = nullptr, *rho = nullptr, *out = nullptr;
    uint16_t *cls[2] = {};
    uint8_t *rowflag[2] = {};  // per class buffer: one byte per (row, 128-cell chunk), see build_class_kernel
    uint8_t *mask = nullptr;
    unsigned long long *flags = nullptr;  // [0] epoch from the slab above, [16] from below (128 B apart)
    int *err_flag = nullptr;
    double *red_sums = nullptr;
    float *red_max = nullptr;
    size_t off_f[2][8] = {};
    size_t off_mx = 0, off_my = 0, off_flags = 0;
    int cls_cur = 0;
    bool cls_pending = false;  // the mask changed while a stream was pending: cls[cls_cur^1] is newer
    // owned rows in which the two class buffers may differ (a paint rebuilds only the rows it touches)
    uint32_t diff_lo = 0, diff_hi = 0;
    bool regimeT = false;
    bool halo_dirty = false;
    float omega = 1.0f;
    int stat = BLBM_CURL;
    uint64_t step = 0, frame = 0;
    int kernel = BLBM_KERNEL_VEC4;
    int vec4_dense = -1;  // bounce-back flavour of the vec4 kernel: -1 auto, 0 sparse, 1 dense
    int vec4_packed = 0;  // 1: collide cell pairs with packed fp32 adds (FADD2): same bits, but measured slower (registers)
    int vec4_index32 = -1;  // 32-bit plane offsets in the vec4 kernel (plane < 2^32 elements): -1 auto, 0, 1
    int vec4_rows = 4;  // rows per block of the vec4 kernel (tuning knob; 4 measured best on the porous case)
    uint64_t launches = 0;
    blbmh::Peer up, dn;
    unsigned long long epoch = 0, waited = 0;
    unsigned int *link_done = nullptr;  // per-face arrival counters of the in-kernel handshake (LinkSync)
    int link_in_kernel = 1;             // fused vec4 launches of linked slabs wait and signal inside the kernel
    unsigned long long wait_timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
    uint64_t *d_pairs = nullptr;
    size_t d_pairs_cap = 0;
    // pinned staging ring for paint lists: blbm_draw_points returns without waiting for the upload (a frame loop of
    // paint -> iterate -> read_output_async never synchronises the host with the stream)
    static constexpr int STAGE_SLOTS = 4;
    static constexpr size_t STAGE_PAIRS = 8192;  // pairs per slot; longer lists take the synchronous pageable path
    uint64_t *stage_host = nullptr;
    cudaEvent_t stage_ev[STAGE_SLOTS] = {};
    bool stage_used[STAGE_SLOTS] = {};
    int stage_next = 0;
    // barrier chains (lazy barrier cells): see aux_kernels.cu
    int lazy_mode = 2;           // 0 never, 1 always, 2 auto (enough barrier cells and enough steps to pay off)
    bool chain_active = false;   // barrier slots of the planes are don't-care, their state is in the table
    bool chain_declined = false; // auto mode looked at the current mask and decided against
    uint32_t *chain_idx = nullptr;
    uint8_t *chain_flag = nullptr;
    float *chain_state = nullptr;
    size_t chain_n = 0, chain_cap = 0;
    uint32_t *chunk_base = nullptr;    // slots in front of each (row, 128-cell chunk); per-chunk counts while building
    uint32_t *scan_scratch = nullptr;  // one word per 1024 chunks
    bool chain_unsettle = false;       // omega changed: settled chains must be recomputed by the next replay
    unsigned long long *chain_counter = nullptr;
    unsigned long long *mailbox_host = nullptr, *mailbox_dev = nullptr;  // mapped pinned word for small results
    // Small lattices are launch-bound (512x256: ~2.5 us of kernel per step): runs of 2, 4, .. 16 fused steps are
    // captured into CUDA graphs - every even length, for both start parities, both class buffers, with / without a
    // class swap pending after the first step (the step after a paint), and without / with the moment-storing step
    // that ends a call (so that the 15 steps of a frame are ONE launch) - all at once, the first time a run is wanted
    // (a few milliseconds, in the caller's first iterate), so that no capture ever lands inside a frame loop; they
    // are re-captured when omega, the kernel shape or the chain table change.
    static constexpr int GRAPH_SIZES = 8;  // every even run length up to 16
    // [start parity][class buffer][swap pending][+ the call's moment-storing last step][length index]
    cudaGraphExec_t graph[2][2][2][2][GRAPH_SIZES] = {};
    unsigned long long graph_sig[4] = {0, 0, 0, 0}, graph_pending_sig[4] = {0, 0, 0, 0};
    bool graphs_primed = false;
    int use_graphs = -1;  // -1 auto (small lattices, no peers), 0 never, 1 always (when legal)
    int use_pdl = -1;     // programmatic dependent launch of the fused vec4 step: -1 auto (where no graphs are), 0, 1
    float *rgb = nullptr;  // colour buffer (rows x W x 3), allocated by the first blbm_color_map
};
typedef blbm_handle blbm;

namespace blbmh {
// n x compute_step() on one slab; store_moments = false skips the moment-storing launch of the call's last step
// (a group advances its slabs in interleaved chunks and wants the moments only once, at the very end)
int slab_steps(blbm *h, uint32_t n, bool store_moments);
// calculate_summary on one slab (no frame counter)
int slab_summary(blbm *h);
}  // namespace blbmh
