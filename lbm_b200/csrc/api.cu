// api.cu — the C ABI of include/blbm.h: handle life cycle, the regime state machine around the fused
// step kernel, barrier painting, read-back, and the halo peers of y-slab decomposition.
//
// State machine.  `step` is the reference's compute_step (lbm.rs:91,1112-1116).
//   S-regime: both population buffers are bit-identical to the reference's data_buffers.
//   T-regime: buffer (step-1)%2 holds T_{step-1} (post-collision, every cell); buffer step%2 holds what
//             the reference's non-live buffer held one step earlier; the stream of step-1 is pending and
//             is executed by the gather half of the next fused launch (or by materialise()).
// iterate(n) leaves the handle in the T-regime; anything that must observe or perturb the reference's
// buffers (read/write population, the public half-steps) materialises first.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <unistd.h>
#include <vector>

#include "handle.cuh"
#include "group.h"

using namespace blbmk;

namespace {

using blbmh::fail;
using blbmh::Peer;
using blbmh::PeerBlob;
using blbmh::PEER_MAGIC;

thread_local char g_err[512] = "";

#define CKH(h)                                                                  \
    do {                                                                        \
        if (!(h)) return fail(BLBM_EINVAL, "null handle");                      \
        CK(cudaSetDevice((h)->device));                                         \
    } while (0)

}  // namespace

int blbmh::fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

SlabGeom geom(const blbm *h)
{
    SlabGeom g;
    g.W = h->W;
    g.P = h->P;
    g.rows = h->rows;
    g.row0 = h->row0;
    g.Hg = h->Hg;
    return g;
}

bool has_up(const blbm *h) { return h->row0 > 0; }
bool has_dn(const blbm *h) { return h->row1 < h->Hg; }

// remote halo-row pointers for a launch that writes buffer `b`
PushTargets push_targets(const blbm *h, int b)
{
    PushTargets t;
    memset(&t, 0, sizeof(t));
    const size_t P = h->P;
    if (h->up.linked) {
        const PeerBlob &u = h->up.info;
        char *base = h->up.base;
        const size_t halo1 = (size_t)(u.rows + 1) * P, halo2 = (size_t)(u.rows + 2) * P;
        t.up_n = reinterpret_cast<float *>(base + u.off_f[b][D_N]) + halo1;
        t.up_ne = reinterpret_cast<float *>(base + u.off_f[b][D_NE]) + halo1;
        t.up_nw = reinterpret_cast<float *>(base + u.off_f[b][D_NW]) + halo1;
        t.up_w = reinterpret_cast<float *>(base + u.off_f[b][D_W]) + halo1;
        t.up_nw2 = reinterpret_cast<float *>(base + u.off_f[b][D_NW]) + halo2;
        t.up_mx = reinterpret_cast<float *>(base + u.off_mx) + halo1;
        t.up_my = reinterpret_cast<float *>(base + u.off_my) + halo1;
    }
    if (h->dn.linked) {
        const PeerBlob &d = h->dn.info;
        char *base = h->dn.base;
        t.dn_s = reinterpret_cast<float *>(base + d.off_f[b][D_S]);
        t.dn_se = reinterpret_cast<float *>(base + d.off_f[b][D_SE]);
        t.dn_sw = reinterpret_cast<float *>(base + d.off_f[b][D_SW]);
        t.dn_mx = reinterpret_cast<float *>(base + d.off_mx);
        t.dn_my = reinterpret_cast<float *>(base + d.off_my);
    }
    return t;
}

bool any_peer(const blbm *h) { return h->up.linked || h->dn.linked; }

// Before a kernel that reads halo rows, or stores into a neighbour's: wait until both neighbours have
// finished the launch that produced the current epoch.
int sync_peers(blbm *h)
{
    if (!any_peer(h) || h->waited >= h->epoch) return BLBM_OK;
    const unsigned long long *fu = h->up.linked ? h->flags : nullptr;
    const unsigned long long *fd = h->dn.linked ? h->flags + 16 : nullptr;
    CK(launch_wait(fu, fd, h->epoch, h->err_flag, h->wait_timeout_ns, h->stream));
    h->launches++;
    h->waited = h->epoch;
    return BLBM_OK;
}

// After a kernel that stored into neighbours' halo rows: publish a new epoch to them.
int signal_peers(blbm *h)
{
    if (!any_peer(h)) return BLBM_OK;
    h->epoch++;
    // our word in the slab above is its "from below" flag, and vice versa
    unsigned long long *ru =
        h->up.linked ? reinterpret_cast<unsigned long long *>(h->up.base + h->up.info.off_flags) + 16 : nullptr;
    unsigned long long *rd =
        h->dn.linked ? reinterpret_cast<unsigned long long *>(h->dn.base + h->dn.info.off_flags) : nullptr;
    CK(launch_signal(ru, rd, h->epoch, h->stream));
    h->launches++;
    return BLBM_OK;
}

int push_all_halos(blbm *h);

ChainPlanes chain_planes(const blbm *h)
{
    ChainPlanes pl;
    for (int d = 0; d < 8; d++) {
        pl.f0[d] = h->f[0][d];
        pl.f1[d] = h->f[1][d];
    }
    pl.R = h->R;
    return pl;
}

ChainTable chain_table(const blbm *h)
{
    ChainTable t;
    t.idx = h->chain_idx;
    t.flag = h->chain_flag;
    t.state = h->chain_state;
    t.chunk_base = h->chunk_base;
    t.n = h->chain_n;
    t.cap = h->chain_cap;
    return t;
}

// table -> planes: afterwards every barrier slot again holds exactly what the reference's buffers hold
int chain_flush(blbm *h)
{
    if (!h->chain_active) return BLBM_OK;
    CK(launch_chain_flush(chain_table(h), chain_planes(h), h->cls[0], h->cls[1], (uint32_t)(h->step % 2), h->stream));
    h->launches++;
    h->chain_active = false;
    return BLBM_OK;
}

// Stream-ordered allocation: cudaMalloc/cudaFree synchronise the whole device, which deadlocks when a
// sibling slab of the same process has a wait kernel queued on this slab's next signal.
cudaError_t stream_alloc(void **p, size_t bytes, cudaStream_t st) { return cudaMallocAsync(p, bytes, st); }
void stream_free(void *p, cudaStream_t st)
{
    if (p) cudaFreeAsync(p, st);
}

// planes -> table, if it pays off.  Never fails the caller: on any shortage the dense path stays in use.
int chain_try_enter(blbm *h, uint32_t steps_left)
{
    if (h->chain_active || h->lazy_mode == 0 || h->chain_declined || h->cls_pending) return BLBM_OK;
    (void)steps_left;  // the build pays for itself within a few steps: enter as soon as the mask qualifies
    if (h->plane >= 0xffffffffull) return BLBM_OK;
    CK(launch_chain_count(h->cls[h->cls_cur], geom(h), h->chunk_base, h->chain_counter, h->mailbox_dev, h->stream));
    h->launches += 2;
    CK(cudaStreamSynchronize(h->stream));
    const unsigned long long n = *reinterpret_cast<volatile unsigned long long *>(h->mailbox_host);
    const unsigned long long cells = (unsigned long long)h->rows * h->W;
    if (n == 0 || (h->lazy_mode == 2 && n * 50ull < cells)) {
        h->chain_declined = true;
        return BLBM_OK;
    }
    if (h->chain_cap < n) {
        stream_free(h->chain_idx, h->stream);
        stream_free(h->chain_flag, h->stream);
        stream_free(h->chain_state, h->stream);
        h->chain_idx = nullptr;
        h->chain_flag = nullptr;
        h->chain_state = nullptr;
        h->chain_cap = 0;
        if (stream_alloc((void **)&h->chain_idx, n * sizeof(uint32_t), h->stream) != cudaSuccess ||
            stream_alloc((void **)&h->chain_flag, n, h->stream) != cudaSuccess ||
            stream_alloc((void **)&h->chain_state, n * CHAIN_ROWS * sizeof(float), h->stream) != cudaSuccess) {
            cudaGetLastError();
            stream_free(h->chain_idx, h->stream);
            stream_free(h->chain_flag, h->stream);
            h->chain_idx = nullptr;
            h->chain_flag = nullptr;
            h->chain_declined = true;  // not enough memory: stay dense
            return BLBM_OK;
        }
        h->chain_cap = (size_t)n;
    }
    h->chain_n = (size_t)n;
    CK(launch_chain_build(h->cls[h->cls_cur], h->cls[h->cls_cur ^ 1], geom(h), chain_planes(h), chain_table(h),
                          h->scan_scratch, (uint32_t)(h->step % 2), h->stream));
    h->launches += 4;
    h->chain_active = true;
    h->chain_unsettle = false;  // fresh entries carry no settled flag
    return BLBM_OK;
}

bool graphs_wanted(const blbm *h);

int launch_step(blbm *h, int mode, int xbuf, int ybuf, bool mom)
{
    StepParams p;
    memset(&p, 0, sizeof(p));
    for (int d = 0; d < 8; d++) {
        p.X[d] = h->f[xbuf][d];
        p.Y[d] = h->f[ybuf][d];
    }
    p.R = h->R;
    p.cls = h->cls[h->cls_cur];
    p.rowflag = h->rowflag[h->cls_cur];
    p.mx = h->mx;
    p.my = h->my;
    p.rho = h->rho;
    p.W = h->W;
    p.P = h->P;
    p.rows = h->rows;
    p.row0 = h->row0;
    p.Hg = h->Hg;
    p.omega = h->omega;
    const bool pushes = mode != MODE_STREAM_ONLY;
    if (pushes) p.push = push_targets(h, ybuf);
    int rc;
    if (h->halo_dirty && any_peer(h) && mode != MODE_COLLIDE_ONLY && (rc = push_all_halos(h)) != BLBM_OK) return rc;
    int k = h->kernel;
    if (mode != MODE_FUSED) k = BLBM_KERNEL_SCALAR;
    // Linked slabs: the fused vec4 kernel does the epoch handshake itself (its face row blocks wait, store, and the
    // last of them publishes the new epoch); every other launch is bracketed by the one-thread wait / signal kernels.
    const bool link_in_kernel = any_peer(h) && k == BLBM_KERNEL_VEC4 && h->link_in_kernel &&
                                vec4_links_in_kernel(h->vec4_rows, h->vec4_packed != 0);
    if (link_in_kernel) {
        p.link.wait_up = h->up.linked ? h->flags : nullptr;
        p.link.wait_dn = h->dn.linked ? h->flags + 16 : nullptr;
        // our word in the slab above is its "from below" flag, and vice versa
        p.link.sig_up =
            h->up.linked ? reinterpret_cast<unsigned long long *>(h->up.base + h->up.info.off_flags) + 16 : nullptr;
        p.link.sig_dn = h->dn.linked ? reinterpret_cast<unsigned long long *>(h->dn.base + h->dn.info.off_flags) : nullptr;
        p.link.wait_epoch = h->epoch;
        p.link.sig_epoch = h->epoch + 1;
        p.link.done = h->link_done;
        p.link.err_flag = h->err_flag;
        p.link.timeout_ns = h->wait_timeout_ns;
    } else {
        rc = sync_peers(h);
        if (rc) return rc;
    }
    cudaError_t e;
    if (mom && h->chain_active) {
        // chain cells' moments (of the collide of buffer `ybuf`, which this launch performs for every other cell)
        // live in the table: the vec4 kernel merges them by slot rank, the scalar kernel wants them in the planes
        if (k == BLBM_KERNEL_SCALAR) {
            CK(launch_chain_scatter_moments(chain_table(h), (uint32_t)ybuf, h->mx, h->my, h->rho, h->stream));
            h->launches++;
        } else {
            p.chunk_base = h->chunk_base;
            p.chain_mom = h->chain_state + (size_t)(CHAIN_ROW_MOM + 3 * ybuf) * h->chain_cap;
            p.chain_cap = (uint32_t)h->chain_cap;
        }
    }
    switch (k) {
    case BLBM_KERNEL_SCALAR: e = launch_step_scalar(p, mode, mom, h->stream); break;
    default:
        // the obstacle-dense flavour with cp.async-staged own rows wherever the chain table is in use (>= 2 % barrier
        // cells), unless overridden
        {
            const int flavour = h->vec4_dense < 0 ? (h->chain_active ? 2 : 0) : h->vec4_dense;
            // 32-bit plane offsets wherever the slab allows it (always on a B200 at the default block shape): one
            // IMAD.WIDE per address, and the variants that then stay at 64 registers without a spill
            const bool index32 = h->vec4_index32 != 0;
            // stream-ordered steps: let the next step's blocks be scheduled while this one drains (4096^2: +1.2 %).
            // Not inside the graphs of small lattices: programmatic edges made their replay slower (4.7 vs 4.1 us
            // per step, profiles/r2/r2u_*)
            const bool pdl = !any_peer(h) && (h->use_pdl < 0 ? !graphs_wanted(h) : h->use_pdl != 0);
            e = launch_step_vec4(p, mode, mom, h->vec4_rows, flavour, h->vec4_packed != 0, index32, pdl, h->stream);
        }
        break;
    }
    if (e != cudaSuccess) return fail(BLBM_ECUDA, "step kernel launch failed: %s", cudaGetErrorString(e));
    h->launches++;
    if (mode == MODE_COLLIDE_ONLY) h->halo_dirty = false;  // this launch pushed the live buffer's boundary rows
    if (link_in_kernel) {
        h->waited = h->epoch;
        h->epoch++;
        return BLBM_OK;
    }
    if (pushes) return signal_peers(h);
    return BLBM_OK;
}

void consume_pending_class(blbm *h)
{
    if (h->cls_pending) {
        h->cls_cur ^= 1;
        h->cls_pending = false;
    }
}

// T-regime -> S-regime: run the pending stream (stream/*.wgsl) so that both buffers equal the reference's
int materialise(blbm *h)
{
    if (!h->regimeT) return BLBM_OK;
    const int x = (int)((h->step + 1) % 2), y = (int)(h->step % 2);
    int rc = launch_step(h, MODE_STREAM_ONLY, x, y, false);
    if (rc) return rc;
    h->regimeT = false;
    consume_pending_class(h);
    // The live buffer now holds S_step in our rows, but our halo rows of that buffer still hold what the
    // neighbours pushed two steps ago.  The next collide refreshes them (it pushes T_step); anything that
    // gathers from the live buffer before that (the public stream half-step) must re-exchange first.
    if (any_peer(h)) h->halo_dirty = true;
    return BLBM_OK;
}

int run_summary(blbm *h)
{
    int rc = sync_peers(h);  // curl reads the moment halo rows
    if (rc) return rc;
    if (h->copy_pending) CK(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));  // the copy still reads `out`
    CK(launch_summary(h->stat, h->mx, h->my, h->rho, h->out, geom(h), h->stream));
    h->launches++;
    // Linked slabs: the summary is part of the epoch protocol.  curl READS the moment halo rows, and the neighbours'
    // next moment-storing launch overwrites them; that launch only waits for our previous *pushing* launch, which is
    // stream-ordered BEFORE this kernel - so a one-step iterate or a public collide half-step right after an iterate
    // could overtake a summary still pending here (seen once in 400 random walks, with four processes time-slicing
    // the GPU: one boundary row of the curl field computed from the next call's moments).  Publishing an epoch after
    // the summary makes every later push into our halo rows wait for it.  (Every slab runs the same calls, so the
    // epochs stay in step.)
    return signal_peers(h);
}

// all even: a call's runs share one start parity; every even length, so that the 14 graph-able steps of a 15-step
// frame are one launch instead of three (8 + 4 + 2)
constexpr uint32_t GRAPH_LEN[blbm::GRAPH_SIZES] = {16, 14, 12, 10, 8, 6, 4, 2};

bool graphs_wanted(const blbm *h)
{
    if (any_peer(h)) return false;  // peers: the epochs of the halo handshake are launch parameters
    if (h->use_graphs >= 0) return h->use_graphs != 0;
    return (unsigned long long)h->rows * h->W <= (4ull << 20);
}

void drop_step_graphs(blbm *h)
{
    for (auto &par : h->graph)
        for (auto &cls : par)
            for (auto &pend : cls)
                for (auto &tail : pend)
                    for (cudaGraphExec_t &g : tail)
                        if (g) {
                            cudaGraphExecDestroy(g);
                            g = nullptr;
                        }
    h->graphs_primed = false;
}

// capture every run length for both start parities and both class buffers under the current kernel configuration,
// without and with the moment-storing step that ends a call (tail; not with the chain table, whose moment rows
// are launch parameters of that step)
int prime_step_graphs(blbm *h)
{
    drop_step_graphs(h);
    const uint64_t step0 = h->step, launches0 = h->launches;
    const int cls0 = h->cls_cur;
    int rc = BLBM_OK;
    for (int par = 0; par < 2 && rc == BLBM_OK; par++)
        for (int cls = 0; cls < 2 && rc == BLBM_OK; cls++)
          for (int pend = 0; pend < 2 && rc == BLBM_OK; pend++)
           for (int tail = 0; tail < (h->chain_active ? 1 : 2) && rc == BLBM_OK; tail++)
            for (int q = 0; q < blbm::GRAPH_SIZES && rc == BLBM_OK; q++) {
                cudaGraph_t graph = nullptr;
                cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
                if (e != cudaSuccess) {
                    rc = fail(BLBM_ECUDA, "stream capture failed: %s", cudaGetErrorString(e));
                    break;
                }
                h->step = (uint64_t)par;
                h->cls_cur = cls;
                for (uint32_t k = 0; k < GRAPH_LEN[q] + (uint32_t)tail && rc == BLBM_OK; k++) {
                    const int x = (int)((h->step + 1) % 2), y = (int)(h->step % 2);
                    rc = launch_step(h, MODE_FUSED, x, y, k == GRAPH_LEN[q]);
                    h->step++;
                    if (pend && k == 0) h->cls_cur ^= 1;  // the pending stream saw the old classification; swap
                }
                e = cudaStreamEndCapture(h->stream, &graph);
                if (rc == BLBM_OK && e != cudaSuccess) rc = fail(BLBM_ECUDA, "stream capture failed: %s", cudaGetErrorString(e));
                if (rc == BLBM_OK) {
                    e = cudaGraphInstantiate(&h->graph[par][cls][pend][tail][q], graph, 0);
                    if (e != cudaSuccess) {
                        h->graph[par][cls][pend][tail][q] = nullptr;
                        rc = fail(BLBM_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
                    }
                }
                if (graph) cudaGraphDestroy(graph);
            }
    h->step = step0;
    h->cls_cur = cls0;
    h->launches = launches0;  // capture enqueues nothing
    if (rc != BLBM_OK) {
        drop_step_graphs(h);
        return rc;
    }
    h->graphs_primed = true;
    return BLBM_OK;
}

// Are the graphs usable for this call?  Captured at once the first time; after a change of configuration (omega,
// kernel shape, chain table) only when the new configuration has been seen by two calls in a row, so that e.g. a
// viscosity slider being dragged costs graph-less frames instead of a re-capture per frame.
int ensure_step_graphs(blbm *h, bool *usable)
{
    unsigned int omega_bits;
    memcpy(&omega_bits, &h->omega, sizeof(omega_bits));
    const unsigned long long sig[4] = {
        1, omega_bits,
        (unsigned long long)h->kernel | ((unsigned long long)h->vec4_rows << 8) |
            ((unsigned long long)(h->vec4_dense + 1) << 16) | ((unsigned long long)h->chain_active << 24) |
            ((unsigned long long)h->vec4_packed << 25) | ((unsigned long long)(h->vec4_index32 + 1) << 26) |
            ((unsigned long long)(h->use_pdl + 1) << 28),
        (unsigned long long)(uintptr_t)h->pool};
    *usable = true;
    if (h->graphs_primed && memcmp(h->graph_sig, sig, sizeof(sig)) == 0) return BLBM_OK;
    if (h->graphs_primed && memcmp(h->graph_pending_sig, sig, sizeof(sig)) != 0) {
        memcpy(h->graph_pending_sig, sig, sizeof(sig));
        *usable = false;
        return BLBM_OK;
    }
    const int rc = prime_step_graphs(h);
    if (rc != BLBM_OK) return rc;
    memcpy(h->graph_sig, sig, sizeof(sig));
    return BLBM_OK;
}

// GRAPH_LEN[q] fused, non-moment-storing steps starting at the current parity (and, with tail, the moment-storing
// step after them), as one graph launch
int run_step_graph(blbm *h, int q, bool tail)
{
    CK(cudaGraphLaunch(h->graph[h->step % 2][h->cls_cur][h->cls_pending ? 1 : 0][tail ? 1 : 0][q], h->stream));
    h->step += GRAPH_LEN[q] + (tail ? 1u : 0u);
    h->launches += GRAPH_LEN[q] + (tail ? 1u : 0u);
    consume_pending_class(h);  // (its first step ran on the old class buffer)
    return BLBM_OK;
}

int do_steps(blbm *h, uint32_t n, bool store_moments = true)
{
    uint32_t left = n;
    bool replayed = false;
    const bool graphs = graphs_wanted(h);
    int graph_state = -1;  // decided once per call, after the chain table had its chance to change the configuration
    while (left) {
        const bool mom = left == 1 && store_moments;
        int rc;
        if (!h->chain_active && (rc = chain_try_enter(h, left)) != BLBM_OK) return rc;
        if (h->chain_active && !replayed) {
            // barrier cells do not depend on anything else: advance their chains through all the steps of
            // this call up front, in registers; this also stores the moments of the call's last collide
            CK(launch_chain_replay(chain_table(h), left, (uint32_t)(h->step % 2), h->omega, h->chain_unsettle,
                                   h->stream));
            h->launches++;
            h->chain_unsettle = false;
            replayed = true;
        }
        // every step but the call's last (it may store moments) can go into a graph; the graphs are made (or re-made
        // after a change of omega / kernel shape / chain table) by the first call that could use one, however short,
        // so that the millisecond of capturing lands in a caller's warm-up and never in the middle of a frame loop
        if (graphs && graph_state < 0) {
            bool usable = false;
            if ((rc = ensure_step_graphs(h, &usable)) != BLBM_OK) return rc;
            graph_state = usable ? 1 : 0;
        }
        if (graph_state == 1 && h->regimeT && left - 1 >= GRAPH_LEN[blbm::GRAPH_SIZES - 1]) {
            int q = 0;
            while (GRAPH_LEN[q] > left - 1) q++;
            // the call ends with this run and its moment-storing step: one launch for both
            const bool tail = store_moments && !h->chain_active && left - 1 == GRAPH_LEN[q];
            if ((rc = run_step_graph(h, q, tail)) != BLBM_OK) return rc;
            left -= GRAPH_LEN[q] + (tail ? 1u : 0u);
            continue;
        }
        if (!h->regimeT) {
            // collide of step `step`, in place on the live buffer (collision/*.wgsl); its stream stays pending
            const int y = (int)(h->step % 2);
            rc = launch_step(h, MODE_COLLIDE_ONLY, y, y, mom);
            if (rc) return rc;
            h->regimeT = true;
        } else {
            // stream of step-1 fused with collide of step
            const int x = (int)((h->step + 1) % 2), y = (int)(h->step % 2);
            rc = launch_step(h, MODE_FUSED, x, y, mom);
            if (rc) return rc;
            consume_pending_class(h);
        }
        h->step++;
        left--;
    }
    return BLBM_OK;
}

// The mask changed in owned rows [lo, hi): bring the class words (and chunk flags) up to date.
// big_change: the number of barrier cells may have changed enough for auto mode to decide afresh.
// Decide which class buffer a rebuild of owned rows [lo, hi) goes to and which rows it must cover (see the comments
// inside); returns false when there is nothing to rebuild.  Updates the handle's bookkeeping: the caller MUST launch.
bool plan_class_rebuild(blbm *h, uint32_t lo, uint32_t hi, bool big_change, int *target_out, uint32_t *blo_out,
                        uint32_t *bhi_out)
{
    if (big_change) h->chain_declined = false;
    if (hi > h->rows) hi = h->rows;
    if (lo >= hi) return false;
    const bool had_diff = h->diff_hi > h->diff_lo;
    const uint32_t ulo = had_diff ? std::min(lo, h->diff_lo) : lo, uhi = had_diff ? std::max(hi, h->diff_hi) : hi;
    int target;
    uint32_t blo = lo, bhi = hi;
    if (h->regimeT) {
        // the pending stream must still see the old classification: build the other buffer
        target = h->cls_cur ^ 1;
        if (!h->cls_pending) {
            // that buffer is stale wherever the two differ: refresh those rows too; afterwards it is current
            // everywhere and the two buffers differ only where this paint changed the mask
            blo = ulo;
            bhi = uhi;
            h->diff_lo = lo;
            h->diff_hi = hi;
        } else {
            // a newer classification is already waiting there: only this paint's rows need rebuilding
            h->diff_lo = ulo;
            h->diff_hi = uhi;
        }
        h->cls_pending = true;
    } else {
        target = h->cls_cur;
        h->diff_lo = ulo;
        h->diff_hi = uhi;
    }
    *target_out = target;
    *blo_out = blo;
    *bhi_out = bhi;
    return true;
}

int rebuild_class(blbm *h, uint32_t lo, uint32_t hi, bool big_change)
{
    int target;
    uint32_t blo, bhi;
    if (!plan_class_rebuild(h, lo, hi, big_change, &target, &blo, &bhi)) return BLBM_OK;
    CK(launch_build_class(h->cls[target], h->mask, geom(h), h->chain_active ? h->cls[h->cls_cur] : nullptr,
                          h->rowflag[target], blo, bhi, h->stream));
    h->launches++;
    return BLBM_OK;
}

int fill_equilibrium(blbm *h, float ux, int single_index)
{
    // Leave the chain table first: the flush also clears CLS_CHAIN in both class planes.  (Dropping the table
    // with the bits still set made the next collide half-step keep stale moments at barrier cells, and a step
    // kernel with the table switched off treat their plane slots as don't-care.)
    {
        const int rcf = chain_flush(h);
        if (rcf) return rcf;
    }
    // set_equil on the host in fp32 with the reference's op order (lbm.rs:611-643); uy = 0, rho = 1
    float v9[9];
    {
        float uxx = ux, uy = 0.0f, rho = 1.0f;
        float ux_2 = uxx * uxx, uy_2 = uy * uy;
        float u_dot = ux_2 + uy_2;
        float uxuy = uxx * uy;
        float pos = u_dot + 2.0f * uxuy, neg = u_dot - 2.0f * uxuy;
        uxx *= 3.0f;
        uy *= 3.0f;
        ux_2 *= 4.5f;
        uy_2 *= 4.5f;
        u_dot *= 1.5f;
        neg *= 4.5f;
        pos *= 4.5f;
        const float r9 = rho / 9.0f, r36 = rho / 36.0f;
        v9[BLBM_NW] = r36 * ((((1.0f - uxx) + uy) + neg) - u_dot);
        v9[BLBM_N] = r9 * (((1.0f + uy) + uy_2) - u_dot);
        v9[BLBM_NE] = r36 * ((((1.0f + uxx) + uy) + pos) - u_dot);
        v9[BLBM_W] = r9 * (((1.0f - uxx) + ux_2) - u_dot);
        v9[BLBM_REST] = (4.0f * r9) * (1.0f - u_dot);
        v9[BLBM_E] = r9 * (((1.0f + uxx) + ux_2) - u_dot);
        v9[BLBM_SW] = r36 * ((((1.0f - uxx) - uy) + pos) - u_dot);
        v9[BLBM_S] = r9 * (((1.0f - uy) - uy_2) - u_dot);
        v9[BLBM_SE] = r36 * ((((1.0f + uxx) - uy) + neg) - u_dot);
    }
    static const int pop_of_dir[8] = {BLBM_NW, BLBM_N, BLBM_NE, BLBM_W, BLBM_E, BLBM_SW, BLBM_S, BLBM_SE};
    float *planes[17];
    float vals[17];
    int n = 0;
    for (int b = 0; b < 2; b++)
        for (int d = 0; d < 8; d++) {
            planes[n] = h->f[b][d];
            vals[n++] = v9[pop_of_dir[d]];
        }
    planes[n] = h->R;
    vals[n++] = v9[BLBM_REST];
    // own rows plus the halo rows that belong to a neighbouring slab (rows outside the lattice stay 0)
    const uint32_t r_begin = has_up(h) ? 0u : 1u;
    uint32_t r_end = h->rows + 1;
    if (has_dn(h)) r_end += (uint32_t)std::min<uint64_t>(2, h->Hg - h->row1);
    CK(launch_fill_rows(planes, vals, n, h->W, h->P, r_begin, r_end, h->stream));
    h->launches++;
    if (single_index >= 0 && single_index <= 8) {
        // set_single_cell, lbm.rs:1482-1500
        const int64_t x = h->W, y = (int64_t)h->Hg;
        int64_t cx = 0, cy = 0;
        switch (single_index) {
        case 0: cx = x - 2; cy = y - 2; break;
        case 1: cx = 3 * x / 4; cy = y - 2; break;
        case 2: cx = x / 3; cy = y - 2; break;
        case 3: cx = x - 2; cy = y / 2; break;
        case 4: cx = 3 * x / 4; cy = y / 2; break;
        case 5: cx = x / 2; cy = y / 2; break;
        case 6: cx = x - 2; cy = 1; break;
        case 7: cx = 3 * x / 4; cy = 1; break;
        default: cx = x / 2; cy = 1; break;
        }
        // the reference indexes the flat array with cx + cy*W; negative or past-the-end indices would
        // panic there, here they are dropped
        const int64_t flat = cx + cy * x;
        if (cx >= 0 && cy >= 0 && flat >= 0 && flat < x * y) {
            const int64_t gy = flat / x, gx = flat % x;
            const int64_t dr = gy - (int64_t)h->row0 + 1;  // device row
            if (dr >= (int64_t)r_begin && dr < (int64_t)r_end) {
                // one-cell fills through the fill kernel: no host buffer, no host synchronisation
                const size_t off = (size_t)dr * h->P + (size_t)gx;
                float *cell[2];
                const float four[2] = {4.0f, 4.0f};
                int nc = 0;
                if (single_index == BLBM_REST) {
                    cell[nc++] = h->R + off;
                } else {
                    int d = 0;
                    for (int q = 0; q < 8; q++)
                        if (pop_of_dir[q] == single_index) d = q;
                    for (int b = 0; b < 2; b++) cell[nc++] = h->f[b][d] + off;
                }
                CK(launch_fill_rows(cell, four, nc, 1, 1, 0, 1, h->stream));
                h->launches++;
            }
        }
    }
    h->step = 0;
    h->frame = 0;
    h->regimeT = false;
    consume_pending_class(h);
    h->halo_dirty = false;  // every slab filled its halo rows with the same constants
    return BLBM_OK;
}

int moments_without_rest(blbm *h)
{
    const uint32_t r_begin = has_up(h) ? 0u : 1u;
    const uint32_t r_end = h->rows + 1 + (has_dn(h) ? 1u : 0u);
    const float *f8[8];
    for (int d = 0; d < 8; d++) f8[d] = h->f[0][d];
    CK(launch_precollision_moments(f8, h->mx, h->my, h->rho, h->W, h->P, r_begin, r_end, h->stream));
    h->launches++;
    return BLBM_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int check_peer_err(blbm *h)
{
    if (!any_peer(h)) return BLBM_OK;
    int e = 0;
    CK(cudaMemcpyAsync(&e, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (e) return fail(BLBM_EPEER, "a neighbouring slab did not reach the expected epoch within the timeout");
    return BLBM_OK;
}

int sync_stream(blbm *h)
{
    CK(cudaStreamSynchronize(h->stream));
    if (h->copy_pending) {
        CK(cudaStreamSynchronize(h->copy_stream));
        h->copy_pending = false;
    }
    return check_peer_err(h);
}

// push the rows neighbours gather from, for both buffers (used right after linking and after
// blbm_write_population)
int push_all_halos(blbm *h)
{
    int rc = sync_peers(h);
    if (rc) return rc;
    const size_t rowb = (size_t)h->P * sizeof(float);
    for (int b = 0; b < 2; b++) {
        PushTargets t = push_targets(h, b);
        const size_t first = row_off(0, h->P), second = row_off(1, h->P), last = row_off(h->rows - 1, h->P);
        if (h->up.linked) {
            CK(cudaMemcpyAsync(t.up_n, h->f[b][D_N] + first, rowb, cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.up_ne, h->f[b][D_NE] + first, rowb, cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.up_nw, h->f[b][D_NW] + first, rowb, cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.up_w, h->f[b][D_W] + first, sizeof(float), cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.up_nw2, h->f[b][D_NW] + second, sizeof(float), cudaMemcpyDefault, h->stream));
        }
        if (h->dn.linked) {
            CK(cudaMemcpyAsync(t.dn_s, h->f[b][D_S] + last, rowb, cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.dn_se, h->f[b][D_SE] + last, rowb, cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.dn_sw, h->f[b][D_SW] + last, rowb, cudaMemcpyDefault, h->stream));
        }
    }
    {
        PushTargets t = push_targets(h, 0);
        const size_t first = row_off(0, h->P), last = row_off(h->rows - 1, h->P);
        if (h->up.linked) {
            CK(cudaMemcpyAsync(t.up_mx, h->mx + first, rowb, cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.up_my, h->my + first, rowb, cudaMemcpyDefault, h->stream));
        }
        if (h->dn.linked) {
            CK(cudaMemcpyAsync(t.dn_mx, h->mx + last, rowb, cudaMemcpyDefault, h->stream));
            CK(cudaMemcpyAsync(t.dn_my, h->my + last, rowb, cudaMemcpyDefault, h->stream));
        }
    }
    h->halo_dirty = false;
    return signal_peers(h);
}

void fill_blob(const blbm *h, PeerBlob *b)
{
    memset(b, 0, sizeof(*b));
    b->magic = PEER_MAGIC;
    b->W = h->W;
    b->P = h->P;
    b->rows = h->rows;
    b->row0 = h->row0;
    b->row1 = h->row1;
    b->Hg = h->Hg;
    b->pool_bytes = h->pool_bytes;
    for (int q = 0; q < 2; q++)
        for (int d = 0; d < 8; d++) b->off_f[q][d] = h->off_f[q][d];
    b->off_mx = h->off_mx;
    b->off_my = h->off_my;
    b->off_flags = h->off_flags;
    b->device = h->device;
    b->pid = (int32_t)getpid();
    b->local_ptr = (uint64_t)(uintptr_t)h->pool;
}

int attach_peer(blbm *h, int side, const PeerBlob &b, char *base, bool ipc_opened)
{
    Peer &p = side == 0 ? h->up : h->dn;
    if (p.linked) return fail(BLBM_ESTATE, "side %d already linked", side);
    if (b.magic != PEER_MAGIC) return fail(BLBM_EINVAL, "bad peer blob");
    if (b.W != h->W || b.Hg != h->Hg) return fail(BLBM_EINVAL, "peer belongs to a different lattice");
    if (side == 0 && b.row1 != h->row0) return fail(BLBM_EINVAL, "peer is not the slab directly above");
    if (side == 1 && b.row0 != h->row1) return fail(BLBM_EINVAL, "peer is not the slab directly below");
    if (h->rows < 2 || b.rows < 2) return fail(BLBM_EINVAL, "linked slabs need at least 2 rows each");
    p.linked = true;
    p.ipc_opened = ipc_opened;
    p.base = base;
    p.info = b;
    return BLBM_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
int blbmh::slab_steps(blbm *h, uint32_t n, bool store_moments)
{
    CKH(h);
    return do_steps(h, n, store_moments);
}

int blbmh::slab_summary(blbm *h)
{
    CKH(h);
    return run_summary(h);
}

// group handles (blbm_create_group): hand the call to the fan-out of the same name in group.cu
#define GRP(h, expr)                                \
    do {                                            \
        if ((h) && (h)->group) return (expr);       \
    } while (0)
#define GRP_EACH(h, call) GRP(h, blbmh::group_each(h, [&](blbm *s) { return call; }))
using namespace blbmh;

extern "C" {

const char *blbm_last_error(void) { return g_err; }
int blbm_abi_version(void) { return 1; }

int blbm_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(BLBM_ENOGPU, "no CUDA device: %s", cudaGetErrorString(e));
    return n;
}

int blbm_create_slab(uint32_t w, uint64_t h_global, uint64_t row_begin, uint64_t row_end, float omega,
                     float inflow_ux, int device, blbm_t **out)
{
    if (!out) return fail(BLBM_EINVAL, "out is null");
    *out = nullptr;
    if (w == 0 || h_global == 0) return fail(BLBM_EINVAL, "lattice must be at least 1x1");
    if (row_begin >= row_end || row_end > h_global) return fail(BLBM_EINVAL, "bad row range");
    if (row_end - row_begin > 0x7ffffff0ull) return fail(BLBM_EINVAL, "slab too tall");
    if (w > 0x7fffff00u) return fail(BLBM_EINVAL, "rows too long");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(BLBM_ENOGPU, "no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fail(BLBM_EINVAL, "device %d out of range [0,%d)", device, ndev);
    CK(cudaSetDevice(device));
    CK(preload_aux_kernels());
    CK(preload_step_kernels());

    blbm *h = new (std::nothrow) blbm();
    if (!h) return fail(BLBM_ENOMEM, "out of host memory");
    h->device = device;
    h->W = w;
    h->P = (w + 31u) / 32u * 32u;
    h->Hg = h_global;
    h->row0 = row_begin;
    h->row1 = row_end;
    h->rows = (uint32_t)(row_end - row_begin);
    h->plane = (size_t)(h->rows + 3) * h->P;
    h->omega = omega;

    // one pool, carved into 1 KiB-aligned planes (a single allocation = a single IPC handle)
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    const size_t pb = h->plane * sizeof(float);
    for (int b = 0; b < 2; b++)
        for (int d = 0; d < 8; d++) h->off_f[b][d] = carve(pb);
    const size_t off_R = carve(pb);
    h->off_mx = carve(pb);
    h->off_my = carve(pb);
    const size_t off_rho = carve(pb), off_out = carve(pb);
    const size_t off_cls0 = carve(h->plane * sizeof(uint16_t)), off_cls1 = carve(h->plane * sizeof(uint16_t));
    const size_t flag_bytes = (size_t)h->rows * ((h->P + CHUNK - 1) / CHUNK);
    const size_t off_rf0 = carve(flag_bytes), off_rf1 = carve(flag_bytes);
    const size_t off_cb = carve(flag_bytes * sizeof(uint32_t));
    const size_t off_ss = carve((flag_bytes / 1024 + 1) * sizeof(uint32_t));
    const size_t off_mask = carve((size_t)(h->rows + 4) * h->P);
    h->off_flags = carve(256);
    const size_t off_err = carve(64), off_red = carve(64), off_cnt = carve(64), off_done = carve(64);
    h->pool_bytes = off;
    e = cudaMalloc(&h->pool, h->pool_bytes);
    if (e != cudaSuccess) {
        delete h;
        return fail(BLBM_ENOMEM, "cudaMalloc of %zu bytes failed: %s", off, cudaGetErrorString(e));
    }
    for (int b = 0; b < 2; b++)
        for (int d = 0; d < 8; d++) h->f[b][d] = reinterpret_cast<float *>(h->pool + h->off_f[b][d]);
    h->R = reinterpret_cast<float *>(h->pool + off_R);
    h->mx = reinterpret_cast<float *>(h->pool + h->off_mx);
    h->my = reinterpret_cast<float *>(h->pool + h->off_my);
    h->rho = reinterpret_cast<float *>(h->pool + off_rho);
    h->out = reinterpret_cast<float *>(h->pool + off_out);
    h->cls[0] = reinterpret_cast<uint16_t *>(h->pool + off_cls0);
    h->cls[1] = reinterpret_cast<uint16_t *>(h->pool + off_cls1);
    h->rowflag[0] = reinterpret_cast<uint8_t *>(h->pool + off_rf0);
    h->rowflag[1] = reinterpret_cast<uint8_t *>(h->pool + off_rf1);
    h->chunk_base = reinterpret_cast<uint32_t *>(h->pool + off_cb);
    h->scan_scratch = reinterpret_cast<uint32_t *>(h->pool + off_ss);
    h->mask = reinterpret_cast<uint8_t *>(h->pool + off_mask);
    h->flags = reinterpret_cast<unsigned long long *>(h->pool + h->off_flags);
    h->err_flag = reinterpret_cast<int *>(h->pool + off_err);
    h->red_sums = reinterpret_cast<double *>(h->pool + off_red);
    h->red_max = reinterpret_cast<float *>(h->pool + off_red + 32);
    h->chain_counter = reinterpret_cast<unsigned long long *>(h->pool + off_cnt);
    h->link_done = reinterpret_cast<unsigned int *>(h->pool + off_done);

    int rc = BLBM_OK;
    do {
        cudaError_t ce;
        if ((ce = cudaHostAlloc((void **)&h->mailbox_host, 64, cudaHostAllocMapped)) != cudaSuccess ||
            (ce = cudaHostGetDevicePointer((void **)&h->mailbox_dev, h->mailbox_host, 0)) != cudaSuccess ||
            (ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (ce = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (ce = cudaEventCreate(&h->ev0)) != cudaSuccess || (ce = cudaEventCreate(&h->ev1)) != cudaSuccess ||
            (ce = cudaEventCreateWithFlags(&h->ev_sum, cudaEventDisableTiming)) != cudaSuccess ||
            (ce = cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming)) != cudaSuccess ||
            (ce = cudaHostAlloc((void **)&h->stage_host, blbm::STAGE_SLOTS * blbm::STAGE_PAIRS * 2 * sizeof(uint64_t),
                                cudaHostAllocDefault)) != cudaSuccess ||
            (ce = cudaMemsetAsync(h->pool, 0, h->pool_bytes, h->stream)) != cudaSuccess) {
            rc = fail(BLBM_ECUDA, "handle setup failed: %s", cudaGetErrorString(ce));
            break;
        }
        for (int q = 0; q < blbm::STAGE_SLOTS && rc == BLBM_OK; q++)
            if ((ce = cudaEventCreateWithFlags(&h->stage_ev[q], cudaEventDisableTiming)) != cudaSuccess)
                rc = fail(BLBM_ECUDA, "handle setup failed: %s", cudaGetErrorString(ce));
        if (rc != BLBM_OK) break;
        if ((rc = fill_equilibrium(h, inflow_ux, -1)) != BLBM_OK) break;
        if ((ce = launch_mask_init(h->mask, geom(h), h->stream)) != cudaSuccess ||
            (ce = launch_build_class(h->cls[0], h->mask, geom(h), nullptr, h->rowflag[0], 0, h->rows, h->stream)) !=
                cudaSuccess ||
            (ce = launch_build_class(h->cls[1], h->mask, geom(h), nullptr, h->rowflag[1], 0, h->rows, h->stream)) !=
                cudaSuccess) {
            rc = fail(BLBM_ECUDA, "mask setup failed: %s", cudaGetErrorString(ce));
            break;
        }
        h->launches += 2;
        if ((ce = cudaStreamSynchronize(h->stream)) != cudaSuccess) {
            rc = fail(BLBM_ECUDA, "handle setup failed: %s", cudaGetErrorString(ce));
            break;
        }
    } while (0);
    if (rc != BLBM_OK) {
        blbm_destroy(h);
        return rc;
    }
    *out = h;
    return BLBM_OK;
}

int blbm_create(uint32_t w, uint32_t hgt, float omega, float inflow_ux, int device, blbm_t **out)
{
    return blbm_create_slab(w, hgt, 0, hgt, omega, inflow_ux, device, out);
}

int blbm_destroy(blbm_t *h)
{
    GRP(h, group_destroy(h));
    if (!h) return BLBM_OK;
    cudaSetDevice(h->device);
    if (h->stream && any_peer(h)) {
        // Neighbours store into this pool from their step kernels.  All linked slabs run the same call sequence, so
        // once both neighbours have published our current epoch they have nothing left in flight that targets us;
        // wait for that (bounded: a neighbour that died early must not stall the destructor for long).
        const unsigned long long keep = h->wait_timeout_ns;
        h->wait_timeout_ns = std::min<unsigned long long>(keep, 2ull * 1000ull * 1000ull * 1000ull);
        h->waited = 0;
        sync_peers(h);
        h->wait_timeout_ns = keep;
    }
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->up.ipc_opened) cudaIpcCloseMemHandle(h->up.base);
    if (h->dn.ipc_opened) cudaIpcCloseMemHandle(h->dn.base);
    if (h->stream) {
        stream_free(h->rgb, h->stream);
        stream_free(h->d_pairs, h->stream);
        stream_free(h->chain_idx, h->stream);
        stream_free(h->chain_flag, h->stream);
        stream_free(h->chain_state, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    drop_step_graphs(h);
    if (h->mailbox_host) cudaFreeHost(h->mailbox_host);
    if (h->stage_host) cudaFreeHost(h->stage_host);
    for (int q = 0; q < blbm::STAGE_SLOTS; q++)
        if (h->stage_ev[q]) cudaEventDestroy(h->stage_ev[q]);
    if (h->pool) cudaFree(h->pool);
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamDestroy(h->copy_stream);
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_sum) cudaEventDestroy(h->ev_sum);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return BLBM_OK;
}

int blbm_iterate(blbm_t *h, uint32_t n)
{
    GRP(h, group_steps(h, n, true, true));
    CKH(h);
    int rc = do_steps(h, n);
    if (rc) return rc;
    rc = run_summary(h);
    if (rc) return rc;
    h->frame++;
    return BLBM_OK;
}

int blbm_advance(blbm_t *h, uint32_t n)
{
    GRP(h, group_steps(h, n, false, false));
    CKH(h);
    return do_steps(h, n);
}

int blbm_iterate_timed(blbm_t *h, uint32_t n, float *elapsed_ms)
{
    GRP(h, group_iterate_timed(h, n, elapsed_ms));
    CKH(h);
    if (!elapsed_ms) return fail(BLBM_EINVAL, "elapsed_ms is null");
    CK(cudaEventRecord(h->ev0, h->stream));
    int rc = do_steps(h, n);
    if (rc) return rc;
    rc = run_summary(h);
    if (rc) return rc;
    h->frame++;
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    return check_peer_err(h);
}

int blbm_timer_start(blbm_t *h)
{
    GRP(h, group_timer_start(h));
    CKH(h);
    CK(cudaEventRecord(h->ev0, h->stream));
    return BLBM_OK;
}

int blbm_timer_stop(blbm_t *h, float *elapsed_ms)
{
    GRP(h, group_timer_stop(h, elapsed_ms));
    CKH(h);
    if (!elapsed_ms) return fail(BLBM_EINVAL, "elapsed_ms is null");
    if (h->copy_pending) CK(cudaStreamWaitEvent(h->stream, h->ev_copy, 0));  // the stopwatch covers the copy
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    return check_peer_err(h);
}

int blbm_collide(blbm_t *h)
{
    GRP_EACH(h, blbm_collide(s));
    CKH(h);
    int rc = chain_flush(h);
    if (rc) return rc;
    rc = materialise(h);
    if (rc) return rc;
    const int y = (int)(h->step % 2);
    return launch_step(h, MODE_COLLIDE_ONLY, y, y, true);
}

int blbm_stream(blbm_t *h)
{
    GRP_EACH(h, blbm_stream(s));
    CKH(h);
    int rc = chain_flush(h);
    if (rc) return rc;
    rc = materialise(h);
    if (rc) return rc;
    const int x = (int)(h->step % 2), y = (int)((h->step + 1) % 2);
    rc = launch_step(h, MODE_STREAM_ONLY, x, y, false);
    if (rc) return rc;
    // Like the summary, the public stream half-step READS halo rows without pushing anything: publish an epoch, so
    // that a neighbour running ahead cannot start the next collide - whose boundary rows go into the halo rows this
    // launch gathers from - before it is done.  (materialise()'s stream-only launch needs none: whatever the next
    // pushing launch of a neighbour is, it writes the other buffer's halo rows or waits for a launch of ours that is
    // stream-ordered after it - which keeps read-backs on one slab alone legal.)
    return signal_peers(h);
}

int blbm_set_summary(blbm_t *h, int stat)
{
    GRP_EACH(h, blbm_set_summary(s, stat));
    if (!h) return fail(BLBM_EINVAL, "null handle");
    if (stat < 0 || stat > 4) return fail(BLBM_EINVAL, "stat %d out of range", stat);
    h->stat = stat;
    return BLBM_OK;
}

int blbm_rerender(blbm_t *h)
{
    GRP(h, group_steps(h, 0, true, true));
    CKH(h);
    int rc = run_summary(h);
    if (rc) return rc;
    h->frame++;
    return BLBM_OK;
}

int blbm_compute_summary(blbm_t *h, int stat)
{
    GRP_EACH(h, blbm_compute_summary(s, stat));
    CKH(h);
    int rc = blbm_set_summary(h, stat);
    if (rc) return rc;
    return run_summary(h);
}

int blbm_set_omega(blbm_t *h, float omega)
{
    GRP_EACH(h, blbm_set_omega(s, omega));
    if (!h) return fail(BLBM_EINVAL, "null handle");
    if (memcmp(&h->omega, &omega, sizeof(float)) != 0) h->chain_unsettle = true;  // settled chains hold for one omega
    h->omega = omega;
    return BLBM_OK;
}

int blbm_custom_speed(blbm_t *h, float ux)
{
    GRP_EACH(h, blbm_custom_speed(s, ux));
    CKH(h);
    // collective over linked slabs: wait until the neighbours' last pushes into our halo rows have landed, refill,
    // then publish a new epoch so that no neighbour's next launch stores into our halo rows before our fill ran
    int rc = sync_peers(h);
    if (rc) return rc;
    rc = fill_equilibrium(h, ux, -1);
    if (rc) return rc;
    rc = moments_without_rest(h);
    if (rc) return rc;
    return signal_peers(h);
}

int blbm_reset_to_equilibrium(blbm_t *h) { return blbm_custom_speed(h, 0.1f); }

int blbm_single_cell(blbm_t *h, uint32_t index)
{
    GRP_EACH(h, blbm_single_cell(s, index));
    CKH(h);
    int rc = sync_peers(h);
    if (rc) return rc;
    rc = fill_equilibrium(h, 0.0f, index <= 8 ? (int)index : 9);
    if (rc) return rc;
    return signal_peers(h);  // see blbm_custom_speed
}

int blbm_draw_points64(blbm_t *h, const uint64_t *pairs, size_t npairs)
{
    GRP_EACH(h, blbm_draw_points64(s, pairs, npairs));
    CKH(h);
    if (npairs == 0) return BLBM_OK;
    if (!pairs) return fail(BLBM_EINVAL, "pairs is null");
    // keep, per location, the last pair (sequential scatter semantics), then upload what touches this slab
    std::vector<std::pair<uint64_t, uint64_t>> idx;  // (location, position)
    std::vector<uint64_t> uniq;
    try {  // no C++ exception may cross the C ABI
        idx.reserve(npairs);
        const uint64_t total = (uint64_t)h->W * h->Hg;
        const uint64_t lo = (h->row0 >= 2 ? h->row0 - 2 : 0) * (uint64_t)h->W;
        const uint64_t hi = std::min<uint64_t>(h->Hg, h->row1 + 2) * (uint64_t)h->W;
        for (size_t p = 0; p < npairs; p++) {
            const uint64_t loc = pairs[2 * p];
            if (loc >= total || loc < lo || loc >= hi) continue;
            idx.emplace_back(loc, (uint64_t)p);
        }
        std::sort(idx.begin(), idx.end());
        uniq.reserve(idx.size() * 2);
        for (size_t q = 0; q < idx.size(); q++) {
            if (q + 1 < idx.size() && idx[q + 1].first == idx[q].first) continue;
            uniq.push_back(idx[q].first);
            uniq.push_back(pairs[2 * idx[q].second + 1]);
        }
    } catch (...) {
        return fail(BLBM_ENOMEM, "out of host memory");
    }
    const size_t nu = uniq.size() / 2;
    // owned rows whose class words depend on the painted cells: a cell at row y is an upstream neighbour of
    // rows y-1..y+1, and through the flat-index wrap of column W-1 also of row y-2
    uint32_t rlo = 0, rhi = 0;
    if (nu) {
        const int64_t ymin = (int64_t)(uniq[0] / h->W) - (int64_t)h->row0;           // uniq is sorted by location
        const int64_t ymax = (int64_t)(uniq[2 * (nu - 1)] / h->W) - (int64_t)h->row0;
        rlo = (uint32_t)std::max<int64_t>(0, ymin - 2);
        rhi = (uint32_t)std::max<int64_t>(0, std::min<int64_t>((int64_t)h->rows, ymax + 3));
    }
    if (nu && nu <= SMALL_PAINT_PAIRS && !h->chain_active) {
        // small stroke, no chain table to evict from: the pairs go as kernel arguments, no upload
        SmallPaint sp;
        memcpy(sp.v, uniq.data(), nu * 2 * sizeof(uint64_t));
        int target;
        uint32_t blo, bhi;
        if (plan_class_rebuild(h, rlo, rhi, false, &target, &blo, &bhi)) {
            // mask update and class rebuild in ONE launch (a paint was two launches of ~5 us on a lattice whose
            // 15-step frame is 60 us)
            CK(launch_paint_small(h->mask, geom(h), sp, (uint32_t)nu, h->cls[target], h->rowflag[target], blo, bhi,
                                  h->stream));
        } else {
            CK(launch_mask_scatter_args(h->mask, geom(h), sp, (uint32_t)nu, h->stream));  // halo rows only
        }
        h->launches++;
        return BLBM_OK;
    }
    if (nu) {
        if (h->d_pairs_cap < nu) {
            stream_free(h->d_pairs, h->stream);
            h->d_pairs = nullptr;
            h->d_pairs_cap = 0;
            const size_t cap = std::max<size_t>(nu, 4096);
            cudaError_t e = stream_alloc((void **)&h->d_pairs, cap * 2 * sizeof(uint64_t), h->stream);
            if (e != cudaSuccess) return fail(BLBM_ENOMEM, "allocating the paint list failed");
            h->d_pairs_cap = cap;
        }
        const bool staged = nu <= blbm::STAGE_PAIRS;
        if (staged) {
            // through the pinned ring: the upload is asynchronous and the caller's buffer is free on return
            const int slot = h->stage_next;
            h->stage_next = (slot + 1) % blbm::STAGE_SLOTS;
            if (h->stage_used[slot]) CK(cudaEventSynchronize(h->stage_ev[slot]));  // long done unless 4 paints are queued
            uint64_t *st = h->stage_host + (size_t)slot * blbm::STAGE_PAIRS * 2;
            memcpy(st, uniq.data(), nu * 2 * sizeof(uint64_t));
            CK(cudaMemcpyAsync(h->d_pairs, st, nu * 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
            CK(cudaEventRecord(h->stage_ev[slot], h->stream));
            h->stage_used[slot] = true;
        } else {
            CK(cudaMemcpyAsync(h->d_pairs, uniq.data(), nu * 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
        }
        if (h->chain_active) {
            // cells about to change leave the chain table first (their state returns to the planes)
            CK(launch_chain_evict(chain_table(h), chain_planes(h), h->cls[h->cls_cur], h->cls[h->cls_cur ^ 1], geom(h),
                                  h->d_pairs, nu, (uint32_t)(h->step % 2), h->stream));
            h->launches++;
        }
        CK(launch_mask_scatter(h->mask, geom(h), h->d_pairs, nu, h->stream));
        h->launches++;
        if (!staged) CK(cudaStreamSynchronize(h->stream));  // uniq is pageable host memory owned by this frame
    }
    return rebuild_class(h, rlo, rhi, nu * 100ull >= (unsigned long long)h->rows * h->W);
}

int blbm_draw_points(blbm_t *h, const uint32_t *pairs, size_t npairs)
{
    if (!h) return fail(BLBM_EINVAL, "null handle");
    if (npairs == 0) return BLBM_OK;
    if (!pairs) return fail(BLBM_EINVAL, "pairs is null");
    std::vector<uint64_t> wide;
    try {
        wide.resize(npairs * 2);
    } catch (...) {
        return fail(BLBM_ENOMEM, "out of host memory");
    }
    for (size_t q = 0; q < npairs * 2; q++) wide[q] = pairs[q];
    return blbm_draw_points64(h, wide.data(), npairs);
}

int blbm_reset_barrier(blbm_t *h)
{
    GRP_EACH(h, blbm_reset_barrier(s));
    CKH(h);
    {
        int rcf = chain_flush(h);
        if (rcf) return rcf;
    }
    CK(launch_mask_init(h->mask, geom(h), h->stream));
    h->launches++;
    return rebuild_class(h, 0, h->rows, true);
}

int blbm_write_barrier_rows(blbm_t *h, uint64_t row_begin, uint64_t nrows, const uint8_t *mask)
{
    GRP_EACH(h, blbm_write_barrier_rows(s, row_begin, nrows, mask));
    CKH(h);
    if (!mask && nrows) return fail(BLBM_EINVAL, "mask is null");
    {
        int rcf = chain_flush(h);
        if (rcf) return rcf;
    }
    // intersect with the mask window [row0-2, row1+2) and with the lattice
    const uint64_t win_lo = h->row0 >= 2 ? h->row0 - 2 : 0;
    const uint64_t win_hi = std::min<uint64_t>(h->Hg, h->row1 + 2);
    const uint64_t lo = std::max<uint64_t>(row_begin, win_lo);
    const uint64_t hi = std::min<uint64_t>(row_begin + nrows, win_hi);
    if (lo < hi) {
        const int64_t lr = (int64_t)lo - (int64_t)h->row0;
        CK(cudaMemcpy2DAsync(h->mask + mask_row_off(lr, h->P), h->P, mask + (size_t)(lo - row_begin) * h->W, h->W,
                             h->W, (size_t)(hi - lo), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));  // the caller's buffer is free again on return
    }
    return rebuild_class(h, 0, h->rows, true);
}

uint64_t blbm_get_compute_num(const blbm_t *h) { return !h ? 0 : h->group ? h->group->slabs.front()->step : h->step; }
uint64_t blbm_get_frame_num(const blbm_t *h) { return !h ? 0 : h->group ? h->group->slabs.front()->frame : h->frame; }

static int copy_plane_to_host(blbm *h, const float *plane, float *dst)
{
    CK(cudaMemcpy2DAsync(dst, (size_t)h->W * sizeof(float), plane + row_off(0, h->P), (size_t)h->P * sizeof(float),
                         (size_t)h->W * sizeof(float), h->rows, cudaMemcpyDeviceToHost, h->stream));
    return BLBM_OK;
}

static float *population_plane(blbm *h, int buffer, int k)
{
    static const int dir_of_pop[9] = {D_NW, D_N, D_NE, D_W, -1, D_E, D_SW, D_S, D_SE};
    if (k == BLBM_REST) return h->R;
    if (buffer < 0) buffer = (int)(h->step % 2);
    return h->f[buffer][dir_of_pop[k]];
}

int blbm_read_population(blbm_t *h, int buffer, int k, float *dst)
{
    GRP(h, group_read_rows(h, GR_POPULATION, buffer, k, dst, nullptr, nullptr));
    CKH(h);
    if (!dst || k < 0 || k > 8 || buffer < -1 || buffer > 1) return fail(BLBM_EINVAL, "bad argument");
    int rc = chain_flush(h);
    if (rc) return rc;
    rc = materialise(h);
    if (rc) return rc;
    rc = copy_plane_to_host(h, population_plane(h, buffer, k), dst);
    if (rc) return rc;
    return sync_stream(h);
}

int blbm_write_population(blbm_t *h, int buffer, int k, const float *src)
{
    GRP(h, group_write_population(h, buffer, k, src));
    CKH(h);
    if (!src || k < 0 || k > 8 || buffer < -1 || buffer > 1) return fail(BLBM_EINVAL, "bad argument");
    int rc = chain_flush(h);
    if (rc) return rc;
    rc = materialise(h);
    if (rc) return rc;
    float *plane = population_plane(h, buffer, k);
    CK(cudaMemcpy2DAsync(plane + row_off(0, h->P), (size_t)h->P * sizeof(float), src, (size_t)h->W * sizeof(float),
                         (size_t)h->W * sizeof(float), h->rows, cudaMemcpyHostToDevice, h->stream));
    h->halo_dirty = true;
    return sync_stream(h);
}

int blbm_read_moments(blbm_t *h, float *mx, float *my, float *rho)
{
    GRP(h, group_read_rows(h, GR_MOMENTS, 0, 0, mx, my, rho));
    CKH(h);
    int rc;
    if (mx && (rc = copy_plane_to_host(h, h->mx, mx))) return rc;
    if (my && (rc = copy_plane_to_host(h, h->my, my))) return rc;
    if (rho && (rc = copy_plane_to_host(h, h->rho, rho))) return rc;
    return sync_stream(h);
}

int blbm_read_output(blbm_t *h, float *dst)
{
    GRP(h, group_read_rows(h, GR_OUTPUT, 0, 0, dst, nullptr, nullptr));
    CKH(h);
    if (!dst) return fail(BLBM_EINVAL, "dst is null");
    int rc = copy_plane_to_host(h, h->out, dst);
    if (rc) return rc;
    return sync_stream(h);
}

int blbm_color_map(blbm_t *h, int map)
{
    GRP_EACH(h, blbm_color_map(s, map));
    CKH(h);
    if (map < 0 || map > 2) return fail(BLBM_EINVAL, "colour map %d out of range", map);
    if (!h->rgb) {
        const size_t bytes = (size_t)h->rows * h->W * 3 * sizeof(float);
        if (stream_alloc((void **)&h->rgb, bytes, h->stream) != cudaSuccess) {
            cudaGetLastError();
            h->rgb = nullptr;
            return fail(BLBM_ENOMEM, "allocating %zu bytes for the colour buffer failed", bytes);
        }
    }
    CK(launch_color_map(h->out, h->mask, h->rgb, geom(h), map, h->stream));
    h->launches++;
    return BLBM_OK;
}

int blbm_read_colors(blbm_t *h, float *rgb)
{
    GRP(h, group_read_rows(h, GR_COLORS, 0, 0, rgb, nullptr, nullptr));
    CKH(h);
    if (!rgb) return fail(BLBM_EINVAL, "rgb is null");
    if (!h->rgb) return fail(BLBM_ESTATE, "blbm_color_map has not been called");
    CK(cudaMemcpyAsync(rgb, h->rgb, (size_t)h->rows * h->W * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    return sync_stream(h);
}

int blbm_read_output_async(blbm_t *h, float *pinned_dst)
{
    GRP(h, group_read_rows(h, GR_OUTPUT_ASYNC, 0, 0, pinned_dst, nullptr, nullptr));
    CKH(h);
    if (!pinned_dst) return fail(BLBM_EINVAL, "dst is null");
    // order the copy after everything enqueued so far, but run it on the copy stream so that the steps
    // enqueued next overlap it; the next summary launch waits for it (it overwrites `out`)
    if (h->copy_pending) CK(cudaStreamWaitEvent(h->copy_stream, h->ev_copy, 0));
    CK(cudaEventRecord(h->ev_sum, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_sum, 0));
    CK(cudaMemcpy2DAsync(pinned_dst, (size_t)h->W * sizeof(float), h->out + row_off(0, h->P),
                         (size_t)h->P * sizeof(float), (size_t)h->W * sizeof(float), h->rows, cudaMemcpyDeviceToHost,
                         h->copy_stream));
    CK(cudaEventRecord(h->ev_copy, h->copy_stream));
    h->copy_pending = true;
    return BLBM_OK;
}

int blbm_synchronize(blbm_t *h)
{
    GRP_EACH(h, blbm_synchronize(s));
    CKH(h);
    return sync_stream(h);
}

int blbm_read_barrier(blbm_t *h, uint32_t *dst)
{
    GRP(h, group_read_rows(h, GR_BARRIER, 0, 0, dst, nullptr, nullptr));
    CKH(h);
    if (!dst) return fail(BLBM_EINVAL, "dst is null");
    std::vector<uint8_t> tmp;
    try {
        tmp.resize((size_t)h->rows * h->W);
    } catch (...) {
        return fail(BLBM_ENOMEM, "out of host memory");
    }
    CK(cudaMemcpy2DAsync(tmp.data(), h->W, h->mask + mask_row_off(0, h->P), h->P, h->W, h->rows,
                         cudaMemcpyDeviceToHost, h->stream));
    int rc = sync_stream(h);
    if (rc) return rc;
    for (size_t q = 0; q < tmp.size(); q++) dst[q] = tmp[q];
    return BLBM_OK;
}

int blbm_read_cell_class(blbm_t *h, uint16_t *dst)
{
    GRP(h, group_read_rows(h, GR_CLASS, 0, 0, dst, nullptr, nullptr));
    CKH(h);
    if (!dst) return fail(BLBM_EINVAL, "dst is null");
    // the kernel-facing class words use an internal encoding; the public word is derived from the (current)
    // mask on demand
    const size_t bytes = (size_t)h->rows * h->W * sizeof(uint16_t);
    uint16_t *tmp = nullptr;
    if (stream_alloc((void **)&tmp, bytes, h->stream) != cudaSuccess) {
        cudaGetLastError();
        return fail(BLBM_ENOMEM, "allocating %zu bytes for the class read-back failed", bytes);
    }
    cudaError_t e = launch_build_public_class(tmp, h->mask, geom(h), h->stream);
    h->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, h->stream);
    stream_free(tmp, h->stream);  // stream-ordered: released after the copy, also on the error path
    if (e != cudaSuccess) return fail(BLBM_ECUDA, "class read-back failed: %s", cudaGetErrorString(e));
    return sync_stream(h);
}

int blbm_reduce_moments(blbm_t *h, double *sum_rho, double *sum_mx, double *sum_my, float *max_abs_output)
{
    GRP(h, group_reduce_moments(h, sum_rho, sum_mx, sum_my, max_abs_output));
    CKH(h);
    CK(launch_reduce(h->mx, h->my, h->rho, h->out, geom(h), h->red_sums, h->red_max, h->stream));
    h->launches++;
    double s[3];
    float m;
    CK(cudaMemcpyAsync(s, h->red_sums, sizeof(s), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&m, h->red_max, sizeof(m), cudaMemcpyDeviceToHost, h->stream));
    int rc = sync_stream(h);
    if (rc) return rc;
    if (sum_rho) *sum_rho = s[0];
    if (sum_mx) *sum_mx = s[1];
    if (sum_my) *sum_my = s[2];
    if (max_abs_output) *max_abs_output = m;
    return BLBM_OK;
}

int blbm_get_geometry(const blbm_t *h, uint32_t *w, uint64_t *h_global, uint64_t *row_begin, uint64_t *row_end,
                      int *device)
{
    GRP(h, group_get_geometry(h, w, h_global, row_begin, row_end, device));
    if (!h) return fail(BLBM_EINVAL, "null handle");
    if (w) *w = h->W;
    if (h_global) *h_global = h->Hg;
    if (row_begin) *row_begin = h->row0;
    if (row_end) *row_end = h->row1;
    if (device) *device = h->device;
    return BLBM_OK;
}

int blbm_export_peer(blbm_t *h, void *blob)
{
    GRP(h, fail(BLBM_ESTATE, "a group handle links its slabs itself"));
    CKH(h);
    if (!blob) return fail(BLBM_EINVAL, "blob is null");
    PeerBlob b;
    fill_blob(h, &b);
    CK(cudaIpcGetMemHandle(&b.ipc, h->pool));
    memset(blob, 0, BLBM_PEER_HANDLE_BYTES);
    memcpy(blob, &b, sizeof(b));
    return BLBM_OK;
}

int blbm_link_peer(blbm_t *h, int side, const void *blob)
{
    GRP(h, fail(BLBM_ESTATE, "a group handle links its slabs itself"));
    CKH(h);
    if (!blob || (side != 0 && side != 1)) return fail(BLBM_EINVAL, "bad argument");
    PeerBlob b;
    memcpy(&b, blob, sizeof(b));
    if (b.magic != PEER_MAGIC) return fail(BLBM_EINVAL, "bad peer blob");
    char *base = nullptr;
    bool opened = false;
    if (b.pid == (int32_t)getpid()) {
        base = reinterpret_cast<char *>((uintptr_t)b.local_ptr);
        if (b.device != h->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(BLBM_EPEER, "cudaDeviceEnablePeerAccess(%d) failed: %s", b.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
    } else {
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(BLBM_EPEER, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
        base = static_cast<char *>(p);
        opened = true;
    }
    int rc = attach_peer(h, side, b, base, opened);
    if (rc != BLBM_OK && opened) cudaIpcCloseMemHandle(base);
    if (rc) return rc;
    h->halo_dirty = true;
    return BLBM_OK;
}

int blbm_link_local(blbm_t *upper, blbm_t *lower)
{
    if ((upper && upper->group) || (lower && lower->group)) return fail(BLBM_ESTATE, "a group handle links its slabs itself");
    if (!upper || !lower) return fail(BLBM_EINVAL, "null handle");
    unsigned char bu[BLBM_PEER_HANDLE_BYTES], bl[BLBM_PEER_HANDLE_BYTES];
    PeerBlob b;
    fill_blob(upper, &b);
    memset(bu, 0, sizeof(bu));
    memcpy(bu, &b, sizeof(b));
    fill_blob(lower, &b);
    memset(bl, 0, sizeof(bl));
    memcpy(bl, &b, sizeof(b));
    int rc = blbm_link_peer(upper, 1, bl);
    if (rc) return rc;
    return blbm_link_peer(lower, 0, bu);
}

int blbm_exchange_halos(blbm_t *h)
{
    GRP_EACH(h, blbm_exchange_halos(s));
    CKH(h);
    if (!any_peer(h)) return BLBM_OK;
    int rc = chain_flush(h);
    if (rc) return rc;
    rc = materialise(h);
    if (rc) return rc;
    // A barrier before the pushes: the rows about to be sent may have been rewritten (blbm_write_population), and the
    // neighbour may still have a stream pass queued that gathers the old ones from its halo rows (materialise()
    // publishes nothing).  Our epoch tells the neighbours that everything WE still had to read from our halo rows is
    // done; push_all_halos then waits for theirs.  (Found by tests/test_epoch_protocol_model.py.)
    rc = signal_peers(h);
    if (rc) return rc;
    return push_all_halos(h);
}

int blbm_set_kernel(blbm_t *h, int kernel)
{
    GRP_EACH(h, blbm_set_kernel(s, kernel));
    CKH(h);
    switch (kernel) {
    case BLBM_KERNEL_AUTO: h->kernel = BLBM_KERNEL_VEC4; break;
    case BLBM_KERNEL_SCALAR:
    case BLBM_KERNEL_VEC4: h->kernel = kernel; break;
    default: return fail(BLBM_EINVAL, "kernel %d not available", kernel);
    }
    return BLBM_OK;
}

int blbm_get_kernel(const blbm_t *h) { return !h ? BLBM_EINVAL : h->group ? h->group->slabs.front()->kernel : h->kernel; }

int blbm_set_tuning(blbm_t *h, int knob, int value)
{
    GRP_EACH(h, blbm_set_tuning(s, knob, value));
    if (!h) return fail(BLBM_EINVAL, "null handle");
    switch (knob) {
    case BLBM_TUNE_VEC4_BLOCK_ROWS:
        if (value != 1 && value != 2 && value != 4 && value != 8 && value != 16)
            return fail(BLBM_EINVAL, "block rows must be 1, 2, 4, 8 or 16");
        h->vec4_rows = value;
        return BLBM_OK;
    case BLBM_TUNE_VEC4_DENSE:
        if (value < -1 || value > 3) return fail(BLBM_EINVAL, "dense flavour must be -1 (auto) or 0..3");
        h->vec4_dense = value;
        return BLBM_OK;
    case BLBM_TUNE_VEC4_PACKED:
        if (value != 0 && value != 1) return fail(BLBM_EINVAL, "packed adds must be 0 or 1");
        h->vec4_packed = value;
        return BLBM_OK;
    case BLBM_TUNE_VEC4_INDEX32:
        if (value < -1 || value > 1) return fail(BLBM_EINVAL, "index32 must be -1 (auto), 0 or 1");
        h->vec4_index32 = value;
        return BLBM_OK;
    case BLBM_TUNE_LINK_IN_KERNEL:
        if (value != 0 && value != 1) return fail(BLBM_EINVAL, "in-kernel handshake must be 0 or 1");
        h->link_in_kernel = value;
        return BLBM_OK;
    case BLBM_TUNE_CUDA_GRAPHS:
        if (value < -1 || value > 1) return fail(BLBM_EINVAL, "graphs must be -1 (auto), 0 or 1");
        h->use_graphs = value;
        return BLBM_OK;
    case BLBM_TUNE_PDL:
        if (value < -1 || value > 1) return fail(BLBM_EINVAL, "dependent launch must be -1 (auto), 0 or 1");
        h->use_pdl = value;
        return BLBM_OK;
    default: return fail(BLBM_EINVAL, "unknown tuning knob %d", knob);
    }
}

int blbm_set_lazy_barriers(blbm_t *h, int mode)
{
    GRP_EACH(h, blbm_set_lazy_barriers(s, mode));
    CKH(h);
    if (mode < 0 || mode > 2) return fail(BLBM_EINVAL, "mode %d out of range", mode);
    if (mode == 0) {
        int rc = chain_flush(h);
        if (rc) return rc;
    }
    h->lazy_mode = mode;
    h->chain_declined = false;
    return BLBM_OK;
}

int blbm_get_lazy_barriers_active(const blbm_t *h)
{
    if (h && h->group) {
        for (const blbm *s : h->group->slabs)
            if (s->chain_active) return 1;
        return 0;
    }
    return h && h->chain_active ? 1 : 0;
}
uint64_t blbm_get_launch_count(const blbm_t *h) { return !h ? 0 : h->group ? group_sum(h, GS_LAUNCHES) : h->launches; }
uint64_t blbm_get_device_bytes(const blbm_t *h) { return !h ? 0 : h->group ? group_sum(h, GS_BYTES) : h->pool_bytes; }

}  // extern "C"
