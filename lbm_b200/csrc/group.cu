// group.cu — blbm_create_group and the fan-out of the C ABI over the slabs of a group handle.
//
// The reference's simulation is ONE value (`LBM::new(&driver, omega, x, y)`, lbm-wgpu/src/lbm.rs:726) that the
// caller steps, paints and reads (lbm.rs:1065, :1337).  For lattices that need more than one B200 the value stays
// one handle: the lattice is cut into y-slabs (rows [r*H/n, (r+1)*H/n), contiguous in every plane because
// i = x + y*W), one per device, linked to their neighbours so that the step kernel stores the three crossing
// populations per face straight into the neighbour's halo rows over NVLink (step_common.cuh), and every entry point
// of blbm.h applied to the group handle is applied to all slabs: mutators in lock-step, read-backs concatenated top
// to bottom.
//
// One host thread drives all slabs.  A slab's stream waits on its neighbours' epoch flags every step, so the slabs'
// launch queues are filled in interleaved chunks (Group::chunk steps per slab per turn): no queue ever holds work
// whose matching neighbour launches have not been enqueued within the same turn, and no slab runs further ahead of
// the host than one chunk.  Invariant kept by every function here: when an entry point returns, all slabs have been
// handed the same call sequence — which is what makes the few host synchronisations inside the slab code
// (read-backs, the barrier-chain decision) safe from this single thread.
#include <new>

#include "group.h"

using blbmh::fail;
using blbmh::Group;

namespace blbmh {

int group_destroy(blbm *g)
{
    Group *G = g->group;
    // nothing of any slab may still be storing into a sibling's halo rows when pools are freed
    for (blbm *s : G->slabs) blbm_synchronize(s);
    for (blbm *s : G->slabs) blbm_destroy(s);
    delete G;
    g->group = nullptr;
    delete g;
    return BLBM_OK;
}

// LBM::iterate (lbm.rs:1065-1074) over all slabs: n steps in interleaved chunks, then calculate_summary.  Only the
// call's very last step stores the moments (they are one step behind the populations, as in the reference).
int group_steps(blbm *g, uint32_t n, bool summary, bool count_frame)
{
    Group *G = g->group;
    uint32_t left = n;
    while (left) {
        const uint32_t c = left < G->chunk ? left : G->chunk;
        const bool last = c == left;
        for (blbm *s : G->slabs) {
            const int rc = slab_steps(s, c, last);
            if (rc != BLBM_OK) return rc;
        }
        left -= c;
    }
    if (summary)
        for (blbm *s : G->slabs) {
            const int rc = slab_summary(s);
            if (rc != BLBM_OK) return rc;
            if (count_frame) s->frame++;
        }
    return BLBM_OK;
}

int group_timer_start(blbm *g)
{
    return group_each(g, [](blbm *s) { return blbm_timer_start(s); });
}

// the slowest slab's stopwatch: every slab brackets the same call sequence on its own stream
int group_timer_stop(blbm *g, float *elapsed_ms)
{
    if (!elapsed_ms) return fail(BLBM_EINVAL, "elapsed_ms is null");
    float worst = 0.0f;
    for (blbm *s : g->group->slabs) {
        float ms = 0.0f;
        const int rc = blbm_timer_stop(s, &ms);
        if (rc != BLBM_OK) return rc;
        worst = ms > worst ? ms : worst;
    }
    *elapsed_ms = worst;
    return BLBM_OK;
}

int group_iterate_timed(blbm *g, uint32_t n, float *elapsed_ms)
{
    if (!elapsed_ms) return fail(BLBM_EINVAL, "elapsed_ms is null");
    int rc = group_timer_start(g);
    if (rc != BLBM_OK) return rc;
    rc = group_steps(g, n, true, true);
    if (rc != BLBM_OK) return rc;
    return group_timer_stop(g, elapsed_ms);
}

// read-backs: every slab writes its own rows into the caller's rows x W array
int group_read_rows(blbm *g, int what, int a, int b, void *p0, void *p1, void *p2)
{
    const Group *G = g->group;
    for (size_t q = 0; q < G->slabs.size(); q++) {
        blbm *s = G->slabs[q];
        const size_t cells = (size_t)G->row0[q] * G->W;  // cells above this slab
        auto at = [cells](void *p, size_t elem) -> void * { return p ? static_cast<char *>(p) + cells * elem : nullptr; };
        int rc;
        switch (what) {
        case GR_POPULATION: rc = blbm_read_population(s, a, b, static_cast<float *>(at(p0, 4))); break;
        case GR_MOMENTS:
            rc = blbm_read_moments(s, static_cast<float *>(at(p0, 4)), static_cast<float *>(at(p1, 4)),
                                   static_cast<float *>(at(p2, 4)));
            break;
        case GR_OUTPUT: rc = blbm_read_output(s, static_cast<float *>(at(p0, 4))); break;
        case GR_OUTPUT_ASYNC: rc = blbm_read_output_async(s, static_cast<float *>(at(p0, 4))); break;
        case GR_BARRIER: rc = blbm_read_barrier(s, static_cast<uint32_t *>(at(p0, 4))); break;
        case GR_CLASS: rc = blbm_read_cell_class(s, static_cast<uint16_t *>(at(p0, 2))); break;
        case GR_COLORS: rc = blbm_read_colors(s, static_cast<float *>(at(p0, 12))); break;
        default: rc = fail(BLBM_EINVAL, "bad read selector");
        }
        if (rc != BLBM_OK) return rc;
    }
    return BLBM_OK;
}

int group_write_population(blbm *g, int buffer, int k, const float *src)
{
    if (!src) return fail(BLBM_EINVAL, "bad argument");
    const Group *G = g->group;
    for (size_t q = 0; q < G->slabs.size(); q++) {
        const int rc = blbm_write_population(G->slabs[q], buffer, k, src + (size_t)G->row0[q] * G->W);
        if (rc != BLBM_OK) return rc;
    }
    return BLBM_OK;
}

int group_reduce_moments(blbm *g, double *sum_rho, double *sum_mx, double *sum_my, float *max_abs_output)
{
    double a = 0, b = 0, c = 0;
    float m = 0;
    for (blbm *s : g->group->slabs) {
        double sa, sb, sc;
        float sm;
        const int rc = blbm_reduce_moments(s, &sa, &sb, &sc, &sm);
        if (rc != BLBM_OK) return rc;
        a += sa;
        b += sb;
        c += sc;
        m = sm > m ? sm : m;
    }
    if (sum_rho) *sum_rho = a;
    if (sum_mx) *sum_mx = b;
    if (sum_my) *sum_my = c;
    if (max_abs_output) *max_abs_output = m;
    return BLBM_OK;
}

int group_get_geometry(const blbm *g, uint32_t *w, uint64_t *h_global, uint64_t *row_begin, uint64_t *row_end,
                       int *device)
{
    const Group *G = g->group;
    if (w) *w = G->W;
    if (h_global) *h_global = G->Hg;
    if (row_begin) *row_begin = 0;
    if (row_end) *row_end = G->Hg;
    if (device) *device = G->slabs.front()->device;
    return BLBM_OK;
}

uint64_t group_sum(const blbm *g, int what)
{
    uint64_t t = 0;
    for (const blbm *s : g->group->slabs) t += what == GS_LAUNCHES ? s->launches : (uint64_t)s->pool_bytes;
    return t;
}

}  // namespace blbmh

extern "C" {

int blbm_create_group(uint32_t w, uint64_t hgt, float omega, float inflow_ux, const int *devices, int ndev,
                      blbm_t **out)
{
    if (!out) return fail(BLBM_EINVAL, "out is null");
    *out = nullptr;
    if (!devices || ndev < 1) return fail(BLBM_EINVAL, "need at least one device");
    if (ndev == 1) return blbm_create_slab(w, hgt, 0, hgt, omega, inflow_ux, devices[0], out);
    if (hgt < 2ull * (uint64_t)ndev) return fail(BLBM_EINVAL, "%d slabs need at least %d rows", ndev, 2 * ndev);
    blbm *g = new (std::nothrow) blbm();
    Group *G = new (std::nothrow) Group();
    if (!g || !G) {
        delete g;
        delete G;
        return fail(BLBM_ENOMEM, "out of host memory");
    }
    g->group = G;
    G->W = w;
    G->Hg = hgt;
    int rc = BLBM_OK;
    try {
        // contiguous row ranges whose sizes differ by at most one row
        const uint64_t base = hgt / (uint64_t)ndev, extra = hgt % (uint64_t)ndev;
        uint64_t r = 0;
        for (int q = 0; q < ndev; q++) {
            G->row0.push_back(r);
            r += base + ((uint64_t)q < extra ? 1 : 0);
        }
        G->row0.push_back(r);
        for (int q = 0; q < ndev && rc == BLBM_OK; q++) {
            blbm *s = nullptr;
            rc = blbm_create_slab(w, hgt, G->row0[q], G->row0[q + 1], omega, inflow_ux, devices[q], &s);
            if (rc == BLBM_OK) G->slabs.push_back(s);
        }
        for (size_t q = 0; q + 1 < G->slabs.size() && rc == BLBM_OK; q++)
            rc = blbm_link_local(G->slabs[q], G->slabs[q + 1]);
    } catch (...) {
        rc = fail(BLBM_ENOMEM, "out of host memory");
    }
    if (rc != BLBM_OK) {
        blbmh::group_destroy(g);
        return rc;
    }
    *out = g;
    return BLBM_OK;
}

int blbm_group_size(const blbm_t *h)
{
    if (!h) return 0;
    return h->group ? (int)h->group->slabs.size() : 1;
}

int blbm_group_slab(blbm_t *h, int index, blbm_t **slab)
{
    if (!h || !slab) return fail(BLBM_EINVAL, "null argument");
    if (!h->group) {
        if (index != 0) return fail(BLBM_EINVAL, "slab %d out of range [0,1)", index);
        *slab = h;
        return BLBM_OK;
    }
    if (index < 0 || index >= (int)h->group->slabs.size())
        return fail(BLBM_EINVAL, "slab %d out of range [0,%d)", index, (int)h->group->slabs.size());
    *slab = h->group->slabs[(size_t)index];
    return BLBM_OK;
}

}  // extern "C"
