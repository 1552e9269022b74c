// group.h — internal: a lattice split into y-slabs over several GPUs behind ONE blbm_t (blbm_create_group).
// Every C-ABI entry point of api.cu hands a group handle to the function of the same name here, which fans the
// call out to the slab handles the group owns.  The reference's boundary is a single `LBM` value
// (lbm-wgpu/src/lbm.rs:32-98, constructor :726, iterate :1065); a group keeps it a single value however many
// B200s the lattice needs.
#pragma once
#include "handle.cuh"

namespace blbmh {

struct Group {
    std::vector<blbm *> slabs;   // top to bottom; slab s owns rows [row0[s], row0[s+1])
    std::vector<uint64_t> row0;  // slabs.size() + 1 entries
    uint32_t W = 0;
    uint64_t Hg = 0;
    uint32_t chunk = 32;  // steps enqueued per slab before moving on to the next slab (see group_steps)
};

int group_destroy(blbm *g);
int group_steps(blbm *g, uint32_t n, bool summary, bool count_frame);
int group_iterate_timed(blbm *g, uint32_t n, float *elapsed_ms);
int group_timer_start(blbm *g);
int group_timer_stop(blbm *g, float *elapsed_ms);
int group_read_rows(blbm *g, int what, int a, int b, void *p0, void *p1, void *p2);
int group_write_population(blbm *g, int buffer, int k, const float *src);
int group_reduce_moments(blbm *g, double *sum_rho, double *sum_mx, double *sum_my, float *max_abs_output);
int group_get_geometry(const blbm *g, uint32_t *w, uint64_t *h_global, uint64_t *row_begin, uint64_t *row_end,
                       int *device);
uint64_t group_sum(const blbm *g, int what);

// what-codes of group_read_rows (rows x W arrays, concatenated top to bottom)
enum { GR_POPULATION, GR_MOMENTS, GR_OUTPUT, GR_OUTPUT_ASYNC, GR_BARRIER, GR_CLASS, GR_COLORS };
// what-codes of group_sum
enum { GS_LAUNCHES, GS_BYTES };

// call f(slab) on every slab, stop at the first failure
template <class F>
int group_each(const blbm *g, F f)
{
    for (blbm *s : g->group->slabs) {
        const int rc = f(s);
        if (rc != BLBM_OK) return rc;
    }
    return BLBM_OK;
}

}  // namespace blbmh
