/*
 * blbm.h — C ABI of the B200-native lattice update that replaces lbm-wgpu's WGSL pipeline.
 *
 * The drop-in boundary is the reference's `pub struct LBM` (lbm-wgpu/src/lbm.rs:32-98).  Every entry
 * point below names the reference method or pass it replaces; `&Driver` (lbm-wgpu/src/driver.rs:1-7,
 * wgpu device + queue + surface) has no counterpart — the handle owns a CUDA device, one stream and
 * all device buffers.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - Every call returns BLBM_OK (0) or a negative blbm_status; blbm_last_error() gives the text of
 *     the last failure on the calling thread.  The reference panics instead (expect/unwrap).
 *   - A handle is not thread-safe (the reference's mutators take &mut self, single-threaded wasm).
 *   - Mutators enqueue on the handle's stream and return; blbm_read_* and blbm_synchronize wait.  Paint lists of up
 *     to 8192 pairs go through a pinned staging ring, so a frame loop of draw_points -> iterate -> read_output_async
 *     never synchronises the host.  Two calls wait for the stream by design: blbm_write_barrier_rows (the caller's
 *     buffer is free on return) and the first blbm_iterate / blbm_advance after creation, a whole-mask rewrite or a
 *     paint of >= 1 % of the cells, which reads back one 8-byte count to decide whether barrier cells move to the
 *     compact chain table (blbm_set_lazy_barriers).
 *   - BLBM_EPEER (a linked neighbour did not reach the expected halo epoch within 20 s) is terminal: steps enqueued
 *     after the time-out ran on stale halo rows and pushed their results on; destroy the linked handles.
 *   - Host buffers passed in are copied before the call returns (like queue.write_buffer,
 *     lbm.rs:1342-1343); the caller keeps ownership.
 *   - Cell index i = x + y*W, y = 0 is the top row, north = -W, east = +1 (lbm.rs:607-609).
 *   - Population order (lbm.rs:632-640): 0 nw, 1 n, 2 ne, 3 w, 4 rest, 5 e, 6 sw, 7 s, 8 se.
 *   - There is no CPU fallback: without a CUDA device every call fails with BLBM_ENOGPU.
 *
 * A handle simulates a *slab*: rows [row_begin, row_end) of a W x H_global lattice on one GPU.
 * blbm_create() makes the whole-lattice slab.  Slabs of one lattice are linked to their neighbours
 * (same process: blbm_link_local; other process: blbm_export_peer / blbm_link_peer) and then exchange
 * the one-row halos of the three crossing populations per face by direct NVLink stores issued from
 * the step kernel itself.  Read-backs and writes address the slab's own rows only.  Linked slabs advance an
 * epoch counter per halo-exchanging launch and per summary: every slab of a lattice must be given the same
 * sequence of stepping, half-step, reset and summary calls (a group handle does that by construction).
 */
#ifndef BLBM_H
#define BLBM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct blbm_handle blbm_t;

typedef enum blbm_status {
    BLBM_OK = 0,
    BLBM_EINVAL = -1, /* bad argument */
    BLBM_ENOMEM = -2, /* host or device allocation failed */
    BLBM_ECUDA = -3,  /* a CUDA call failed; see blbm_last_error() */
    BLBM_ENOGPU = -4, /* no usable CUDA device */
    BLBM_ESTATE = -5, /* call not valid in the handle's current state */
    BLBM_EPEER = -6   /* halo peer missing / timed out */
} blbm_status;

/* SummaryStat, lbm.rs:10-16 (same order) */
typedef enum blbm_stat { BLBM_CURL = 0, BLBM_UX = 1, BLBM_UY = 2, BLBM_RHO = 3, BLBM_SPEED = 4 } blbm_stat;

/* index into data_buffers[b][k], lbm.rs:632-640 */
typedef enum blbm_pop {
    BLBM_NW = 0, BLBM_N = 1, BLBM_NE = 2, BLBM_W = 3, BLBM_REST = 4,
    BLBM_E = 5, BLBM_SW = 6, BLBM_S = 7, BLBM_SE = 8
} blbm_pop;

/* step-kernel implementations (all bit-identical; selectable for A/B measurement) */
typedef enum blbm_kernel {
    BLBM_KERNEL_AUTO = 0,
    BLBM_KERNEL_SCALAR = 1, /* one cell per thread, 32-bit accesses */
    BLBM_KERNEL_VEC4 = 2    /* four cells per thread, 128-bit accesses, shuffle-realigned x+-1 gathers, the
                               bounce-back's own rows staged in shared memory with cp.async (default) */
    /* 3 was a TMA-staged variant (persistent CTAs, cp.async.bulk.tensor tiles behind an mbarrier ring): bit-identical
       but 0.72-0.80 of the HBM peak against 0.95-1.0 for VEC4 on every workload (profiles/r1/tma_ab_*.json,
       ncu_step_tma_channel16384.txt) — retired in round 2; blbm_set_kernel(3) fails with BLBM_EINVAL */
} blbm_kernel;

#define BLBM_PEER_HANDLE_BYTES 512

const char *blbm_last_error(void);
int blbm_abi_version(void);
/* number of CUDA devices visible, or a negative blbm_status */
int blbm_device_count(void);

/* ---- life cycle ------------------------------------------------------------------------------ */

/* LBM::new(&driver, omega, x, y), lbm.rs:726 — populations = set_equil(inflow_ux, 0, 1) (lbm.rs:611-643;
 * the reference hard-codes inflow_ux = 0.1, lbm.rs:740), barrier = init_barrier (rows 0 and H-1,
 * lbm.rs:595-605), moments/output zero, compute_step = 0, summary = Curl. */
int blbm_create(uint32_t w, uint32_t h, float omega, float inflow_ux, int device, blbm_t **out);

/* Same, for rows [row_begin, row_end) of a w x h_global lattice (y-slab decomposition; new — the
 * reference is single-device).  h_global may exceed 2^32 / w cells in total. */
int blbm_create_slab(uint32_t w, uint64_t h_global, uint64_t row_begin, uint64_t row_end, float omega,
                     float inflow_ux, int device, blbm_t **out);

/* LBM::new for lattices that need more than one GPU (lbm.rs:726; the reference is single-device): the lattice is
 * cut into ndev y-slabs (contiguous row ranges whose sizes differ by at most one row, top slab on devices[0]), one
 * per listed device (a device may be listed more than once), linked to their neighbours, and returned as ONE handle.
 * Every entry point of this header accepts it: mutators are applied to all slabs in lock-step (one host thread
 * enqueues the slabs' work in interleaved chunks), read-backs return the whole lattice, rows x W, top to bottom;
 * blbm_timer_stop / blbm_iterate_timed report the slowest slab.  ndev == 1 is blbm_create on devices[0].  Peer access
 * between the listed devices must be possible (NVLink / NVSwitch on a B200 node).  Group handles cannot be linked
 * further (blbm_export_peer / blbm_link_* fail with BLBM_ESTATE). */
int blbm_create_group(uint32_t w, uint64_t h, float omega, float inflow_ux, const int *devices, int ndev,
                      blbm_t **out);
/* number of slabs behind a handle (1 for plain handles) and access to one of them, e.g. for per-slab tuning or
 * geometry; the slab stays owned by the group */
int blbm_group_size(const blbm_t *h);
int blbm_group_slab(blbm_t *h, int index, blbm_t **slab);

/* Linked slabs (and group handles) must not be destroyed while a neighbour still has steps in flight that store into
 * them: destroy waits (bounded) for both neighbours to reach this slab's epoch before it frees the pool. */
int blbm_destroy(blbm_t *h);

/* ---- the per-timestep update ----------------------------------------------------------------- */

/* LBM::iterate(&driver, n) minus colour map and render, lbm.rs:1065-1074: n x (collide; stream;
 * compute_step += 1), then calculate_summary with the current stat; frame_number += 1. */
int blbm_iterate(blbm_t *h, uint32_t n);

/* n x compute_step() (lbm.rs:1112-1116) and nothing else: no summary, frame_number unchanged.  The
 * moments of the last step are stored, as after iterate. */
int blbm_advance(blbm_t *h, uint32_t n);

/* Same work as blbm_iterate, bracketed by CUDA events on the handle's stream; waits for completion and
 * returns the device time in milliseconds.  (Measurement hook; not in the reference.) */
int blbm_iterate_timed(blbm_t *h, uint32_t n, float *elapsed_ms);

/* Device stopwatch on the handle's stream: start records an event, stop records another, waits for it
 * and returns the milliseconds between them — everything enqueued in between, copies included. */
int blbm_timer_start(blbm_t *h);
int blbm_timer_stop(blbm_t *h, float *elapsed_ms);

/* pub fn collide / pub fn stream, lbm.rs:1118-1134: one half-step on the live buffer pair; neither
 * touches compute_step (only compute_step() does, lbm.rs:1112-1116). */
int blbm_collide(blbm_t *h);
int blbm_stream(blbm_t *h);

/* LBM::set_summary, lbm.rs:1061 */
int blbm_set_summary(blbm_t *h, int stat);
/* LBM::rerender minus colour map / render, lbm.rs:1104-1110: calculate_summary only; frame_number += 1 */
int blbm_rerender(blbm_t *h);
/* set_summary + calculate_summary without touching frame_number (convenience) */
int blbm_compute_summary(blbm_t *h, int stat);

/* LBM::update_omega_buffer, lbm.rs:1358 — takes effect at the next collide */
int blbm_set_omega(blbm_t *h, float omega);

/* LBM::reset_to_equilibrium (lbm.rs:1076) and LBM::custom_speed (lbm.rs:1090): both population buffers
 * <- set_equil(0.1 | ux, 0, 1), compute_step = frame_number = 0, then the two pre-collision passes only
 * (moments WITHOUT the rest population). */
int blbm_reset_to_equilibrium(blbm_t *h);
int blbm_custom_speed(blbm_t *h, float ux);

/* LBM::single_cell(index), lbm.rs:1482-1515: set_equil(0,0,1) and population `index` = 4.0 at a fixed
 * cell; counters reset; moments untouched.  index > 8 only resets. */
int blbm_single_cell(blbm_t *h, uint32_t index);

/* ---- barrier mask ---------------------------------------------------------------------------- */

/* LBM::draw_shape -> draw_barrier_updates + barrier_draw.wgsl, lbm.rs:1337-1356: pairs is
 * [loc0, val0, loc1, val1, ...] as produced by get_points_vector (merge_shapes.rs:12-22); loc is a
 * GLOBAL cell index, val 1 = barrier, anything else = fluid (stored as 0: barrier_draw.wgsl stores the raw value, but
 * every consumer — the stream passes, the colour maps — only tests == 1, and the reference's own callers only ever
 * produce 0 / 1; blbm_read_barrier therefore returns 0 / 1 words); out-of-range locations are dropped.  Duplicate
 * locations: last pair wins (the reference leaves it to the GPU's scatter order; Blob keys are unique).
 * npairs == 0 is a no-op (the reference underflows, lbm.rs:1343; its callers guard). */
int blbm_draw_points(blbm_t *h, const uint32_t *loc_val_pairs, size_t npairs);
/* 64-bit locations for lattices of 2^32 cells and more */
int blbm_draw_points64(blbm_t *h, const uint64_t *loc_val_pairs, size_t npairs);
/* LBM::reset_barrier, lbm.rs:1362-1365 */
int blbm_reset_barrier(blbm_t *h);
/* Whole-row mask upload — what reset_barrier does with queue.write_buffer(&barrier_buffer, ..)
 * (lbm.rs:1362-1365), for an arbitrary mask: rows [row_begin, row_begin + nrows) of the GLOBAL lattice,
 * nrows x W bytes, 1 = barrier.  Rows outside this slab's window (own rows +-2) are ignored, so every
 * slab may be handed the same data or just its own window. */
int blbm_write_barrier_rows(blbm_t *h, uint64_t row_begin, uint64_t nrows, const uint8_t *mask);

/* ---- barrier shapes: the host-side rasteriser in front of draw_points (barrier_shapes/line.rs) ------------ */

/* Line::new (erase = 0, line.rs:22-54) / Line::new_erased (erase != 0, line.rs:56-87, 30 wide) without a handle:
 * the distinct cells of the thick line between two end points on an xdim x ydim lattice, as (x, y) pairs in
 * xy[0 .. 2*min(count, capacity)), ordered by (x, y).  BLBM_EINVAL when an end point lies outside the lattice
 * (the reference returns Err there).  Pure host code: works without a GPU. */
int blbm_rasterize_line(int64_t x1, int64_t y1, int64_t x2, int64_t y2, int64_t xdim, int64_t ydim, int erase,
                        int64_t *xy, size_t capacity, size_t *count);
/* draw_shape(&Line::new(..)) / draw_shape(&Line::new_erased(..)) */
int blbm_draw_line(blbm_t *h, int64_t x1, int64_t y1, int64_t x2, int64_t y2);
int blbm_erase_line(blbm_t *h, int64_t x1, int64_t y1, int64_t x2, int64_t y2);
/* LBM::curl_barrier / chaos_barrier / welcome_barrier, lbm.rs:1367-1480 (callers reset_barrier first, lib.rs:120-128) */
typedef enum blbm_preset { BLBM_PRESET_CURL = 0, BLBM_PRESET_CHAOS = 1, BLBM_PRESET_WELCOME = 2 } blbm_preset;
/* The end points (x1, y1, x2, y2 per line, in the order the reference creates them) of the thick lines a preset
 * consists of on an xdim x ydim lattice; count = number of lines.  Pure host code: works without a GPU. */
int blbm_preset_lines(int preset, int64_t xdim, int64_t ydim, int64_t *xyxy, size_t capacity, size_t *count);
int blbm_curl_barrier(blbm_t *h);
int blbm_chaos_barrier(blbm_t *h);
int blbm_welcome_barrier(blbm_t *h);

/* ---- counters, lbm.rs:1166-1172 ---------------------------------------------------------------- */
uint64_t blbm_get_compute_num(const blbm_t *h);
uint64_t blbm_get_frame_num(const blbm_t *h);

/* ---- read-back / restore (absent in the reference, which never reads back; required by "read back
 *      macroscopic fields" and by the parity tests).  All arrays are rows x W of the slab's own rows,
 *      densely packed.  dst/src are HOST pointers. ------------------------------------------------ */

/* data_buffers[buffer][k]; buffer 0/1, or -1 = the live one (compute_step % 2).  k = 4 (rest) always
 * addresses the single rest array (the reference binds data_buffers[0][4] only, lbm.rs:775-778).
 * Contents are bit-identical to what the reference's buffers hold after the same call sequence,
 * including the stale values at cells the stream passes skip. */
int blbm_read_population(blbm_t *h, int buffer, int k, float *dst);
int blbm_write_population(blbm_t *h, int buffer, int k, const float *src);
/* density_bg = (momentum x, momentum y, density) of the state before the last collide; any may be NULL */
int blbm_read_moments(blbm_t *h, float *mx, float *my, float *rho);
/* output_bg, the array the colour map reads */
int blbm_read_output(blbm_t *h, float *dst);
/* barrier_buffer as the reference's u32 0/1 words */
int blbm_read_barrier(blbm_t *h, uint32_t *dst);
/* per-cell classification: bit0 barrier, bit1 skipped by stream (barrier | x==0 | y>=H-1), bits 2..9
 * "upstream neighbour is a barrier" for nw n ne w e sw s se */
int blbm_read_cell_class(blbm_t *h, uint16_t *dst);

/* ColorMap, lbm.rs:18-24 (same order) */
typedef enum blbm_colormap { BLBM_INFERNO = 0, BLBM_VIRIDIS = 1, BLBM_JET = 2 } blbm_colormap;
/* LBM::color_map, lbm.rs:1299-1335 (color_map/{inferno,viridis,jet}.wgsl): piece-wise linear LUT of the output
 * field into RGB, barrier cells black.  The colours stay on the device (rows x W x 3 fp32, densely packed — the
 * reference's array<vec3<f32>> has a 16-byte stride in a buffer sized for 12, lbm.rs:193; not reproduced). */
int blbm_color_map(blbm_t *h, int map);
int blbm_read_colors(blbm_t *h, float *rgb);

/* Asynchronous variants for frame pipelines: dst must be page-locked (cudaHostAlloc / pinned torch
 * tensor); the copy is ordered after everything enqueued so far and overlaps later work.  Complete
 * after blbm_synchronize(). */
int blbm_read_output_async(blbm_t *h, float *pinned_dst);
int blbm_synchronize(blbm_t *h);

/* Global reductions over the slab's rows (warp-shuffle + one atomic per block): sum of density, of
 * both momenta, and max |output|.  sums are double.  (Diagnostics; not in the reference.) */
int blbm_reduce_moments(blbm_t *h, double *sum_rho, double *sum_mx, double *sum_my, float *max_abs_output);

/* ---- slab geometry and halo peers ------------------------------------------------------------ */
int blbm_get_geometry(const blbm_t *h, uint32_t *w, uint64_t *h_global, uint64_t *row_begin, uint64_t *row_end,
                      int *device);

/* Link two slabs living in the SAME process: `upper` owns the rows just above `lower`. */
int blbm_link_local(blbm_t *upper, blbm_t *lower);
/* Different processes (one per GPU): export this slab's halo window as an opaque blob
 * (BLBM_PEER_HANDLE_BYTES bytes, CUDA IPC inside), ship it with any transport, and link the blob of the
 * slab above (side = 0) or below (side = 1). */
int blbm_export_peer(blbm_t *h, void *blob);
int blbm_link_peer(blbm_t *h, int side, const void *blob);
/* Re-send the boundary rows of both population buffers and of the moments to the linked neighbours
 * (collective: every slab of the lattice calls it).  Needed only after blbm_write_population; steps,
 * resets and paints keep halos current by themselves. */
int blbm_exchange_halos(blbm_t *h);

/* ---- tuning / measurement -------------------------------------------------------------------- */
int blbm_set_kernel(blbm_t *h, int kernel); /* blbm_kernel */
int blbm_get_kernel(const blbm_t *h);       /* the resolved implementation (never AUTO) */
/* launch-shape knobs for A/B measurement; results never depend on them */
typedef enum blbm_tune {
    BLBM_TUNE_VEC4_BLOCK_ROWS = 0, /* rows of 128 cells per block: 1, 2, 4 (default), 8, 16 */
    /* 1..3 belonged to the retired TMA-staged kernel */
    BLBM_TUNE_VEC4_DENSE = 4,      /* bounce-back fix-up flavour: -1 auto (default), 0 sparse, 1 dense obstacles,
                                      2 dense + own-row vectors staged in shared memory with cp.async (what auto
                                      picks wherever the barrier-chain table is active), 3 as 2 with the class
                                      words read without the chunk-flag test */
    BLBM_TUNE_CUDA_GRAPHS = 5,     /* replay runs of 2..16 steps (and a call's moment-storing last step) per CUDA-graph launch: -1 auto (lattices <= 4 Mi cells), 0, 1 */
    BLBM_TUNE_VEC4_PACKED = 6,     /* collide cell pairs with packed fp32 adds (sm_100 FADD2): 0 (default) or 1; same bits,
                                      measured slower (register pressure) */
    BLBM_TUNE_VEC4_INDEX32 = 7,    /* 32-bit plane offsets where a plane has < 2^32 elements: -1 auto (default, = 1), 0, 1 */
    BLBM_TUNE_PDL = 9,             /* programmatic dependent launch of the fused vec4 step (its blocks are scheduled while the
                                      previous step drains): -1 auto (default: on, except inside the CUDA graphs of small lattices), 0, 1;
                                      unlinked handles only */
    BLBM_TUNE_LINK_IN_KERNEL = 8   /* linked slabs: 1 (default) the fused step kernel waits for / publishes the halo
                                      epochs itself (face row blocks first, interior rows overlap the exchange);
                                      0 one-thread wait and signal kernels around every launch */
} blbm_tune;
int blbm_set_tuning(blbm_t *h, int knob, int value);
/* Barrier cells are isolated (nothing reads them; the reference merely keeps colliding their stale
 * copies; the collision shaders have no mask test), so their state can live in a compact side table that is advanced in
 * registers, which removes their memory traffic from the step kernel.  Results are bit-identical either
 * way.  mode: 0 never, 1 always, 2 auto (default: when >= 2 % of the slab's cells are barriers). */
int blbm_set_lazy_barriers(blbm_t *h, int mode);
int blbm_get_lazy_barriers_active(const blbm_t *h);
/* kernels launched by this handle since creation (the bench's gpu_launches claim) */
uint64_t blbm_get_launch_count(const blbm_t *h);
/* device bytes held by the handle */
uint64_t blbm_get_device_bytes(const blbm_t *h);

#ifdef __cplusplus
}
#endif
#endif /* BLBM_H */
