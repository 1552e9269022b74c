// blbm_internal.cuh — device-side data layout, kernel parameter blocks and the BGK collision core
// shared by every step-kernel implementation.
//
// Arithmetic contract (normative semantics block): fp32, every binary op individually
// rounded, IEEE division, association exactly as the reference WGSL writes it.  This translation unit
// is compiled with -fmad=false and the core below additionally spells every op with __f*_rn
// intrinsics, which the compiler never contracts into FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace blbmk {

// Moving populations in the order of the reference's data_buffers minus the rest slot
// (lbm-wgpu/src/lbm.rs:632-640): nw n ne w e sw s se.
enum Dir { D_NW = 0, D_N = 1, D_NE = 2, D_W = 3, D_E = 4, D_SW = 5, D_S = 6, D_SE = 7 };
// travel vector of each population, y grows downwards (north = -W): stream/*.wgsl index helpers
__host__ __device__ constexpr int dir_dx(int d) { return d == D_NW || d == D_W || d == D_SW ? -1 : (d == D_N || d == D_S ? 0 : 1); }
__host__ __device__ constexpr int dir_dy(int d) { return d <= D_NE ? -1 : (d <= D_E ? 0 : 1); }
__host__ __device__ constexpr int dir_opp(int d) { return 7 - d; }

// Internal class word per cell, laid out for the step kernel (the public word of blbm_read_cell_class —
// bit0 barrier, bit1 skipped, bits 2..9 upstream-is-barrier — is derived from the mask on demand):
//   bits 0..7  upstream neighbour of population d is a barrier -> bounce back; ZERO for skipped cells,
//              so the step kernel tests one bit per (cell, direction)
//   bit 8      skipped by stream (barrier | x == 0 | y >= H-1)
//   bit 9      state lives in the barrier-chain table; plane slots are don't-care
//   bit 10     barrier
//   bit 11     scratch mark while a paint evicts cells from the chain table
constexpr uint16_t CLS_UP_MASK = 0xffu;
constexpr uint16_t CLS_SKIP = 1u << 8;
constexpr uint16_t CLS_CHAIN = 1u << 9;
constexpr uint16_t CLS_BARRIER = 1u << 10;
constexpr uint16_t CLS_DIRTY = 1u << 11;
constexpr uint32_t CHAIN_DEAD = 0xffffffffu;
__host__ __device__ constexpr uint16_t cls_upstream_bit(int d) { return (uint16_t)(1u << d); }
// public word (blbm.h)
constexpr uint16_t PUB_BARRIER = 1u, PUB_SKIP = 2u;
__host__ __device__ constexpr uint16_t pub_upstream_bit(int d) { return (uint16_t)(4u << d); }
// cells per row-chunk flag: one byte per 128 cells of a row says "some class word here is non-zero"
constexpr uint32_t CHUNK = 128;

// Where the cells that a neighbouring slab gathers from get mirrored (direct stores into the peer
// GPU's halo rows).  Pointers address x = 0 of the destination row; null = no neighbour on that side.
struct PushTargets {
    // to the slab above: our first row -> its first halo row below; our second row -> its second
    float *up_n, *up_ne, *up_nw;  // full rows
    float *up_w;                  // only x = 0 is ever read there (flat-index wrap of column W-1)
    float *up_nw2;                // second halo row, x = 0 only
    // to the slab below: our last row -> its halo row above
    float *dn_s, *dn_se, *dn_sw;
    // moment rows for the curl stencil (only written by a moment-storing launch)
    float *up_mx, *up_my, *dn_mx, *dn_my;
};

struct StepParams {
    const float *X[8];  // source buffer: post-collision populations T_{k-1} (rows incl. halos)
    float *Y[8];        // destination buffer
    float *R;           // rest population (single array, read-modify-write)
    const uint16_t *cls;
    const uint8_t *rowflag;  // [rows][ceil(P/128)]: 0 = every class word of that 128-cell chunk is 0
    float *mx, *my, *rho;
    uint32_t W;       // cells per row
    uint32_t P;       // row pitch in elements (multiple of 32)
    uint32_t rows;    // rows owned by this slab
    uint32_t pad_;
    uint64_t row0;    // global y of the first owned row
    uint64_t Hg;      // global lattice height
    float omega;
    PushTargets push;
};

// device row index of owned row r in [0, rows): one halo/guard row above
__host__ __device__ inline size_t row_off(uint32_t r, uint32_t P) { return (size_t)(r + 1) * P; }

// ---- BGK collision of one cell, op-for-op the four collide passes -------------------------------
// pre_collision/corner_pre_collision.wgsl:19-21, cardinal_pre_collision.wgsl:19-21,
// collision/corner_collision.wgsl:24-42, cardinal_collision.wgsl:25-43.
// f[] in Dir order, updated in place; rest updated in place; mx/my/rho are the values the reference
// leaves in density_bg (momentum, momentum, density incl. rest).
__device__ __forceinline__ void collide_cell(float (&f)[8], float &rest, const float omega, float &mx,
                                             float &my, float &rho)
{
    const float nw = f[D_NW], n = f[D_N], ne = f[D_NE], w = f[D_W], e = f[D_E], sw = f[D_SW], s = f[D_S],
                se = f[D_SE];
    // corner pre-collision
    float m_x = __fsub_rn(__fsub_rn(__fadd_rn(ne, se), nw), sw);
    float m_y = __fsub_rn(__fsub_rn(__fadd_rn(ne, nw), se), sw);
    float r = __fadd_rn(__fadd_rn(__fadd_rn(ne, se), nw), sw);
    // cardinal pre-collision
    m_x = __fadd_rn(m_x, __fsub_rn(e, w));
    m_y = __fadd_rn(m_y, __fsub_rn(n, s));
    r = __fadd_rn(r, __fadd_rn(__fadd_rn(__fadd_rn(e, n), s), w));
    // corner collision: rho += origin
    r = __fadd_rn(r, rest);
    mx = m_x;
    my = m_y;
    rho = r;
    const float ux = __fdiv_rn(m_x, r);
    const float uy = __fdiv_rn(m_y, r);
    const float k36 = __fmul_rn(1.0f / 36.0f, r);
    const float k9 = __fmul_rn(1.0f / 9.0f, r);
    const float k49 = __fmul_rn(4.0f / 9.0f, r);
    const float ux3 = __fmul_rn(3.0f, ux);
    const float uy3 = __fmul_rn(3.0f, uy);
    const float ux2 = __fmul_rn(ux, ux);
    const float uy2 = __fmul_rn(uy, uy);
    const float uxuy2 = __fmul_rn(__fmul_rn(2.0f, ux), uy);
    const float u2 = __fadd_rn(ux2, uy2);
    const float u215 = __fmul_rn(1.5f, u2);
    const float one_p_ux3 = __fadd_rn(1.0f, ux3);
    const float one_m_ux3 = __fsub_rn(1.0f, ux3);
    const float q_pos = __fmul_rn(4.5f, __fadd_rn(u2, uxuy2));
    const float q_neg = __fmul_rn(4.5f, __fsub_rn(u2, uxuy2));
#define BLBM_RELAX(fi, k, poly) __fadd_rn(fi, __fmul_rn(omega, __fsub_rn(__fmul_rn(k, __fsub_rn(poly, u215)), fi)))
    f[D_NE] = BLBM_RELAX(ne, k36, __fadd_rn(__fadd_rn(one_p_ux3, uy3), q_pos));
    f[D_SE] = BLBM_RELAX(se, k36, __fadd_rn(__fsub_rn(one_p_ux3, uy3), q_neg));
    f[D_NW] = BLBM_RELAX(nw, k36, __fadd_rn(__fadd_rn(one_m_ux3, uy3), q_neg));
    f[D_SW] = BLBM_RELAX(sw, k36, __fadd_rn(__fsub_rn(one_m_ux3, uy3), q_pos));
    // cardinal collision
    rest = __fadd_rn(rest, __fmul_rn(omega, __fsub_rn(__fmul_rn(k49, __fsub_rn(1.0f, u215)), rest)));
    const float ax = __fmul_rn(4.5f, ux2);
    const float ay = __fmul_rn(4.5f, uy2);
    f[D_E] = BLBM_RELAX(e, k9, __fadd_rn(one_p_ux3, ax));
    f[D_W] = BLBM_RELAX(w, k9, __fadd_rn(one_m_ux3, ax));
    f[D_N] = BLBM_RELAX(n, k9, __fadd_rn(__fadd_rn(1.0f, uy3), ay));
    f[D_S] = BLBM_RELAX(s, k9, __fadd_rn(__fsub_rn(1.0f, uy3), ay));
#undef BLBM_RELAX
}

// moments exactly as the two pre-collision passes leave them (no rest term): reset_to_equilibrium /
// custom_speed, lbm.rs:1076-1102
__device__ __forceinline__ void precollision_moments(const float (&f)[8], float &mx, float &my, float &rho)
{
    const float nw = f[D_NW], n = f[D_N], ne = f[D_NE], w = f[D_W], e = f[D_E], sw = f[D_SW], s = f[D_S],
                se = f[D_SE];
    float m_x = __fsub_rn(__fsub_rn(__fadd_rn(ne, se), nw), sw);
    float m_y = __fsub_rn(__fsub_rn(__fadd_rn(ne, nw), se), sw);
    float r = __fadd_rn(__fadd_rn(__fadd_rn(ne, se), nw), sw);
    mx = __fadd_rn(m_x, __fsub_rn(e, w));
    my = __fadd_rn(m_y, __fsub_rn(n, s));
    rho = __fadd_rn(r, __fadd_rn(__fadd_rn(__fadd_rn(e, n), s), w));
}

enum StepMode { MODE_FUSED = 0, MODE_COLLIDE_ONLY = 1, MODE_STREAM_ONLY = 2 };

// force the lazy module load of every kernel of the library (see aux_kernels.cu)
cudaError_t preload_aux_kernels();
cudaError_t preload_step_kernels();
cudaError_t preload_tma_kernels();

// launchers (kernels.cu)
cudaError_t launch_step_scalar(const StepParams &p, int mode, bool store_moments, cudaStream_t st);
cudaError_t launch_step_vec4(const StepParams &p, int mode, bool store_moments, int block_rows, bool dense_obstacles,
                             cudaStream_t st);

// TMA-staged variant (tma_kernel.cu).  The tensor maps are opaque 128-byte blobs owned by the handle:
// 16 population maps (buffer-major, Dir order) and one for the rest plane, encoded for `tile_rows`.
bool tma_available();
bool tma_encode_planes(void *maps16x128, float *const f[2][8], const float *R, void *mapR, uint32_t P,
                       uint32_t dev_rows, int tile_rows);
cudaError_t launch_step_tma(const StepParams &p, int mode, bool store_moments, const void *maps16x128,
                            const void *mapR, int xbuf, int tile_rows, int stages, int ctas_per_sm, cudaStream_t st);

// ---- auxiliary kernels (aux_kernels.cu) ----------------------------------------------------------
// Slab geometry shared by the auxiliary launchers.  Population/moment/class planes have rows+3 device
// rows (1 halo above, 2 below); the barrier mask has rows+4 (2 above, 2 below) so that the class word
// of every owned cell can be derived locally.
struct SlabGeom {
    uint32_t W, P, rows;
    uint64_t row0, Hg;
};
__host__ __device__ inline size_t mask_row_off(int64_t r, uint32_t P) { return (size_t)(r + 2) * P; }

cudaError_t launch_fill_rows(float *const *planes, const float *values, int nplanes, uint32_t W, uint32_t P,
                             uint32_t dev_row_begin, uint32_t dev_row_end, cudaStream_t st);
cudaError_t launch_mask_init(uint8_t *mask, const SlabGeom &g, cudaStream_t st);
cudaError_t launch_mask_scatter(uint8_t *mask, const SlabGeom &g, const uint64_t *pairs, size_t npairs,
                                cudaStream_t st);
// keep_chain (may alias cls, may be null): class words whose CLS_CHAIN bit is carried over.
// rowflag (may be null): the per-chunk "any non-zero class word" bytes are rebuilt alongside.
// Only owned rows [row_begin, row_end) are rebuilt (a paint touches a few rows).
cudaError_t launch_build_class(uint16_t *cls, const uint8_t *mask, const SlabGeom &g, const uint16_t *keep_chain,
                               uint8_t *rowflag, uint32_t row_begin, uint32_t row_end, cudaStream_t st);
// the public class words of blbm_read_cell_class, densely packed rows x W
cudaError_t launch_build_public_class(uint16_t *dst, const uint8_t *mask, const SlabGeom &g, cudaStream_t st);
cudaError_t launch_precollision_moments(const float *const *f8, float *mx, float *my, float *rho, uint32_t W,
                                        uint32_t P, uint32_t dev_row_begin, uint32_t dev_row_end,
                                        cudaStream_t st);
cudaError_t launch_summary(int stat, const float *mx, const float *my, const float *rho, float *out,
                           const SlabGeom &g, cudaStream_t st);
cudaError_t launch_color_map(const float *out, const uint8_t *mask, float *rgb, const SlabGeom &g, int map,
                             cudaStream_t st);
cudaError_t launch_reduce(const float *mx, const float *my, const float *rho, const float *out,
                          const SlabGeom &g, double *sums3, float *maxabs, cudaStream_t st);
// barrier chains (see aux_kernels.cu)
struct ChainPlanes {
    float *f0[8], *f1[8], *R;
};
cudaError_t launch_chain_count(const uint16_t *cls, const SlabGeom &g, unsigned long long *count,
                               unsigned long long *host_mailbox, cudaStream_t st);
cudaError_t launch_chain_build(uint16_t *cls, uint16_t *cls_other, const SlabGeom &g, const ChainPlanes &pl,
                               uint32_t *idx, float *state, size_t cap, unsigned long long *cursor, cudaStream_t st);
cudaError_t launch_chain_flush(const uint32_t *idx, const float *state, size_t n, size_t cap,
                               const ChainPlanes &pl, uint16_t *cls0, uint16_t *cls1, cudaStream_t st);
// a paint is about to change the mask at `pairs` (global location, value): move those cells' chains back
// into the planes and drop them from the table
cudaError_t launch_chain_evict(uint32_t *idx, const float *state, size_t n, size_t cap, const ChainPlanes &pl,
                               uint16_t *cls_cur, uint16_t *cls_other, const SlabGeom &g, const uint64_t *pairs,
                               size_t npairs, cudaStream_t st);
cudaError_t launch_chain_replay(const uint32_t *idx, float *state, size_t n, size_t cap, uint32_t nsteps,
                                uint32_t parity0, float omega, float *mx, float *my, float *rho, cudaStream_t st);

cudaError_t launch_signal(unsigned long long *remote_up, unsigned long long *remote_dn,
                          unsigned long long epoch, cudaStream_t st);
cudaError_t launch_wait(const unsigned long long *from_up, const unsigned long long *from_dn,
                        unsigned long long epoch, int *err_flag, unsigned long long timeout_ns,
                        cudaStream_t st);

}  // namespace blbmk
