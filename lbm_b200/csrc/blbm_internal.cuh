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
//   bit 12     the cell owns a slot of the barrier-chain table (live: bit 9 set too; evicted by a paint: bit 9
//              clear).  The table is ordered by plane offset, so the slot of a cell is the number of slot bits in
//              front of it: a per-chunk base (one u32 per 128 cells) plus a prefix count inside the chunk.
constexpr uint16_t CLS_UP_MASK = 0xffu;
constexpr uint16_t CLS_SKIP = 1u << 8;
constexpr uint16_t CLS_CHAIN = 1u << 9;
constexpr uint16_t CLS_BARRIER = 1u << 10;
constexpr uint16_t CLS_SLOT = 1u << 12;
// per-entry flag byte of the chain table
constexpr uint8_t CHAIN_F_DEAD = 1u;     // evicted by a paint: the slot stays (ranks must not shift), the state is gone
constexpr uint8_t CHAIN_F_SETTLED = 2u;  // bitwise period-2 cycle reached: nothing left to compute until omega changes
// rows of the chain table (each `cap` floats long): 0..7 buffer-0 populations (Dir order), 8..15 buffer-1,
// 16/17 the rest population as it stands before a collide of the buffer-0 / buffer-1 copy,
// 18..20 / 21..23 (mx, my, rho) of the latest collide of the buffer-0 / buffer-1 copy
constexpr int CHAIN_ROW_REST = 16, CHAIN_ROW_MOM = 18, CHAIN_ROWS = 24;
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

// Epoch handshake with the neighbouring slabs, executed INSIDE the fused vec4 step kernel (sig_epoch != 0): only the
// row blocks at a slab face read halo rows or store into a neighbour's; they wait (one thread spins on our flag word
// with ld.acquire.sys) until that neighbour has published wait_epoch, and after their last store they arrive on a
// per-face counter whose last arrival publishes sig_epoch to the neighbour with st.release.sys.  Those row blocks
// are scheduled first, so the flags go out at the start of a launch and interior rows overlap the exchange.
struct LinkSync {
    const unsigned long long *wait_up, *wait_dn;  // our flag words, written by the slab above / below (null: none)
    unsigned long long *sig_up, *sig_dn;          // the neighbours' flag words for us
    unsigned long long wait_epoch, sig_epoch;
    unsigned int *done;  // [0] arrivals of the up-face blocks of this launch, [1] of the down-face blocks
    int *err_flag;
    unsigned long long timeout_ns;
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
    // barrier-chain table, needed by a moment-storing launch only: slot base per (row, 128-cell chunk) and the
    // (mx, my, rho) rows (each chain_cap long) of the collide that the launch's moments belong to
    const uint32_t *chunk_base;
    const float *chain_mom;
    uint32_t chain_cap;
    LinkSync link;
};

// device row index of owned row r in [0, rows): one halo/guard row above
__host__ __device__ inline size_t row_off(uint32_t r, uint32_t P) { return (size_t)(r + 1) * P; }

// ---- contraction twin (-DBLBM_CONTRACT -> libblbm_contract.so) -----------------------------------
// The default build rounds every multiply and every add separately (the defined parity semantics; what lavapipe
// does).  Should a WebGPU backend that fuses ever be the parity target, the twin applies the rule an LLVM-style
// backend applies to each shader after CSE — a multiply with exactly one use is fused into the add/sub that uses
// it (left operand first when both are multiplies) — with explicit __fmaf_rn (still -fmad=false: nothing else
// may fuse):
//   corner_collision.wgsl    u2 = fma(ux, ux, uy*uy);  f += w*(k*(p-u215)-f)  ->  fma(w, fma(k, p-u215, -f), f)
//   cardinal_collision.wgsl  the relaxations as above; u2 stays a plain sum there (ux2, uy2 have two uses each)
//   rho.wgsl  fma(4, clamp(..), -0.5);  speed.wgsl  sqrt(fma(mx, mx, my*my));  color maps  fma(lw, A, rw*B)
// ux3, uy3, uxuy2, u215, 4.5*(..), 4.5*ux2 have several uses after CSE and keep their own rounding.
// tests/test_contract_twin.py runs the parity suite on the twin pair.
#ifdef BLBM_CONTRACT
__device__ __forceinline__ float mul_add(const float a, const float b, const float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float mul_sub(const float a, const float b, const float c) { return __fmaf_rn(a, b, -c); }
#else
__device__ __forceinline__ float mul_add(const float a, const float b, const float c) { return __fadd_rn(__fmul_rn(a, b), c); }
__device__ __forceinline__ float mul_sub(const float a, const float b, const float c) { return __fsub_rn(__fmul_rn(a, b), c); }
#endif

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
#ifdef BLBM_CONTRACT
    // corner shader: ux2 and uy2 have one use each there, so u2 = fma(ux, ux, uy2); the cardinal shader uses both
    // squares twice and keeps the plain sum above
    const float u2c = __fmaf_rn(ux, ux, uy2);
    const float u215c = __fmul_rn(1.5f, u2c);
#else
    const float u2c = u2, u215c = u215;
#endif
    const float one_p_ux3 = __fadd_rn(1.0f, ux3);
    const float one_m_ux3 = __fsub_rn(1.0f, ux3);
    const float q_pos = __fmul_rn(4.5f, __fadd_rn(u2c, uxuy2));
    const float q_neg = __fmul_rn(4.5f, __fsub_rn(u2c, uxuy2));
    // f += omega * (k * (poly - u215) - f): two multiplies, each consumed once (mul_sub / mul_add above)
#define BLBM_RELAX(fi, k, poly, u215x) mul_add(omega, mul_sub(k, __fsub_rn(poly, u215x), fi), fi)
    f[D_NE] = BLBM_RELAX(ne, k36, __fadd_rn(__fadd_rn(one_p_ux3, uy3), q_pos), u215c);
    f[D_SE] = BLBM_RELAX(se, k36, __fadd_rn(__fsub_rn(one_p_ux3, uy3), q_neg), u215c);
    f[D_NW] = BLBM_RELAX(nw, k36, __fadd_rn(__fadd_rn(one_m_ux3, uy3), q_neg), u215c);
    f[D_SW] = BLBM_RELAX(sw, k36, __fadd_rn(__fsub_rn(one_m_ux3, uy3), q_pos), u215c);
    // cardinal collision
    rest = BLBM_RELAX(rest, k49, 1.0f, u215);
    const float ax = __fmul_rn(4.5f, ux2);
    const float ay = __fmul_rn(4.5f, uy2);
    f[D_E] = BLBM_RELAX(e, k9, __fadd_rn(one_p_ux3, ax), u215);
    f[D_W] = BLBM_RELAX(w, k9, __fadd_rn(one_m_ux3, ax), u215);
    f[D_N] = BLBM_RELAX(n, k9, __fadd_rn(__fadd_rn(1.0f, uy3), ay), u215);
    f[D_S] = BLBM_RELAX(s, k9, __fadd_rn(__fsub_rn(1.0f, uy3), ay), u215);
#undef BLBM_RELAX
}

// ---- the same collision on TWO cells at once with Blackwell's packed fp32 adds -------------------------
// sm_100 has add/sub/mul/fma on f32x2 register pairs (SASS FADD2/FMUL2/FFMA2): one issue slot, two individually
// rounded fp32 results.  Every add and sub of collide_cell() is done packed here (62 of the 94 fp32 ops per
// cell), which removes a quarter of the step kernel's instructions at bit-identical results.  The multiplies
// stay scalar on purpose: ptxas 12.9 contracts mul.rn.f32x2 (and even fma.rn.f32x2 with a -0.0 addend)
// followed by add/sub.rn.f32x2 into one FFMA2 although the roundings are explicit, which would break the
// "every op individually rounded" contract; scalar FMUL feeding FADD2 is never contracted (the build checks
// the SASS of the step kernels for FFMA2/FMUL2: `make sass-check`).  lo half = first cell, hi half = second.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(const float lo, const float hi)
{
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(const f32x2_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t add2(const f32x2_t a, const f32x2_t b)
{
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t sub2(const f32x2_t a, const f32x2_t b)
{
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t mul2(const f32x2_t a, const f32x2_t b)
{
    float al, ah, bl, bh;
    upk2(a, al, ah);
    upk2(b, bl, bh);
    return pk2(__fmul_rn(al, bl), __fmul_rn(ah, bh));
}
__device__ __forceinline__ f32x2_t mulc2(const float c, const f32x2_t a)
{
    float al, ah;
    upk2(a, al, ah);
    return pk2(__fmul_rn(c, al), __fmul_rn(c, ah));
}
__device__ __forceinline__ f32x2_t div2(const f32x2_t a, const f32x2_t b)
{
    float al, ah, bl, bh;
    upk2(a, al, ah);
    upk2(b, bl, bh);
    return pk2(__fdiv_rn(al, bl), __fdiv_rn(ah, bh));
}

// collide_cell() on the cell pair held in the lo/hi halves; same operations in the same association
__device__ __forceinline__ void collide_pair(f32x2_t (&f)[8], f32x2_t &rest, const float omega, f32x2_t &mx,
                                             f32x2_t &my, f32x2_t &rho)
{
    const f32x2_t nw = f[D_NW], n = f[D_N], ne = f[D_NE], w = f[D_W], e = f[D_E], sw = f[D_SW], s = f[D_S],
                  se = f[D_SE];
    const f32x2_t one = pk2(1.0f, 1.0f);
    f32x2_t m_x = sub2(sub2(add2(ne, se), nw), sw);
    f32x2_t m_y = sub2(sub2(add2(ne, nw), se), sw);
    f32x2_t r = add2(add2(add2(ne, se), nw), sw);
    m_x = add2(m_x, sub2(e, w));
    m_y = add2(m_y, sub2(n, s));
    r = add2(r, add2(add2(add2(e, n), s), w));
    r = add2(r, rest);
    mx = m_x;
    my = m_y;
    rho = r;
    const f32x2_t ux = div2(m_x, r), uy = div2(m_y, r);
    const f32x2_t k36 = mulc2(1.0f / 36.0f, r), k9 = mulc2(1.0f / 9.0f, r), k49 = mulc2(4.0f / 9.0f, r);
    const f32x2_t ux3 = mulc2(3.0f, ux), uy3 = mulc2(3.0f, uy);
    const f32x2_t ux2 = mul2(ux, ux), uy2 = mul2(uy, uy);
    const f32x2_t uxuy2 = mul2(mulc2(2.0f, ux), uy);
    const f32x2_t u2 = add2(ux2, uy2);
    const f32x2_t u215 = mulc2(1.5f, u2);
    const f32x2_t one_p_ux3 = add2(one, ux3), one_m_ux3 = sub2(one, ux3);
    const f32x2_t q_pos = mulc2(4.5f, add2(u2, uxuy2)), q_neg = mulc2(4.5f, sub2(u2, uxuy2));
#define BLBM_RELAX2(fi, k, poly) add2(fi, mulc2(omega, sub2(mul2(k, sub2(poly, u215)), fi)))
    f[D_NE] = BLBM_RELAX2(ne, k36, add2(add2(one_p_ux3, uy3), q_pos));
    f[D_SE] = BLBM_RELAX2(se, k36, add2(sub2(one_p_ux3, uy3), q_neg));
    f[D_NW] = BLBM_RELAX2(nw, k36, add2(add2(one_m_ux3, uy3), q_neg));
    f[D_SW] = BLBM_RELAX2(sw, k36, add2(sub2(one_m_ux3, uy3), q_pos));
    rest = add2(rest, mulc2(omega, sub2(mul2(k49, sub2(one, u215)), rest)));
    const f32x2_t ax = mulc2(4.5f, ux2), ay = mulc2(4.5f, uy2);
    f[D_E] = BLBM_RELAX2(e, k9, add2(one_p_ux3, ax));
    f[D_W] = BLBM_RELAX2(w, k9, add2(one_m_ux3, ax));
    f[D_N] = BLBM_RELAX2(n, k9, add2(add2(one, uy3), ay));
    f[D_S] = BLBM_RELAX2(s, k9, add2(sub2(one, uy3), ay));
#undef BLBM_RELAX2
}

// four consecutive cells g[q][d], q = 0..3: two packed pair collisions, results back in the scalar arrays
__device__ __forceinline__ void collide_quad_packed(float (&g)[4][8], float (&rr)[4], const float omega,
                                                    float (&mx)[4], float (&my)[4], float (&rho)[4])
{
#ifdef BLBM_CONTRACT
    // the twin has no packed flavour (ptxas would have to keep add.f32x2 and fma apart): same results, scalar
#pragma unroll
    for (int q = 0; q < 4; q++) collide_cell(g[q], rr[q], omega, mx[q], my[q], rho[q]);
#else
#pragma unroll
    for (int h = 0; h < 2; h++) {
        f32x2_t f[8], rest = pk2(rr[2 * h], rr[2 * h + 1]), pmx, pmy, prho;
#pragma unroll
        for (int d = 0; d < 8; d++) f[d] = pk2(g[2 * h][d], g[2 * h + 1][d]);
        collide_pair(f, rest, omega, pmx, pmy, prho);
#pragma unroll
        for (int d = 0; d < 8; d++) upk2(f[d], g[2 * h][d], g[2 * h + 1][d]);
        upk2(rest, rr[2 * h], rr[2 * h + 1]);
        upk2(pmx, mx[2 * h], mx[2 * h + 1]);
        upk2(pmy, my[2 * h], my[2 * h + 1]);
        upk2(prho, rho[2 * h], rho[2 * h + 1]);
    }
#endif
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

// exclusive prefix sum over the lanes of a (fully converged) warp; *total = sum over the warp
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t *total)
{
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (total) *total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

// how many of four class words have `bit` set
__device__ __forceinline__ uint32_t count4(const ushort4 c, const uint16_t bit)
{
    return ((c.x & bit) ? 1u : 0u) + ((c.y & bit) ? 1u : 0u) + ((c.z & bit) ? 1u : 0u) + ((c.w & bit) ? 1u : 0u);
}

// force the lazy module load of every kernel of the library (see aux_kernels.cu)
cudaError_t preload_aux_kernels();
cudaError_t preload_step_kernels();

// launchers (kernels.cu)
cudaError_t launch_step_scalar(const StepParams &p, int mode, bool store_moments, cudaStream_t st);
// dense_obstacles: 0 sparse fix-up (one branch per direction), 1 branch-free, 2 branch-free + cp.async staging
// index32: use 32-bit plane offsets where the slab allows it
// p.link.sig_epoch != 0 selects the variant with the in-kernel neighbour handshake (fused mode only; needs
// vec4_links_in_kernel(block_rows, packed))
// pdl: launch with the programmatic-serialization attribute (unlinked launches only; see pdl_wait in kernels.cu)
cudaError_t launch_step_vec4(const StepParams &p, int mode, bool store_moments, int block_rows, int dense_obstacles,
                             bool packed, bool index32, bool pdl, cudaStream_t st);
bool vec4_links_in_kernel(int block_rows, bool packed);

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
// a paint of at most SMALL_PAINT_PAIRS cells: the pairs travel as kernel arguments instead of through an upload
constexpr uint32_t SMALL_PAINT_PAIRS = 64;
struct SmallPaint {
    uint64_t v[2 * SMALL_PAINT_PAIRS];
};
cudaError_t launch_mask_scatter_args(uint8_t *mask, const SlabGeom &g, const SmallPaint &pairs, uint32_t npairs,
                                     cudaStream_t st);
// the same mask update and the rebuild of the class words of owned rows [row_begin, row_end) in one launch (no chain
// table: there is nothing to keep)
cudaError_t launch_paint_small(uint8_t *mask, const SlabGeom &g, const SmallPaint &pairs, uint32_t npairs, uint16_t *cls,
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
struct ChainTable {
    uint32_t *idx;         // plane offset of every slot, ascending
    uint8_t *flag;         // CHAIN_F_*
    float *state;          // CHAIN_ROWS x cap
    uint32_t *chunk_base;  // slots in front of each (row, chunk), row-major over the owned rows
    size_t n, cap;
};
// counts the barrier cells of every (row, chunk) into chunk_cnt and their total into *count, then posts the
// total to a mapped host word
cudaError_t launch_chain_count(const uint16_t *cls, const SlabGeom &g, uint32_t *chunk_cnt, unsigned long long *count,
                               unsigned long long *host_mailbox, cudaStream_t st);
// chunk_cnt -> exclusive prefix in place (scratch: one u32 per 1024 chunks), then planes -> table in plane order
cudaError_t launch_chain_build(uint16_t *cls, uint16_t *cls_other, const SlabGeom &g, const ChainPlanes &pl,
                               const ChainTable &t, uint32_t *scan_scratch, uint32_t parity, cudaStream_t st);
cudaError_t launch_chain_flush(const ChainTable &t, const ChainPlanes &pl, uint16_t *cls0, uint16_t *cls1,
                               uint32_t parity, cudaStream_t st);
// a paint is about to change the mask at `pairs` (global location, value): move those cells' chains back
// into the planes and mark their slots dead
cudaError_t launch_chain_evict(const ChainTable &t, const ChainPlanes &pl, uint16_t *cls_cur, uint16_t *cls_other,
                               const SlabGeom &g, const uint64_t *pairs, size_t npairs, uint32_t parity, cudaStream_t st);
cudaError_t launch_chain_replay(const ChainTable &t, uint32_t nsteps, uint32_t parity0, float omega, bool unsettle,
                                cudaStream_t st);
// table moments -> moment planes (only in front of a moment-storing launch of the scalar kernel)
cudaError_t launch_chain_scatter_moments(const ChainTable &t, uint32_t last_parity, float *mx, float *my, float *rho,
                                         cudaStream_t st);

cudaError_t launch_signal(unsigned long long *remote_up, unsigned long long *remote_dn,
                          unsigned long long epoch, cudaStream_t st);
cudaError_t launch_wait(const unsigned long long *from_up, const unsigned long long *from_dn,
                        unsigned long long epoch, int *err_flag, unsigned long long timeout_ns,
                        cudaStream_t st);

}  // namespace blbmk
