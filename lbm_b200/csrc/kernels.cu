// kernels.cu — the fused stream+collide step kernels (scalar and 128-bit vectorised) for sm_100a.
//
// What one launch computes ("pull-T" formulation).  Between steps the device
// holds T_{k-1} = the post-collision populations of step k-1 in buffer X = (k-1)%2 at EVERY cell — the
// same thing the reference's collide passes leave in place (collision/*.wgsl run over all cells with
// no mask test).  Step k of the reference is
//     stream  (stream/*.wgsl, 4 passes)   X -> Y at active cells only; skipped cells keep Y's old data
//     collide (pre_collision/*.wgsl + collision/*.wgsl, 4 passes) in place on Y, every cell
// and this kernel does both in one pass: an active cell gathers S_k from X (pull streaming with
// half-way bounce-back, e_w_stream.wgsl:51-62 etc.), a skipped cell (barrier | x==0 | y>=H-1,
// e_w_stream.wgsl:31-45) re-reads its own stale copy from Y, then every cell is collided and written
// to Y.  9 fp32 loads + 9 fp32 stores per cell = 72 B, the algorithmic minimum of a two-lattice D2Q9.
//
// Chain cells (class bit CLS_CHAIN): barrier cells are isolated — no other cell ever reads their
// populations — so their two collide-only chains can be kept in a compact side table and replayed in
// registers (aux_kernels.cu, barrier-chain kernels).  The step kernels treat the plane slots of such
// cells as don't-care: they neither re-read them from Y (the sector-granular re-read costs ~30 B/cell
// on a 15 % porous mask) nor store anything meaningful there, and a moment-storing launch keeps the
// moments the chain kernel already put there.
#include "step_common.cuh"

namespace blbmk {

// ------------------------------------------------------------------------------------------------
// scalar kernel: one cell per thread.  Used for the collide-only and stream-only passes, for tiny
// lattices and as the in-GPU cross-check of the vectorised kernels.
// ------------------------------------------------------------------------------------------------
template <int MODE, bool MOM>
__global__ void __launch_bounds__(128) step_scalar_kernel(const StepParams p)
{
    const uint32_t nbx = (p.W + 127u) / 128u;
    const uint32_t bx = blockIdx.x % nbx;
    const uint32_t r = blockIdx.x / nbx;  // owned row index
    const uint32_t x = bx * 128u + threadIdx.x;
    if (x >= p.W) return;
    const size_t i = row_off(r, p.P) + x;
    const uint16_t c = p.cls[i];
    const bool skipped = (c & CLS_SKIP) != 0;
    const bool lazy_barrier = (c & CLS_CHAIN) != 0;

    float f[8];
    if (MODE == MODE_STREAM_ONLY && skipped) return;  // destination untouched, like the WGSL early return
    if (MODE == MODE_COLLIDE_ONLY || (MODE == MODE_FUSED && skipped)) {
        if (lazy_barrier && MODE == MODE_FUSED) {
#pragma unroll
            for (int d = 0; d < 8; d++) f[d] = 0.0625f;  // don't-care slot: no load
        } else {
#pragma unroll
            for (int d = 0; d < 8; d++) f[d] = p.Y[d][i];
        }
    } else {
#pragma unroll
        for (int d = 0; d < 8; d++) {
            if (c & cls_upstream_bit(d))
                f[d] = p.X[dir_opp(d)][i];  // half-way bounce-back
            else
                f[d] = p.X[d][pull_src(i, x, d, p.W, p.P)];
        }
    }
    if (MODE == MODE_STREAM_ONLY) {
#pragma unroll
        for (int d = 0; d < 8; d++) p.Y[d][i] = f[d];
        return;
    }

    float rest = p.R[i], mx, my, rho;
    collide_cell(f, rest, p.omega, mx, my, rho);
    p.R[i] = rest;
#pragma unroll
    for (int d = 0; d < 8; d++) p.Y[d][i] = f[d];
    if (MOM) {
        if (lazy_barrier) {  // the chain kernel already stored this cell's moments
            mx = p.mx[i];
            my = p.my[i];
        } else {
            p.mx[i] = mx;
            p.my[i] = my;
            p.rho[i] = rho;
        }
    }
    // mirror the cells a neighbouring slab gathers from into its halo rows (NVLink stores)
    if (r == 0 && p.push.up_n) {
        p.push.up_n[x] = f[D_N];
        p.push.up_ne[x] = f[D_NE];
        p.push.up_nw[x] = f[D_NW];
        if (x == 0) p.push.up_w[0] = f[D_W];
        if (MOM) {
            p.push.up_mx[x] = mx;
            p.push.up_my[x] = my;
        }
    }
    if (r == 1 && x == 0 && p.push.up_nw2) p.push.up_nw2[0] = f[D_NW];
    if (r == p.rows - 1 && p.push.dn_s) {
        p.push.dn_s[x] = f[D_S];
        p.push.dn_se[x] = f[D_SE];
        p.push.dn_sw[x] = f[D_SW];
        if (MOM) {
            p.push.dn_mx[x] = mx;
            p.push.dn_my[x] = my;
        }
    }
}

cudaError_t launch_step_scalar(const StepParams &p, int mode, bool mom, cudaStream_t st)
{
    const uint32_t nbx = (p.W + 127u) / 128u;
    const uint64_t nblocks = (uint64_t)nbx * p.rows;
    if (nblocks == 0 || nblocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)nblocks), block(128);
    if (mode == MODE_FUSED) {
        if (mom) step_scalar_kernel<MODE_FUSED, true><<<grid, block, 0, st>>>(p);
        else step_scalar_kernel<MODE_FUSED, false><<<grid, block, 0, st>>>(p);
    } else if (mode == MODE_COLLIDE_ONLY) {
        if (mom) step_scalar_kernel<MODE_COLLIDE_ONLY, true><<<grid, block, 0, st>>>(p);
        else step_scalar_kernel<MODE_COLLIDE_ONLY, false><<<grid, block, 0, st>>>(p);
    } else {
        step_scalar_kernel<MODE_STREAM_ONLY, false><<<grid, block, 0, st>>>(p);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// vec4 kernel: four consecutive cells per thread, every global access an aligned 128-bit transaction.
// The six populations that move in x are realigned by one float through warp shuffles; the lane at
// the warp edge fetches the single missing float itself.  A block is 32 x 8 threads = 128 x 8 cells.
// Cells that need anything but a plain pull (class word != 0, column W-1, ragged row end) take a
// per-cell fix-up path after the vector loads.
// ------------------------------------------------------------------------------------------------
// rows per block is a tuning knob (blbm_set_tuning): 1, 2, 4 (default), 8 or 16
// DENSE selects the flavour of the bounce-back fix-up: branch-free over the directions (best where obstacles
// are dense, e.g. porous media: +2.5 %) or one branch per direction (best where they are sparse: the clean
// path then compiles to 64 registers without any spill, +5 % on an empty channel).  Same results either way.
// PACKED collides cell pairs with Blackwell's packed fp32 adds (FADD2): same bits, a quarter fewer instructions.
template <bool MOM, int V4_ROWS, bool DENSE, bool PACKED>
__global__ void __launch_bounds__(32 * V4_ROWS, DENSE ? 1024 / (32 * V4_ROWS) : 0) step_vec4_kernel(const StepParams p)
{
    const uint32_t nbx = (p.P + 127u) / 128u;
    const uint32_t bx = blockIdx.x % nbx;
    const uint32_t r = (blockIdx.x / nbx) * V4_ROWS + threadIdx.y;
    if (r >= p.rows) return;  // whole warps leave together (a warp is one row)
    const uint32_t lane = threadIdx.x;
    const uint32_t x4 = bx * 128u + lane * 4u;
    const bool valid = x4 < p.W;  // P is a multiple of 32, so x4 + 3 < P always
    const uint32_t P = p.P;
    const size_t i = row_off(r, P) + x4;
    const unsigned FULL = 0xffffffffu;

    float g[4][8];  // [cell][direction], gathered pre-collision state
    ushort4 c4 = make_ushort4(0, 0, 0, 0);
    float4 vn, vs, ve, vw, vne, vnw, vse, vsw, vr;
    vn = vs = ve = vw = vne = vnw = vse = vsw = vr = make_float4(0.f, 0.f, 0.f, 0.f);
    // one flag byte per (row, 128-cell chunk) — the same address for the whole warp — tells whether any
    // class word of the chunk is non-zero; clean chunks never touch the class plane
    const bool chunk_dirty = p.rowflag[(size_t)r * nbx + bx] != 0;
    if (valid) {
        if (chunk_dirty) c4 = *reinterpret_cast<const ushort4 *>(p.cls + i);
        vn = ldg4(p.X[D_N] + i + P);
        vne = ldg4(p.X[D_NE] + i + P);
        vnw = ldg4(p.X[D_NW] + i + P);
        ve = ldg4(p.X[D_E] + i);
        vw = ldg4(p.X[D_W] + i);
        vs = ldg4(p.X[D_S] + i - P);
        vse = ldg4(p.X[D_SE] + i - P);
        vsw = ldg4(p.X[D_SW] + i - P);
        vr = ldg4(p.R + i);
    }
    // element x4-1 of the east-moving populations, element x4+4 of the west-moving ones
    float le = __shfl_up_sync(FULL, ve.w, 1), lne = __shfl_up_sync(FULL, vne.w, 1),
          lse = __shfl_up_sync(FULL, vse.w, 1);
    float rw = __shfl_down_sync(FULL, vw.x, 1), rnw = __shfl_down_sync(FULL, vnw.x, 1),
          rsw = __shfl_down_sync(FULL, vsw.x, 1);
    if (valid && lane == 0 && x4 != 0) {
        le = p.X[D_E][i - 1];
        lne = p.X[D_NE][i + P - 1];
        lse = p.X[D_SE][i - P - 1];
    }
    if (valid && lane == 31) {  // x4+4 <= P: still inside the row pitch (or x=0 of the next row: unused)
        rw = p.X[D_W][i + 4];
        rnw = p.X[D_NW][i + P + 4];
        rsw = p.X[D_SW][i - P + 4];
    }
    if (!valid) return;

    g[0][D_N] = vn.x; g[1][D_N] = vn.y; g[2][D_N] = vn.z; g[3][D_N] = vn.w;
    g[0][D_S] = vs.x; g[1][D_S] = vs.y; g[2][D_S] = vs.z; g[3][D_S] = vs.w;
    g[0][D_E] = le; g[1][D_E] = ve.x; g[2][D_E] = ve.y; g[3][D_E] = ve.z;
    g[0][D_NE] = lne; g[1][D_NE] = vne.x; g[2][D_NE] = vne.y; g[3][D_NE] = vne.z;
    g[0][D_SE] = lse; g[1][D_SE] = vse.x; g[2][D_SE] = vse.y; g[3][D_SE] = vse.z;
    g[0][D_W] = vw.y; g[1][D_W] = vw.z; g[2][D_W] = vw.w; g[3][D_W] = rw;
    g[0][D_NW] = vnw.y; g[1][D_NW] = vnw.z; g[2][D_NW] = vnw.w; g[3][D_NW] = rnw;
    g[0][D_SW] = vsw.y; g[1][D_SW] = vsw.z; g[2][D_SW] = vsw.w; g[3][D_SW] = rsw;

    const uint32_t c0 = c4.x, c1 = c4.y, c2 = c4.z, c3 = c4.w;
    const uint32_t cany = c0 | c1 | c2 | c3;
    if (cany & CLS_UP_MASK) {
        // half-way bounce-back: population d of a cell whose upstream neighbour is a barrier is the cell's
        // own opposite population (skipped cells carry no upstream bits, so one bit test decides)
        if (DENSE) {
            // Branch-free over the directions (where obstacles are dense every direction is needed by some lane
            // of the warp anyway; where they are sparse few threads come here at all), in two batches of four to
            // keep the live registers — and with them the occupancy of the clean path — unchanged.
    #pragma unroll
            for (int half = 0; half < 2; half++) {
                float4 own[4];
    #pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int d = half * 4 + q;
                    if (dir_opp(d) == D_E) own[q] = ve;
                    else if (dir_opp(d) == D_W) own[q] = vw;
                    else own[q] = ldg4(p.X[dir_opp(d)] + i);
                }
    #pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int d = half * 4 + q;
                    const uint32_t bit = cls_upstream_bit(d);
                    g[0][d] = (c0 & bit) ? own[q].x : g[0][d];
                    g[1][d] = (c1 & bit) ? own[q].y : g[1][d];
                    g[2][d] = (c2 & bit) ? own[q].z : g[2][d];
                    g[3][d] = (c3 & bit) ? own[q].w : g[3][d];
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < 8; d++) {
                const uint32_t bit = cls_upstream_bit(d);
                if (cany & bit) {
                    float4 own;
                    if (dir_opp(d) == D_E) own = ve;
                    else if (dir_opp(d) == D_W) own = vw;
                    else own = ldg4(p.X[dir_opp(d)] + i);
                    if (c0 & bit) g[0][d] = own.x;
                    if (c1 & bit) g[1][d] = own.y;
                    if (c2 & bit) g[2][d] = own.z;
                    if (c3 & bit) g[3][d] = own.w;
                }
            }
        }
    }
    finish_group<MOM, PACKED>(p, i, x4, r, g, c0, c1, c2, c3, vr);
}

template <int V4_ROWS, bool DENSE, bool PACKED>
static cudaError_t launch_vec4_rows(const StepParams &p, bool mom, cudaStream_t st)
{
    const uint32_t nbx = (p.P + 127u) / 128u;
    const uint64_t nblocks = (uint64_t)nbx * ((p.rows + V4_ROWS - 1) / V4_ROWS);
    if (nblocks == 0 || nblocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)nblocks), block(32, V4_ROWS);
    if (mom) step_vec4_kernel<true, V4_ROWS, DENSE, PACKED><<<grid, block, 0, st>>>(p);
    else step_vec4_kernel<false, V4_ROWS, DENSE, PACKED><<<grid, block, 0, st>>>(p);
    return cudaGetLastError();
}

template <bool DENSE, bool PACKED>
static cudaError_t launch_vec4_flavour(const StepParams &p, bool mom, int block_rows, cudaStream_t st)
{
    switch (block_rows) {
    case 1: return launch_vec4_rows<1, DENSE, PACKED>(p, mom, st);
    case 2: return launch_vec4_rows<2, DENSE, PACKED>(p, mom, st);
    case 8: return launch_vec4_rows<8, DENSE, PACKED>(p, mom, st);
    case 16: return launch_vec4_rows<16, DENSE, PACKED>(p, mom, st);
    default: return launch_vec4_rows<4, DENSE, PACKED>(p, mom, st);
    }
}

cudaError_t launch_step_vec4(const StepParams &p, int mode, bool mom, int block_rows, bool dense_obstacles,
                             bool packed, cudaStream_t st)
{
    if (mode != MODE_FUSED) return launch_step_scalar(p, mode, mom, st);
    if (dense_obstacles)
        return packed ? launch_vec4_flavour<true, true>(p, mom, block_rows, st)
                      : launch_vec4_flavour<true, false>(p, mom, block_rows, st);
    return packed ? launch_vec4_flavour<false, true>(p, mom, block_rows, st)
                  : launch_vec4_flavour<false, false>(p, mom, block_rows, st);
}

// see preload_aux_kernels(): force the (lazy) load of every step-kernel instantiation
#define BLBM_TOUCH(...)                                                  \
    do {                                                                 \
        cudaFuncAttributes a__;                                          \
        cudaError_t e__ = cudaFuncGetAttributes(&a__, __VA_ARGS__);      \
        if (e__ != cudaSuccess) return e__;                              \
    } while (0)

template <int ROWS>
static cudaError_t touch_vec4_rows()
{
    BLBM_TOUCH(step_vec4_kernel<false, ROWS, false, false>);
    BLBM_TOUCH(step_vec4_kernel<true, ROWS, false, false>);
    BLBM_TOUCH(step_vec4_kernel<false, ROWS, true, false>);
    BLBM_TOUCH(step_vec4_kernel<true, ROWS, true, false>);
    BLBM_TOUCH(step_vec4_kernel<false, ROWS, false, true>);
    BLBM_TOUCH(step_vec4_kernel<true, ROWS, false, true>);
    BLBM_TOUCH(step_vec4_kernel<false, ROWS, true, true>);
    BLBM_TOUCH(step_vec4_kernel<true, ROWS, true, true>);
    return cudaSuccess;
}

cudaError_t preload_step_kernels()
{
    BLBM_TOUCH(step_scalar_kernel<MODE_FUSED, false>);
    BLBM_TOUCH(step_scalar_kernel<MODE_FUSED, true>);
    BLBM_TOUCH(step_scalar_kernel<MODE_COLLIDE_ONLY, false>);
    BLBM_TOUCH(step_scalar_kernel<MODE_COLLIDE_ONLY, true>);
    BLBM_TOUCH(step_scalar_kernel<MODE_STREAM_ONLY, false>);
    cudaError_t e;
    if ((e = touch_vec4_rows<1>()) != cudaSuccess) return e;
    if ((e = touch_vec4_rows<2>()) != cudaSuccess) return e;
    if ((e = touch_vec4_rows<4>()) != cudaSuccess) return e;
    if ((e = touch_vec4_rows<8>()) != cudaSuccess) return e;
    return touch_vec4_rows<16>();
}
#undef BLBM_TOUCH

}  // namespace blbmk
