// tma_kernel.cu — the fused step with TMA-staged tiles (BLBM_KERNEL_TMA), sm_100a.
//
// Same arithmetic and the same finish_group() as the register/shuffle kernel; only the gather differs.
// A persistent CTA walks tiles of 128 x TY cells.  One producer thread issues, per tile, nine
// cp.async.bulk.tensor.2d loads (SASS: UTMALDG) — the eight moving populations as boxes that already
// contain both the cells each population is pulled from (one row / one float4 to the side) and the tile's
// own cells (needed by half-way bounce-back), plus the rest population — into a STAGES-deep ring of shared
// memory, each stage guarded by a full/empty mbarrier pair.  TY consumer warps (one tile row each, four
// cells per lane) wait on the full barrier, read their operands with LDS.128 + one LDS.32 per x-moving
// population, release the stage, collide and store with STG.128.  Loads are thus issued ~STAGES tiles ahead
// of use by a single thread, address arithmetic is done by the TMA unit, out-of-range rows/columns are
// zero-filled by the hardware (the reference's "out-of-range reads return 0").
//
// Box of population d for the tile at (x0, device row ry0), TY rows:
//     n : cols [x0, x0+128)   rows [ry0,   ry0+TY+1)    pull row j+1, own row j
//     s : cols [x0, x0+128)   rows [ry0-1, ry0+TY)      pull row j,   own row j+1
//     e : cols [x0-4, x0+128) rows [ry0,   ry0+TY)      pull col lx+3, own col lx+4
//     w : cols [x0, x0+132)   rows [ry0,   ry0+TY)      pull col lx+1, own col lx
//     ne/nw/se/sw: the corresponding combination (132 cols, TY+1 rows)
#include <cuda.h>

#include "step_common.cuh"

namespace blbmk {

constexpr int TMA_TX = 128;

template <int TY>
struct TmaLayout {
    // byte sizes of the per-population boxes, each rounded up to 128 B so every box starts 128-B aligned
    static constexpr uint32_t round128(uint32_t v) { return (v + 127u) / 128u * 128u; }
    static constexpr uint32_t bytes_ns = round128((TY + 1) * 128 * 4);
    static constexpr uint32_t bytes_ew = round128(TY * 132 * 4);
    static constexpr uint32_t bytes_diag = round128((TY + 1) * 132 * 4);
    static constexpr uint32_t bytes_r = round128(TY * 128 * 4);
    // offsets in Dir order nw n ne w e sw s se, then rest
    static constexpr uint32_t off_nw = 0;
    static constexpr uint32_t off_n = off_nw + bytes_diag;
    static constexpr uint32_t off_ne = off_n + bytes_ns;
    static constexpr uint32_t off_w = off_ne + bytes_diag;
    static constexpr uint32_t off_e = off_w + bytes_ew;
    static constexpr uint32_t off_sw = off_e + bytes_ew;
    static constexpr uint32_t off_s = off_sw + bytes_diag;
    static constexpr uint32_t off_se = off_s + bytes_ns;
    static constexpr uint32_t off_r = off_se + bytes_diag;
    static constexpr uint32_t stage_bytes = off_r + bytes_r;
    // bytes the TMA unit reports per stage (un-rounded box sizes)
    static constexpr uint32_t tx_bytes =
        2 * (TY + 1) * 128 * 4 + 2 * TY * 132 * 4 + 4 * (TY + 1) * 132 * 4 + TY * 128 * 4;
};

struct TmaMaps {
    CUtensorMap X[8];  // source-buffer populations, Dir order
    CUtensorMap R;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

template <bool MOM, int TY, int STAGES>
__global__ void __launch_bounds__((TY + 1) * 32)
    step_tma_kernel(const __grid_constant__ TmaMaps maps, const StepParams p, const uint32_t ntx, const uint32_t ntiles)
{
    using L = TmaLayout<TY>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);  // [STAGES]
    uint64_t *empty = full + STAGES;                       // [STAGES]
    uint8_t *stages = smem + 128;                          // barriers occupy the first 128 bytes (STAGES <= 8)
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);    // the producer's arrive.expect_tx; the data arrives as transaction bytes
            mbar_init(&empty[s], TY);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == TY) {
        // ===== producer: one thread feeds the ring =====
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);  // fresh barriers pass a wait on the previous phase
                const int x0 = (int)((t % ntx) * TMA_TX);
                const int ry0 = (int)((t / ntx) * TY) + 1;  // device row of the tile's first row
                uint8_t *st = stages + (size_t)s * L::stage_bytes;
                mbar_arrive_expect_tx(&full[s], L::tx_bytes);
                tma_load_2d(st + L::off_n, &maps.X[D_N], x0, ry0, &full[s]);
                tma_load_2d(st + L::off_ne, &maps.X[D_NE], x0 - 4, ry0, &full[s]);
                tma_load_2d(st + L::off_nw, &maps.X[D_NW], x0, ry0, &full[s]);
                tma_load_2d(st + L::off_e, &maps.X[D_E], x0 - 4, ry0, &full[s]);
                tma_load_2d(st + L::off_w, &maps.X[D_W], x0, ry0, &full[s]);
                tma_load_2d(st + L::off_s, &maps.X[D_S], x0, ry0 - 1, &full[s]);
                tma_load_2d(st + L::off_se, &maps.X[D_SE], x0 - 4, ry0 - 1, &full[s]);
                tma_load_2d(st + L::off_sw, &maps.X[D_SW], x0, ry0 - 1, &full[s]);
                tma_load_2d(st + L::off_r, &maps.R, x0, ry0, &full[s]);
            }
        }
        return;
    }

    // ===== consumers: warp j owns row j of every tile, lane l the cells 4l..4l+3 =====
    const uint32_t j = warp;
    const uint32_t lx = lane * 4u;
    // Class words come straight from global memory, and only for chunks that have any (row-chunk flags).
    // A persistent warp would expose both dependent latencies (flag, then words) once per tile, so they are
    // software-pipelined: the flag two tiles ahead, the words one tile ahead.
    auto tile_flag = [&](uint32_t t) -> uint32_t {
        if (t >= ntiles) return 0u;
        const uint32_t r = (t / ntx) * TY + j;
        return r < p.rows ? (uint32_t)p.rowflag[(size_t)r * ntx + (t % ntx)] : 0u;
    };
    auto tile_class = [&](uint32_t t, uint32_t flag) -> ushort4 {
        ushort4 c = make_ushort4(0, 0, 0, 0);
        if (flag && t < ntiles) {
            const uint32_t r = (t / ntx) * TY + j, x4 = (t % ntx) * TMA_TX + lx;
            if (r < p.rows && x4 < p.W) c = *reinterpret_cast<const ushort4 *>(p.cls + row_off(r, p.P) + x4);
        }
        return c;
    };
    ushort4 c4_next = tile_class(blockIdx.x, tile_flag(blockIdx.x));
    uint32_t flag_next = tile_flag(blockIdx.x + gridDim.x);
    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
        const uint32_t tx = t % ntx, ty = t / ntx;
        const uint32_t r = ty * TY + j;
        const uint32_t x4 = tx * TMA_TX + lx;
        const bool valid = r < p.rows && x4 < p.W;
        const size_t i = row_off(r, p.P) + x4;

        const ushort4 c4 = c4_next;
        c4_next = tile_class(t + gridDim.x, flag_next);   // consumed one tile later
        flag_next = tile_flag(t + 2u * gridDim.x);        // consumed one tile later to gate that load

        mbar_wait(&full[s], ph);
        const uint8_t *st = stages + (size_t)s * L::stage_bytes;
        const float *sN = reinterpret_cast<const float *>(st + L::off_n);
        const float *sS = reinterpret_cast<const float *>(st + L::off_s);
        const float *sE = reinterpret_cast<const float *>(st + L::off_e);
        const float *sW = reinterpret_cast<const float *>(st + L::off_w);
        const float *sNE = reinterpret_cast<const float *>(st + L::off_ne);
        const float *sNW = reinterpret_cast<const float *>(st + L::off_nw);
        const float *sSE = reinterpret_cast<const float *>(st + L::off_se);
        const float *sSW = reinterpret_cast<const float *>(st + L::off_sw);
        const float *sR = reinterpret_cast<const float *>(st + L::off_r);

        float g[4][8];
        // pulls
        const float4 vn = lds4(sN + (j + 1) * 128 + lx);
        const float4 vs = lds4(sS + j * 128 + lx);
        const float4 ve = lds4(sE + j * 132 + lx + 4);   // own e; pull = one column to the left
        const float le = sE[j * 132 + lx + 3];
        const float4 vw = lds4(sW + j * 132 + lx);       // own w; pull = one column to the right
        const float rw = sW[j * 132 + lx + 4];
        const float4 vne = lds4(sNE + (j + 1) * 132 + lx + 4);
        const float lne = sNE[(j + 1) * 132 + lx + 3];
        const float4 vnw = lds4(sNW + (j + 1) * 132 + lx);
        const float rnw = sNW[(j + 1) * 132 + lx + 4];
        const float4 vse = lds4(sSE + j * 132 + lx + 4);
        const float lse = sSE[j * 132 + lx + 3];
        const float4 vsw = lds4(sSW + j * 132 + lx);
        const float rsw = sSW[j * 132 + lx + 4];
        const float4 vr = lds4(sR + j * 128 + lx);

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
            // half-way bounce-back from the tile's own cells, which the boxes already hold
#define BLBM_BOUNCE(d, own)                                   \
    if (cany & cls_upstream_bit(d)) {                         \
        const float4 o = (own);                               \
        if (c0 & cls_upstream_bit(d)) g[0][d] = o.x;          \
        if (c1 & cls_upstream_bit(d)) g[1][d] = o.y;          \
        if (c2 & cls_upstream_bit(d)) g[2][d] = o.z;          \
        if (c3 & cls_upstream_bit(d)) g[3][d] = o.w;          \
    }
            BLBM_BOUNCE(D_N, lds4(sS + (j + 1) * 128 + lx))          // own s
            BLBM_BOUNCE(D_S, lds4(sN + j * 128 + lx))                // own n
            BLBM_BOUNCE(D_E, vw)                                     // own w
            BLBM_BOUNCE(D_W, ve)                                     // own e
            BLBM_BOUNCE(D_NE, lds4(sSW + (j + 1) * 132 + lx))        // own sw
            BLBM_BOUNCE(D_SW, lds4(sNE + j * 132 + lx + 4))          // own ne
            BLBM_BOUNCE(D_NW, lds4(sSE + (j + 1) * 132 + lx + 4))    // own se
            BLBM_BOUNCE(D_SE, lds4(sNW + j * 132 + lx))              // own nw
#undef BLBM_BOUNCE
        }
        // every operand is in registers: hand the stage back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);

        if (valid) finish_group<MOM>(p, i, x4, r, g, c0, c1, c2, c3, vr);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved at run time so that the library has no link-time dependency on libcuda
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool encode_plane(CUtensorMap *m, const float *plane, uint32_t P, uint32_t dev_rows, uint32_t box_w,
                         uint32_t box_h)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {P, dev_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)P * sizeof(float)};
    const cuuint32_t box[2] = {box_w, box_h};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(plane), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// box (w, h) of population d for TY rows
static void box_of(int d, int ty, uint32_t *w, uint32_t *h)
{
    const bool xmove = dir_dx(d) != 0, ymove = dir_dy(d) != 0;
    *w = xmove ? 132u : 128u;
    *h = (uint32_t)(ymove ? ty + 1 : ty);
}

bool tma_available() { return encode_fn() != nullptr; }

bool tma_encode_planes(void *maps16x128, float *const f[2][8], const float *R, void *mapR, uint32_t P,
                       uint32_t dev_rows, int ty)
{
    CUtensorMap *m = static_cast<CUtensorMap *>(maps16x128);
    for (int b = 0; b < 2; b++)
        for (int d = 0; d < 8; d++) {
            uint32_t w, h;
            box_of(d, ty, &w, &h);
            if (!encode_plane(&m[b * 8 + d], f[b][d], P, dev_rows, w, h)) return false;
        }
    return encode_plane(static_cast<CUtensorMap *>(mapR), R, P, dev_rows, 128u, (uint32_t)ty);
}

template <int TY, int STAGES>
static cudaError_t launch_tma_cfg(const StepParams &p, const TmaMaps &maps, bool mom, int ctas_per_sm, cudaStream_t st)
{
    using L = TmaLayout<TY>;
    const uint32_t ntx = (p.P + TMA_TX - 1) / TMA_TX;
    const uint64_t nt = (uint64_t)ntx * ((p.rows + TY - 1) / TY);
    if (nt == 0 || nt > 0xffffffffull) return cudaErrorInvalidConfiguration;
    const size_t smem = 128 + (size_t)STAGES * L::stage_bytes;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t grid = (uint64_t)sms * ctas_per_sm;
    if (grid > nt) grid = nt;
    cudaError_t e;
    if (mom) {
        e = cudaFuncSetAttribute(step_tma_kernel<true, TY, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        step_tma_kernel<true, TY, STAGES><<<(unsigned)grid, (TY + 1) * 32, smem, st>>>(maps, p, ntx, (uint32_t)nt);
    } else {
        e = cudaFuncSetAttribute(step_tma_kernel<false, TY, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        step_tma_kernel<false, TY, STAGES><<<(unsigned)grid, (TY + 1) * 32, smem, st>>>(maps, p, ntx, (uint32_t)nt);
    }
    return cudaGetLastError();
}

// maps16x128: the 16 population maps made by tma_encode_planes (buffer-major), mapR the rest map
cudaError_t launch_step_tma(const StepParams &p, int mode, bool mom, const void *maps16x128, const void *mapR,
                            int xbuf, int ty, int stages, int ctas_per_sm, cudaStream_t st)
{
    if (mode != MODE_FUSED) return launch_step_scalar(p, mode, mom, st);
    TmaMaps maps;
    memcpy(maps.X, static_cast<const CUtensorMap *>(maps16x128) + xbuf * 8, sizeof(maps.X));
    memcpy(&maps.R, mapR, sizeof(CUtensorMap));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ty == 8) {
        if (stages >= 4) return launch_tma_cfg<8, 4>(p, maps, mom, ctas_per_sm, st);
        return launch_tma_cfg<8, 2>(p, maps, mom, ctas_per_sm, st);
    }
    if (stages >= 4) return launch_tma_cfg<4, 4>(p, maps, mom, ctas_per_sm, st);
    if (stages == 3) return launch_tma_cfg<4, 3>(p, maps, mom, ctas_per_sm, st);
    return launch_tma_cfg<4, 2>(p, maps, mom, ctas_per_sm, st);
}

template <int TY, int STAGES>
static cudaError_t touch_tma()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, step_tma_kernel<false, TY, STAGES>);
    if (e != cudaSuccess) return e;
    return cudaFuncGetAttributes(&a, step_tma_kernel<true, TY, STAGES>);
}

// see preload_aux_kernels()
cudaError_t preload_tma_kernels()
{
    cudaError_t e;
    if ((e = touch_tma<8, 4>()) != cudaSuccess) return e;
    if ((e = touch_tma<8, 2>()) != cudaSuccess) return e;
    if ((e = touch_tma<4, 4>()) != cudaSuccess) return e;
    if ((e = touch_tma<4, 3>()) != cudaSuccess) return e;
    return touch_tma<4, 2>();
}

}  // namespace blbmk
