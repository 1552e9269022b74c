// aux_kernels.cu — everything around the step kernel: initial conditions, barrier mask and cell
// classification, the summary-stat kernels that feed the renderer, reductions, and the
// neighbour-slab flag handshake.  All rare or cheap relative to the step kernel.
#include "blbm_internal.cuh"

namespace blbmk {

// ---- uniform fill of row ranges (set_equil broadcast, lbm.rs:611-643 / :1078-1081) ---------------
struct FillArgs {
    float *plane[20];
    float value[20];
    int n;
    uint32_t W, P, row_begin, row_end;
};

__global__ void fill_rows_kernel(const FillArgs a)
{
    const uint32_t rows = a.row_end - a.row_begin;
    const size_t per_plane = (size_t)rows * a.W;
    const size_t total = per_plane * a.n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(t / per_plane);
        const size_t rem = t - (size_t)k * per_plane;
        const uint32_t r = (uint32_t)(rem / a.W);
        const uint32_t x = (uint32_t)(rem - (size_t)r * a.W);
        a.plane[k][(size_t)(a.row_begin + r) * a.P + x] = a.value[k];
    }
}

cudaError_t launch_fill_rows(float *const *planes, const float *values, int nplanes, uint32_t W, uint32_t P,
                             uint32_t dev_row_begin, uint32_t dev_row_end, cudaStream_t st)
{
    if (dev_row_end <= dev_row_begin || nplanes <= 0) return cudaSuccess;
    if (nplanes > 20) return cudaErrorInvalidValue;
    FillArgs a;
    a.n = nplanes;
    for (int k = 0; k < nplanes; k++) {
        a.plane[k] = planes[k];
        a.value[k] = values[k];
    }
    a.W = W;
    a.P = P;
    a.row_begin = dev_row_begin;
    a.row_end = dev_row_end;
    fill_rows_kernel<<<148 * 8, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---- barrier mask --------------------------------------------------------------------------------
// init_barrier, lbm.rs:595-605: rows 0 and H-1 are wall, everything else fluid.  The mask plane keeps
// two halo rows on each side; rows outside the lattice stay 0 (out-of-range reads return 0).
__global__ void mask_init_kernel(uint8_t *mask, const SlabGeom g)
{
    const size_t total = (size_t)(g.rows + 4) * g.P;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t dr = (uint32_t)(t / g.P);
        const uint32_t x = (uint32_t)(t - (size_t)dr * g.P);
        const int64_t gy = (int64_t)g.row0 + (int64_t)dr - 2;
        uint8_t v = 0;
        if (x < g.W && gy >= 0 && gy < (int64_t)g.Hg && (gy == 0 || gy == (int64_t)g.Hg - 1)) v = 1;
        mask[t] = v;
    }
}

cudaError_t launch_mask_init(uint8_t *mask, const SlabGeom &g, cudaStream_t st)
{
    mask_init_kernel<<<148 * 4, 256, 0, st>>>(mask, g);
    return cudaGetLastError();
}

// barrier_draw.wgsl:11-18: barrier[location] = value.  pairs are (global location, value), already
// de-duplicated on the host (last writer wins); locations outside this slab's mask window are ignored.
__global__ void mask_scatter_kernel(uint8_t *mask, const SlabGeom g, const uint64_t *pairs, size_t npairs)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < npairs;
         t += (size_t)gridDim.x * blockDim.x) {
        const uint64_t loc = pairs[2 * t], val = pairs[2 * t + 1];
        const uint64_t gy = loc / g.W;
        const uint32_t x = (uint32_t)(loc - gy * g.W);
        if (gy >= g.Hg) continue;
        const int64_t lr = (int64_t)gy - (int64_t)g.row0;  // local row, halo rows are -2,-1 and rows,rows+1
        if (lr < -2 || lr >= (int64_t)g.rows + 2) continue;
        mask[mask_row_off(lr, g.P) + x] = (val == 1u) ? 1 : 0;
    }
}

cudaError_t launch_mask_scatter(uint8_t *mask, const SlabGeom &g, const uint64_t *pairs, size_t npairs,
                                cudaStream_t st)
{
    if (npairs == 0) return cudaSuccess;
    size_t nb = (npairs + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    mask_scatter_kernel<<<(unsigned)nb, 256, 0, st>>>(mask, g, pairs, npairs);
    return cudaGetLastError();
}

// Cell classification from the mask: which cells the stream passes skip (e_w_stream.wgsl:31-45) and, per
// moving population, whether the cell it is pulled from is a barrier (e_w_stream.wgsl:51-62), with the
// reference's flat-index neighbours (column W-1's east side is column 0 of the next row) and
// out-of-range reads returning 0.
__device__ __forceinline__ uint32_t mask_at(const uint8_t *mask, const SlabGeom &g, int64_t gx, int64_t gy)
{
    if (gx == (int64_t)g.W) {
        gx = 0;
        gy += 1;
    } else if (gx < 0) {
        gx = (int64_t)g.W - 1;
        gy -= 1;
    }
    if (gy < 0 || gy >= (int64_t)g.Hg) return 0;
    const int64_t lr = gy - (int64_t)g.row0;
    if (lr < -2 || lr >= (int64_t)g.rows + 2) return 0;  // cannot happen for owned cells
    return mask[mask_row_off(lr, g.P) + (size_t)gx];
}

// One warp per 128-cell chunk of a row (4 cells per lane); PUBLIC selects the word of blbm_read_cell_class
// (densely packed rows x W) instead of the kernel-facing word (plane layout) and skips the chunk flags.
template <bool PUBLIC>
__device__ __forceinline__ void build_class_chunk(uint16_t *cls, const uint8_t *mask, const SlabGeom &g,
                                                  const uint16_t *keep_chain, uint8_t *rowflag, const uint32_t r,
                                                  const uint32_t chunk, const uint32_t nchunk, const uint32_t lane)
{
    const int64_t gy = (int64_t)g.row0 + r;
    uint32_t any = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t x = chunk * CHUNK + lane * 4u + q;
        if (x >= g.P) continue;
        uint16_t c = 0;
        if (x < g.W) {
            const bool bar = mask[mask_row_off(r, g.P) + x] == 1;
            const bool skip = bar || x == 0 || gy >= (int64_t)g.Hg - 1;
            uint32_t up = 0;
#pragma unroll
            for (int d = 0; d < 8; d++)
                if (mask_at(mask, g, (int64_t)x - dir_dx(d), gy - dir_dy(d)) == 1) up |= 1u << d;
            if (PUBLIC) {
                c = (uint16_t)((bar ? PUB_BARRIER : 0) | (skip ? PUB_SKIP : 0) | (up << 2));
            } else {
                c = (uint16_t)((bar ? CLS_BARRIER : 0) | (skip ? CLS_SKIP : (uint16_t)up));
                // cells whose state lives in the chain table stay there (a paint evicts the cells it
                // touches before the mask changes, so a kept bit belongs to a cell that is still a barrier)
                if (keep_chain) c |= keep_chain[row_off(r, g.P) + x] & (CLS_CHAIN | CLS_SLOT);
            }
        }
        if (PUBLIC) {
            if (x < g.W) cls[(size_t)r * g.W + x] = c;
        } else {
            cls[row_off(r, g.P) + x] = c;
            any |= c;
        }
    }
    if (!PUBLIC && rowflag) {
        any = __reduce_or_sync(0xffffffffu, any);
        if (lane == 0) rowflag[(size_t)r * nchunk + chunk] = any ? 1 : 0;
    }
}

template <bool PUBLIC>
__global__ void build_class_kernel(uint16_t *cls, const uint8_t *mask, const SlabGeom g, const uint16_t *keep_chain,
                                   uint8_t *rowflag, const uint32_t row_begin, const uint32_t row_end)
{
    const uint32_t nchunk = (g.P + CHUNK - 1) / CHUNK;
    const size_t nwarps_total = (size_t)(row_end - row_begin) * nchunk;
    const uint32_t lane = threadIdx.x & 31u;
    for (size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < nwarps_total;
         wid += ((size_t)gridDim.x * blockDim.x) >> 5)
        build_class_chunk<PUBLIC>(cls, mask, g, keep_chain, rowflag, row_begin + (uint32_t)(wid / nchunk),
                                  (uint32_t)(wid % nchunk), nchunk, lane);
}

// A small paint: the (location, value) pairs travel as kernel arguments, which saves the upload of the general path
// (small lattices are launch-bound: a 15-step frame of the 512 x 256 cylinder is 60 us, the upload 3 us of it).
// (Fusing the class rebuild into the same SINGLE-BLOCK launch was measured and rejected: one SM rebuilding 13 rows
// takes longer than a second launch spread over the chip, 14.4 vs 11.4 us per paint; paint_small_kernel below fuses
// them across many blocks instead.)
__global__ void mask_scatter_args_kernel(uint8_t *mask, const SlabGeom g, const SmallPaint pairs, const uint32_t npairs)
{
    if (threadIdx.x >= npairs) return;
    const uint64_t loc = pairs.v[2 * threadIdx.x], val = pairs.v[2 * threadIdx.x + 1];
    const uint64_t gy = loc / g.W;
    const int64_t lr = (int64_t)gy - (int64_t)g.row0;
    if (gy < g.Hg && lr >= -2 && lr < (int64_t)g.rows + 2)
        mask[mask_row_off(lr, g.P) + (uint32_t)(loc - gy * g.W)] = (val == 1u) ? 1 : 0;
}

cudaError_t launch_mask_scatter_args(uint8_t *mask, const SlabGeom &g, const SmallPaint &pairs, uint32_t npairs,
                                     cudaStream_t st)
{
    if (npairs == 0) return cudaSuccess;
    if (npairs > SMALL_PAINT_PAIRS) return cudaErrorInvalidValue;
    mask_scatter_args_kernel<<<1, SMALL_PAINT_PAIRS, 0, st>>>(mask, g, pairs, npairs);
    return cudaGetLastError();
}

static unsigned class_grid(const SlabGeom &g, uint32_t nrows);

// Small paint and class rebuild in one launch, without a grid-wide barrier between the two: EVERY block applies the
// whole (at most 64-pair, already de-duplicated) stroke to the mask itself - all blocks store the same bytes to the
// same places - and, after a block barrier, rebuilds its share of the class words from a mask that holds the stroke
// wherever it looks: its own stores are visible to it, and anybody else's store to those bytes has the same value.
__global__ void __launch_bounds__(256) paint_small_kernel(uint8_t *mask, const SlabGeom g, const SmallPaint pairs,
                                                          const uint32_t npairs, uint16_t *cls, uint8_t *rowflag,
                                                          const uint32_t row_begin, const uint32_t row_end)
{
    if (threadIdx.x < npairs) {
        const uint64_t loc = pairs.v[2 * threadIdx.x], val = pairs.v[2 * threadIdx.x + 1];
        const uint64_t gy = loc / g.W;
        const int64_t lr = (int64_t)gy - (int64_t)g.row0;
        if (gy < g.Hg && lr >= -2 && lr < (int64_t)g.rows + 2)
            mask[mask_row_off(lr, g.P) + (uint32_t)(loc - gy * g.W)] = (val == 1u) ? 1 : 0;
    }
    __syncthreads();
    const uint32_t nchunk = (g.P + CHUNK - 1) / CHUNK;
    const size_t nwarps_total = (size_t)(row_end - row_begin) * nchunk;
    const uint32_t lane = threadIdx.x & 31u;
    for (size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < nwarps_total;
         wid += ((size_t)gridDim.x * blockDim.x) >> 5)
        build_class_chunk<false>(cls, mask, g, nullptr, rowflag, row_begin + (uint32_t)(wid / nchunk),
                                 (uint32_t)(wid % nchunk), nchunk, lane);
}

cudaError_t launch_paint_small(uint8_t *mask, const SlabGeom &g, const SmallPaint &pairs, uint32_t npairs, uint16_t *cls,
                               uint8_t *rowflag, uint32_t row_begin, uint32_t row_end, cudaStream_t st)
{
    if (npairs == 0 || npairs > SMALL_PAINT_PAIRS || row_end <= row_begin) return cudaErrorInvalidValue;
    paint_small_kernel<<<class_grid(g, row_end - row_begin), 256, 0, st>>>(mask, g, pairs, npairs, cls, rowflag, row_begin,
                                                                           row_end);
    return cudaGetLastError();
}

static unsigned class_grid(const SlabGeom &g, uint32_t nrows)
{
    const size_t warps = (size_t)nrows * ((g.P + CHUNK - 1) / CHUNK);
    size_t nb = (warps + 7) / 8;  // 8 warps per block
    if (nb > 148 * 8) nb = 148 * 8;
    return (unsigned)(nb ? nb : 1);
}

cudaError_t launch_build_class(uint16_t *cls, const uint8_t *mask, const SlabGeom &g, const uint16_t *keep_chain,
                               uint8_t *rowflag, uint32_t row_begin, uint32_t row_end, cudaStream_t st)
{
    if (row_end <= row_begin) return cudaSuccess;
    build_class_kernel<false><<<class_grid(g, row_end - row_begin), 256, 0, st>>>(cls, mask, g, keep_chain, rowflag,
                                                                                  row_begin, row_end);
    return cudaGetLastError();
}

cudaError_t launch_build_public_class(uint16_t *dst, const uint8_t *mask, const SlabGeom &g, cudaStream_t st)
{
    build_class_kernel<true><<<class_grid(g, g.rows), 256, 0, st>>>(dst, mask, g, nullptr, nullptr, 0, g.rows);
    return cudaGetLastError();
}

// ---- moments without the rest term (reset_to_equilibrium / custom_speed, lbm.rs:1082-1087) ----------
struct MomArgs {
    const float *f[8];
    float *mx, *my, *rho;
    uint32_t W, P, row_begin, row_end;
};

__global__ void precollision_moments_kernel(const MomArgs a)
{
    const size_t total = (size_t)(a.row_end - a.row_begin) * a.W;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / a.W);
        const uint32_t x = (uint32_t)(t - (size_t)r * a.W);
        const size_t i = (size_t)(a.row_begin + r) * a.P + x;
        float f[8], mx, my, rho;
#pragma unroll
        for (int d = 0; d < 8; d++) f[d] = a.f[d][i];
        precollision_moments(f, mx, my, rho);
        a.mx[i] = mx;
        a.my[i] = my;
        a.rho[i] = rho;
    }
}

cudaError_t launch_precollision_moments(const float *const *f8, float *mx, float *my, float *rho, uint32_t W,
                                        uint32_t P, uint32_t dev_row_begin, uint32_t dev_row_end,
                                        cudaStream_t st)
{
    if (dev_row_end <= dev_row_begin) return cudaSuccess;
    MomArgs a;
    for (int d = 0; d < 8; d++) a.f[d] = f8[d];
    a.mx = mx;
    a.my = my;
    a.rho = rho;
    a.W = W;
    a.P = P;
    a.row_begin = dev_row_begin;
    a.row_end = dev_row_end;
    precollision_moments_kernel<<<148 * 8, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---- summary statistics, summary_stats/{curl,ux,uy,rho,speed}.wgsl ------------------------------------
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

template <int STAT>
__global__ void summary_kernel(const float *__restrict__ mx, const float *__restrict__ my,
                               const float *__restrict__ rho, float *__restrict__ out, const SlabGeom g)
{
    const size_t total = (size_t)g.rows * g.W;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / g.W);
        const uint32_t x = (uint32_t)(t - (size_t)r * g.W);
        const size_t i = row_off(r, g.P) + x;
        if (STAT == 0) {
            // curl.wgsl:37-49 — skips column 0 and rows >= H-1 (their output keeps its old value);
            // index+1 at column W-1 is column 0 of the next row (flat index)
            if (x == 0) continue;
            if (g.row0 + r >= g.Hg - 1) continue;
            const float a = (x + 1 < g.W) ? my[i + 1] : my[i - x + g.P];
            const float b = my[i - 1];
            const float c = mx[i - g.P];
            const float d = mx[i + g.P];
            out[i] = __fdiv_rn(__fmul_rn(10.0f, __fadd_rn(__fsub_rn(__fsub_rn(a, b), c), d)), rho[i]);
        } else if (STAT == 1) {
            out[i] = mx[i];
        } else if (STAT == 2) {
            out[i] = my[i];
        } else if (STAT == 3) {
            out[i] = mul_sub(4.0f, clamp01(__fmul_rn(0.15f, rho[i])), 0.5f);
        } else {
            const float s2 = mul_add(mx[i], mx[i], __fmul_rn(my[i], my[i]));
            out[i] = __fsub_rn(clamp01(__fmul_rn(5.0f, __fsqrt_rn(s2))), 0.5f);
        }
    }
}

// curl.wgsl, four consecutive cells per thread with 128-bit accesses (a warp = one 128-cell chunk of a row, a block
// = 8 rows); groups that touch column 0, column W-1 or the row end take the per-cell path of summary_kernel<0>.
constexpr uint32_t CURL_ROWS = 8, CURL_GRID_Y = 32768;
__global__ void __launch_bounds__(32 * CURL_ROWS) curl_vec4_kernel(const float *__restrict__ mx,
                                                                   const float *__restrict__ my,
                                                                   const float *__restrict__ rho,
                                                                   float *__restrict__ out, const SlabGeom g)
{
    const uint32_t r = (blockIdx.z * CURL_GRID_Y + blockIdx.y) * CURL_ROWS + threadIdx.y;
    const uint32_t x4 = (blockIdx.x * 32u + threadIdx.x) * 4u;
    if (r >= g.rows || x4 >= g.W) return;
    if (g.row0 + r >= g.Hg - 1) return;  // rows >= H-1 keep their old output
    const size_t i = row_off(r, g.P) + x4;
    if (x4 != 0 && x4 + 4 < g.W) {
        const float4 m = *reinterpret_cast<const float4 *>(my + i);
        const float4 up = *reinterpret_cast<const float4 *>(mx + i - g.P);
        const float4 dn = *reinterpret_cast<const float4 *>(mx + i + g.P);
        const float4 d = *reinterpret_cast<const float4 *>(rho + i);
        const float l = my[i - 1], rgt = my[i + 4];
        float4 o;
        o.x = __fdiv_rn(__fmul_rn(10.0f, __fadd_rn(__fsub_rn(__fsub_rn(m.y, l), up.x), dn.x)), d.x);
        o.y = __fdiv_rn(__fmul_rn(10.0f, __fadd_rn(__fsub_rn(__fsub_rn(m.z, m.x), up.y), dn.y)), d.y);
        o.z = __fdiv_rn(__fmul_rn(10.0f, __fadd_rn(__fsub_rn(__fsub_rn(m.w, m.y), up.z), dn.z)), d.z);
        o.w = __fdiv_rn(__fmul_rn(10.0f, __fadd_rn(__fsub_rn(__fsub_rn(rgt, m.z), up.w), dn.w)), d.w);
        *reinterpret_cast<float4 *>(out + i) = o;
        return;
    }
    for (uint32_t q = 0; q < 4; q++) {
        const uint32_t x = x4 + q;
        if (x == 0 || x >= g.W) continue;  // column 0 keeps its old output
        const size_t j = i + q;
        const float a = (x + 1 < g.W) ? my[j + 1] : my[j - x + g.P];  // flat index: (0, y+1) at column W-1
        const float b = my[j - 1];
        out[j] = __fdiv_rn(__fmul_rn(10.0f, __fadd_rn(__fsub_rn(__fsub_rn(a, b), mx[j - g.P]), mx[j + g.P])), rho[j]);
    }
}

cudaError_t launch_summary(int stat, const float *mx, const float *my, const float *rho, float *out,
                           const SlabGeom &g, cudaStream_t st)
{
    if (stat == 0 && g.rows > 0 && g.W > 0) {
        const uint32_t nbx = (g.W + 127u) / 128u, nrb = (g.rows + CURL_ROWS - 1) / CURL_ROWS;
        dim3 grid(nbx, nrb < CURL_GRID_Y ? nrb : CURL_GRID_Y, (nrb + CURL_GRID_Y - 1) / CURL_GRID_Y), block(32, CURL_ROWS);
        if (grid.z <= 65535u) {
            curl_vec4_kernel<<<grid, block, 0, st>>>(mx, my, rho, out, g);
            return cudaGetLastError();
        }
    }
    const size_t total = (size_t)g.rows * g.W;
    size_t nb = (total + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    if (nb == 0) return cudaSuccess;
    const unsigned b = (unsigned)nb;
    switch (stat) {
    case 0: summary_kernel<0><<<b, 256, 0, st>>>(mx, my, rho, out, g); break;
    case 1: summary_kernel<1><<<b, 256, 0, st>>>(mx, my, rho, out, g); break;
    case 2: summary_kernel<2><<<b, 256, 0, st>>>(mx, my, rho, out, g); break;
    case 3: summary_kernel<3><<<b, 256, 0, st>>>(mx, my, rho, out, g); break;
    case 4: summary_kernel<4><<<b, 256, 0, st>>>(mx, my, rho, out, g); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ---- global reductions: warp shuffles, one atomic per block ---------------------------------------
__global__ void reduce_kernel(const float *__restrict__ mx, const float *__restrict__ my,
                              const float *__restrict__ rho, const float *__restrict__ out, const SlabGeom g,
                              double *sums3, float *maxabs)
{
    double s_rho = 0.0, s_mx = 0.0, s_my = 0.0;
    float m = 0.0f;
    const size_t total = (size_t)g.rows * g.W;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / g.W);
        const uint32_t x = (uint32_t)(t - (size_t)r * g.W);
        const size_t i = row_off(r, g.P) + x;
        s_rho += (double)rho[i];
        s_mx += (double)mx[i];
        s_my += (double)my[i];
        m = fmaxf(m, fabsf(out[i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_rho += __shfl_xor_sync(0xffffffffu, s_rho, o);
        s_mx += __shfl_xor_sync(0xffffffffu, s_mx, o);
        s_my += __shfl_xor_sync(0xffffffffu, s_my, o);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    __shared__ double sh[3][8];
    __shared__ float shm[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[0][warp] = s_rho;
        sh[1][warp] = s_mx;
        sh[2][warp] = s_my;
        shm[warp] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        float mm = 0.f;
        for (int w = 0; w < 8; w++) {
            a += sh[0][w];
            b += sh[1][w];
            c += sh[2][w];
            mm = fmaxf(mm, shm[w]);
        }
        atomicAdd(&sums3[0], a);
        atomicAdd(&sums3[1], b);
        atomicAdd(&sums3[2], c);
        atomicMax(reinterpret_cast<int *>(maxabs), __float_as_int(mm));  // non-negative floats order as ints
    }
}

cudaError_t launch_reduce(const float *mx, const float *my, const float *rho, const float *out,
                          const SlabGeom &g, double *sums3, float *maxabs, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(sums3, 0, 3 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(maxabs, 0, sizeof(float), st);
    if (e != cudaSuccess) return e;
    reduce_kernel<<<148 * 4, 256, 0, st>>>(mx, my, rho, out, g, sums3, maxabs);
    return cudaGetLastError();
}

// ---- neighbour handshake: one 64-bit epoch word per face, in the receiver's memory ------------------
__global__ void signal_kernel(unsigned long long *remote_up, unsigned long long *remote_dn,
                              unsigned long long epoch)
{
    // everything the preceding kernels on this stream stored into the neighbours' halo rows is ordered
    // before the flag by the kernel boundary plus this system-scope fence
    __threadfence_system();
    if (remote_up) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote_up), "l"(epoch) : "memory");
    if (remote_dn) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote_dn), "l"(epoch) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void wait_kernel(const unsigned long long *from_up, const unsigned long long *from_dn,
                            unsigned long long epoch, int *err_flag, unsigned long long timeout_ns)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (*reinterpret_cast<volatile int *>(err_flag)) return;  // a previous wait already gave up
    for (;;) {
        const bool up_ok = !from_up || ld_acquire_sys(from_up) >= epoch;
        const bool dn_ok = !from_dn || ld_acquire_sys(from_dn) >= epoch;
        if (up_ok && dn_ok) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {  // never hang the GPU on a dead neighbour: flag and carry on
            *err_flag = 1;
            break;
        }
        __nanosleep(200);
    }
}

cudaError_t launch_signal(unsigned long long *remote_up, unsigned long long *remote_dn,
                          unsigned long long epoch, cudaStream_t st)
{
    signal_kernel<<<1, 1, 0, st>>>(remote_up, remote_dn, epoch);
    return cudaGetLastError();
}

cudaError_t launch_wait(const unsigned long long *from_up, const unsigned long long *from_dn,
                        unsigned long long epoch, int *err_flag, unsigned long long timeout_ns,
                        cudaStream_t st)
{
    wait_kernel<<<1, 1, 0, st>>>(from_up, from_dn, epoch, err_flag, timeout_ns);
    return cudaGetLastError();
}

}  // namespace blbmk

// ================================================================================================
// Barrier chains.  A barrier cell is never read by any other cell (its neighbours bounce back instead
// of pulling from it) and is never written by a stream pass, so all the reference does to it is
// collide the stale copy in buffer step%2 — in place, every step, sharing one rest population
// (collision/*.wgsl have no mask test).  That is a closed 17-float recurrence per cell.  The chain
// table holds those 17 floats compactly; chain_replay_kernel advances them n steps in registers with
// the very same collide_cell() the step kernels use (bit-identical by construction).
//
// The table is ORDERED by plane offset: the slot of a cell is chunk_base[row][chunk] plus the number of slot
// bits (CLS_SLOT) in front of it inside its 128-cell chunk.  That makes every access of the replay coalesced —
// including the moments of the latest collide, which stay in the table (rows CHAIN_ROW_MOM..) instead of being
// scattered into the moment planes — lets the one moment-storing step launch of a call pick them up by rank,
// and lets a paint find the slots of the cells it touches without scanning the table.
//
// Settled entries: once a two-step block (one collide of each copy) reproduces all 17 floats bit for bit, the
// chain is on an exact period-2 cycle: both copies are fixed and the rest population alternates between two
// values.  The table then holds both rest values and the moments of both collides, and the entry costs one flag
// byte per call until omega changes.
// ================================================================================================
namespace blbmk {

// class words of the four cells lane `lane` owns in chunk `chunk` of owned row r (zero beyond the pitch)
__device__ __forceinline__ ushort4 chunk_words(const uint16_t *cls, const SlabGeom &g, uint32_t r, uint32_t chunk,
                                               uint32_t lane, size_t *i0)
{
    const uint32_t x = chunk * CHUNK + lane * 4u;
    *i0 = row_off(r, g.P) + x;
    if (x >= g.P) return make_ushort4(0, 0, 0, 0);
    return *reinterpret_cast<const ushort4 *>(cls + *i0);
}

// one warp per (row, chunk): barrier cells of the chunk -> chunk_cnt, grand total -> *count
__global__ void chain_count_kernel(const uint16_t *cls, const SlabGeom g, uint32_t *chunk_cnt, unsigned long long *count)
{
    const uint32_t nchunk = (g.P + CHUNK - 1) / CHUNK;
    const size_t nwarps_total = (size_t)g.rows * nchunk;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long local = 0;
    for (size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < nwarps_total;
         wid += ((size_t)gridDim.x * blockDim.x) >> 5) {
        size_t i0;
        const ushort4 c = chunk_words(cls, g, (uint32_t)(wid / nchunk), (uint32_t)(wid % nchunk), lane, &i0);
        const uint32_t sum = __reduce_add_sync(0xffffffffu, count4(c, CLS_BARRIER));
        if (lane == 0) {
            chunk_cnt[wid] = sum;
            local += sum;
        }
    }
    if (lane == 0 && local) atomicAdd(count, local);
}

__global__ void mailbox_kernel(unsigned long long *host_mapped, const unsigned long long *src)
{
    *host_mapped = *src;
    __threadfence_system();
}

// host_mailbox: mapped pinned host word the result is posted to by a store from the GPU — a D2H memcpy
// would queue behind a large asynchronous read-back on the same copy engine
cudaError_t launch_chain_count(const uint16_t *cls, const SlabGeom &g, uint32_t *chunk_cnt, unsigned long long *count,
                               unsigned long long *host_mailbox, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    chain_count_kernel<<<148 * 8, 256, 0, st>>>(cls, g, chunk_cnt, count);
    mailbox_kernel<<<1, 1, 0, st>>>(host_mailbox, count);
    return cudaGetLastError();
}

// ---- exclusive prefix sum of the per-chunk counts (three small kernels; tiles of 1024) ------------------
constexpr uint32_t SCAN_TILE = 1024;

// exclusive scan of up to 1024 values held four per thread by a 256-thread block; returns the block total
__device__ __forceinline__ uint32_t block_excl_scan4(uint32_t (&v)[4], uint32_t *sh /* 8 words */)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t wtotal;
    const uint32_t in_warp = warp_excl_scan(mine, lane, &wtotal);
    if (lane == 0) sh[warp] = wtotal;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; w++) {
        if (w < warp) before += sh[w];
        total += sh[w];
    }
    __syncthreads();
    uint32_t run = before + in_warp;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t t = v[q];
        v[q] = run;
        run += t;
    }
    return total;
}

__global__ void __launch_bounds__(256) scan_tiles_kernel(uint32_t *a, size_t m, uint32_t *tile_sums)
{
    __shared__ uint32_t sh[8];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4u;
    uint32_t v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = base + q < m ? a[base + q] : 0u;
    const uint32_t total = block_excl_scan4(v, sh);
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (base + q < m) a[base + q] = v[q];
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block walks all tile sums (a few thousand at most) with a running carry
__global__ void __launch_bounds__(256) scan_sums_kernel(uint32_t *sums, size_t t)
{
    __shared__ uint32_t sh[8];
    uint32_t carry = 0;
    for (size_t base = 0; base < t; base += SCAN_TILE) {
        const size_t b = base + threadIdx.x * 4u;
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = b + q < t ? sums[b + q] : 0u;
        const uint32_t total = block_excl_scan4(v, sh);
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (b + q < t) sums[b + q] = v[q] + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(256) scan_add_kernel(uint32_t *a, size_t m, const uint32_t *tile_sums)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4u;
    const uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (base + q < m) a[base + q] += add;
}

// one warp per (row, chunk): barrier cells -> consecutive slots, planes -> table
__global__ void chain_fill_kernel(uint16_t *cls, uint16_t *cls_other, const SlabGeom g, const ChainPlanes pl,
                                  const ChainTable t, const uint32_t parity)
{
    const uint32_t nchunk = (g.P + CHUNK - 1) / CHUNK;
    const size_t nwarps_total = (size_t)g.rows * nchunk;
    const uint32_t lane = threadIdx.x & 31u;
    for (size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < nwarps_total;
         wid += ((size_t)gridDim.x * blockDim.x) >> 5) {
        size_t i0;
        const ushort4 c = chunk_words(cls, g, (uint32_t)(wid / nchunk), (uint32_t)(wid % nchunk), lane, &i0);
        const uint32_t mine = count4(c, CLS_BARRIER);
        uint32_t total;
        uint32_t e = t.chunk_base[wid] + warp_excl_scan(mine, lane, &total);
        if (total == 0) continue;
        const uint16_t cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (!(cw[q] & CLS_BARRIER)) continue;
            const size_t i = i0 + q;
            if (e < t.cap) {
                cls[i] = cw[q] | CLS_CHAIN | CLS_SLOT;
                cls_other[i] |= CLS_CHAIN | CLS_SLOT;  // the other class buffer may be stale elsewhere, never about these bits
                t.idx[e] = (uint32_t)i;
                t.flag[e] = 0;
#pragma unroll
                for (int d = 0; d < 8; d++) {
                    t.state[(size_t)d * t.cap + e] = pl.f0[d][i];
                    t.state[(size_t)(8 + d) * t.cap + e] = pl.f1[d][i];
                }
                t.state[(size_t)(CHAIN_ROW_REST + parity) * t.cap + e] = pl.R[i];
            }
            e++;
        }
    }
}

cudaError_t launch_chain_build(uint16_t *cls, uint16_t *cls_other, const SlabGeom &g, const ChainPlanes &pl,
                               const ChainTable &t, uint32_t *scan_scratch, uint32_t parity, cudaStream_t st)
{
    const size_t m = (size_t)g.rows * ((g.P + CHUNK - 1) / CHUNK);
    const size_t tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    if (m == 0 || tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    scan_tiles_kernel<<<(unsigned)tiles, 256, 0, st>>>(t.chunk_base, m, scan_scratch);
    scan_sums_kernel<<<1, 256, 0, st>>>(scan_scratch, tiles);
    scan_add_kernel<<<(unsigned)tiles, 256, 0, st>>>(t.chunk_base, m, scan_scratch);
    chain_fill_kernel<<<148 * 8, 256, 0, st>>>(cls, cls_other, g, pl, t, parity & 1u);
    return cudaGetLastError();
}

// table -> planes for one entry; parity = buffer the NEXT collide acts on (its rest value is the current one)
__device__ __forceinline__ void chain_store_entry(const ChainTable &t, size_t e, const ChainPlanes &pl, size_t i,
                                                  uint32_t parity)
{
#pragma unroll
    for (int d = 0; d < 8; d++) {
        pl.f0[d][i] = t.state[(size_t)d * t.cap + e];
        pl.f1[d][i] = t.state[(size_t)(8 + d) * t.cap + e];
    }
    pl.R[i] = t.state[(size_t)(CHAIN_ROW_REST + parity) * t.cap + e];
}

__global__ void chain_flush_kernel(const ChainTable t, const ChainPlanes pl, uint16_t *cls0, uint16_t *cls1,
                                   const uint32_t parity)
{
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < t.n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = t.idx[e];
        if (!(t.flag[e] & CHAIN_F_DEAD)) chain_store_entry(t, e, pl, i, parity);
        cls0[i] &= (uint16_t)~(CLS_CHAIN | CLS_SLOT);
        cls1[i] &= (uint16_t)~(CLS_CHAIN | CLS_SLOT);
    }
}

cudaError_t launch_chain_flush(const ChainTable &t, const ChainPlanes &pl, uint16_t *cls0, uint16_t *cls1,
                               uint32_t parity, cudaStream_t st)
{
    if (t.n == 0) return cudaSuccess;
    size_t nb = (t.n + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    chain_flush_kernel<<<(unsigned)nb, 256, 0, st>>>(t, pl, cls0, cls1, parity & 1u);
    return cudaGetLastError();
}

// ---- a paint touches some chain cells: move exactly those back into the planes ------------------------
// One warp per painted location: the slot is found by rank (chunk base + slot bits in front of the cell).
__global__ void chain_evict_kernel(const ChainTable t, const ChainPlanes pl, uint16_t *cls_cur, uint16_t *cls_other,
                                   const SlabGeom g, const uint64_t *pairs, size_t npairs, const uint32_t parity)
{
    const uint32_t nchunk = (g.P + CHUNK - 1) / CHUNK;
    const uint32_t lane = threadIdx.x & 31u;
    for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < npairs;
         w += ((size_t)gridDim.x * blockDim.x) >> 5) {
        const uint64_t loc = pairs[2 * w];
        const uint64_t gy = loc / g.W;
        if (gy < g.row0 || gy >= g.row0 + g.rows) continue;
        const uint32_t r = (uint32_t)(gy - g.row0), x = (uint32_t)(loc - gy * g.W);
        const size_t i = row_off(r, g.P) + x;
        if (!(cls_cur[i] & CLS_CHAIN)) continue;  // warp-uniform
        const uint32_t chunk = x / CHUNK;
        size_t i0;
        const ushort4 c = chunk_words(cls_cur, g, r, chunk, lane, &i0);
        const uint16_t cw[4] = {c.x, c.y, c.z, c.w};
        uint32_t before = 0;
#pragma unroll
        for (uint32_t q = 0; q < 4; q++)
            if ((cw[q] & CLS_SLOT) && chunk * CHUNK + lane * 4u + q < x) before++;
        const size_t e = (size_t)t.chunk_base[(size_t)r * nchunk + chunk] + __reduce_add_sync(0xffffffffu, before);
        if (e >= t.n) continue;
        if (lane < 8) pl.f0[lane][i] = t.state[(size_t)lane * t.cap + e];
        else if (lane < 16) pl.f1[lane - 8][i] = t.state[(size_t)lane * t.cap + e];
        else if (lane == 16) pl.R[i] = t.state[(size_t)(CHAIN_ROW_REST + parity) * t.cap + e];
        else if (lane == 17) t.flag[e] = CHAIN_F_DEAD;
        else if (lane == 18) cls_cur[i] &= (uint16_t)~CLS_CHAIN;
        else if (lane == 19) cls_other[i] &= (uint16_t)~CLS_CHAIN;
    }
}

cudaError_t launch_chain_evict(const ChainTable &t, const ChainPlanes &pl, uint16_t *cls_cur, uint16_t *cls_other,
                               const SlabGeom &g, const uint64_t *pairs, size_t npairs, uint32_t parity, cudaStream_t st)
{
    if (t.n == 0 || npairs == 0) return cudaSuccess;
    size_t nb = (npairs + 7) / 8;  // 8 warps per block
    if (nb > 148 * 16) nb = 148 * 16;
    chain_evict_kernel<<<(unsigned)nb, 256, 0, st>>>(t, pl, cls_cur, cls_other, g, pairs, npairs, parity & 1u);
    return cudaGetLastError();
}

__device__ __forceinline__ bool same_bits17(const float (&a)[8], const float (&b)[8], float r, const float (&a0)[8],
                                            const float (&b0)[8], float r0)
{
    bool same = __float_as_uint(r) == __float_as_uint(r0);
#pragma unroll
    for (int d = 0; d < 8; d++)
        same = same && __float_as_uint(a[d]) == __float_as_uint(a0[d]) &&
               __float_as_uint(b[d]) == __float_as_uint(b0[d]);
    return same;
}

// Advance every live, unsettled chain by nsteps collides: step s collides the copy in buffer (parity0 + s) % 2.
// Exact shortcut: once a pair of steps leaves all 17 floats bit-identical, every later pair does too (same
// deterministic map, same omega) — the entry is marked settled with both rest values and both moment triples in
// the table and is skipped from then on.  unsettle: omega changed since the last replay, recompute everything.
__global__ void __launch_bounds__(128) chain_replay_kernel(const ChainTable t, uint32_t nsteps, uint32_t parity0,
                                                           float omega, bool unsettle)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= t.n || nsteps == 0) return;
    const uint8_t fl = t.flag[e];
    if (fl & CHAIN_F_DEAD) return;  // evicted by a paint
    if ((fl & CHAIN_F_SETTLED) && !unsettle) return;
    const size_t cap = t.cap;
    float a[8], b[8];
#pragma unroll
    for (int d = 0; d < 8; d++) {
        a[d] = t.state[(size_t)d * cap + e];
        b[d] = t.state[(size_t)(8 + d) * cap + e];
    }
    uint32_t par = parity0 & 1u;  // buffer the next collide acts on
    float R = t.state[(size_t)(CHAIN_ROW_REST + par) * cap + e];
    float ma[3] = {0.f, 0.f, 0.f}, mb[3] = {0.f, 0.f, 0.f};
    bool did_a = false, did_b = false, changed = false;
    uint32_t left = nsteps;
    while (left >= 2) {
        float a0[8], b0[8];
        const float R0 = R;
#pragma unroll
        for (int d = 0; d < 8; d++) {
            a0[d] = a[d];
            b0[d] = b[d];
        }
        float Rmid;
        if (par == 0) {
            collide_cell(a, R, omega, ma[0], ma[1], ma[2]);
            Rmid = R;
            collide_cell(b, R, omega, mb[0], mb[1], mb[2]);
        } else {
            collide_cell(b, R, omega, mb[0], mb[1], mb[2]);
            Rmid = R;
            collide_cell(a, R, omega, ma[0], ma[1], ma[2]);
        }
        left -= 2;
        if (same_bits17(a, b, R, a0, b0, R0)) {
            // exact period-2 cycle: rest is R before a collide of buffer `par`, Rmid before one of the other
            if (changed) {
#pragma unroll
                for (int d = 0; d < 8; d++) {
                    t.state[(size_t)d * cap + e] = a[d];
                    t.state[(size_t)(8 + d) * cap + e] = b[d];
                }
            }
            t.state[(size_t)(CHAIN_ROW_REST + par) * cap + e] = R;
            t.state[(size_t)(CHAIN_ROW_REST + (par ^ 1u)) * cap + e] = Rmid;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                t.state[(size_t)(CHAIN_ROW_MOM + q) * cap + e] = ma[q];
                t.state[(size_t)(CHAIN_ROW_MOM + 3 + q) * cap + e] = mb[q];
            }
            t.flag[e] = CHAIN_F_SETTLED;
            return;
        }
        changed = true;
        did_a = did_b = true;
    }
    if (left) {
        if (par == 0) {
            collide_cell(a, R, omega, ma[0], ma[1], ma[2]);
            did_a = true;
        } else {
            collide_cell(b, R, omega, mb[0], mb[1], mb[2]);
            did_b = true;
        }
        par ^= 1u;
    }
#pragma unroll
    for (int d = 0; d < 8; d++) {
        t.state[(size_t)d * cap + e] = a[d];
        t.state[(size_t)(8 + d) * cap + e] = b[d];
    }
    t.state[(size_t)(CHAIN_ROW_REST + par) * cap + e] = R;
    if (did_a) {
#pragma unroll
        for (int q = 0; q < 3; q++) t.state[(size_t)(CHAIN_ROW_MOM + q) * cap + e] = ma[q];
    }
    if (did_b) {
#pragma unroll
        for (int q = 0; q < 3; q++) t.state[(size_t)(CHAIN_ROW_MOM + 3 + q) * cap + e] = mb[q];
    }
    if (fl & CHAIN_F_SETTLED) t.flag[e] = 0;
}

cudaError_t launch_chain_replay(const ChainTable &t, uint32_t nsteps, uint32_t parity0, float omega, bool unsettle,
                                cudaStream_t st)
{
    if (t.n == 0 || nsteps == 0) return cudaSuccess;
    const size_t nb = (t.n + 127) / 128;
    if (nb > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    chain_replay_kernel<<<(unsigned)nb, 128, 0, st>>>(t, nsteps, parity0, omega, unsettle);
    return cudaGetLastError();
}

// moments of the latest collide of buffer `last_parity`, table -> planes (scattered 4-byte stores: only the
// scalar kernel's moment-storing launches need the planes to hold them beforehand)
__global__ void chain_scatter_moments_kernel(const ChainTable t, const uint32_t last_parity, float *mx, float *my,
                                             float *rho)
{
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < t.n; e += (size_t)gridDim.x * blockDim.x) {
        if (t.flag[e] & CHAIN_F_DEAD) continue;
        const size_t i = t.idx[e];
        const float *m = t.state + (size_t)(CHAIN_ROW_MOM + 3 * last_parity) * t.cap + e;
        mx[i] = m[0];
        my[i] = m[t.cap];
        rho[i] = m[2 * t.cap];
    }
}

cudaError_t launch_chain_scatter_moments(const ChainTable &t, uint32_t last_parity, float *mx, float *my, float *rho,
                                         cudaStream_t st)
{
    if (t.n == 0) return cudaSuccess;
    size_t nb = (t.n + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    chain_scatter_moments_kernel<<<(unsigned)nb, 256, 0, st>>>(t, last_parity & 1u, mx, my, rho);
    return cudaGetLastError();
}

}  // namespace blbmk

// ================================================================================================
// Colour maps: rewritten_shaders/color_map/{jet,viridis,inferno}.wgsl.
// Piece-wise linear LUT of the output field, barrier cells black.  rgb is rows x W x 3 floats, dense.
// ================================================================================================
namespace blbmk {

__constant__ float c_cmap_nodes[3][9][3] = {
    // Inferno = 0
    {{0.98828125f, 1.0f, 0.64453125f}, {0.97265625f, 0.55859375f, 0.0390625f}, {0.73828125f, 0.21875f, 0.33203125f},
     {0.34375f, 0.06640625f, 0.43359375f}, {0.0f, 0.0f, 0.01853125f}},
    // Viridis = 1
    {{0.9921875f, 0.90625f, 0.1484375f}, {0.3671875f, 0.7890625f, 0.3828125f}, {0.1328125f, 0.56640625f, 0.55078125f},
     {0.23046875f, 0.32421875f, 0.546875f}, {0.265625f, 0.0078125f, 0.33203125f}},
    // Jet = 2
    {{0.0f, 0.0f, 0.5f}, {0.0f, 0.0f, 1.0f}, {0.0f, 0.5f, 1.0f}, {0.0f, 1.0f, 1.0f}, {0.5f, 1.0f, 0.5f},
     {1.0f, 1.0f, 0.0f}, {1.0f, 0.5f, 0.0f}, {1.0f, 0.0f, 0.0f}, {0.5f, 0.0f, 0.0f}}};

__global__ void color_map_kernel(const float *__restrict__ out, const uint8_t *__restrict__ mask, float *__restrict__ rgb,
                                 const SlabGeom g, const int map)
{
    const int nseg = map == 2 ? 8 : 4;
    const float scale = map == 2 ? 20.0f : 15.0f;
    const float lo = (float)(-nseg / 2), hi = (float)(nseg / 2);
    const int ilo = -nseg / 2;
    const size_t total = (size_t)g.rows * g.W;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / g.W);
        const uint32_t x = (uint32_t)(t - (size_t)r * g.W);
        float c = __fmul_rn(scale, out[row_off(r, g.P) + x]);
        c = fminf(fmaxf(c, lo), hi);
        const int block = (int)floorf(c);
        float cr, cg, cb;
        if (block >= ilo && block < ilo + nseg) {
            const float rw = __fadd_rn((float)(-block), c);
            const float lw = __fsub_rn(1.0f, rw);
            const float *A = c_cmap_nodes[map][block - ilo], *B = c_cmap_nodes[map][block - ilo + 1];
            cr = mul_add(lw, A[0], __fmul_rn(rw, B[0]));
            cg = mul_add(lw, A[1], __fmul_rn(rw, B[1]));
            cb = mul_add(lw, A[2], __fmul_rn(rw, B[2]));
        } else {
            cr = c_cmap_nodes[map][nseg][0];
            cg = c_cmap_nodes[map][nseg][1];
            cb = c_cmap_nodes[map][nseg][2];
        }
        if (mask[mask_row_off(r, g.P) + x] == 1) cr = cg = cb = 0.0f;
        rgb[3 * t + 0] = cr;
        rgb[3 * t + 1] = cg;
        rgb[3 * t + 2] = cb;
    }
}

cudaError_t launch_color_map(const float *out, const uint8_t *mask, float *rgb, const SlabGeom &g, int map,
                             cudaStream_t st)
{
    const size_t total = (size_t)g.rows * g.W;
    size_t nb = (total + 255) / 256;
    if (nb > 148 * 16) nb = 148 * 16;
    if (nb == 0) return cudaSuccess;
    color_map_kernel<<<(unsigned)nb, 256, 0, st>>>(out, mask, rgb, g, map);
    return cudaGetLastError();
}

}  // namespace blbmk

// ---- eager module loading ---------------------------------------------------------------------------
// CUDA loads kernels lazily, and loading one may wait for the device to go idle.  Linked slabs keep a
// spinning wait kernel on the device until the neighbour signals, so a first-use load on the host thread that
// must enqueue that very signal would deadlock (until the wait times out).  Touch every kernel once up front.
namespace blbmk {
#define BLBM_TOUCH(k)                                                    \
    do {                                                                 \
        cudaFuncAttributes a__;                                          \
        cudaError_t e__ = cudaFuncGetAttributes(&a__, k);                \
        if (e__ != cudaSuccess) return e__;                              \
    } while (0)

cudaError_t preload_aux_kernels()
{
    BLBM_TOUCH(fill_rows_kernel);
    BLBM_TOUCH(mask_init_kernel);
    BLBM_TOUCH(mask_scatter_kernel);
    BLBM_TOUCH(build_class_kernel<false>);
    BLBM_TOUCH(build_class_kernel<true>);
    BLBM_TOUCH(mask_scatter_args_kernel);
    BLBM_TOUCH(paint_small_kernel);
    BLBM_TOUCH(precollision_moments_kernel);
    BLBM_TOUCH(curl_vec4_kernel);
    BLBM_TOUCH(summary_kernel<0>);
    BLBM_TOUCH(summary_kernel<1>);
    BLBM_TOUCH(summary_kernel<2>);
    BLBM_TOUCH(summary_kernel<3>);
    BLBM_TOUCH(summary_kernel<4>);
    BLBM_TOUCH(reduce_kernel);
    BLBM_TOUCH(signal_kernel);
    BLBM_TOUCH(wait_kernel);
    BLBM_TOUCH(chain_count_kernel);
    BLBM_TOUCH(mailbox_kernel);
    BLBM_TOUCH(scan_tiles_kernel);
    BLBM_TOUCH(scan_sums_kernel);
    BLBM_TOUCH(scan_add_kernel);
    BLBM_TOUCH(chain_fill_kernel);
    BLBM_TOUCH(chain_flush_kernel);
    BLBM_TOUCH(chain_evict_kernel);
    BLBM_TOUCH(chain_replay_kernel);
    BLBM_TOUCH(chain_scatter_moments_kernel);
    BLBM_TOUCH(color_map_kernel);
    return cudaSuccess;
}
#undef BLBM_TOUCH
}  // namespace blbmk
