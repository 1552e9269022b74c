// step_common.cuh — the second half of the four-cells-per-thread step kernel (register/shuffle gather in
// kernels.cu): everything after the gather of one group of four
// consecutive cells — own-copy reload of skipped cells, the flat-index wrap at column W-1, collision,
// stores, moments, halo mirroring.
#pragma once
#include "blbm_internal.cuh"

namespace blbmk {

__device__ __forceinline__ float4 ldg4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void stg4(float *p, const float4 v) { *reinterpret_cast<float4 *>(p) = v; }
// (st.global.cs — the "streaming" hint — on these stores was measured: porous 3.12 vs 3.08 ms, channel no change.)

// Offset of the cell population d is pulled from, for the cell at (x, device-row offset `i`).
// Column W-1 follows the reference's flat indexing: i+1 is (0, y+1)  (SURVEY.md section 8 a-4).
__device__ __forceinline__ size_t pull_src(const size_t i, const uint32_t x, const int d, const uint32_t W,
                                           const uint32_t P)
{
    const int dx = dir_dx(d), dy = dir_dy(d);
    if (dx < 0 && x == W - 1) {
        // source = flat index i + 1 - dy*W  ->  column 0 of row (y + 1 - dy)
        return i - x + (size_t)((1 - dy)) * P;
    }
    return (size_t)((ptrdiff_t)i - dx - (ptrdiff_t)dy * (ptrdiff_t)P);
}

// does a cell with class word c continue from its own copy in the destination buffer?
__device__ __forceinline__ bool reloads_own(const uint32_t c)
{
    return (c & (CLS_SKIP | CLS_CHAIN)) == CLS_SKIP;
}

// g[q][d]: gathered (pulled, bounce-back already applied) populations of cells x4+q, q = 0..3, at plane
// offset i; c0..c3 their class words; vr their rest populations.
// PACKED: collide the four cells as two pairs with packed fp32 adds (collide_pair, blbm_internal.cuh).
// IDX: type of the plane offset i — uint32_t where a plane has fewer than 2^32 elements (one IMAD.WIDE per
// address instead of a 64-bit add pair), size_t otherwise.
template <bool MOM, bool PACKED = false, typename IDX = size_t>
__device__ __forceinline__ void finish_group(const StepParams &p, const IDX i, const uint32_t x4, const uint32_t r,
                                             float (&g)[4][8], const uint32_t c0, const uint32_t c1,
                                             const uint32_t c2, const uint32_t c3, const float4 vr,
                                             const uint32_t slot_e)
{
    const uint32_t P = p.P, W = p.W;
    const uint32_t cany = c0 | c1 | c2 | c3;
    const bool ragged = x4 + 4 > W - 1;  // group holds column W-1 or cells beyond the row end
    if ((cany & CLS_SKIP) || ragged) {
        const bool o0 = reloads_own(c0), o1 = reloads_own(c1), o2 = reloads_own(c2), o3 = reloads_own(c3);
        if (o0 | o1 | o2 | o3) {
            // skipped cells continue from their own stale copy in the destination buffer
#pragma unroll
            for (int d = 0; d < 8; d++) {
                const float4 o = ldg4(p.Y[d] + i);
                if (o0) g[0][d] = o.x;
                if (o1) g[1][d] = o.y;
                if (o2) g[2][d] = o.z;
                if (o3) g[3][d] = o.w;
            }
        }
        if (ragged) {
            const uint32_t j = W - 1 - x4;  // position of column W-1 inside the group (0..3), if present
            if (j < 4) {
                const uint32_t cj = j == 0 ? c0 : (j == 1 ? c1 : (j == 2 ? c2 : c3));
                if (!(cj & CLS_SKIP)) {
                    const size_t ij = (size_t)i + j;
#pragma unroll
                    for (int d = 0; d < 8; d++) {
                        if (dir_dx(d) < 0 && !(cj & cls_upstream_bit(d))) {
                            const float v = p.X[d][pull_src(ij, W - 1, d, W, P)];
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if (q == (int)j) g[q][d] = v;
                        }
                    }
                }
            }
        }
    }

    float rr[4] = {vr.x, vr.y, vr.z, vr.w};
    if (MOM) {
        // The moments of the pre-collision state — exactly the sums collide_cell() forms first — are computed and
        // stored BEFORE the collision, so that twelve values do not stay live across it (the collision below forms
        // the same sums again; the empty asm keeps the compiler from carrying them over).  64 registers, like the
        // launches that store no moments.
        float mx[4], my[4], rho[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            precollision_moments(g[q], mx[q], my[q], rho[q]);
            rho[q] = __fadd_rn(rho[q], rr[q]);
        }
        if (cany & CLS_CHAIN) {
            // chain cells: their plane slots are don't-care, the moments of their latest collide are in the table,
            // slot_e onwards in plane order (dead slots still count)
            const uint32_t cq[4] = {c0, c1, c2, c3};
            uint32_t e = slot_e;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (cq[q] & CLS_CHAIN) {
                    mx[q] = p.chain_mom[e];
                    my[q] = p.chain_mom[(size_t)p.chain_cap + e];
                    rho[q] = p.chain_mom[2 * (size_t)p.chain_cap + e];
                }
                if (cq[q] & CLS_SLOT) e++;
            }
        }
        stg4(p.mx + i, make_float4(mx[0], mx[1], mx[2], mx[3]));
        stg4(p.my + i, make_float4(my[0], my[1], my[2], my[3]));
        stg4(p.rho + i, make_float4(rho[0], rho[1], rho[2], rho[3]));
        // moment rows a neighbouring slab's curl stencil reads
        if (r == 0 && p.push.up_n) {
            stg4(p.push.up_mx + x4, make_float4(mx[0], mx[1], mx[2], mx[3]));
            stg4(p.push.up_my + x4, make_float4(my[0], my[1], my[2], my[3]));
        }
        if (r == p.rows - 1 && p.push.dn_s) {
            stg4(p.push.dn_mx + x4, make_float4(mx[0], mx[1], mx[2], mx[3]));
            stg4(p.push.dn_my + x4, make_float4(my[0], my[1], my[2], my[3]));
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            asm volatile("" : "+f"(rr[q]));
#pragma unroll
            for (int d = 0; d < 8; d++) asm volatile("" : "+f"(g[q][d]));
        }
    }
    {
        float mx[4], my[4], rho[4];  // formed again by the collision; dead in every launch
        if (PACKED) {
            collide_quad_packed(g, rr, p.omega, mx, my, rho);
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) collide_cell(g[q], rr[q], p.omega, mx[q], my[q], rho[q]);
        }
    }

    // cells beyond the row end (ragged W) are padding: written, never read
    stg4(p.R + i, make_float4(rr[0], rr[1], rr[2], rr[3]));
#pragma unroll
    for (int d = 0; d < 8; d++) stg4(p.Y[d] + i, make_float4(g[0][d], g[1][d], g[2][d], g[3][d]));
    // mirror the cells a neighbouring slab gathers from into its halo rows (NVLink stores)
    if (r == 0 && p.push.up_n) {
        stg4(p.push.up_n + x4, make_float4(g[0][D_N], g[1][D_N], g[2][D_N], g[3][D_N]));
        stg4(p.push.up_ne + x4, make_float4(g[0][D_NE], g[1][D_NE], g[2][D_NE], g[3][D_NE]));
        stg4(p.push.up_nw + x4, make_float4(g[0][D_NW], g[1][D_NW], g[2][D_NW], g[3][D_NW]));
        if (x4 == 0) p.push.up_w[0] = g[0][D_W];
    }
    if (r == 1 && x4 == 0 && p.push.up_nw2) p.push.up_nw2[0] = g[0][D_NW];
    if (r == p.rows - 1 && p.push.dn_s) {
        stg4(p.push.dn_s + x4, make_float4(g[0][D_S], g[1][D_S], g[2][D_S], g[3][D_S]));
        stg4(p.push.dn_se + x4, make_float4(g[0][D_SE], g[1][D_SE], g[2][D_SE], g[3][D_SE]));
        stg4(p.push.dn_sw + x4, make_float4(g[0][D_SW], g[1][D_SW], g[2][D_SW], g[3][D_SW]));
    }
}

}  // namespace blbmk
