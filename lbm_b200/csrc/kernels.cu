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
        if (lazy_barrier) {  // the planes already hold this cell's moments (chain_scatter_moments_kernel)
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
// the warp edge fetches the single missing float itself.  A block is 32 x V4_ROWS threads = 128 x V4_ROWS
// cells (V4_ROWS = 4 by default; 1, 2, 8, 16 through blbm_set_tuning).  Cells that need anything but a plain
// pull (class word != 0, column W-1, ragged row end) take a per-cell fix-up path after the vector loads.
//
// DENSE selects the flavour of the bounce-back fix-up (same results in every flavour):
//   0  sparse: one branch per direction, and the class words are read only after the pulls have arrived, inside
//      a warp-uniform branch on the chunk flag (best where obstacles are rare: the clean path is 64 registers,
//      no spill; empty 16384^2 channel 2.89 ms);
//   1  dense: branch-free selects over the directions (every direction is needed by some lane anyway);
//   2  dense + staged (what the handle picks wherever the barrier-chain table is active): the six own-row vectors
//      the bounce-back needs (the cell's own n, s, ne, nw, se, sw) are fetched into shared memory with cp.async
//      together with the pulls, so they cost no registers while in flight and the fix-up no longer pays a second,
//      dependent round trip to L2 after the class words arrive (porous 16384^2: 3.41 -> 3.26 ms);
//   3  as 2, with the class words read up front without consulting the chunk flag (+2 B/cell in clean chunks;
//      within 0.3 % of flavour 2 on a fully porous lattice).
// PACKED collides cell pairs with Blackwell's packed fp32 adds (FADD2): same bits, measured slower (registers).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- in-kernel epoch handshake of linked slabs (LinkSync, blbm_internal.cuh) ---------------------------------
__device__ __forceinline__ void link_wait(const unsigned long long *from_up, const unsigned long long *from_dn,
                                       const unsigned long long epoch, int *err_flag, const unsigned long long timeout_ns)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned long long u = epoch, d = epoch;
        if (from_up) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(u) : "l"(from_up) : "memory");
        if (from_dn) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(d) : "l"(from_dn) : "memory");
        if (u >= epoch && d >= epoch) return;
        if (*reinterpret_cast<volatile int *>(err_flag)) return;  // an earlier wait already gave up
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {  // never hang the GPU on a dead neighbour: flag and carry on
            *err_flag = 1;
            return;
        }
        __nanosleep(100);
    }
}

// the last of `target` blocks to arrive publishes the epoch to the neighbour and re-arms the counter
__device__ __forceinline__ void link_arrive(unsigned int *counter, const unsigned int target, unsigned long long *remote_flag,
                                         const unsigned long long epoch)
{
    __threadfence_system();
    if (atomicAdd(counter, 1u) == target - 1u) {
        *counter = 0u;  // the next launch on this stream starts after this one has drained
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote_flag), "l"(epoch) : "memory");
    }
}
// Programmatic dependent launch: a step launched with the programmatic-serialization attribute may have its blocks
// scheduled while the previous kernel on the stream drains (that kernel's blocks have all passed pdl_trigger or
// exited); nothing of the previous kernel's results - nor anything it still reads - may be touched before pdl_wait
// returns, which is when that kernel has completed and its stores are visible.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// own-row plane of staging slot q (the opposite of the directions that move in y)
__host__ __device__ constexpr int stage_dir(int q) { return q == 0 ? D_NW : q == 1 ? D_N : q == 2 ? D_NE : q == 3 ? D_SW : q == 4 ? D_S : D_SE; }
__host__ __device__ constexpr int stage_slot(int d) { return d == D_NW ? 0 : d == D_N ? 1 : d == D_NE ? 2 : d == D_SW ? 3 : d == D_S ? 4 : 5; }

// The moment-storing variant (one launch per call) may use 85 registers instead of spilling at 64.
// IDX: uint32_t plane offsets where a plane has fewer than 2^32 elements (every slab that fits a B200 at
// rows-per-block 4), size_t otherwise.  The grid is (chunks along x, row blocks [, overflow of row blocks]) so that
// no thread divides a linear block index.
constexpr uint32_t GRID_Y = 32768;
// LINKED: the epoch handshake with the neighbouring slabs runs inside the kernel (LinkSync); a separate instantiation
// so that the unlinked kernel keeps its 64-register budget.
template <bool MOM, int V4_ROWS, int DENSE, bool PACKED, typename IDX, bool LINKED = false>
__global__ void __launch_bounds__(32 * V4_ROWS, (DENSE || LINKED || MOM) ? 1024 / (32 * V4_ROWS) : 0) step_vec4_kernel(const StepParams p)
{
    constexpr bool STAGED = DENSE >= 2;
    constexpr bool EAGER_CLS = DENSE == 3;  // class words read up front, without the chunk-flag test
    // sparse flavour: class words read only after the pulls have arrived, inside a warp-uniform branch on the chunk
    // flag - a predicated load up front makes the in-order issue wait for the flag byte before the remaining pulls
    constexpr bool LATE_CLS = DENSE == 0;
    __shared__ float4 own_s[STAGED ? 6 : 1][STAGED ? 32 * V4_ROWS : 1];
    if (!LINKED) pdl_trigger();  // the next step's blocks may be scheduled behind this kernel's last wave
    const uint32_t tid = threadIdx.y * 32u + threadIdx.x;
    const uint32_t nbx = gridDim.x;  // = ceil(P / 128)
    const uint32_t bx = blockIdx.x;
    uint32_t rb = blockIdx.z * GRID_Y + blockIdx.y;  // row block
    bool face_up = false, face_dn = false;
    if (LINKED) {
        // linked slab, handshake in-kernel: the row blocks at the two faces come first in launch order ...
        const uint32_t nrb = (p.rows + V4_ROWS - 1) / V4_ROWS;
        if (nrb >= 4) rb = rb == 0 ? 0u : rb == 1 ? nrb - 1 : rb == 2 ? nrb - 2 : rb - 2;
        // ... and are the only ones that touch halo rows: row 0 gathers from the halo above and rows 0-1 store into
        // the slab above; rows-1 gathers from the halo below (rows-2 too, through the flat-index wrap of column
        // W-1) and stores into the slab below
        face_up = p.link.sig_up != nullptr && rb == 0;
        face_dn = p.link.sig_dn != nullptr && (rb == (p.rows - 1) / V4_ROWS || rb == (p.rows - 2) / V4_ROWS);
        if (face_up || face_dn) {
            if (tid == 0)
                link_wait(face_up ? p.link.wait_up : nullptr, face_dn ? p.link.wait_dn : nullptr, p.link.wait_epoch,
                          p.link.err_flag, p.link.timeout_ns);
            __syncthreads();  // every warp of the block is still here
        }
    }
    const uint32_t r = rb * V4_ROWS + threadIdx.y;
    if (r >= p.rows) return;  // whole warps leave together (a warp is one row)
    const uint32_t lane = threadIdx.x;
    const uint32_t x4 = bx * 128u + lane * 4u;
    const bool valid = x4 < p.W;  // P is a multiple of 32, so x4 + 3 < P always
    const uint32_t P = p.P;
    const IDX i = (IDX)row_off(r, P) + x4;
    const unsigned FULL = 0xffffffffu;

    float g[4][8];  // [cell][direction], gathered pre-collision state
    ushort4 c4 = make_ushort4(0, 0, 0, 0);
    float4 vn, vs, ve, vw, vne, vnw, vse, vsw, vr;
    vn = vs = ve = vw = vne = vnw = vse = vsw = vr = make_float4(0.f, 0.f, 0.f, 0.f);
    // one flag byte per (row, 128-cell chunk) — the same address for the whole warp — tells whether any
    // class word of the chunk is non-zero; clean chunks never touch the class plane.  Issue order matters (the SM
    // issues in order and a predicated load waits for its predicate): the flag byte is requested first and
    // consumed LAST, after every independent load of the thread is in flight; the warp-edge floats are requested
    // with the pulls, not after the shuffles that wait for the pulls.
    if (!LINKED) pdl_wait();  // first access to anything an earlier kernel wrote (or still reads)
    const uint8_t flag = EAGER_CLS ? (uint8_t)1 : p.rowflag[(size_t)r * nbx + bx];
    float le = 0.f, lne = 0.f, lse = 0.f, rw = 0.f, rnw = 0.f, rsw = 0.f;
    const bool edge_l = valid && lane == 0 && x4 != 0;
    const bool edge_r = valid && lane == 31;  // x4+4 <= P: still inside the row pitch (or x=0 of the next row: unused)
    if (valid) {
        if (EAGER_CLS) c4 = *reinterpret_cast<const ushort4 *>(p.cls + i);
        vn = ldg4(p.X[D_N] + i + P);
        vne = ldg4(p.X[D_NE] + i + P);
        vnw = ldg4(p.X[D_NW] + i + P);
        ve = ldg4(p.X[D_E] + i);
        vw = ldg4(p.X[D_W] + i);
        vs = ldg4(p.X[D_S] + i - P);
        vse = ldg4(p.X[D_SE] + i - P);
        vsw = ldg4(p.X[D_SW] + i - P);
        vr = ldg4(p.R + i);
        if (edge_l) {
            le = p.X[D_E][i - 1];
            lne = p.X[D_NE][i + P - 1];
            lse = p.X[D_SE][i - P - 1];
        }
        if (edge_r) {
            rw = p.X[D_W][i + 4];
            rnw = p.X[D_NW][i + P + 4];
            rsw = p.X[D_SW][i - P + 4];
        }
        if (STAGED) {
#pragma unroll
            for (int q = 0; q < 6; q++) cp_async16(&own_s[q][tid], p.X[stage_dir(q)] + i);
        } else {
            asm volatile("" ::: "memory");  // keep the class-word load below behind the loads above
        }
        if (!EAGER_CLS && !LATE_CLS && flag != 0) c4 = *reinterpret_cast<const ushort4 *>(p.cls + i);
    }
    // element x4-1 of the east-moving populations, element x4+4 of the west-moving ones: from the neighbouring
    // lane, except at the warp edges
    {
        const float se_ = __shfl_up_sync(FULL, ve.w, 1), sne_ = __shfl_up_sync(FULL, vne.w, 1),
                    sse_ = __shfl_up_sync(FULL, vse.w, 1);
        const float sw_ = __shfl_down_sync(FULL, vw.x, 1), snw_ = __shfl_down_sync(FULL, vnw.x, 1),
                    ssw_ = __shfl_down_sync(FULL, vsw.x, 1);
        if (!edge_l) {
            le = se_;
            lne = sne_;
            lse = sse_;
        }
        if (!edge_r) {
            rw = sw_;
            rnw = snw_;
            rsw = ssw_;
        }
    }
    if (LATE_CLS && valid && flag != 0) c4 = *reinterpret_cast<const ushort4 *>(p.cls + i);
    // moment-storing launch with the barrier-chain table active: the table slot of this lane's first slot-owning
    // cell = chunk base + slot bits in front of it (the table is ordered by plane offset); the whole warp takes part
    uint32_t slot_e = 0;
    if (MOM && p.chain_mom != nullptr) {
        uint32_t total;
        const uint32_t before = warp_excl_scan(count4(c4, CLS_SLOT), lane, &total);
        if (total) slot_e = p.chunk_base[(size_t)r * nbx + bx] + before;
    }
    if (valid) {  // (lanes past the row end issued no cp.async and have nothing to store)
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
        if (STAGED) cp_async_wait_all();
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
                    else if (STAGED) own[q] = own_s[stage_slot(dir_opp(d))][tid];
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
    finish_group<MOM, PACKED, IDX>(p, i, x4, r, g, c0, c1, c2, c3, vr, slot_e);
    if (STAGED) cp_async_wait_all();  // nothing may still be landing in shared memory when the block retires
    }
    if (LINKED && (face_up || face_dn)) {
        // every thread orders its own stores into the neighbour before the block arrives; the barrier counts the
        // live warps only (rows past the slab end have left)
        __threadfence_system();
        const uint32_t live = (p.rows - rb * V4_ROWS < (uint32_t)V4_ROWS ? p.rows - rb * V4_ROWS : (uint32_t)V4_ROWS) * 32u;
        asm volatile("bar.sync 1, %0;" ::"r"(live) : "memory");
        if (tid == 0) {
            const uint32_t ndn = (p.rows - 1) / V4_ROWS != (p.rows - 2) / V4_ROWS ? 2u : 1u;
            if (face_up) link_arrive(&p.link.done[0], nbx, p.link.sig_up, p.link.sig_epoch);
            if (face_dn) link_arrive(&p.link.done[1], nbx * ndn, p.link.sig_dn, p.link.sig_epoch);
        }
    }
}

// every element offset a launch forms (rows+3 device rows, one float4 past the last) fits 32 bits
static bool plane_fits_u32(const StepParams &p) { return ((uint64_t)p.rows + 3u) * p.P + 8u < (1ull << 32); }

bool vec4_links_in_kernel(int block_rows, bool packed) { return block_rows == 4 && !packed; }

template <int V4_ROWS, int DENSE, bool PACKED, typename IDX, bool LINKED = false>
static cudaError_t launch_vec4_idx(const StepParams &p, bool mom, bool pdl, cudaStream_t st)
{
    const uint32_t nbx = (p.P + 127u) / 128u;
    const uint32_t nrb = (p.rows + V4_ROWS - 1) / V4_ROWS;
    if (nbx == 0 || nrb == 0) return cudaErrorInvalidConfiguration;
    dim3 grid(nbx, nrb < GRID_Y ? nrb : GRID_Y, (nrb + GRID_Y - 1) / GRID_Y), block(32, V4_ROWS);
    if (grid.z > 65535u) return cudaErrorInvalidConfiguration;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && !LINKED) ? 1 : 0;
    if (mom) return cudaLaunchKernelEx(&cfg, step_vec4_kernel<true, V4_ROWS, DENSE, PACKED, IDX, LINKED>, p);
    return cudaLaunchKernelEx(&cfg, step_vec4_kernel<false, V4_ROWS, DENSE, PACKED, IDX, LINKED>, p);
}


// The default block shape (4 rows) exists in every flavour; the other shapes (A/B knob) only with scalar adds and
// 64-bit offsets.
template <int DENSE>
static cudaError_t launch_vec4_flavour(const StepParams &p, bool mom, int block_rows, bool packed, bool index32,
                                       bool pdl, cudaStream_t st)
{
    if (p.link.sig_epoch != 0) {
        // in-kernel handshake: default block shape, scalar adds (vec4_links_in_kernel)
        if (block_rows != 4 || packed) return cudaErrorInvalidConfiguration;
        if (index32 && plane_fits_u32(p)) return launch_vec4_idx<4, DENSE, false, uint32_t, true>(p, mom, false, st);
        return launch_vec4_idx<4, DENSE, false, size_t, true>(p, mom, false, st);
    }
    switch (block_rows) {
    case 1: return launch_vec4_idx<1, DENSE, false, size_t>(p, mom, pdl, st);
    case 2: return launch_vec4_idx<2, DENSE, false, size_t>(p, mom, pdl, st);
    case 8: return launch_vec4_idx<8, DENSE, false, size_t>(p, mom, pdl, st);
    case 16: return launch_vec4_idx<16, DENSE, false, size_t>(p, mom, pdl, st);
    default: break;
    }
    if (index32 && plane_fits_u32(p))
        return packed ? launch_vec4_idx<4, DENSE, true, uint32_t>(p, mom, pdl, st)
                      : launch_vec4_idx<4, DENSE, false, uint32_t>(p, mom, pdl, st);
    return packed ? launch_vec4_idx<4, DENSE, true, size_t>(p, mom, pdl, st)
                  : launch_vec4_idx<4, DENSE, false, size_t>(p, mom, pdl, st);
}

cudaError_t launch_step_vec4(const StepParams &p, int mode, bool mom, int block_rows, int dense_obstacles,
                             bool packed, bool index32, bool pdl, cudaStream_t st)
{
    if (mode != MODE_FUSED) return launch_step_scalar(p, mode, mom, st);
    switch (dense_obstacles) {
    case 3: return launch_vec4_flavour<3>(p, mom, block_rows, packed, index32, pdl, st);
    case 2: return launch_vec4_flavour<2>(p, mom, block_rows, packed, index32, pdl, st);
    case 1: return launch_vec4_flavour<1>(p, mom, block_rows, packed, index32, pdl, st);
    default: return launch_vec4_flavour<0>(p, mom, block_rows, packed, index32, pdl, st);
    }
}

// see preload_aux_kernels(): force the (lazy) load of every step-kernel instantiation
#define BLBM_TOUCH(...)                                                  \
    do {                                                                 \
        cudaFuncAttributes a__;                                          \
        cudaError_t e__ = cudaFuncGetAttributes(&a__, __VA_ARGS__);      \
        if (e__ != cudaSuccess) return e__;                              \
    } while (0)

template <int ROWS, int DENSE, bool PACKED, typename IDX, bool LINKED = false>
static cudaError_t touch_vec4()
{
    BLBM_TOUCH(step_vec4_kernel<false, ROWS, DENSE, PACKED, IDX, LINKED>);
    BLBM_TOUCH(step_vec4_kernel<true, ROWS, DENSE, PACKED, IDX, LINKED>);
    return cudaSuccess;
}

template <int DENSE>
static cudaError_t touch_vec4_flavour()
{
    cudaError_t e;
    if ((e = touch_vec4<1, DENSE, false, size_t>()) != cudaSuccess) return e;
    if ((e = touch_vec4<2, DENSE, false, size_t>()) != cudaSuccess) return e;
    if ((e = touch_vec4<8, DENSE, false, size_t>()) != cudaSuccess) return e;
    if ((e = touch_vec4<16, DENSE, false, size_t>()) != cudaSuccess) return e;
    if ((e = touch_vec4<4, DENSE, false, size_t>()) != cudaSuccess) return e;
    if ((e = touch_vec4<4, DENSE, true, size_t>()) != cudaSuccess) return e;
    if ((e = touch_vec4<4, DENSE, false, uint32_t>()) != cudaSuccess) return e;
    if ((e = touch_vec4<4, DENSE, false, uint32_t, true>()) != cudaSuccess) return e;
    if ((e = touch_vec4<4, DENSE, false, size_t, true>()) != cudaSuccess) return e;
    return touch_vec4<4, DENSE, true, uint32_t>();
}

cudaError_t preload_step_kernels()
{
    BLBM_TOUCH(step_scalar_kernel<MODE_FUSED, false>);
    BLBM_TOUCH(step_scalar_kernel<MODE_FUSED, true>);
    BLBM_TOUCH(step_scalar_kernel<MODE_COLLIDE_ONLY, false>);
    BLBM_TOUCH(step_scalar_kernel<MODE_COLLIDE_ONLY, true>);
    BLBM_TOUCH(step_scalar_kernel<MODE_STREAM_ONLY, false>);
    cudaError_t e;
    if ((e = touch_vec4_flavour<0>()) != cudaSuccess) return e;
    if ((e = touch_vec4_flavour<1>()) != cudaSuccess) return e;
    if ((e = touch_vec4_flavour<2>()) != cudaSuccess) return e;
    return touch_vec4_flavour<3>();
}
#undef BLBM_TOUCH

}  // namespace blbmk
