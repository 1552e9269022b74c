/*
 * lbm_oracle.c — CPU restatement of lbm-wgpu's per-timestep lattice update.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lbm_b200/, include/) may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU arm.
 *
 * PARITY STATUS: the reference (PPLUSCHT/LBM) ships no tests or golden vectors and has no read-back path; its WebGPU
 * half cannot run here (Rust -> wasm32 + browser WebGPU; no cargo / Vulkan / lavapipe in the image), so no output of
 * an actual wgpu execution exists.  What pins this oracle:
 *   (a) OUTPUTS OF THE REFERENCE ITSELF for its host-side rows: oracle/wasm_mini.py executes functions of the
 *       reference's shipped binary (lbm-wgpu/pkg/lbm_wgpu_bg.wasm): set_equil, LBM::single_cell, LBM::draw_shape, LBM::iterate,
 *       Line::new, Line::new_erased, Curve::add/erase_segment -> tests/golden/wasm_golden.npz; this file's initial
 *       populations and single_cell state must equal them bit for bit (tests/test_wasm_pin.py);
 *   (b) the reference's own WGSL shader text (verified to be embedded byte for byte in that binary), executed
 *       invocation by invocation by oracle/wgsl_interp.py (tests/golden/wgsl_golden.npz: random API scripts, config-1
 *       miniatures, porous mask, single_cell presets, paints on every special cell class, every summary statistic
 *       and colour map) and, all invocations at once, by oracle/wgsl_simt.py at the reference's own sizes
 *       (wgsl_config1.npz: BASELINE configs[0] in full, 512 x 256 for 10,000 steps; wgsl_wide.npz: configs[1] at
 *       4096^2 for 1,000 steps, a 16384-wide porous strip, colour maps at 700 x 300); this file must reproduce
 *       every buffer / digest bit for bit (tests/test_wgsl_pin.py), so arithmetic, association, guards and index
 *       helpers are pinned to the reference's source; the host-side dispatch order is checked against the command
 *       stream of the binary's own LBM::iterate run on a mock wgpu context (a); only the bind-group contents are
 *       restated from lbm.rs;
 *   (c) an independent numpy restatement (oracle/lbm_numpy.py) that must agree bit for bit; (d) analytic KATs.
 * Still UNPINNED: what a particular WebGPU backend does where WGSL leaves room (FMA contraction, out-of-range
 * access) -- see the semantics below.
 *
 * The reference's structure is preserved on purpose: 8 passes per step over 2x9 fp32 SoA arrays,
 * three moment arrays and a u32 barrier mask, every intermediate going through fp32 memory exactly
 * where the WGSL stores it.  Compile with -O2 -ffp-contract=off (no FMA contraction), no fast-math.
 * Semantics where WebGPU is implementation-defined: out-of-range reads return 0 / 0.0f, out-of-range
 * writes are dropped (robust-buffer-access behaviour: what a Vulkan device with robustBufferAccess2 - lavapipe, the
 * backend SURVEY.md section 8 names - does).  lbm_oracle_set_oob_clamp(o, 1) selects the other behaviour the WebGPU
 * specification allows and browsers implement (Tint's robustness transform, naga's `Restrict` policy): the index is
 * clamped, min(u32(index), length - 1).  It is NOT a corner case: cell (W-1, H-2) pulls its north-west-moving
 * population from index W*H every step - 0.0 under the first policy, a bounce-back off the wall cell (W-1, H-1) under
 * the second - and the difference spreads over the whole field.  The mode exists so that the
 * difference can be measured and a future parity target that clamps has an oracle; the CUDA path implements the
 * first policy only.
 *
 * Reference files followed (paths relative to /root/reference/lbm-wgpu/src):
 *   lbm.rs:595-609     init_barrier, index_pre_init
 *   lbm.rs:611-643     set_equil
 *   lbm.rs:726-791     LBM::new — buffers, which array is bound where (rest population from buffer 0 only,
 *                      :775-778)
 *   lbm.rs:1051-1134   calculate_summary, iterate, reset_to_equilibrium, custom_speed, compute_step,
 *                      collide, stream
 *   lbm.rs:1174-1252   bind-group / ping-pong selection per pass (compute_step % 2)
 *   lbm.rs:1337-1365   draw_barrier_updates, update_omega_buffer, reset_barrier
 *   lbm.rs:1482-1515   set_single_cell, single_cell
 *   rewritten_shaders/pre_collision/{corner,cardinal}_pre_collision.wgsl
 *   rewritten_shaders/collision/{corner,cardinal}_collision.wgsl
 *   rewritten_shaders/stream/{e_w,n_s,ne_sw,se_nw}_stream.wgsl
 *   rewritten_shaders/summary_stats/{curl,ux,uy,rho,speed}.wgsl
 *   rewritten_shaders/update_barrier/barrier_draw.wgsl
 *   barrier_shapes/merge_shapes.rs:12-22  (the [location, value] u32 pair format)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Contraction twin (-DLBM_CONTRACT -> liblbm_oracle_contract.so).  WGSL lets a backend fuse a multiply into the add
 * or subtract that consumes it; naga emits no NoContraction.  The default build defines "no contraction" (what
 * lavapipe's LLVM pipeline does: its NIR options lower ffma and gallivm sets no fast-math flags).  Should a backend
 * that fuses ever be the parity target, this twin states the rule an LLVM-style backend applies per shader after
 * CSE — a multiply with exactly ONE use is fused into the add/sub using it (left operand first when both are
 * multiplies) — and libblbm_contract.so (csrc -DBLBM_CONTRACT) follows the same rule, so the pair is ready:
 *   corner_collision.wgsl   u2 = fma(ux,ux,uy*uy);  f += w*(k*(p-u215)-f)  ->  fma(w, fma(k, p-u215, -f), f)
 *   cardinal_collision.wgsl the relaxations as above; u2 stays unfused there (ux2, uy2 have two uses each)
 *   rho.wgsl                fma(4, clamp(..), -0.5);   speed.wgsl  sqrt(fma(mx,mx,my*my))
 *   color_map/{jet,viridis,inferno}.wgsl   fma(lw, A, rw*B)
 * ux3, uy3, uxuy2, u215, 4.5*(..), 4.5*ux2 have several uses after CSE and stay separate roundings. */
#ifdef LBM_CONTRACT
#define MULADD(a, b, c) fmaf((a), (b), (c))
#define MULSUB(a, b, c) fmaf((a), (b), -(c))
#else
#define MULADD(a, b, c) ((a) * (b) + (c))
#define MULSUB(a, b, c) ((a) * (b) - (c))
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

/* population order of data_buffers[b][k], lbm.rs:632-640 */
enum { NW = 0, N_ = 1, NE = 2, W_ = 3, REST = 4, E_ = 5, SW = 6, S_ = 7, SE = 8 };
enum { STAT_CURL = 0, STAT_UX = 1, STAT_UY = 2, STAT_RHO = 3, STAT_SPEED = 4 }; /* lbm.rs:10-16 */

typedef struct lbm_oracle {
    uint32_t w, h;
    int64_t n;            /* w*h */
    float *f[2][9];       /* data_buffers; f[1][REST] is allocated and initialised but never bound (dead) */
    uint32_t *bar;        /* barrier_buffer */
    float *mx, *my, *rho; /* density_bg: momentum x, momentum y, density */
    float *out;           /* output_bg */
    float omega;
    uint64_t step;        /* compute_step */
    int stat;             /* summary_stat */
    int oob_clamp;        /* 0 (default): out-of-range reads return 0; 1: they return the last element (see below) */
} lbm_oracle;

/* ---- lbm.rs:611-643 set_equil: nine uniform equilibrium values, fp32, op order as written ---- */
void lbm_oracle_set_equil(float ux, float uy, float rho, float out9[9])
{
    float ux_2 = ux * ux;
    float uy_2 = uy * uy;
    float u_dot = ux_2 + uy_2;
    float uxuy = ux * uy;
    float pos = u_dot + 2.0f * uxuy;
    float neg = u_dot - 2.0f * uxuy;
    ux *= 3.0f;
    uy *= 3.0f;
    ux_2 *= 4.5f;
    uy_2 *= 4.5f;
    u_dot *= 1.5f;
    neg *= 4.5f;
    pos *= 4.5f;
    float r9 = rho / 9.0f;
    float r36 = rho / 36.0f;
    out9[NW] = r36 * ((((1.0f - ux) + uy) + neg) - u_dot);
    out9[N_] = r9 * (((1.0f + uy) + uy_2) - u_dot);
    out9[NE] = r36 * ((((1.0f + ux) + uy) + pos) - u_dot);
    out9[W_] = r9 * (((1.0f - ux) + ux_2) - u_dot);
    out9[REST] = (4.0f * r9) * (1.0f - u_dot);
    out9[E_] = r9 * (((1.0f + ux) + ux_2) - u_dot);
    out9[SW] = r36 * ((((1.0f - ux) - uy) + pos) - u_dot);
    out9[S_] = r9 * (((1.0f - uy) - uy_2) - u_dot); /* sic: "- uy_2", lbm.rs:639 */
    out9[SE] = r36 * ((((1.0f + ux) - uy) + neg) - u_dot);
}

static void fill_all(lbm_oracle *o, const float v[9])
{
    for (int b = 0; b < 2; b++)
        for (int k = 0; k < 9; k++) {
            float *p = o->f[b][k];
            const float x = v[k];
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < o->n; i++) p[i] = x;
        }
}

/* lbm.rs:595-605 */
static void init_barrier(lbm_oracle *o)
{
    memset(o->bar, 0, sizeof(uint32_t) * (size_t)o->n);
    for (uint32_t x = 0; x < o->w; x++) {
        o->bar[(int64_t)x] = 1u;
        o->bar[(int64_t)x + (int64_t)(o->h - 1) * o->w] = 1u;
    }
}

/* lbm.rs:726-791: LBM::new always starts from set_equil(0.1, 0, 1); inflow is a parameter here so
 * that tests can start from custom_speed()'s state directly. */
lbm_oracle *lbm_oracle_create(uint32_t w, uint32_t h, float omega, float inflow_ux)
{
    if (w == 0 || h == 0) return NULL;
    lbm_oracle *o = (lbm_oracle *)calloc(1, sizeof(*o));
    if (!o) return NULL;
    o->w = w;
    o->h = h;
    o->n = (int64_t)w * (int64_t)h;
    size_t nb = sizeof(float) * (size_t)o->n;
    for (int b = 0; b < 2; b++)
        for (int k = 0; k < 9; k++) o->f[b][k] = (float *)malloc(nb);
    o->bar = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)o->n);
    o->mx = (float *)calloc((size_t)o->n, sizeof(float)); /* zero_vec, lbm.rs:774 */
    o->my = (float *)calloc((size_t)o->n, sizeof(float));
    o->rho = (float *)calloc((size_t)o->n, sizeof(float));
    o->out = (float *)calloc((size_t)o->n, sizeof(float));
    float v[9];
    lbm_oracle_set_equil(inflow_ux, 0.0f, 1.0f, v);
    fill_all(o, v);
    init_barrier(o);
    o->omega = omega;
    o->step = 0;
    o->stat = STAT_CURL;
    return o;
}

void lbm_oracle_destroy(lbm_oracle *o)
{
    if (!o) return;
    for (int b = 0; b < 2; b++)
        for (int k = 0; k < 9; k++) free(o->f[b][k]);
    free(o->bar);
    free(o->mx);
    free(o->my);
    free(o->rho);
    free(o->out);
    free(o);
}

/* ---- the four collide passes, lbm.rs:1118-1125 ---- */

/* corner_pre_collision.wgsl:19-21 */
static void pre_collide_corner(lbm_oracle *o, int c)
{
    const float *ne = o->f[c][NE], *se = o->f[c][SE], *nw = o->f[c][NW], *sw = o->f[c][SW];
    float *mx = o->mx, *my = o->my, *rho = o->rho;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; i++) {
        mx[i] = ((ne[i] + se[i]) - nw[i]) - sw[i];
        my[i] = ((ne[i] + nw[i]) - se[i]) - sw[i];
        rho[i] = ((ne[i] + se[i]) + nw[i]) + sw[i];
    }
}

/* cardinal_pre_collision.wgsl:19-21 */
static void pre_collide_cardinal(lbm_oracle *o, int c)
{
    const float *n = o->f[c][N_], *s = o->f[c][S_], *e = o->f[c][E_], *w = o->f[c][W_];
    float *mx = o->mx, *my = o->my, *rho = o->rho;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; i++) {
        mx[i] = mx[i] + (e[i] - w[i]);
        my[i] = my[i] + (n[i] - s[i]);
        rho[i] = rho[i] + (((e[i] + n[i]) + s[i]) + w[i]);
    }
}

/* corner_collision.wgsl:24-42; origin is data_buffers[0][4] whatever c is (lbm.rs:775-778) */
static void collide_corner(lbm_oracle *o, int c)
{
    float *ne = o->f[c][NE], *se = o->f[c][SE], *nw = o->f[c][NW], *sw = o->f[c][SW];
    const float *origin = o->f[0][REST];
    const float *mx = o->mx, *my = o->my;
    float *rho = o->rho;
    const float omega = o->omega;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; i++) {
        rho[i] = rho[i] + origin[i];
        float thisrho = rho[i];
        float thisux = mx[i] / thisrho;
        float thisuy = my[i] / thisrho;
        float one36thrho = (1.0f / 36.0f) * thisrho;
        float ux3 = 3.0f * thisux;
        float uy3 = 3.0f * thisuy;
        float ux2 = thisux * thisux;
        float uy2 = thisuy * thisuy;
        float uxuy2 = (2.0f * thisux) * thisuy;
        float u2 = MULADD(thisux, thisux, uy2); /* default build: ux2 + uy2 */
        (void)ux2;
        float u215 = 1.5f * u2;
        ne[i] = MULADD(omega, MULSUB(one36thrho, (((1.0f + ux3) + uy3) + 4.5f * (u2 + uxuy2)) - u215, ne[i]), ne[i]);
        se[i] = MULADD(omega, MULSUB(one36thrho, (((1.0f + ux3) - uy3) + 4.5f * (u2 - uxuy2)) - u215, se[i]), se[i]);
        nw[i] = MULADD(omega, MULSUB(one36thrho, (((1.0f - ux3) + uy3) + 4.5f * (u2 - uxuy2)) - u215, nw[i]), nw[i]);
        sw[i] = MULADD(omega, MULSUB(one36thrho, (((1.0f - ux3) - uy3) + 4.5f * (u2 + uxuy2)) - u215, sw[i]), sw[i]);
    }
}

/* cardinal_collision.wgsl:25-43 */
static void collide_cardinal(lbm_oracle *o, int c)
{
    float *n = o->f[c][N_], *s = o->f[c][S_], *e = o->f[c][E_], *w = o->f[c][W_];
    float *origin = o->f[0][REST];
    const float *mx = o->mx, *my = o->my, *rho = o->rho;
    const float omega = o->omega;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; i++) {
        float thisrho = rho[i];
        float thisux = mx[i] / thisrho;
        float thisuy = my[i] / thisrho;
        float one9thrho = (1.0f / 9.0f) * thisrho;
        float ux3 = 3.0f * thisux;
        float uy3 = 3.0f * thisuy;
        float ux2 = thisux * thisux;
        float uy2 = thisuy * thisuy;
        float u2 = ux2 + uy2;
        float u215 = 1.5f * u2;
        origin[i] = MULADD(omega, MULSUB((4.0f / 9.0f) * thisrho, 1.0f - u215, origin[i]), origin[i]);
        e[i] = MULADD(omega, MULSUB(one9thrho, ((1.0f + ux3) + 4.5f * ux2) - u215, e[i]), e[i]);
        w[i] = MULADD(omega, MULSUB(one9thrho, ((1.0f - ux3) + 4.5f * ux2) - u215, w[i]), w[i]);
        n[i] = MULADD(omega, MULSUB(one9thrho, ((1.0f + uy3) + 4.5f * uy2) - u215, n[i]), n[i]);
        s[i] = MULADD(omega, MULSUB(one9thrho, ((1.0f - uy3) + 4.5f * uy2) - u215, s[i]), s[i]);
    }
}

void lbm_oracle_set_oob_clamp(lbm_oracle *o, int clamp) { o->oob_clamp = clamp != 0; }

void lbm_oracle_collide(lbm_oracle *o)
{
    int c = (int)(o->step % 2);
    pre_collide_corner(o, c);
    pre_collide_cardinal(o, c);
    collide_corner(o, c);
    collide_cardinal(o, c);
}

/* ---- the four stream passes, lbm.rs:1127-1134; one generic body for the four WGSL files ---- */
static inline uint32_t bar_at(const lbm_oracle *o, int64_t j)
{
    if (j < 0 || j >= o->n) return o->oob_clamp ? o->bar[o->n - 1] : 0u; /* (a negative index is a huge u32) */
    return o->bar[j];
}
static inline float pop_at(const lbm_oracle *o, const float *p, int64_t j)
{
    if (j < 0 || j >= o->n) return o->oob_clamp ? p[o->n - 1] : 0.0f;
    return p[j];
}

/* population `a` travels by +off per step, `b` by -off (e_w_stream.wgsl:25-64 with a=e, b=w, off=+1) */
static void stream_pair(lbm_oracle *o, int a, int b, int64_t off)
{
    int c = (int)(o->step % 2), d = 1 - c;
    const float *src_a = o->f[c][a], *src_b = o->f[c][b];
    float *dst_a = o->f[d][a], *dst_b = o->f[d][b];
    const int64_t row = o->w, col = o->h;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; i++) {
        if (o->bar[i] == 1u) continue;
        if (i % row == 0) continue;
        if (i / row >= col - 1) continue;
        int64_t ip = i + off, im = i - off;
        if (bar_at(o, ip) == 1u)
            dst_b[i] = src_a[i];
        else
            dst_b[i] = pop_at(o, src_b, ip);
        if (bar_at(o, im) == 1u)
            dst_a[i] = src_b[i];
        else
            dst_a[i] = pop_at(o, src_a, im);
    }
}

void lbm_oracle_stream(lbm_oracle *o)
{
    const int64_t row = o->w;
    stream_pair(o, E_, W_, +1);       /* e_w_stream.wgsl */
    stream_pair(o, S_, N_, +row);     /* n_s_stream.wgsl: n_index = i - row */
    stream_pair(o, SE, NW, +1 + row); /* se_nw_stream.wgsl */
    stream_pair(o, NE, SW, +1 - row); /* ne_sw_stream.wgsl */
}

/* ---- summary_stats/{curl,ux,uy,rho,speed}.wgsl ---- */
static inline float clamp01(float v)
{
    /* WGSL clamp(e, low, high) = min(max(e, low), high) */
    float t = v > 0.0f ? v : 0.0f;
    return t < 1.0f ? t : 1.0f;
}

void lbm_oracle_summary(lbm_oracle *o, int stat)
{
    const float *mx = o->mx, *my = o->my, *rho = o->rho;
    float *out = o->out;
    const int64_t row = o->w, col = o->h, n = o->n;
    o->stat = stat;
    switch (stat) {
    case STAT_CURL: /* curl.wgsl:37-49 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) {
            if (i % row == 0) continue;
            if (i / row >= col - 1) continue;
            float a = pop_at(o, my, i + 1);
            float b = pop_at(o, my, i - 1);
            float c = pop_at(o, mx, i - row);
            float d = pop_at(o, mx, i + row);
            out[i] = (10.0f * (((a - b) - c) + d)) / rho[i];
        }
        break;
    case STAT_UX:
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) out[i] = mx[i];
        break;
    case STAT_UY:
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) out[i] = my[i];
        break;
    case STAT_RHO: /* rho.wgsl:24 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) out[i] = MULSUB(4.0f, clamp01(0.15f * rho[i]), 0.5f);
        break;
    case STAT_SPEED: /* speed.wgsl:25 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++)
            out[i] = clamp01(5.0f * sqrtf(MULADD(mx[i], mx[i], my[i] * my[i]))) - 0.5f;
        break;
    default:
        break;
    }
}

/* lbm.rs:1112-1116 */
void lbm_oracle_step(lbm_oracle *o)
{
    lbm_oracle_collide(o);
    lbm_oracle_stream(o);
    o->step += 1;
}

/* lbm.rs:1065-1074 without colour map / render */
void lbm_oracle_iterate(lbm_oracle *o, uint32_t nsteps)
{
    for (uint32_t s = 0; s < nsteps; s++) lbm_oracle_step(o);
    lbm_oracle_summary(o, o->stat);
}

/* barrier_draw.wgsl:11-18 through lbm.rs:1341-1356; pairs = [loc0,val0,loc1,val1,...] */
void lbm_oracle_draw_points(lbm_oracle *o, const uint32_t *pairs, size_t npairs)
{
    for (size_t p = 0; p < npairs; p++) {
        uint32_t loc = pairs[2 * p], val = pairs[2 * p + 1];
        if ((int64_t)loc < o->n) o->bar[loc] = val;
    }
}

void lbm_oracle_reset_barrier(lbm_oracle *o) { init_barrier(o); }       /* lbm.rs:1362-1365 */
void lbm_oracle_set_omega(lbm_oracle *o, float omega) { o->omega = omega; } /* lbm.rs:1358-1360 */
void lbm_oracle_set_summary(lbm_oracle *o, int stat) { o->stat = stat; }  /* lbm.rs:1061-1063 */

/* lbm.rs:1090-1102 (and :1076-1088 with ux = 0.1) */
void lbm_oracle_custom_speed(lbm_oracle *o, float ux)
{
    float v[9];
    lbm_oracle_set_equil(ux, 0.0f, 1.0f, v);
    fill_all(o, v);
    o->step = 0;
    pre_collide_corner(o, 0);
    pre_collide_cardinal(o, 0);
}
void lbm_oracle_reset_to_equilibrium(lbm_oracle *o) { lbm_oracle_custom_speed(o, 0.1f); }

/* lbm.rs:1482-1515 */
void lbm_oracle_single_cell(lbm_oracle *o, uint32_t index)
{
    float v[9];
    lbm_oracle_set_equil(0.0f, 0.0f, 1.0f, v);
    fill_all(o, v);
    const int64_t x = o->w, y = o->h;
    int64_t cx = -1, cy = -1;
    switch (index) {
    case 0: cx = x - 2; cy = y - 2; break;
    case 1: cx = 3 * x / 4; cy = y - 2; break;
    case 2: cx = x / 3; cy = y - 2; break;
    case 3: cx = x - 2; cy = y / 2; break;
    case 4: cx = 3 * x / 4; cy = y / 2; break;
    case 5: cx = x / 2; cy = y / 2; break;
    case 6: cx = x - 2; cy = 1; break;
    case 7: cx = 3 * x / 4; cy = 1; break;
    case 8: cx = x / 2; cy = 1; break;
    default: break;
    }
    if (cx >= 0) {
        int64_t i = cx + cy * x;
        if (i >= 0 && i < o->n) {
            o->f[0][index][i] = 4.0f;
            o->f[1][index][i] = 4.0f;
        }
    }
    o->step = 0;
}

/* Per-cell classification word used by the "bit-exact classification" parity check:
 * bit0 barrier, bit1 skipped by stream, bits 2..9 "upstream neighbour is a barrier" for the moving
 * populations in order nw n ne w e sw s se (upstream of a population travelling by +c is cell i-c). */
void lbm_oracle_cell_class(const lbm_oracle *o, uint16_t *dst)
{
    const int64_t row = o->w, col = o->h;
    const int64_t trav[8] = { -1 - row, -row, 1 - row, -1, +1, -1 + row, +row, +1 + row };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; i++) {
        uint16_t c = 0;
        if (o->bar[i] == 1u) c |= 1u;
        if (o->bar[i] == 1u || i % row == 0 || i / row >= col - 1) c |= 2u;
        for (int d = 0; d < 8; d++)
            if (bar_at(o, i - trav[d]) == 1u) c |= (uint16_t)(4u << d);
        dst[i] = c;
    }
}

/* ---- accessors for ctypes ---- */
float *lbm_oracle_population(lbm_oracle *o, int buffer, int k) { return o->f[buffer][k]; }
uint32_t *lbm_oracle_barrier(lbm_oracle *o) { return o->bar; }
float *lbm_oracle_mx(lbm_oracle *o) { return o->mx; }
float *lbm_oracle_my(lbm_oracle *o) { return o->my; }
float *lbm_oracle_rho(lbm_oracle *o) { return o->rho; }
float *lbm_oracle_output(lbm_oracle *o) { return o->out; }
uint64_t lbm_oracle_compute_num(const lbm_oracle *o) { return o->step; }
float lbm_oracle_omega(const lbm_oracle *o) { return o->omega; }
/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU arm of bench.py runs on rank 0 alone and
   asks for all host cores explicitly */
void lbm_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int lbm_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- colour maps (row N3 of SURVEY.md section 8f): rewritten_shaders/color_map/{jet,viridis,inferno}.wgsl
 * through lbm.rs:1299-1335.  rgb is n x 3 floats, densely packed (the reference's array<vec3<f32>> has a
 * 16-byte stride in a buffer sized for 12, lbm.rs:193 vs jet.wgsl:1 — a latent bug that is not restated).
 * ColorMap order, lbm.rs:18-24: Inferno = 0, Viridis = 1, Jet = 2. ---- */
static const float JET_NODES[9][3] = {{0.0f, 0.0f, 0.5f}, {0.0f, 0.0f, 1.0f}, {0.0f, 0.5f, 1.0f},
                                      {0.0f, 1.0f, 1.0f}, {0.5f, 1.0f, 0.5f}, {1.0f, 1.0f, 0.0f},
                                      {1.0f, 0.5f, 0.0f}, {1.0f, 0.0f, 0.0f}, {0.5f, 0.0f, 0.0f}};
static const float VIRIDIS_NODES[5][3] = {{0.9921875f, 0.90625f, 0.1484375f},
                                          {0.3671875f, 0.7890625f, 0.3828125f},
                                          {0.1328125f, 0.56640625f, 0.55078125f},
                                          {0.23046875f, 0.32421875f, 0.546875f},
                                          {0.265625f, 0.0078125f, 0.33203125f}};
static const float INFERNO_NODES[5][3] = {{0.98828125f, 1.0f, 0.64453125f},
                                          {0.97265625f, 0.55859375f, 0.0390625f},
                                          {0.73828125f, 0.21875f, 0.33203125f},
                                          {0.34375f, 0.06640625f, 0.43359375f},
                                          {0.0f, 0.0f, 0.01853125f}};

void lbm_oracle_color_map(const lbm_oracle *o, int map, float *rgb)
{
    const float(*nodes)[3];
    int nseg;
    float scale;
    if (map == 2) {
        nodes = JET_NODES; nseg = 8; scale = 20.0f;      /* jet.wgsl:16: clamp(20 v, -4, 4) */
    } else if (map == 1) {
        nodes = VIRIDIS_NODES; nseg = 4; scale = 15.0f;  /* viridis.wgsl:16: clamp(15 v, -2, 2) */
    } else {
        nodes = INFERNO_NODES; nseg = 4; scale = 15.0f;
    }
    const float lo = (float)(-nseg / 2), hi = (float)(nseg / 2);
    const int ilo = -nseg / 2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; i++) {
        float c = scale * o->out[i];
        c = c > lo ? c : lo; /* clamp = min(max(e, low), high) */
        c = c < hi ? c : hi;
        const int block = (int)floorf(c);
        float r, g, b;
        if (block >= ilo && block < ilo + nseg) {
            const float rw = (float)(-block) + c; /* "4.0 + color", "3.0 + color", ... "-3.0 + color" */
            const float lw = 1.0f - rw;
            const float *A = nodes[block - ilo], *B = nodes[block - ilo + 1];
            r = MULADD(lw, A[0], rw * B[0]);
            g = MULADD(lw, A[1], rw * B[1]);
            b = MULADD(lw, A[2], rw * B[2]);
        } else { /* default: the last node */
            r = nodes[nseg][0]; g = nodes[nseg][1]; b = nodes[nseg][2];
        }
        if (o->bar[i] == 1u) r = g = b = 0.0f;
        rgb[3 * i + 0] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
    }
}

/* 1 in the contraction twin (liblbm_oracle_contract.so), 0 in the default build */
int lbm_oracle_contract(void)
{
#ifdef LBM_CONTRACT
    return 1;
#else
    return 0;
#endif
}
