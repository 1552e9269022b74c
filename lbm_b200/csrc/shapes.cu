// shapes.cu — host-side barrier rasteriser: the step before the hot path.
// Pure host code (no kernels): restates lbm-wgpu/src/barrier_shapes/line.rs (thick line = three Bresenham
// segments plus a fill cell on every diagonal step; the 30-wide eraser) and the presets of
// lbm.rs:1367-1480, producing the [location, value] pairs blbm_draw_points consumes (merge_shapes.rs:12-22).
//
// The reference's Bresenham comes from the un-vendored crate `line_drawing` 1.0.0 (Cargo.lock:679-680): the
// octant-transform formulation — map the segment into the first octant, walk x with error = dy - dx,
// step y when error >= 0 — restated here from its published algorithm.  In the first octant that walk emits
// y_k = floor(k*dy/dx), which is the closed form the tests check this walk against.
// Pinned: Line::new / Line::new_erased as compiled into the reference's shipped wasm binary (the crate's Bresenham
// included) were executed on 99 end-point pairs; blbm_rasterize_line reproduces every cell (tests/test_wasm_pin.py).
// The presets' end-point arithmetic (lbm.rs:1367-1480) is inlined into the binary's event loop and stays a restatement.
#include <algorithm>
#include <cstdint>
#include <set>
#include <utility>
#include <vector>

#include "../../include/blbm.h"

namespace {

typedef std::pair<int64_t, int64_t> Pt;

struct Octant {
    int value;
    Octant(Pt start, Pt end)
    {
        value = 0;
        int64_t dx = end.first - start.first, dy = end.second - start.second;
        if (dy < 0) {
            dx = -dx;
            dy = -dy;
            value += 4;
        }
        if (dx < 0) {
            const int64_t tmp = dx;
            dx = dy;
            dy = -tmp;
            value += 2;
        }
        if (dx < dy) value += 1;
    }
    Pt to(Pt p) const
    {
        switch (value) {
        case 0: return Pt(p.first, p.second);
        case 1: return Pt(p.second, p.first);
        case 2: return Pt(p.second, -p.first);
        case 3: return Pt(-p.first, p.second);
        case 4: return Pt(-p.first, -p.second);
        case 5: return Pt(-p.second, -p.first);
        case 6: return Pt(-p.second, p.first);
        default: return Pt(p.first, -p.second);
        }
    }
    Pt from(Pt p) const
    {
        switch (value) {
        case 0: return Pt(p.first, p.second);
        case 1: return Pt(p.second, p.first);
        case 2: return Pt(-p.second, p.first);
        case 3: return Pt(-p.first, p.second);
        case 4: return Pt(-p.first, -p.second);
        case 5: return Pt(-p.second, -p.first);
        case 6: return Pt(p.second, -p.first);
        default: return Pt(p.first, -p.second);
        }
    }
};

// line_drawing::Bresenham::new(start, end): every point from start to end inclusive
void bresenham(Pt start, Pt end, std::vector<Pt> *out)
{
    const Octant oct(start, end);
    Pt p = oct.to(start);
    const Pt e = oct.to(end);
    const int64_t dx = e.first - p.first, dy = e.second - p.second;
    int64_t error = dy - dx;
    while (p.first <= e.first) {
        out->push_back(oct.from(p));
        if (error >= 0) {
            p.second += 1;
            error -= dx;
        }
        p.first += 1;
        error += dy;
    }
}

bool validate(Pt a, Pt b, int64_t xdim, int64_t ydim)  // line.rs:151-153
{
    return a.first >= 0 && b.first >= 0 && a.second >= 0 && b.second >= 0 && a.first < xdim && b.first < xdim &&
           a.second < ydim && b.second < ydim;
}

std::pair<Pt, Pt> order(Pt p1, Pt p2)  // line.rs:142-147: the endpoint with the larger x first
{
    if (p1.first > p2.first) return std::make_pair(p1, p2);
    return std::make_pair(p2, p1);
}

// line.rs:94-114 (thickness 1: three segments) and :116-140 (eraser: 1 + 2*(thickness-1) segments)
std::vector<std::pair<Pt, Pt>> endpoints(Pt e1, Pt e2, int thickness)
{
    std::vector<std::pair<Pt, Pt>> out;
    out.push_back(order(e1, e2));
    const Pt a = out[0].first, b = out[0].second;
    const int64_t s = a.second > b.second ? -1 : +1;
    for (int64_t i = 1; i <= thickness; i++) {
        // left: (a.x, a.y -/+ i) and (a.x + i, a.y); right: (b.x - i, b.y) and (b.x, b.y - i)
        out.push_back(std::make_pair(Pt(a.first, a.second + s * i), Pt(b.first - i, b.second)));
        out.push_back(std::make_pair(Pt(a.first + i, a.second), Pt(b.first, b.second - i)));
    }
    return out;
}

// Line::new (erase = false, line.rs:22-54) / Line::new_erased (erase = true, :56-87); false = invalid endpoints
bool line_points(Pt e1, Pt e2, int64_t xdim, int64_t ydim, bool erase, std::set<Pt> *pts)
{
    if (!validate(e1, e2, xdim, ydim)) return false;
    // generate_endpoints == generate_endpoints_variable(.., 2); the eraser passes thickness 30 -> i in 1..30
    const std::vector<std::pair<Pt, Pt>> segs = endpoints(e1, e2, erase ? 29 : 1);
    std::vector<Pt> walk;
    for (size_t q = 0; q < segs.size(); q++) {
        if (!validate(segs[q].first, segs[q].second, xdim, ydim)) continue;
        walk.clear();
        bresenham(segs[q].first, segs[q].second, &walk);
        Pt prev = segs[q].first;
        for (size_t k = 0; k < walk.size(); k++) {
            const Pt i = walk[k];
            pts->insert(i);
            if (prev.first - i.first != 0 && prev.second - i.second != 0) {  // diagonal_step, line.rs:89-92
                pts->insert(Pt(prev.first, i.second));
                pts->insert(Pt(i.first, prev.second));
            }
            prev = i;
        }
    }
    return true;
}

int draw_set(blbm_t *h, const std::set<Pt> &pts, uint64_t w, uint64_t val)
{
    if (pts.empty()) return BLBM_OK;
    std::vector<uint64_t> pairs;
    pairs.reserve(pts.size() * 2);
    for (std::set<Pt>::const_iterator it = pts.begin(); it != pts.end(); ++it) {
        pairs.push_back((uint64_t)it->first + (uint64_t)it->second * w);  // get_index, merge_shapes.rs:16-18
        pairs.push_back(val);
    }
    return blbm_draw_points64(h, pairs.data(), pairs.size() / 2);
}

int geometry(blbm_t *h, int64_t *x, int64_t *y)
{
    uint32_t w = 0;
    uint64_t hg = 0;
    int rc = blbm_get_geometry(h, &w, &hg, nullptr, nullptr, nullptr);
    *x = (int64_t)w;
    *y = (int64_t)hg;
    return rc;
}

}  // namespace

extern "C" {

int blbm_rasterize_line(int64_t x1, int64_t y1, int64_t x2, int64_t y2, int64_t xdim, int64_t ydim, int erase,
                        int64_t *xy, size_t capacity, size_t *count)
{
    std::set<Pt> pts;
    if (!line_points(Pt(x1, y1), Pt(x2, y2), xdim, ydim, erase != 0, &pts)) return BLBM_EINVAL;
    if (count) *count = pts.size();
    size_t q = 0;
    for (std::set<Pt>::const_iterator it = pts.begin(); it != pts.end() && q < capacity; ++it, ++q) {
        xy[2 * q] = it->first;
        xy[2 * q + 1] = it->second;
    }
    return BLBM_OK;
}

int blbm_draw_line(blbm_t *h, int64_t x1, int64_t y1, int64_t x2, int64_t y2)
{
    int64_t x, y;
    int rc = geometry(h, &x, &y);
    if (rc) return rc;
    std::set<Pt> pts;
    if (!line_points(Pt(x1, y1), Pt(x2, y2), x, y, false, &pts)) return BLBM_EINVAL;
    return draw_set(h, pts, (uint64_t)x, 1);
}

int blbm_erase_line(blbm_t *h, int64_t x1, int64_t y1, int64_t x2, int64_t y2)
{
    int64_t x, y;
    int rc = geometry(h, &x, &y);
    if (rc) return rc;
    std::set<Pt> pts;
    if (!line_points(Pt(x1, y1), Pt(x2, y2), x, y, true, &pts)) return BLBM_EINVAL;
    return draw_set(h, pts, (uint64_t)x, 0);
}

// LBM::curl_barrier, lbm.rs:1367-1370
int blbm_curl_barrier(blbm_t *h)
{
    int64_t x, y;
    int rc = geometry(h, &x, &y);
    if (rc) return rc;
    return blbm_draw_line(h, 4 * x / 10, y / 4, 4 * x / 10, y / 2);
}

// LBM::chaos_barrier, lbm.rs:1372-1386: four separate draw_shape calls
int blbm_chaos_barrier(blbm_t *h)
{
    int64_t x, y;
    int rc = geometry(h, &x, &y);
    if (rc) return rc;
    if ((rc = blbm_draw_line(h, x / 2, 9 * y / 20, x / 2, 0))) return rc;
    if ((rc = blbm_draw_line(h, x / 2, 11 * y / 20, x / 2, y - 1))) return rc;
    if ((rc = blbm_draw_line(h, 3 * x / 5, y / 2, 3 * x / 4, 3 * y / 4))) return rc;
    return blbm_draw_line(h, 3 * x / 5, y / 2, 3 * x / 4, y / 4);
}

// LBM::welcome_barrier, lbm.rs:1388-1480: "Welcome!" out of thick lines, joined into one Blob, drawn once
int blbm_welcome_barrier(blbm_t *h)
{
    int64_t X, Y;
    int rc = geometry(h, &X, &Y);
    if (rc) return rc;
    std::set<Pt> blob;
    bool ok = true;
    auto line = [&](int64_t ax, int64_t ay, int64_t bx, int64_t by) {
        ok = ok && line_points(Pt(ax, ay), Pt(bx, by), X, Y, false, &blob);
    };
    const int64_t height = -1 * (Y / 4), bottom = Y / 2, space = X / 50, lw = X / 13;
    int64_t cx = X / 5;
    // W
    line(cx, bottom + height, cx, bottom);
    line(cx, bottom, cx + lw / 2, bottom + height / 2);
    cx += lw / 2;
    line(cx, bottom + height / 2, cx + lw / 2, bottom);
    cx += lw / 2;
    line(cx, bottom + height, cx, bottom);
    cx += space;
    // e
    auto letter_e = [&]() {
        line(cx, bottom + height / 2, cx, bottom);
        line(cx, bottom, cx + lw, bottom);
        line(cx, bottom + height / 4, cx + lw, bottom + height / 4);
        line(cx + lw, bottom + height / 2, cx + lw, bottom + height / 4);
        line(cx, bottom + height / 2, cx + lw, bottom + height / 2);
        cx += lw + space;
    };
    letter_e();
    // l
    line(cx, bottom, cx, bottom + height);
    cx += space;
    // c
    line(cx, bottom + height / 2, cx, bottom);
    line(cx, bottom, cx + lw, bottom);
    line(cx, bottom + height / 2, cx + lw, bottom + height / 2);
    cx += lw + space;
    // o
    line(cx, bottom + height / 2, cx, bottom);
    line(cx, bottom, cx + lw, bottom);
    line(cx, bottom + height / 2, cx + lw, bottom + height / 2);
    line(cx + lw, bottom + height / 2, cx + lw, bottom);
    cx += lw + space;
    // m
    line(cx, bottom + height / 2, cx, bottom);
    line(cx, bottom + height / 2, cx + lw, bottom + height / 2);
    line(cx + lw, bottom + height / 2, cx + lw, bottom);
    line(cx + lw / 2, bottom, cx + lw / 2, bottom + height / 2);
    cx += lw + space;
    // e
    letter_e();
    // !
    line(cx, height / 10 + bottom, cx, bottom);
    line(cx, height / 5 + bottom, cx, bottom + height);
    if (!ok) return BLBM_EINVAL;  // Line::new(..).unwrap() would panic
    return draw_set(h, blob, (uint64_t)X, 1);
}

}  // extern "C"
