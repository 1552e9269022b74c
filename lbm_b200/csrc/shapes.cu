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
// The presets (lbm.rs:1367-1480) are inlined into the binary's event-loop closure; their `match` arms were executed
// in place (fragment execution from the arm's first instruction): blbm_preset_lines reproduces every Line::new
// argument and the painted masks equal the point sets the binary hands to draw_shape (tests/golden/wasm_presets.npz).
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

// The end points of the thick lines a preset is made of, in the order the reference creates them
// (LBM::curl_barrier lbm.rs:1367-1370, chaos_barrier :1372-1386, welcome_barrier :1388-1480; isize arithmetic:
// `/` truncates toward zero).  Pinned, argument for argument, against the Line::new calls of the reference's shipped
// binary executing these very `match` arms of its event loop (tests/golden/wasm_presets.npz).
int blbm_preset_lines(int preset, int64_t X, int64_t Y, int64_t *xyxy, size_t capacity, size_t *count)
{
    std::vector<int64_t> v;
    try {
        auto line = [&](int64_t ax, int64_t ay, int64_t bx, int64_t by) {
            v.push_back(ax);
            v.push_back(ay);
            v.push_back(bx);
            v.push_back(by);
        };
        if (preset == BLBM_PRESET_CURL) {
            line(4 * X / 10, Y / 4, 4 * X / 10, Y / 2);
        } else if (preset == BLBM_PRESET_CHAOS) {
            line(X / 2, 9 * Y / 20, X / 2, 0);
            line(X / 2, 11 * Y / 20, X / 2, Y - 1);
            line(3 * X / 5, Y / 2, 3 * X / 4, 3 * Y / 4);
            line(3 * X / 5, Y / 2, 3 * X / 4, Y / 4);
        } else if (preset == BLBM_PRESET_WELCOME) {
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
            auto letter_e = [&]() {
                line(cx, bottom + height / 2, cx, bottom);
                line(cx, bottom, cx + lw, bottom);
                line(cx, bottom + height / 4, cx + lw, bottom + height / 4);
                line(cx, bottom + height / 2, cx + lw, bottom + height / 2);
                line(cx + lw, bottom + height / 2, cx + lw, bottom + height / 4);
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
            line(cx + lw / 2, bottom, cx + lw / 2, bottom + height / 2);
            line(cx, bottom + height / 2, cx + lw, bottom + height / 2);
            line(cx + lw, bottom + height / 2, cx + lw, bottom);
            cx += lw + space;
            letter_e();
            // !
            line(cx, height / 10 + bottom, cx, bottom);
            line(cx, height / 5 + bottom, cx, bottom + height);
        } else {
            return BLBM_EINVAL;
        }
    } catch (...) {
        return BLBM_ENOMEM;
    }
    if (count) *count = v.size() / 4;
    for (size_t q = 0; q < v.size() && q < 4 * capacity; q++) xyxy[q] = v[q];
    return BLBM_OK;
}

static int preset_end_points(blbm_t *h, int preset, int64_t *X, int64_t *Y, std::vector<int64_t> *v)
{
    int rc = geometry(h, X, Y);
    if (rc) return rc;
    size_t n = 0;
    if ((rc = blbm_preset_lines(preset, *X, *Y, nullptr, 0, &n))) return rc;
    try {
        v->resize(4 * n);
    } catch (...) {
        return BLBM_ENOMEM;
    }
    return blbm_preset_lines(preset, *X, *Y, v->data(), n, &n);
}

// LBM::curl_barrier (one draw_shape) and LBM::chaos_barrier (four separate draw_shape calls)
static int draw_preset_lines(blbm_t *h, int preset)
{
    int64_t X, Y;
    std::vector<int64_t> v;
    int rc = preset_end_points(h, preset, &X, &Y, &v);
    for (size_t q = 0; q + 3 < v.size() && rc == BLBM_OK; q += 4) rc = blbm_draw_line(h, v[q], v[q + 1], v[q + 2], v[q + 3]);
    return rc;
}

int blbm_curl_barrier(blbm_t *h) { return draw_preset_lines(h, BLBM_PRESET_CURL); }
int blbm_chaos_barrier(blbm_t *h) { return draw_preset_lines(h, BLBM_PRESET_CHAOS); }

// LBM::welcome_barrier: "Welcome!" out of thick lines, joined into one Blob, drawn once
int blbm_welcome_barrier(blbm_t *h)
{
    int64_t X, Y;
    std::vector<int64_t> v;
    int rc = preset_end_points(h, BLBM_PRESET_WELCOME, &X, &Y, &v);
    if (rc) return rc;
    std::set<Pt> blob;
    for (size_t q = 0; q + 3 < v.size(); q += 4)
        if (!line_points(Pt(v[q], v[q + 1]), Pt(v[q + 2], v[q + 3]), X, Y, false, &blob))
            return BLBM_EINVAL;  // Line::new(..).unwrap() would panic
    return draw_set(h, blob, (uint64_t)X, 1);
}

}  // extern "C"
