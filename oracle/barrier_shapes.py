"""Independent restatement of lbm-wgpu's barrier rasteriser (TEST INFRASTRUCTURE ONLY, see lbm_oracle.c).

Follows lbm-wgpu/src/barrier_shapes/line.rs:22-155 (thick line = three segments + a fill cell on every
diagonal step; eraser = 59 segments) and the presets of lbm.rs:1367-1480.  The Bresenham walk of the
un-vendored crate `line_drawing` 1.0.0 (Cargo.lock:679-680) is NOT transcribed here: in the first octant its
"error = dy - dx, step y when error >= 0" loop emits y_k = floor(k*dy/dx), and this file uses that closed form
(exact integer arithmetic) together with the crate's octant transforms — a different code path from the
product's, so a slip in either shows up.  Pinned: reproduces, cell for cell, Line::new / Line::new_erased as
compiled (with the crate) into the reference's shipped wasm binary, on 99 end-point pairs (tests/test_wasm_pin.py,
tests/golden/wasm_golden.npz).
"""


def _octant(start, end):
    v = 0
    dx, dy = end[0] - start[0], end[1] - start[1]
    if dy < 0:
        dx, dy, v = -dx, -dy, v + 4
    if dx < 0:
        dx, dy, v = dy, -dx, v + 2
    if dx < dy:
        v += 1
    return v


_TO = {0: lambda x, y: (x, y), 1: lambda x, y: (y, x), 2: lambda x, y: (y, -x), 3: lambda x, y: (-x, y),
       4: lambda x, y: (-x, -y), 5: lambda x, y: (-y, -x), 6: lambda x, y: (-y, x), 7: lambda x, y: (x, -y)}
_FROM = {0: lambda x, y: (x, y), 1: lambda x, y: (y, x), 2: lambda x, y: (-y, x), 3: lambda x, y: (-x, y),
         4: lambda x, y: (-x, -y), 5: lambda x, y: (-y, -x), 6: lambda x, y: (y, -x), 7: lambda x, y: (x, -y)}


def bresenham(start, end):
    o = _octant(start, end)
    sx, sy = _TO[o](*start)
    ex, ey = _TO[o](*end)
    dx, dy = ex - sx, ey - sy
    assert 0 <= dy <= dx
    return [_FROM[o](sx + k, sy + (k * dy) // dx if dx else sy) for k in range(dx + 1)]


def _valid(a, b, xdim, ydim):
    return min(a[0], a[1], b[0], b[1]) >= 0 and a[0] < xdim and b[0] < xdim and a[1] < ydim and b[1] < ydim


def line_points(p1, p2, xdim, ydim, erase=False):
    """Set of (x, y) of Line::new / Line::new_erased; None when an end point is outside the lattice."""
    if not _valid(p1, p2, xdim, ydim):
        return None
    a, b = (p1, p2) if p1[0] > p2[0] else (p2, p1)
    s = -1 if a[1] > b[1] else 1
    segs = [(a, b)]
    for i in range(1, 30 if erase else 2):
        segs.append(((a[0], a[1] + s * i), (b[0] - i, b[1])))
        segs.append(((a[0] + i, a[1]), (b[0], b[1] - i)))
    pts = set()
    for q0, q1 in segs:
        if not _valid(q0, q1, xdim, ydim):
            continue
        prev = q0
        for p in bresenham(q0, q1):
            pts.add(p)
            if prev[0] != p[0] and prev[1] != p[1]:
                pts.add((prev[0], p[1]))
                pts.add((p[0], prev[1]))
            prev = p
    return pts


def _tdiv(a, b):
    """Rust / C integer division: truncation toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def curl_barrier(x, y):
    return line_points((4 * x // 10, y // 4), (4 * x // 10, y // 2), x, y)


def chaos_barrier(x, y):
    out = set()
    for p1, p2 in (((x // 2, 9 * y // 20), (x // 2, 0)), ((x // 2, 11 * y // 20), (x // 2, y - 1)),
                   ((3 * x // 5, y // 2), (3 * x // 4, 3 * y // 4)), ((3 * x // 5, y // 2), (3 * x // 4, y // 4))):
        out |= line_points(p1, p2, x, y)
    return out


def welcome_barrier(x, y):
    blob = set()

    def line(ax, ay, bx, by):
        blob.update(line_points((ax, ay), (bx, by), x, y))

    height, bottom, space, lw = -(y // 4), y // 2, x // 50, x // 13
    h2, h4, w2 = _tdiv(height, 2), _tdiv(height, 4), lw // 2
    cx = x // 5
    line(cx, bottom + height, cx, bottom); line(cx, bottom, cx + w2, bottom + h2); cx += w2
    line(cx, bottom + h2, cx + w2, bottom); cx += w2
    line(cx, bottom + height, cx, bottom); cx += space

    def letter_e(cx):
        line(cx, bottom + h2, cx, bottom); line(cx, bottom, cx + lw, bottom)
        line(cx, bottom + h4, cx + lw, bottom + h4); line(cx, bottom + h2, cx + lw, bottom + h2)
        line(cx + lw, bottom + h2, cx + lw, bottom + h4)
        return cx + lw + space

    cx = letter_e(cx)
    line(cx, bottom, cx, bottom + height); cx += space
    line(cx, bottom + h2, cx, bottom); line(cx, bottom, cx + lw, bottom); line(cx, bottom + h2, cx + lw, bottom + h2)
    cx += lw + space
    line(cx, bottom + h2, cx, bottom); line(cx, bottom, cx + lw, bottom); line(cx, bottom + h2, cx + lw, bottom + h2)
    line(cx + lw, bottom + h2, cx + lw, bottom); cx += lw + space
    line(cx, bottom + h2, cx, bottom); line(cx + w2, bottom, cx + w2, bottom + h2)
    line(cx, bottom + h2, cx + lw, bottom + h2); line(cx + lw, bottom + h2, cx + lw, bottom); cx += lw + space
    cx = letter_e(cx)
    line(cx, _tdiv(height, 10) + bottom, cx, bottom); line(cx, _tdiv(height, 5) + bottom, cx, bottom + height)
    return blob
