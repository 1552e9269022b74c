"""Row N2: the host-side barrier rasteriser of the C-ABI library against the independent
restatement in oracle/barrier_shapes.py.  Pure host code: no GPU needed.  UNPINNED against the un-vendored
`line_drawing` 1.0.0 crate the reference uses (Cargo.lock:679-680)."""
import numpy as np
import pytest

from lbm_b200 import rasterize_line
from oracle import barrier_shapes as bs


def as_set(a):
    return {(int(x), int(y)) for x, y in a}


def test_bresenham_octants_include_both_end_points_and_are_connected():
    for end in [(9, 0), (9, 3), (9, 9), (3, 9), (0, 9), (-3, 9), (-9, 9), (-9, 3), (-9, 0), (-9, -3), (-9, -9),
                (-3, -9), (0, -9), (3, -9), (9, -9), (9, -3), (0, 0)]:
        pts = bs.bresenham((2, 5), (2 + end[0], 5 + end[1]))
        assert pts[0] == (2, 5) and pts[-1] == (2 + end[0], 5 + end[1])
        assert len(pts) == max(abs(end[0]), abs(end[1])) + 1
        for a, b in zip(pts, pts[1:]):
            assert max(abs(a[0] - b[0]), abs(a[1] - b[1])) == 1


@pytest.mark.parametrize("erase", [False, True])
def test_library_rasteriser_matches_independent_restatement(erase):
    rng = np.random.default_rng(11)
    xdim, ydim = 97, 61
    cases = [((10, 10), (10, 10)), ((0, 0), (96, 60)), ((96, 0), (0, 60)), ((5, 30), (90, 30)), ((40, 2), (40, 58)),
             ((0, 60), (96, 60)), ((96, 60), (96, 0))]
    cases += [(tuple(rng.integers(0, [xdim, ydim])), tuple(rng.integers(0, [xdim, ydim]))) for _ in range(200)]
    for p1, p2 in cases:
        got = rasterize_line(p1, p2, xdim, ydim, erase)
        want = bs.line_points(tuple(map(int, p1)), tuple(map(int, p2)), xdim, ydim, erase)
        assert got is not None and want is not None
        assert as_set(got) == want, (p1, p2)
        assert len(got) == len(want)  # the library returns distinct cells


def test_invalid_end_points_are_rejected_like_line_new():
    assert rasterize_line((0, 0), (97, 3), 97, 61) is None       # x == xdim
    assert rasterize_line((-1, 0), (5, 3), 97, 61) is None
    assert bs.line_points((0, 0), (97, 3), 97, 61) is None


def test_thick_line_shape():
    # Line::new of a horizontal segment: the segment itself plus two slightly slanted companions that start one
    # cell below / one cell right of the larger-x end and finish one cell left / one cell above the other end
    # (line.rs:94-114), with fill cells at their single diagonal step
    pts = as_set(rasterize_line((10, 20), (30, 20), 64, 64))
    assert {(x, 20) for x in range(10, 31)} <= pts
    assert {(30, 21), (9, 20), (31, 20), (10, 19)} <= pts
    assert all(19 <= y <= 21 and 9 <= x <= 31 for x, y in pts)
    # a single point still gets its two companions
    assert as_set(rasterize_line((5, 5), (5, 5), 64, 64)) >= {(5, 5)}


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(300, 120), (640, 360)])
def test_presets_paint_the_restated_masks(size):
    from lbm_b200 import LBM
    x, y = size
    for name in ("curl_barrier", "chaos_barrier", "welcome_barrier"):
        lbm = LBM(1.25, x, y)
        getattr(lbm, name)()
        want = np.zeros((y, x), np.uint32)
        want[0] = want[y - 1] = 1
        for px, py in getattr(bs, name)(x, y):
            want[py, px] = 1
        assert (lbm.read_barrier() == want).all(), name
        lbm.iterate(20)  # and the painted lattice steps
        lbm.close()


@pytest.mark.gpu
def test_draw_and_erase_line_through_the_abi():
    from lbm_b200 import LBM
    from oracle.lbm_oracle import Oracle
    from tests.util import compare_state
    x, y = 200, 90
    lbm, ora = LBM(1.25, x, y), Oracle(1.25, x, y)
    for p1, p2, erase in (((20, 20), (150, 70), False), ((60, 10), (60, 80), False), ((40, 40), (120, 50), True)):
        (lbm.erase_line if erase else lbm.draw_line)(p1, p2)
        pts = sorted(bs.line_points(p1, p2, x, y, erase))
        ora.draw_points(np.array([[px + py * x, 0 if erase else 1] for px, py in pts], np.uint32))
        lbm.iterate(25)
        ora.iterate(25)
    compare_state(lbm, ora, "lines")
    lbm.close()
