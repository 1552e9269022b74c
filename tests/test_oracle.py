"""CPU tests of the oracle itself (run everywhere): the C restatement against the independent numpy
restatement, against the committed golden fixtures, and against analytic known answers.

The reference ships no tests or vectors, so these are what pins the oracle —
PARITY UNPINNED against the real WGSL pipeline, which cannot run in this environment.
"""
import os

import numpy as np
import pytest

from oracle import lbm_numpy
from oracle.lbm_oracle import Oracle, set_equil
from tests.golden.make_golden import (case_config1_digests, case_small_cylinder, digest, omega_from_viscosity,
                                      state_arrays)
from tests.util import assert_same_bits, disc_pairs, random_script

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class NumpyAdapter:
    """Gives oracle/lbm_numpy.NumpyLBM the Oracle's method names."""

    def __init__(self, omega, w, h, inflow_ux=0.1):
        self.m = lbm_numpy.NumpyLBM(omega, w, h, inflow_ux)
        self.w, self.h = w, h

    def iterate(self, n):
        self.m.iterate(n)

    def draw_points(self, p):
        self.m.draw_points(p)

    def update_omega_buffer(self, om):
        self.m.update_omega_buffer(om)

    def compute_summary(self, s):
        self.m.summary(s)

    def collide(self):
        self.m.collide()

    def stream(self):
        self.m.stream()

    def custom_speed(self, u):
        self.m.custom_speed(u)

    def reset_barrier(self):
        self.m.reset_barrier()

    def single_cell(self, i):
        self.m.single_cell(i)

    def reset_to_equilibrium(self):
        self.m.reset_to_equilibrium()


def compare_oracles(o, a, tag):
    sh = (o.h, o.w)
    for b in (0, 1):
        for k in range(9):
            assert_same_bits(o.population(b, k), a.m.population(b, k), f"{tag} f[{b}][{k}]")
    mx, my, rho = o.moments()
    assert_same_bits(mx, a.m.mx.reshape(sh), f"{tag} mx")
    assert_same_bits(my, a.m.my.reshape(sh), f"{tag} my")
    assert_same_bits(rho, a.m.rho.reshape(sh), f"{tag} rho")
    assert_same_bits(o.output(), a.m.out.reshape(sh), f"{tag} out")
    assert_same_bits(o.barrier(), a.m.bar.reshape(sh), f"{tag} bar")
    assert o.get_compute_num() == a.m.step_no


@pytest.mark.parametrize("size", [(64, 32), (37, 19), (8, 5), (2, 3), (1, 1), (5, 1)])
def test_c_oracle_equals_numpy_restatement(size):
    w, h = size
    rng = np.random.default_rng(1000 * w + h)
    script = random_script(rng, w, h, phases=5, max_steps=25)
    om = omega_from_viscosity(0.02)
    o, a = Oracle(om, w, h), NumpyAdapter(om, w, h)
    checks = 0
    for op in script:
        if op[0] == "compare":
            checks += 1
            compare_oracles(o, a, f"{w}x{h} #{checks}")
            continue
        name = {"draw": "draw_points", "omega": "update_omega_buffer", "summary": "compute_summary"}.get(op[0], op[0])
        getattr(o, name)(*op[1:])
        getattr(a, name)(*op[1:])
    assert checks >= 8


def test_set_equil_matches_numpy_and_known_values():
    for u in (0.0, 0.05, 0.1, 0.3):
        assert_same_bits(set_equil(u, 0.0, 1.0), lbm_numpy.set_equil(u, 0.0, 1.0), f"set_equil({u})")
    v = set_equil(0.0, 0.0, 1.0)
    # rest fluid: 1/36, 1/9, 1/36, 1/9, 4/9, ... exactly as fp32 divisions
    assert v[4] == np.float32(4.0) * (np.float32(1.0) / np.float32(9.0))
    assert v[1] == np.float32(1.0) / np.float32(9.0)
    assert v[0] == np.float32(1.0) / np.float32(36.0)
    # weights sum to ~1 at rest
    assert abs(float(v.astype(np.float64).sum()) - 1.0) < 1e-6


def test_golden_small_cylinder_full_arrays():
    want = np.load(os.path.join(GOLDEN, "small_cylinder_64x32.npz"))
    got = case_small_cylinder()
    assert sorted(want.files) == sorted(got)
    for name in want.files:
        assert_same_bits(got[name], want[name], f"golden small_cylinder {name}")


def test_golden_config1_digests():
    want = np.load(os.path.join(GOLDEN, "config1_512x256_digests.npz"))
    got = case_config1_digests()
    assert sorted(want.files) == sorted(got)
    for name in want.files:
        assert str(got[name]) == str(want[name]), f"golden config1 {name}"


def test_uniform_equilibrium_is_a_bit_exact_fixed_point():
    """u0 = 0.1, omega = 1.25 (the reference's defaults, lib.rs:39, lbm.rs:740): the inlet column and, with no
    obstacle, the first interior columns never change."""
    o = Oracle(1.25, 96, 40)
    first = [o.population(-1, k).copy() for k in range(9)]
    o.iterate(300)
    for k in range(9):
        assert_same_bits(o.population(-1, k)[:, 0], first[k][:, 0], f"inlet pop {k}")


def test_single_cell_packet_translation_and_wall_reflection():
    """single_cell(1): 4.0 in `n` at (3W/4, H-2) (lbm.rs:1488); omega = 0 makes collision the identity."""
    w, h = 64, 24
    o = Oracle(0.0, w, h)
    o.single_cell(1)
    x0 = 3 * w // 4
    base = o.population(-1, 1)[5, 5]
    for k in range(1, h - 2):
        o.iterate(1)
        n = o.population(-1, 1)
        assert n[h - 2 - k, x0] == 4.0
        assert np.count_nonzero(n != base) == 1
    o.iterate(1)
    assert o.population(-1, 7)[1, x0] == 4.0
    assert np.count_nonzero(o.population(-1, 1) != base) == 0


def test_single_cell_north_east_packet_deflects_south_west():
    """tutorial.ts:186-191,706: the ne packet 'deflects to the southwest when it hits the top boundary'."""
    w, h = 90, 30
    o = Oracle(0.0, w, h)
    o.single_cell(2)
    x0, y0 = w // 3, h - 2
    for k in range(1, y0):
        o.iterate(1)
        assert o.population(-1, 2)[y0 - k, x0 + k] == 4.0
    o.iterate(1)  # at y = 1 under the wall: comes back as sw in the same cell
    assert o.population(-1, 6)[1, x0 + y0 - 1] == 4.0


def test_mask_scatter_last_writer_wins_and_drops_out_of_range():
    o = Oracle(1.0, 16, 8)
    o.draw_points(np.array([[20, 1], [20, 0], [21, 0], [21, 1], [16 * 8, 1], [2 ** 32 - 1, 1]], np.uint32))
    b = o.barrier()
    assert b[1, 4] == 0 and b[1, 5] == 1
    assert b.sum() == 2 * 16 + 1


def test_cell_classes():
    w, h = 12, 6
    o = Oracle(1.0, w, h)
    o.draw_points(np.array([[2 * w + 5, 1]], np.uint32))
    c = o.cell_class()
    assert c[2, 5] & 1 and c[2, 5] & 2
    assert (c[:, 0] & 2).all() and (c[h - 1] & 2).all() and (c[0] & 3 == 3).all()
    assert not c[2, 4] & 2
    # the cell west of the barrier pulls its w population from the barrier: upstream bit of w (index 3)
    assert c[2, 4] & (4 << 3)
    # the cell east of it pulls e (index 4) from the barrier
    assert c[2, 6] & (4 << 4)
    # column W-1 looks east through the flat index: (0, y+1) for w, which is fluid here
    assert not c[2, w - 1] & (4 << 3)
    # ... and row H-2's sw/s/se upstream... its n-moving populations come from the bottom wall
    assert c[h - 2, 3] & (4 << 1)


def test_state_arrays_and_digest_are_nan_canonical():
    a = np.array([np.nan, 1.0], np.float32)
    b = a.copy()
    b.view(np.uint32)[0] = 0xFFC00001
    assert digest(a) == digest(b)
    o = Oracle(1.0, 8, 4)
    assert len(state_arrays(o)) == 24


def _numpy_color_map(out, bar, cmap):
    """Independent vectorised restatement of color_map/{inferno,viridis,jet}.wgsl."""
    F = np.float32
    jet = [(0, 0, .5), (0, 0, 1), (0, .5, 1), (0, 1, 1), (.5, 1, .5), (1, 1, 0), (1, .5, 0), (1, 0, 0), (.5, 0, 0)]
    viridis = [(0.9921875, 0.90625, 0.1484375), (0.3671875, 0.7890625, 0.3828125),
               (0.1328125, 0.56640625, 0.55078125), (0.23046875, 0.32421875, 0.546875),
               (0.265625, 0.0078125, 0.33203125)]
    inferno = [(0.98828125, 1.0, 0.64453125), (0.97265625, 0.55859375, 0.0390625),
               (0.73828125, 0.21875, 0.33203125), (0.34375, 0.06640625, 0.43359375), (0.0, 0.0, 0.01853125)]
    nodes = np.array({0: inferno, 1: viridis, 2: jet}[cmap], F)
    nseg = len(nodes) - 1
    scale, lo, hi = (F(20), F(-4), F(4)) if cmap == 2 else (F(15), F(-2), F(2))
    c = np.minimum(np.maximum(scale * out, lo), hi)
    block = np.floor(c).astype(np.int64)
    idx = np.clip(block + nseg // 2, 0, nseg - 1)
    rw = (-block).astype(F) + c
    lw = F(1) - rw
    rgb = lw[..., None] * nodes[idx] + rw[..., None] * nodes[idx + 1]
    rgb = np.where((block >= nseg // 2)[..., None], nodes[nseg], rgb)
    rgb[bar == 1] = 0
    return rgb.astype(F)


@pytest.mark.parametrize("cmap", [0, 1, 2])
def test_color_maps_match_numpy_and_known_colours(cmap):
    w, h = 64, 32
    o = Oracle(omega_from_viscosity(0.02), w, h)
    o.draw_points(disc_pairs(w, 16, 16, 4).astype(np.uint32))
    o.iterate(150)
    for stat in range(5):
        o.compute_summary(stat)
        got = o.color_map(cmap)
        want = _numpy_color_map(o.output(), o.barrier(), cmap)
        assert_same_bits(got.reshape(h, -1), want.reshape(h, -1), f"colour map {cmap} stat {stat}")
    assert (got[0] == 0).all() and (got[16, 16] == 0).all()  # walls and the disc are black (jet.wgsl:56-58)
    # value 0 sits on the middle node; saturated values on the end nodes
    o.output()[5, 5], o.output()[5, 6], o.output()[5, 7] = 0.0, 10.0, -10.0
    c = o.color_map(cmap)
    mid = {0: (0.73828125, 0.21875, 0.33203125), 1: (0.1328125, 0.56640625, 0.55078125), 2: (0.5, 1.0, 0.5)}[cmap]
    top = {0: (0.0, 0.0, 0.01853125), 1: (0.265625, 0.0078125, 0.33203125), 2: (0.5, 0.0, 0.0)}[cmap]
    bot = {0: (0.98828125, 1.0, 0.64453125), 1: (0.9921875, 0.90625, 0.1484375), 2: (0.0, 0.0, 0.5)}[cmap]
    assert tuple(c[5, 5]) == tuple(np.float32(mid))
    assert tuple(c[5, 6]) == tuple(np.float32(top))
    assert tuple(c[5, 7]) == tuple(np.float32(bot))


def test_cpu_arm_can_claim_all_cores_under_torchrun(monkeypatch):
    """torchrun exports OMP_NUM_THREADS=1 to every rank; bench.py --impl reference (rank 0 alone) must still time the
    oracle on all host cores"""
    import os
    from oracle import lbm_oracle
    before = lbm_oracle.threads()
    lbm_oracle.lib().lbm_oracle_set_threads(1)
    assert lbm_oracle.threads() == 1
    n = lbm_oracle.use_all_cores()
    assert n == len(os.sched_getaffinity(0)) >= 1
    lbm_oracle.lib().lbm_oracle_set_threads(before)
