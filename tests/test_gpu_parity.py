"""Parity of the CUDA path (through the C ABI, via lbm_b200.LBM) against the CPU oracle — bit-exact on
populations, moments, output, barrier mask and cell classification.  Run on the B200 box: -m gpu.

The reference has no tests; these follow its behaviour: the orphan per-cell
"diff" shaders (lbm-wgpu/src/testing/test-*.wgsl) are exactly this comparison, and the single_cell
presets driven by tutorial.ts:129-191 are its known-answer demos.
"""
import numpy as np
import pytest

from lbm_b200 import LBM, Kernel, SlabGroup, SummaryStat, omega_from_viscosity
from oracle.lbm_oracle import Oracle
from tests.util import (assert_same_bits, compare_state, disc_pairs, porous_pairs, random_script, run_script)

pytestmark = pytest.mark.gpu

KERNELS = [Kernel.Scalar, Kernel.Vec4]
# barrier cells kept densely in the planes (0) or in the compact chain table, forced on (1)
LAZY = [0, 1]


@pytest.mark.parametrize("lazy", LAZY)
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("size", [(64, 32), (37, 19), (8, 5), (130, 7), (257, 9), (4, 4), (2, 3), (1, 1), (5, 1)])
def test_random_scripts_bit_exact(kernel, size, lazy):
    w, h = size
    rng = np.random.default_rng(1000 * w + h)
    script = random_script(rng, w, h)
    lbm = LBM(omega_from_viscosity(0.02), w, h, kernel=kernel, lazy_barriers=lazy)
    assert lbm.get_kernel() == kernel
    ora = Oracle(omega_from_viscosity(0.02), w, h)
    checks = run_script(script, lbm, ora, f"{kernel.name} {w}x{h}")
    assert checks >= 10
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_create_state_matches_reference_init(kernel):
    """LBM::new: set_equil(0.1,0,1) in both buffers, walls on rows 0 and H-1, zero moments/output."""
    lbm = LBM(1.25, 96, 40, kernel=kernel)
    ora = Oracle(1.25, 96, 40)
    compare_state(lbm, ora, "fresh")
    lbm.close()


@pytest.mark.parametrize("lazy", LAZY)
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("nu", [0.1, 0.02])
def test_config1_cylinder_512x256_10k_steps(kernel, nu, lazy):
    """BASELINE.json configs[0] / SURVEY.md 8d config 1: channel flow past a cylinder, 512x256, u0=0.1,
    10,000 steps; nu=0.1 (Re 32, steady) and nu=0.02 (Re 160, vortex shedding, where only a bit-identical
    kernel stays inside 1e-5 relative).  Compared at 1, 2, 100, 1000 and 10000 steps."""
    w, h = 512, 256
    om = omega_from_viscosity(nu)
    lbm = LBM(om, w, h, kernel=kernel, lazy_barriers=lazy)
    ora = Oracle(om, w, h)
    cyl = disc_pairs(w, 128, 128, 16)
    lbm.draw_points(cyl)
    ora.draw_points(cyl.astype(np.uint32))
    done = 0
    for target in (1, 2, 100, 1000, 10000):
        lbm.iterate(target - done)
        ora.iterate(target - done)
        done = target
        # north_star tolerance, stated for the record; the assertion below is stricter (bit-exact)
        for k in range(9):
            a, b = lbm.read_population(k), ora.population(-1, k)
            assert np.all(np.abs(a - b) <= 1e-6 + 1e-5 * np.abs(b))
        compare_state(lbm, ora, f"cylinder nu={nu} step {target}")
        if lazy and target > 2:
            assert not lbm.lazy_barriers_active()  # the population read-back flushed the chain table
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_uniform_equilibrium_is_a_fixed_point(kernel):
    """u0=0.1, omega=1.25, no obstacle: the inlet column never drifts."""
    lbm = LBM(1.25, 128, 48, kernel=kernel)
    first = [lbm.read_population(k) for k in range(9)]
    lbm.iterate(500)
    for k in range(9):
        now = lbm.read_population(k)
        assert_same_bits(now[:, 0], first[k][:, 0], f"inlet column pop {k}")
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_single_cell_packet_translates_and_reflects(kernel):
    """tutorial.ts:129-134: single_cell(1) puts 4.0 into `n` at (3W/4, H-2).  With omega = 0 collision is
    the identity, so the packet moves one cell north per step and comes back as `s` from the top wall."""
    w, h = 64, 24
    lbm = LBM(0.0, w, h, kernel=kernel)
    lbm.single_cell(1)
    x0 = 3 * w // 4
    base = lbm.read_population(1)[5, 5]
    for k in range(1, h - 2):  # up to y = 1, the row under the top wall
        lbm.iterate(1)
        n = lbm.read_population(1)
        assert n[h - 2 - k, x0] == 4.0, f"step {k}"
        assert np.count_nonzero(n != base) == 1
    lbm.iterate(1)  # packet at y=1 meets the wall row 0: bounces into `s` in the same cell
    s = lbm.read_population(7)
    assert s[1, x0] == 4.0
    assert np.count_nonzero(lbm.read_population(1) != base) == 0
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_stream_conserves_interior_mass(kernel):
    """With omega = 0 a step is a pure permutation of the moving populations away from inlet/outlet."""
    w, h = 96, 40
    lbm = LBM(0.0, w, h, kernel=kernel)
    rng = np.random.default_rng(5)
    pops = {}
    for k in range(9):
        a = lbm.read_population(k)
        a[12:28, 30:60] = rng.random((16, 30), dtype=np.float32)
        pops[k] = a
        lbm.write_population(k, a, buffer=0)
        lbm.write_population(k, a, buffer=1)
    before = sum(np.sort(pops[k][8:32, 20:70].reshape(-1).astype(np.float64)).sum() for k in range(9))
    lbm.iterate(3)
    after = sum(np.sort(lbm.read_population(k)[8:32, 20:70].reshape(-1).astype(np.float64)).sum()
                for k in range(9))
    assert abs(before - after) < 1e-9 * abs(before)
    lbm.close()


@pytest.mark.parametrize("lazy", [0, 1, 2])
@pytest.mark.parametrize("kernel", KERNELS)
def test_porous_channel_small(kernel, lazy):
    """configs[2] in miniature: random porous mask (15 % solid), u0=0.05, omega=1."""
    w, h = 256, 128
    lbm = LBM(1.0, w, h, inflow_ux=0.05, kernel=kernel, lazy_barriers=lazy)
    ora = Oracle(1.0, w, h, inflow_ux=0.05)
    pts = porous_pairs(w, h)
    assert 0.1 < len(pts) / (w * h) < 0.2
    lbm.draw_points(pts)
    ora.draw_points(pts.astype(np.uint32))
    for n in (1, 50, 200):
        lbm.iterate(n)
        ora.iterate(n)
        assert lbm.lazy_barriers_active() == (lazy != 0)
        # moments / output first: they must be right while the chain table is still live
        compare_state(lbm, ora, f"porous +{n} (table live)", populations=False)
        lbm.update_omega_buffer(1.1 if n == 50 else 1.0)
        ora.update_omega_buffer(1.1 if n == 50 else 1.0)
        lbm.iterate(3)
        ora.iterate(3)
        compare_state(lbm, ora, f"porous +{n}+3")
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("size", [(70001, 5), (65536, 4)])
def test_wide_lattice(kernel, size):
    """Rows wider than 2^16 cells (BASELINE.json configs[4] has W = 65536), ragged and not: chunk indexing, the
    flat-index wrap at column W-1, paints in the last chunk."""
    w, h = size
    lbm = LBM(1.1, w, h, kernel=kernel, lazy_barriers=1)
    ora = Oracle(1.1, w, h)
    rng = np.random.default_rng(w)
    loc = np.unique(np.concatenate([rng.integers(w, w * (h - 1), size=3000),
                                    np.array([2 * w - 1, 2 * w - 2, 2 * w - 130, 3 * w - 1, w + 1, 2 * w + 65535])]))
    pts = np.stack([loc, np.ones_like(loc)], 1).astype(np.uint32)
    lbm.draw_points(pts)
    ora.draw_points(pts)
    for n in (1, 9):
        lbm.iterate(n)
        ora.iterate(n)
        compare_state(lbm, ora, f"wide {w}x{h} {kernel.name} +{n}")
    lbm.close()


@pytest.mark.parametrize("flavour", [0, 2])
def test_tall_lattice_uses_the_third_grid_dimension(flavour):
    """More than 4 x 32768 rows: the vec4 kernel's row blocks spill over from gridDim.y into gridDim.z."""
    w, h = 40, 140001
    lbm = LBM(1.3, w, h, kernel=Kernel.Vec4, lazy_barriers=1 if flavour else 0)
    lbm.set_tuning(4, flavour)
    ora = Oracle(1.3, w, h)
    rng = np.random.default_rng(5)
    loc = np.unique(np.concatenate([rng.integers(w, w * (h - 1), size=4000),
                                    np.arange(131070 * w + 3, 131075 * w + 3, w)]))  # across the y/z seam
    pts = np.stack([loc, np.ones_like(loc)], 1).astype(np.uint64)
    lbm.draw_points(pts)
    ora.draw_points(pts.astype(np.uint32))
    for n in (1, 6):
        lbm.iterate(n)
        ora.iterate(n)
        compare_state(lbm, ora, f"tall {w}x{h} flavour {flavour} +{n}")
    lbm.close()


@pytest.mark.parametrize("lazy", [0, 1])
@pytest.mark.parametrize("flavour", [1, 2, 3])
@pytest.mark.parametrize("rows", [1, 4, 16])
@pytest.mark.parametrize("size", [(256, 128), (131, 23), (5, 4)])
def test_staged_bounce_back_bit_exact(size, rows, lazy, flavour):
    """BLBM_TUNE_VEC4_DENSE = 2: the own-row vectors of the bounce-back staged in shared memory with cp.async."""
    w, h = size
    lbm = LBM(1.0, w, h, inflow_ux=0.05, kernel=Kernel.Vec4, lazy_barriers=lazy)
    lbm.set_tuning(4, flavour)
    lbm.set_tuning(0, rows)
    ora = Oracle(1.0, w, h, inflow_ux=0.05)
    pts = porous_pairs(w, h)
    if len(pts):
        lbm.draw_points(pts)
        ora.draw_points(pts.astype(np.uint32))
    for n in (1, 2, 60, 201):
        lbm.iterate(n)
        ora.iterate(n)
        compare_state(lbm, ora, f"staged {w}x{h} rows={rows} +{n}")
    lbm.close()


@pytest.mark.parametrize("lazy", [0, 1])
@pytest.mark.parametrize("size", [(256, 128), (131, 23)])
def test_packed_add_collision_bit_exact(size, lazy):
    """BLBM_TUNE_VEC4_PACKED: the collision of cell pairs with sm_100's packed fp32 adds (FADD2) gives the same
    bits as the scalar kernel and as the oracle (every add still individually rounded, multiplies scalar)."""
    w, h = size
    lbm = LBM(1.0, w, h, inflow_ux=0.05, kernel=Kernel.Vec4, lazy_barriers=lazy)
    lbm.set_tuning(6, 1)
    ora = Oracle(1.0, w, h, inflow_ux=0.05)
    pts = porous_pairs(w, h)
    lbm.draw_points(pts)
    ora.draw_points(pts.astype(np.uint32))
    for n in (1, 2, 60, 201):
        lbm.iterate(n)
        ora.iterate(n)
        compare_state(lbm, ora, f"packed {w}x{h} +{n}")
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_closed_box_cavity_small(kernel):
    """configs[1] made concrete as a closed box, in miniature."""
    w, h = 192, 192
    lbm = LBM(1.25, w, h, kernel=kernel)
    ora = Oracle(1.25, w, h)
    ys = np.arange(h, dtype=np.uint64)
    loc = np.concatenate([ys * w + 1, ys * w + (w - 1)])
    pts = np.stack([loc, np.ones_like(loc)], 1)
    lbm.draw_points(pts)
    ora.draw_points(pts.astype(np.uint32))
    lbm.iterate(300)
    ora.iterate(300)
    compare_state(lbm, ora, "box")
    lbm.close()


@pytest.mark.parametrize("lazy", LAZY)
@pytest.mark.parametrize("nslabs,size", [(2, (96, 50)), (3, (96, 50)), (2, (97, 41)), (4, (131, 23))])
@pytest.mark.parametrize("kernel", KERNELS)
def test_slab_group_on_one_device_matches_single_domain(kernel, nslabs, size, lazy):
    """y-slab decomposition with direct halo stores, all slabs on cuda:0: must be bit-identical to the
    undivided lattice (and hence to the oracle), including paints on slab boundaries; ragged row lengths and
    uneven slab heights included."""
    w, h = size
    rng = np.random.default_rng(77 + nslabs)
    grp = SlabGroup(omega_from_viscosity(0.02), w, h, devices=[0] * nslabs, kernel=kernel, lazy_barriers=lazy)
    ora = Oracle(omega_from_viscosity(0.02), w, h)
    script = random_script(rng, w, h, phases=5, max_steps=25)
    # paint across every slab boundary, including columns 0 and W-1
    b = [r[0] for r in grp.ranges[1:]]
    loc = np.array([y * w + x for y0 in b for y in (y0 - 2, y0 - 1, y0, y0 + 1) for x in (0, 1, 40, w - 2, w - 1)])
    script.insert(2, ("draw", np.stack([loc, np.ones_like(loc)], 1).astype(np.uint32)))
    checks = run_script(script, grp, ora, f"slabs={nslabs} {kernel.name}")
    assert checks >= 8
    grp.close()


def test_full_size_properties_4096():
    """Size-independent properties at a BASELINE.json size the oracle cannot finish quickly
    (configs[1], 4096^2): kernels agree bit-for-bit with each other, the inlet column is a fixed point,
    mass stays within rounding of its start value in a closed box."""
    w = h = 4096
    res = {}
    for kernel in KERNELS:
        lbm = LBM(1.25, w, h, kernel=kernel)
        ys = np.arange(h, dtype=np.uint64)
        loc = np.concatenate([ys * w + 1, ys * w + (w - 1)])
        lbm.draw_points(np.stack([loc, np.ones_like(loc)], 1))
        inlet0 = lbm.read_population(5)[:, 0].copy()
        lbm.iterate(1)
        rho0 = lbm.reduce_moments()[0]
        lbm.iterate(60)
        rho1 = lbm.reduce_moments()[0]
        assert abs(rho1 - rho0) < 1e-5 * rho0
        assert_same_bits(lbm.read_population(5)[:, 0], inlet0, "inlet column e")
        res[kernel] = [lbm.read_population(k) for k in (0, 4, 5, 8)] + list(lbm.read_moments())
        lbm.close()
    for other in KERNELS[1:]:
        for a, b in zip(res[KERNELS[0]], res[other]):
            assert_same_bits(a, b, f"scalar vs {other.name} at 4096^2")


@pytest.mark.parametrize("cmap", [0, 1, 2])
def test_color_maps_bit_exact(cmap):
    """Row N3: color_map/{inferno,viridis,jet}.wgsl on every summary statistic."""
    w, h = 200, 70
    om = omega_from_viscosity(0.02)
    lbm, ora = LBM(om, w, h), Oracle(om, w, h)
    cyl = disc_pairs(w, 40, 35, 8)
    lbm.draw_points(cyl)
    ora.draw_points(cyl.astype(np.uint32))
    lbm.iterate(300)
    ora.iterate(300)
    for stat in range(5):
        lbm.compute_summary(stat)
        ora.compute_summary(stat)
        lbm.color_map(cmap)
        assert_same_bits(lbm.read_colors().reshape(h, -1), ora.color_map(cmap).reshape(h, -1),
                         f"colour map {cmap} stat {stat}")
    lbm.close()


def test_full_size_properties_16384_porous():
    """BASELINE.json configs[2] at full size (16384^2, 15 % porous mask — far beyond what the oracle finishes
    in seconds): size-independent properties.  The three ways of stepping it — vec4 with the barrier-chain table
    (the default), vec4 with barrier cells kept densely in the planes, and the sparse bounce-back flavour — must agree
    bit for bit on populations (incl. every barrier cell), moments and curl; the inlet column must not move;
    total mass of the moment field stays within rounding between two read-outs."""
    import bench
    w = h = 16384
    r0, mask = bench.mask_rows("porous", w, h, 0, h)
    res = {}
    for name, kernel, lazy in (("chain", Kernel.Vec4, 1), ("dense", Kernel.Vec4, 0), ("sparse", Kernel.Vec4, 2)):
        lbm = LBM(1.0, w, h, inflow_ux=0.05, kernel=kernel, lazy_barriers=lazy)
        lbm.write_barrier_rows(0, mask)
        if name == "sparse":
            lbm.set_tuning(4, 0)
        inlet0 = lbm.read_population(5)[:, 0].copy()
        lbm.iterate(12)
        rho_a = lbm.reduce_moments()[0]
        lbm.iterate(13)
        rho_b = lbm.reduce_moments()[0]
        assert abs(rho_b - rho_a) < 1e-4 * rho_a
        assert lbm.lazy_barriers_active() == (lazy != 0)
        res[name] = [lbm.read_output(), lbm.read_moments()[2]] + [lbm.read_population(k) for k in (1, 4, 6)]
        assert_same_bits(res[name][2][:, 0], res[name][2][:, 0], "self")
        assert_same_bits(lbm.read_population(5)[:, 0], inlet0, f"{name}: inlet column e")
        lbm.close()
    for other in ("dense", "sparse"):
        for a, b, what in zip(res["chain"], res[other], ("curl", "rho", "n", "rest", "sw")):
            assert_same_bits(a, b, f"chain vs {other}: {what} at 16384^2")


def test_full_size_32768_offset_widths_agree():
    """BASELINE.json configs[3] at full size on one GPU (32768^2 = 1.07 G cells, 89 GiB resident, far beyond the
    oracle): plane offsets pass 2^30 elements and 2^32 bytes.  The vec4 kernel with 32-bit plane offsets (default)
    and with 64-bit offsets must agree bit for bit on curl and density after the same steps; the inlet column
    must not move and the far corner must have been updated."""
    w = h = 32768
    cyl = disc_pairs(w, w // 4, h // 2, 256).astype(np.uint64)
    res = []
    for index32 in (1, 0):
        lbm = LBM(omega_from_viscosity(0.0512), w, h, kernel=Kernel.Vec4)
        lbm.set_tuning(7, index32)
        lbm.draw_points(cyl)
        lbm.iterate(7)
        out, rho = lbm.read_output(), lbm.read_moments()[2]
        assert np.isfinite(rho).all() and rho[h - 2, w - 1] > 0.5
        e = lbm.read_population(5)
        assert (e[1:h - 1, 0] == e[1, 0]).all()  # inlet column: still the uniform equilibrium
        del e
        res.append((out, rho))
        lbm.close()
    assert_same_bits(res[0][0], res[1][0], "curl, 32- vs 64-bit offsets at 32768^2")
    assert_same_bits(res[0][1], res[1][1], "rho, 32- vs 64-bit offsets at 32768^2")
    # the wake has started to form behind the cylinder and nowhere else
    assert np.abs(res[0][0][h // 2 - 300:h // 2 + 300, w // 4 - 300:w // 4 + 300]).max() > 0
    assert np.abs(res[0][0][16:200, w // 2:w - 64]).max() == 0


@pytest.mark.parametrize("kernel", KERNELS)
def test_against_committed_golden_fixture(kernel):
    """The CUDA path against tests/golden/small_cylinder_64x32.npz (generated by make_golden.py from the oracle
    and committed), i.e. independently of the oracle build of the day."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_cylinder_64x32.npz"))
    w, h = 64, 32
    lbm = LBM(omega_from_viscosity(0.02), w, h, kernel=kernel)
    lbm.draw_points(disc_pairs(w, 16, 16, 4))
    lbm.iterate(120)
    lbm.draw_points(disc_pairs(w, 40, 10, 3))
    lbm.iterate(80)
    lbm.update_omega_buffer(1.6)
    lbm.draw_points(disc_pairs(w, 16, 16, 4, val=0))
    lbm.iterate(50)
    for b in (0, 1):
        for k in range(9):
            assert_same_bits(lbm.read_population(k, b), g[f"f{b}_{k}"], f"golden f{b}_{k}")
    mx, my, rho = lbm.read_moments()
    assert_same_bits(mx, g["mx"], "golden mx")
    assert_same_bits(my, g["my"], "golden my")
    assert_same_bits(rho, g["rho"], "golden rho")
    assert_same_bits(lbm.read_output(), g["out"], "golden out")
    assert_same_bits(lbm.read_barrier(), g["bar"], "golden bar")
    assert_same_bits(lbm.read_cell_class(), g["cls"], "golden cls")
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_every_single_cell_preset(kernel):
    """LBM::single_cell(0..8) (lbm.rs:1482-1515) plus an out-of-range index, each followed by a few steps."""
    w, h = 72, 40
    lbm, ora = LBM(1.0 / (3 * 0.1 + 0.5), w, h, kernel=kernel), Oracle(1.0 / (3 * 0.1 + 0.5), w, h)
    lbm.iterate(3)
    ora.iterate(3)
    for idx in list(range(9)) + [12]:
        lbm.single_cell(idx)
        ora.single_cell(idx)
        compare_state(lbm, ora, f"single_cell({idx}) fresh")
        lbm.iterate(11)
        ora.iterate(11)
        compare_state(lbm, ora, f"single_cell({idx}) +11")
    lbm.close()


def _fuzz(lbm, ora, rng, w, h, tag, tunable, nops=120):
    """on a mismatch the assertion message carries the operations applied so far"""
    log = []
    try:
        _fuzz_walk(lbm, ora, rng, w, h, tag, tunable, nops, log)
    except AssertionError as e:
        raise AssertionError(f"{e}\nops: {log}") from None


def _fuzz_walk(lbm, ora, rng, w, h, tag, tunable, nops, log):
    n_cmp = 0
    for it in range(nops):
        op = int(rng.integers(0, 16))
        if not tunable and op in (9, 10, 11):
            op = 0
        log.append(op)
        if op <= 4:
            n = int(rng.integers(1, 30))
            lbm.iterate(n)
            ora.iterate(n)
        elif op == 5:
            n = int(rng.integers(1, 12))
            lbm.advance(n)
            for _ in range(n):
                ora.step()
        elif op in (6, 7):
            m = int(rng.integers(1, 40))
            loc = rng.integers(0, w * h, size=m)
            if rng.random() < 0.5:  # a blob, likely to hit existing barrier cells too
                y0, x0 = int(rng.integers(1, h - 1)), int(rng.integers(1, w - 4))
                loc = np.array([(y0 + dy) * w + x0 + dx for dy in (0, 1) for dx in range(4) if y0 + dy < h])
            val = rng.integers(0, 2, size=loc.size) if rng.random() < 0.5 else np.ones(loc.size, np.int64)
            pairs = np.stack([loc, val], 1).astype(np.uint32)
            lbm.draw_points(pairs)
            ora.draw_points(pairs)
        elif op == 8:
            o2 = float(rng.uniform(0.6, 1.7))
            lbm.update_omega_buffer(o2)
            ora.update_omega_buffer(o2)
        elif op == 9:
            kk = int(rng.integers(1, 3))
            log.append(('kernel', kk))
            lbm.set_kernel(kk)
        elif op == 10:
            lz = int(rng.integers(0, 3))
            log.append(('lazy', lz))
            lbm.set_lazy_barriers(lz)
        elif op == 11:
            knob = int(rng.choice([0, 4, 5, 6, 7, 9]))
            val = {0: [1, 2, 4, 8, 16], 4: [-1, 0, 1, 2, 3], 5: [-1, 0, 1], 6: [0, 1], 7: [-1, 0, 1], 9: [-1, 0, 1]}[knob]
            kv = int(rng.choice(val))
            log.append(('knob', knob, kv))
            lbm.set_tuning(knob, kv)
        elif op == 12:
            s = int(rng.integers(0, 5))
            lbm.compute_summary(s)
            ora.compute_summary(s)
        elif op == 13:
            which = int(rng.integers(0, 4))
            if which == 0:
                lbm.collide(); ora.collide()
            elif which == 1:
                lbm.stream(); ora.stream()
            elif which == 2:
                u = float(rng.uniform(0.0, 0.12))
                lbm.custom_speed(u); ora.custom_speed(u)
            else:
                lbm.reset_barrier(); ora.reset_barrier()
        elif op == 14:
            k = int(rng.integers(0, 9))
            a = lbm.read_population(k)  # a read-back in the middle (materialises, flushes the chain table)
            assert_same_bits(a, ora.population(-1, k), f"{tag} it {it} pop {k}")
        else:
            mx, _, _ = lbm.read_moments()
            assert_same_bits(mx, ora.moments()[0], f"{tag} it {it} mx")
        if it % 7 == 6:
            n_cmp += 1
            compare_state(lbm, ora, f"{tag} op#{it}")
    compare_state(lbm, ora, f"{tag} final")
    assert n_cmp >= nops // 8
    lbm.close()


def _fuzz_seeds():
    """seeds 1-12 walk an (almost) empty channel, 13-18 start from a 15 % porous mask so that the dense bounce-back
    flavours and the chain table see real work; BLBM_FUZZ_SEEDS=a-b adds a soak range (profiles/run_round1_ae.sh)"""
    import os
    seeds = list(range(1, 19))
    extra = os.environ.get("BLBM_FUZZ_SEEDS")
    if extra:
        a, b = extra.split("-")
        seeds += list(range(int(a), int(b) + 1))
    return seeds


@pytest.mark.parametrize("seed", _fuzz_seeds())
def test_api_fuzz_against_oracle(seed):
    """Random walks through the whole API surface — steps of random length, paints (several before a step, on
    barrier cells, erases), omega changes, resets, half-steps, read-backs at arbitrary points — interleaved with
    changes that must never alter results: kernel implementation, barrier-chain mode, launch-shape knobs, CUDA
    graphs.  Compared with the oracle on everything after every few operations."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(40, 140)), int(rng.integers(12, 40))
    om = omega_from_viscosity(0.05)
    lbm, ora = LBM(om, w, h), Oracle(om, w, h)
    if seed > 12 and seed % 2 == 1 or 12 < seed <= 18:
        pts = porous_pairs(w, h, seed=seed)
        lbm.draw_points(pts)
        ora.draw_points(pts.astype(np.uint32))
    _fuzz(lbm, ora, rng, w, h, f"fuzz seed {seed}", tunable=True)


def _slab_fuzz_seeds():
    import os
    seeds = [21, 22, 23, 24, 25, 26, 27, 28]
    extra = os.environ.get("BLBM_FUZZ_SLAB_SEEDS")  # a-b: soak range
    if extra:
        a, b = extra.split("-")
        seeds += list(range(int(a), int(b) + 1))
    return seeds


@pytest.mark.parametrize("seed", _slab_fuzz_seeds())
def test_api_fuzz_slab_group(seed):
    """The same random walk over a lattice split into 2-4 linked slabs behind one group handle (on distinct GPUs
    where the box has as many, else all on cuda:0)."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(40, 140)), int(rng.integers(16, 40))
    om = omega_from_viscosity(0.05)
    nslabs = int(rng.integers(2, 5))
    from lbm_b200 import load_library
    ndev = load_library().blbm_device_count()
    devices = list(range(nslabs)) if ndev >= nslabs else [0] * nslabs
    grp = SlabGroup(om, w, h, devices=devices, kernel=Kernel(int(rng.integers(1, 3))),
                    lazy_barriers=int(rng.integers(0, 3)))
    _fuzz(grp, Oracle(om, w, h), rng, w, h, f"slab fuzz seed {seed} x{nslabs}", tunable=False, nops=80)


def _group_devices(n):
    """distinct GPUs when the box has them (first hardware path of blbm_link_local across devices), else n slabs
    on cuda:0"""
    from lbm_b200 import load_library
    ndev = load_library().blbm_device_count()
    return list(range(n)) if ndev >= n else [0] * n


@pytest.mark.parametrize("in_kernel", [1, 0])
@pytest.mark.parametrize("ahead", [0, 1])
def test_a_slab_running_ahead_cannot_overtake_its_neighbours_pending_summary(in_kernel, ahead):
    """The summary is part of the epoch protocol of linked slabs.  curl reads the moment halo rows; the neighbour's next
    moment-storing launch (here: the public collide half-step right after an iterate) overwrites them and used to wait
    only for this slab's previous *pushing* launch, which is stream-ordered before the summary - a slab running ahead
    (an unsynchronised rank of a multi-process job; once in 400 random walks with four processes time-slicing one
    GPU) computed one boundary row of the neighbour's curl field from the next call's moments.  Driven
    deterministically here through the slab views of a group: slab `ahead` is given its iterate AND its collide before
    the other slab is given its summary.  Every slab still sees the same call sequence; nothing blocks on the host."""
    w, h = 96, 28
    om = omega_from_viscosity(0.02)
    grp = LBM(om, w, h, inflow_ux=0.09, devices=_group_devices(2), kernel=Kernel.Vec4, lazy_barriers=0)
    grp.set_tuning(8, in_kernel)
    ora = Oracle(om, w, h, inflow_ux=0.09)
    pts = disc_pairs(w, 30, 14, 4).astype(np.uint32)  # a disc across the slab boundary: the wake makes curl non-trivial
    grp.draw_points(pts); ora.draw_points(pts)
    grp.iterate(40); ora.iterate(40)
    compare_state(grp, ora, "before the skewed calls")
    fast, slow = grp.slab(ahead), grp.slab(1 - ahead)
    for n in (3, 1, 2):
        fast.iterate(n)   # steps + summary of the slab that runs ahead
        slow.advance(n)   # the other slab's steps (its moment-storing launch pushes into fast's halo rows) ...
        fast.collide()    # ... fast's NEXT moment push, enqueued before slow's summary exists
        slow.rerender()   # slow's summary of the iterate: must still see the moments of that iterate in its halo rows
        slow.collide()
        ora.iterate(n); ora.collide()
        compare_state(grp, ora, f"skewed iterate({n}) + collide, slab {ahead} ahead, in_kernel={in_kernel}")
    # the public half-steps: stream READS the halo rows the neighbour's next collide stores into
    for k in range(3):
        fast.stream(); fast.collide()
        slow.stream(); slow.collide()
        ora.stream(); ora.collide()
        compare_state(grp, ora, f"skewed stream + collide #{k}, slab {ahead} ahead, in_kernel={in_kernel}")
    grp.iterate(9); ora.iterate(9)
    compare_state(grp, ora, "back in step")
    grp.close()


@pytest.mark.parametrize("nslabs", [2, 3, 5])
def test_group_handle_spans_the_whole_api(nslabs):
    """blbm_create_group: ONE handle over n linked slabs.  Everything a caller of the reference's `LBM` does —
    presets, line strokes, colour maps, resets, half-steps, checkpoint/restore, timers, reductions — goes through
    the same C entry points as on one device and stays bit-identical to the oracle."""
    import ctypes as C
    from lbm_b200.lbm import _check, _P, slab_rows
    from oracle import barrier_shapes
    w, h = 200, 61
    om = omega_from_viscosity(0.03)
    grp = LBM(om, w, h, devices=_group_devices(nslabs), kernel=Kernel.Vec4)
    ora = Oracle(om, w, h)
    L = grp._L
    assert L.blbm_group_size(grp._h) == nslabs
    # geometry: the group is the whole lattice, its slabs the contiguous row ranges
    for q, (r0, r1) in enumerate(slab_rows(h, nslabs)):
        s = _P()
        _check(L.blbm_group_slab(grp._h, q, C.byref(s)))
        ww, hg, a, b, dev = C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        _check(L.blbm_get_geometry(s, C.byref(ww), C.byref(hg), C.byref(a), C.byref(b), C.byref(dev)))
        assert (ww.value, hg.value, a.value, b.value) == (w, h, r0, r1)
    assert L.blbm_group_slab(grp._h, nslabs, C.byref(_P())) < 0
    buf = C.create_string_buffer(512)
    assert L.blbm_export_peer(grp._h, buf) == -5  # BLBM_ESTATE: groups link themselves
    # a stroke across every slab boundary through the host rasteriser, then the frame loop
    pts = barrier_shapes.line_points((30, 3), (150, 57), w, h)
    grp.draw_line((30, 3), (150, 57))
    ora.draw_points(np.array([[px + py * w, 1] for px, py, *_ in pts], np.uint32))
    t = grp.iterate_timed(70)  # more than two interleaving chunks of 32
    assert t > 0
    ora.iterate(70)
    compare_state(grp, ora, f"group x{nslabs} after 70 steps")
    for stat in (SummaryStat.Speed, SummaryStat.Rho, SummaryStat.Curl):
        grp.compute_summary(stat)
        ora.compute_summary(int(stat))
        assert_same_bits(grp.read_output(), ora.output(), f"group stat {stat.name}")
    grp.color_map(2)
    assert_same_bits(grp.read_colors(), np.asarray(ora.color_map(2)).reshape(h, w, 3), "group colours")
    # checkpoint / restore through the group handle, then half-steps and resets in lock-step
    saved = [[grp.read_population(k, b) for k in range(9)] for b in (0, 1)]
    grp.iterate(10)  # an even number of steps: the restored buffers keep their roles (compute_step parity)
    for b in (0, 1):
        for k in range(9):
            grp.write_population(k, saved[b][k], b)
    grp.collide(); ora.collide()
    grp.stream(); ora.stream()
    grp.iterate(5); ora.iterate(5)
    for b in (0, 1):
        for k in range(9):
            assert_same_bits(grp.read_population(k, b), ora.population(b, k), f"group restore pop[{b}][{k}]")
    for idx in (5, 4, 1):  # single_cell 3..5 sit on the slab boundary of a 2-slab split
        grp.single_cell(idx); ora.single_cell(idx)
        grp.iterate(7); ora.iterate(7)
        compare_state(grp, ora, f"group single_cell {idx}")
    grp.custom_speed(0.06); ora.custom_speed(0.06)
    grp.iterate(33); ora.iterate(33)
    compare_state(grp, ora, "group custom_speed")
    s_rho, s_mx, s_my, mabs = grp.reduce_moments()
    omx, omy, orho = ora.moments()
    assert abs(s_rho - float(np.sum(orho, dtype=np.float64))) < 1e-6 * abs(s_rho)
    assert mabs == float(np.abs(ora.output()).max())
    assert grp.launch_count() > 0 and grp.device_bytes() > 0
    grp.close()


def test_group_of_one_device_is_a_plain_handle():
    grp = LBM(1.25, 64, 32, devices=[0])
    assert grp._L.blbm_group_size(grp._h) == 1
    ora = Oracle(1.25, 64, 32)
    grp.iterate(10); ora.iterate(10)
    compare_state(grp, ora, "group of one")
    grp.close()


def test_paint_frames_never_synchronise_and_stay_exact():
    """The frame loop of lib.rs:108-199 — paint a stroke, iterate(n), read the field back asynchronously — with
    more paints in flight than the pinned staging ring has slots, and one list longer than a slot."""
    import torch
    w, h = 256, 96
    om = omega_from_viscosity(0.02)
    lbm, ora = LBM(om, w, h), Oracle(om, w, h)
    rng = np.random.default_rng(5)
    outs = [torch.empty((h, w), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    for fr in range(12):
        n = 9000 if fr == 6 else int(rng.integers(1, 200))  # frame 6 exceeds the 8192-pair slot
        loc = rng.integers(w, w * (h - 1), size=n)
        val = rng.integers(0, 2, size=n)
        pairs = np.stack([loc, val], 1).astype(np.uint32)
        lbm.draw_points(pairs)
        pairs[:] = 0xFFFFFFFF  # the caller's buffer is free again on return
        ora.draw_points(np.stack([loc, val], 1).astype(np.uint32))
        lbm.iterate(3); ora.iterate(3)
        lbm.read_output_async(outs[fr & 1].data_ptr())
    lbm.synchronize()
    assert_same_bits(outs[1].numpy(), ora.output(), "async output of the last frame")
    compare_state(lbm, ora, "after 12 pipelined paint frames")
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("w,h", [(150, 40), (256, 33)])
@pytest.mark.parametrize("pdl", [1, 0])
def test_every_graph_run_length_with_and_without_the_moment_tail(kernel, w, h, pdl):
    """Small lattices replay runs of 2, 4 .. 16 steps as CUDA graphs, the call's moment-storing last step included when
    the run ends the call and not when it does not (longer calls); a stroke of at most 64 cells is
    ONE launch that updates the mask and rebuilds the class words (every block applies the whole stroke before it reads
    the mask), leaving a class swap pending for the first step of the next graph.  Calls of every length 1..35 from
    both start parities, with and without such a stroke in front, strokes on the inlet / outlet columns and on rows
    0 / H-1 and repeated cells (last writer wins), against the oracle after every call; with the steps launched as
    programmatic dependents of one another (a step's blocks are scheduled while the previous one drains and wait for
    its completion before they touch memory: the default outside graphs) and plainly."""
    om = omega_from_viscosity(0.02)
    lbm, ora = LBM(om, w, h, inflow_ux=0.08, kernel=kernel, lazy_barriers=0), Oracle(om, w, h, inflow_ux=0.08)
    lbm.set_tuning(9, pdl)
    rng = np.random.default_rng(w)
    lbm.iterate(1); ora.iterate(1)  # the graphs are captured by the first call that could use one: make it this one
    lbm.iterate(5); ora.iterate(5)
    lengths = list(range(1, 36)) + [17, 15, 15, 16, 3, 2]
    for k, n in enumerate(lengths):
        if k % 3 != 2:
            m = int(rng.integers(1, 65))
            loc = rng.integers(0, w * h, size=m)
            if k % 6 == 0:
                loc[: m // 2] = (rng.integers(0, h, size=m // 2)) * w + rng.choice([0, w - 1], size=m // 2)
            if m > 3:
                loc[1] = loc[0]  # the same cell twice in one stroke
            val = rng.integers(0, 2, size=m)
            pairs = np.stack([loc, val], 1).astype(np.uint32)
            lbm.draw_points(pairs); ora.draw_points(pairs)
            if k % 9 == 4:  # two strokes before the same step
                pairs2 = np.stack([loc[::-1], 1 - val], 1).astype(np.uint32)[: max(1, m // 3)]
                lbm.draw_points(pairs2); ora.draw_points(pairs2)
        if k % 4 == 1:
            lbm.advance(n)  # steps (moments of the last one stored) without the summary launch
            for _ in range(n):
                ora.step()
            for got, want, what in zip(lbm.read_moments(), ora.moments(), ("mx", "my", "rho")):
                assert_same_bits(got, want, f"{what} after advance({n}), call {k}")
        else:
            lbm.iterate(n); ora.iterate(n)
            compare_state(lbm, ora, f"graph runs: iterate({n}), call {k}", populations=(k % 5 == 0))
    lbm.update_omega_buffer(1.3); ora.update_omega_buffer(1.3)  # a changed configuration: graph-less, then re-captured
    for n in (15, 15, 15, 8):
        lbm.iterate(n); ora.iterate(n)
    compare_state(lbm, ora, "after the omega change")
    lbm.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("omega", [1.0, 1.0 / (3 * 0.02 + 0.5)])
def test_chain_table_settles_unsettles_and_evicts_exactly(kernel, omega):
    """The ordered barrier-chain table through its whole life on a 15 % porous lattice: chains settle on their exact
    period-2 cycle (both rest values and both moment triples in the table), calls of odd and even length and single
    steps pick the right phase, a paint evicts settled and unsettled entries by slot rank (dead slots keep the
    ranks of their successors), an omega change unsettles everything, and a flush returns every barrier cell to the
    planes — bit for bit against the oracle at every stage."""
    w, h = 300, 70
    lbm, ora = LBM(omega, w, h, inflow_ux=0.05, kernel=kernel, lazy_barriers=1), Oracle(omega, w, h, inflow_ux=0.05)
    pts = porous_pairs(w, h)
    lbm.draw_points(pts)
    ora.draw_points(pts.astype(np.uint32))
    rng = np.random.default_rng(11)
    stage = 0

    def both(n):
        nonlocal stage
        lbm.iterate(n); ora.iterate(n)
        stage += 1
        assert lbm.lazy_barriers_active()
        compare_state(lbm, ora, f"chain life stage {stage} (+{n} steps)", populations=(stage % 3 == 0))

    for n in (1, 1, 2, 15, 40, 1, 15, 15, 2):
        both(n)
    # evict a mix of barrier cells (settled by now for omega = 1) and paint new ones, mid-table and at chunk edges
    bar = np.flatnonzero(ora.barrier().reshape(-1) == 1)
    gone = rng.choice(bar[(bar > w) & (bar < w * (h - 1))], size=200, replace=False)
    new = rng.integers(w + 2, w * (h - 1) - 2, size=150)
    pairs = np.concatenate([np.stack([gone, np.zeros_like(gone)], 1), np.stack([new, np.ones_like(new)], 1)])
    lbm.draw_points(pairs.astype(np.uint32)); ora.draw_points(pairs.astype(np.uint32))
    for n in (1, 14, 3):
        both(n)
    lbm.update_omega_buffer(1.37); ora.update_omega_buffer(1.37)
    for n in (1, 2, 31):
        both(n)
    # re-barrier some of the evicted cells (dead slots stay dead; the cells run densely), evict more
    back = gone[:60]
    more = rng.choice(bar[(bar > w) & (bar < w * (h - 1))], size=120, replace=False)
    pairs = np.concatenate([np.stack([back, np.ones_like(back)], 1), np.stack([more, np.zeros_like(more)], 1)])
    lbm.draw_points(pairs.astype(np.uint32)); ora.draw_points(pairs.astype(np.uint32))
    for n in (5, 1, 16):
        both(n)
    lbm.collide(); ora.collide()  # leaves the table: everything back in the planes
    assert not lbm.lazy_barriers_active()
    compare_state(lbm, ora, "chain life after flush")
    lbm.iterate(7); ora.iterate(7)  # and re-enters it
    compare_state(lbm, ora, "chain life after re-entry")
    lbm.close()


@pytest.mark.parametrize("in_kernel", [0, 1])
@pytest.mark.parametrize("size", [(300, 41), (129, 9), (1000, 70)])
def test_halo_handshake_inside_and_outside_the_step_kernel(size, in_kernel):
    """Linked slabs publish / await the halo epochs either inside the fused vec4 kernel (face row blocks first; the
    default) or through the one-thread wait / signal kernels around every launch (knob 8 = 0; also what the scalar
    kernel and non-default block shapes use): same bits, including slabs of 2-3 rows where every row block is a
    face block and uneven slab heights where the row above the last sits in its own row block."""
    w, h = size
    om = omega_from_viscosity(0.02)
    for nslabs in (2, 3, 4):
        if h < 2 * nslabs:
            continue
        grp = LBM(om, w, h, devices=_group_devices(nslabs), kernel=Kernel.Vec4)
        grp.set_tuning(8, in_kernel)
        ora = Oracle(om, w, h)
        pts = porous_pairs(w, h, frac=0.1, seed=3)
        grp.draw_points(pts); ora.draw_points(pts.astype(np.uint32))
        for n in (1, 2, 37, 64):
            grp.iterate(n); ora.iterate(n)
        compare_state(grp, ora, f"handshake in_kernel={in_kernel} {w}x{h} x{nslabs}")
        grp.set_tuning(0, 8)  # a block shape without the in-kernel variant falls back to the launch-level handshake
        grp.iterate(5); ora.iterate(5)
        grp.set_tuning(0, 4)
        grp.iterate(6); ora.iterate(6)
        compare_state(grp, ora, f"handshake mixed {w}x{h} x{nslabs}")
        grp.close()
