"""The scenarios whose results were produced by interpreting the reference's own WGSL shader text
(oracle/wgsl_interp.py) and committed as tests/golden/wgsl_*.npz by tests/golden/make_wgsl_golden.py.
Each case is a deterministic call script (tests/util.py: random_script and hand-written ones); a snapshot of
every buffer the reference keeps is taken at each ("compare",) entry.  The same scripts are replayed on the
C oracle (CPU tests) and on the CUDA path (GPU tests) and compared bit for bit with the snapshots."""
import numpy as np

from tests.util import disc_pairs, porous_pairs, random_script

STATE_KEYS = [f"f{b}_{k}" for b in range(2) for k in range(9)] + ["mx", "my", "rho", "out", "barrier"]


def _cylinder(w, h, nu):
    return [("draw", disc_pairs(w, w // 4, h // 2, 3).astype(np.uint32)), ("iterate", 1), ("compare",),
            ("iterate", 1), ("compare",), ("iterate", 38), ("compare",),
            ("summary", 4), ("compare",), ("summary", 3), ("compare",)]


def _porous(w, h):
    return [("draw", porous_pairs(w, h).astype(np.uint32)), ("iterate", 25), ("compare",), ("omega", 1.1),
            ("iterate", 6), ("compare",)]


def _presets():
    s = []
    for idx in (1, 2, 4, 6, 9):
        s += [("single_cell", idx), ("iterate", 9), ("compare",)]
    return s


def _edge_paints(w, h):
    """paints on every special cell class: wall rows erased (u32 underflow of i-W), inlet and outlet columns,
    the cell whose south-east neighbour index is W*H, out-of-range locations"""
    loc = [0, 3, w - 1, w, 2 * w - 1, (h - 2) * w + w - 1, (h - 1) * w + 2, w * h - 1, w * h, w * h + 5,
           (h - 2) * w, w + 1, 3 * w + 4, 3 * w + 5]
    val = [0, 0, 0, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 1]
    return [("iterate", 3), ("draw", np.stack([loc, val], 1).astype(np.uint32)), ("iterate", 1), ("compare",),
            ("iterate", 14), ("compare",), ("collide",), ("compare",), ("stream",), ("compare",),
            ("custom_speed", 0.03), ("compare",), ("iterate", 5), ("compare",), ("reset_barrier",), ("iterate", 2),
            ("compare",)]


# name -> (omega, w, h, inflow_ux, script)
def cases():
    out = {}
    for w, h in ((24, 13), (17, 9)):
        rng = np.random.default_rng(7000 * w + h)
        out[f"random_{w}x{h}"] = (1.0 / (3 * 0.02 + 0.5), w, h, 0.1, random_script(rng, w, h, max_steps=24))
    out["cylinder_40x20_nu002"] = (1.0 / (3 * 0.02 + 0.5), 40, 20, 0.1, _cylinder(40, 20, 0.02))
    out["cylinder_40x20_nu01"] = (1.25, 40, 20, 0.1, _cylinder(40, 20, 0.1))
    out["porous_33x12"] = (1.0, 33, 12, 0.05, _porous(33, 12))
    out["presets_16x12"] = (1.0 / (3 * 0.02 + 0.5), 16, 12, 0.1, _presets())
    out["edges_9x7"] = (1.7, 9, 7, 0.1, _edge_paints(9, 7))
    out["tiny_2x3"] = (1.25, 2, 3, 0.1, [("iterate", 4), ("compare",)])
    return out


def replay(script, sim, snapshot):
    """run the script on `sim` (WgslLBM, Oracle or LBM share these method names); snapshot(sim) at compares"""
    shots = []
    for op in script:
        name = op[0]
        if name == "compare":
            shots.append(snapshot(sim))
        elif name == "draw":
            sim.draw_points(op[1])
        elif name == "omega":
            sim.update_omega_buffer(op[1])
        elif name == "summary":
            sim.compute_summary(op[1])
        elif len(op) > 1:
            getattr(sim, name)(op[1])
        else:
            getattr(sim, name)()
    return shots


# ---------------------------------------------------------------------------------------------------------------
# Larger lattices, executed from the reference's WGSL text by the SIMT executor (oracle/wgsl_simt.py) and stored as
# sha256 digests per buffer (tests/golden/wgsl_wide.npz, tests/golden/make_wgsl_wide.py): sizes at which the CUDA
# kernels' structure is exercised — many 128-cell chunks per row, many row blocks, ragged row lengths, rows wider
# than 4096 cells, the barrier-chain table with paints and erases while a stream is pending.
def _porous_wide(w, h):
    disc = disc_pairs(w, w // 3, h // 2, 9)
    return [("draw", porous_pairs(w, h).astype(np.uint32)), ("iterate", 1), ("compare",), ("iterate", 60),
            ("compare",), ("draw", disc.astype(np.uint32)), ("iterate", 40), ("compare",),
            ("draw", disc_pairs(w, w // 3, h // 2, 9, val=0).astype(np.uint32)), ("omega", 1.2), ("iterate", 20),
            ("compare",), ("summary", 4), ("compare",), ("summary", 3), ("compare",)]


def _box(w, h, steps=(1, 99, 200)):
    ys = np.arange(h, dtype=np.uint64)
    loc = np.concatenate([ys * np.uint64(w) + np.uint64(1), ys * np.uint64(w) + np.uint64(w - 1)])
    return [("draw", np.stack([loc, np.ones_like(loc)], 1).astype(np.uint32)), ("iterate", steps[0]), ("compare",),
            ("iterate", steps[1]), ("compare",), ("iterate", steps[2]), ("compare",), ("summary", 1), ("compare",)]


def _porous_strip(w, h):
    return [("draw", porous_pairs(w, h).astype(np.uint32)), ("iterate", 1), ("compare",), ("iterate", 39),
            ("compare",)]


SLOW_CASE = "box_4096x4096_1000steps"
COLOR_CASE = "colormaps_700x300"


def wide_cases():
    out = {}
    out["porous_1024x512"] = (1.0, 1024, 512, 0.05, _porous_wide(1024, 512))
    out["box_512x512"] = (1.25, 512, 512, 0.1, _box(512, 512))
    # BASELINE.json configs[1] at FULL SIZE (4096^2 closed box) and a 256-row strip of
    # configs[2] at its full row length (16384 cells, 15 % porous)
    out["box_4096x4096"] = (1.25, 4096, 4096, 0.1, _box(4096, 4096, steps=(1, 19, 40)))
    out["porous_16384x256"] = (1.0, 16384, 256, 0.05, _porous_strip(16384, 256))
    # SURVEY.md 8d config 2 as specified: the 4096^2 closed box for 1,000 steps (the oracle needs ~3 minutes for it,
    # so its CPU test only runs with BLBM_SLOW_ORACLE=1; the GPU test always runs)
    out[SLOW_CASE] = (1.25, 4096, 4096, 0.1, _box(4096, 4096, steps=(100, 200, 700)))
    # row N3: after the script, every summary statistic through every colour map (digests under <name>/colors/..)
    out[COLOR_CASE] = (1.0 / (3 * 0.02 + 0.5), 700, 300, 0.1,
                       [("draw", disc_pairs(700, 175, 150, 20).astype(np.uint32)), ("iterate", 400), ("compare",)])
    for w, h, steps in ((300, 170, 40), (1001, 37, 30), (4100, 5, 12)):
        rng = np.random.default_rng(9000 * w + h)
        out[f"random_{w}x{h}"] = (1.0 / (3 * 0.02 + 0.5), w, h, 0.1, random_script(rng, w, h, max_steps=steps))
    return out


def digest_array(a):
    import hashlib
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        a = a.copy()
        a[np.isnan(a)] = np.float32(np.nan)
    return hashlib.sha256(a.tobytes()).hexdigest()


def digest_snapshot(st):
    """sha256 per buffer of a snapshot dict (NaN payloads canonicalised: they are not part of the contract)"""
    import hashlib
    out = {}
    for k in STATE_KEYS:
        a = np.ascontiguousarray(st[k])
        if a.dtype == np.float32:
            a = a.copy()
            a[np.isnan(a)] = np.float32(np.nan)
        elif k == "barrier":
            a = a.astype(np.uint32)
        out[k] = hashlib.sha256(a.tobytes()).hexdigest()
    return out
