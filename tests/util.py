"""Shared helpers of the parity tests: scenario scripts that drive the CUDA path and the oracle with the
same calls, and bit-exact comparison of everything the reference keeps in buffers."""
import numpy as np

POPS = range(9)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_same_bits(a, b, what):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype == np.float32:
        # bit-exact, except that NaN payload/sign bits are not compared: IEEE 754 (and WGSL) leave them
        # implementation-defined, and x86 SSE and the GPU pick different quiet-NaN patterns
        bad = (bits(a) != bits(b)) & ~(np.isnan(a) & np.isnan(b))
    else:
        bad = a != b
    if bad.any():
        idx = np.argwhere(bad)
        first = tuple(idx[0])
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} cells differ; first at (y,x)={first}: "
                             f"got {a[first]!r} want {b[first]!r}")


def compare_state(lbm, oracle, tag="", populations=True):
    """lbm: lbm_b200.LBM or SlabGroup; oracle: oracle.lbm_oracle.Oracle.  Bit-exact on every array the
    reference keeps: 2x9 populations, three moments, output, barrier; plus the derived class word."""
    if populations:
        for buf in (0, 1):
            for k in POPS:
                assert_same_bits(lbm.read_population(k, buf), oracle.population(buf, k),
                                 f"{tag} population[{buf}][{k}]")
    mx, my, rho = lbm.read_moments()
    omx, omy, orho = oracle.moments()
    assert_same_bits(mx, omx, f"{tag} momentum x")
    assert_same_bits(my, omy, f"{tag} momentum y")
    assert_same_bits(rho, orho, f"{tag} density")
    assert_same_bits(lbm.read_output(), oracle.output(), f"{tag} output")
    assert_same_bits(lbm.read_barrier(), oracle.barrier(), f"{tag} barrier")
    assert_same_bits(lbm.read_cell_class(), oracle.cell_class(), f"{tag} cell class")
    assert lbm.get_compute_num() == oracle.get_compute_num()


def disc_pairs(w, cx, cy, r, val=1):
    """All cells with (x-cx)^2 + (y-cy)^2 <= r^2 as sorted [loc, val] pairs."""
    ys, xs = np.mgrid[cy - r:cy + r + 1, cx - r:cx + r + 1]
    m = (xs - cx) ** 2 + (ys - cy) ** 2 <= r * r
    loc = np.sort((xs[m].astype(np.uint64) + ys[m].astype(np.uint64) * np.uint64(w)))
    return np.stack([loc, np.full_like(loc, val)], 1)


def splitmix64(x):
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def porous_pairs(w, h, frac=0.15, seed=0x5EED, row_begin=0, row_end=None):
    """cells with splitmix64(seed ^ idx) < frac*2^64, 2 <= x <= W-2, interior rows."""
    row_end = h if row_end is None else row_end
    ys = np.arange(max(row_begin, 1), min(row_end, h - 1), dtype=np.uint64)
    xs = np.arange(2, w - 1, dtype=np.uint64)
    idx = (ys[:, None] * np.uint64(w) + xs[None, :]).reshape(-1)
    hsh = splitmix64(idx ^ np.uint64(seed))
    thr = np.uint64(int(frac * 2.0 ** 64))
    loc = idx[hsh < thr]
    return np.stack([loc, np.ones_like(loc)], 1)


def random_script(rng, w, h, phases=6, max_steps=40, max_pts=30):
    """A reproducible list of API calls exercising paints on every special cell class, erases, omega
    changes, resets, half-steps and read-backs at arbitrary steps."""
    special = [0, w - 1, w, 2 * w - 1, (h - 2) * w + w - 1, (h - 1) * w, w * h - 1, w * h, w * h + 2,
               max((h - 2) * w, 0), 1, w + 1]
    script = []
    for ph in range(phases):
        script.append(("iterate", int(rng.integers(1, max_steps))))
        script.append(("compare",))
        m = int(rng.integers(1, max_pts))
        loc = np.concatenate([rng.integers(0, w * h, size=m), rng.choice(special, size=min(3, m))])
        val = rng.integers(0, 2, size=loc.size)
        script.append(("draw", np.stack([loc, val], 1).astype(np.uint32)))
        if ph == 1:
            script.append(("omega", 1.6))
        if ph == 2:
            script += [("summary", s) for s in (1, 2, 3, 4, 0)]
            script.append(("compare",))
        if ph == 3:
            script += [("collide",), ("compare",), ("stream",), ("compare",), ("iterate", 3), ("compare",)]
        if ph == 4:
            script += [("custom_speed", 0.07), ("compare",)]
    script += [("iterate", 5), ("compare",), ("reset_barrier",), ("iterate", 4), ("compare",),
               ("single_cell", 2), ("iterate", 7), ("compare",), ("reset_to_equilibrium",), ("iterate", 2),
               ("compare",)]
    return script


def run_script(script, lbm, oracle, tag=""):
    """Apply the same call sequence to both; returns the number of full-state comparisons made."""
    n = 0
    for op in script:
        name = op[0]
        if name == "iterate":
            lbm.iterate(op[1])
            oracle.iterate(op[1])
        elif name == "draw":
            lbm.draw_points(op[1])
            oracle.draw_points(op[1])  # applied in order: last writer wins
        elif name == "omega":
            lbm.update_omega_buffer(op[1])
            oracle.update_omega_buffer(op[1])
        elif name == "summary":
            lbm.compute_summary(op[1])
            oracle.compute_summary(op[1])
        elif name == "collide":
            lbm.collide()
            oracle.collide()
        elif name == "stream":
            lbm.stream()
            oracle.stream()
        elif name == "custom_speed":
            lbm.custom_speed(op[1])
            oracle.custom_speed(op[1])
        elif name == "reset_barrier":
            lbm.reset_barrier()
            oracle.reset_barrier()
        elif name == "single_cell":
            lbm.single_cell(op[1])
            oracle.single_cell(op[1])
        elif name == "reset_to_equilibrium":
            lbm.reset_to_equilibrium()
            oracle.reset_to_equilibrium()
        elif name == "compare":
            n += 1
            compare_state(lbm, oracle, f"{tag} check#{n} step {oracle.get_compute_num()}")
        else:
            raise ValueError(name)
    return n
