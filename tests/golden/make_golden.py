"""Generates tests/golden/*.npz from the CPU oracle.

The reference itself cannot be run anywhere in this environment (Rust -> wasm32 + browser WebGPU, no
read-back path), and ships no golden vectors, so these fixtures are ORACLE-generated:
they pin the oracle against accidental change and give the GPU tests a committed target that does not
depend on the oracle build of the day.  PARITY UNPINNED against the real WGSL pipeline.

    python tests/golden/make_golden.py          # rewrites the fixtures (commit the result)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.lbm_oracle import Oracle  # noqa: E402
from tests.util import disc_pairs  # noqa: E402


def omega_from_viscosity(nu):
    f = np.float32
    return float(f(1.0) / (f(3.0) * f(nu) + f(0.5)))


def digest(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:  # canonicalise NaNs: their payload bits are not part of the contract
        a = a.copy()
        a[np.isnan(a)] = np.float32(np.nan)
    return hashlib.sha256(a.tobytes()).hexdigest()


def state_arrays(o):
    d = {}
    for b in (0, 1):
        for k in range(9):
            d[f"f{b}_{k}"] = o.population(b, k).copy()
    mx, my, rho = o.moments()
    d["mx"], d["my"], d["rho"] = mx.copy(), my.copy(), rho.copy()
    d["out"] = o.output().copy()
    d["bar"] = o.barrier().copy()
    d["cls"] = o.cell_class()
    return d


def case_small_cylinder():
    """64x32, nu=0.02, disc r=4 at (16,16); 120 steps; paint a second disc; 80 steps; omega change; 50 steps."""
    w, h = 64, 32
    o = Oracle(omega_from_viscosity(0.02), w, h)
    o.draw_points(disc_pairs(w, 16, 16, 4).astype(np.uint32))
    o.iterate(120)
    o.draw_points(disc_pairs(w, 40, 10, 3).astype(np.uint32))
    o.iterate(80)
    o.update_omega_buffer(1.6)
    o.draw_points(disc_pairs(w, 16, 16, 4, val=0).astype(np.uint32))
    o.iterate(50)
    return state_arrays(o)


def case_config1_digests():
    """BASELINE.json configs[0]: 512x256 cylinder r=16 at (128,128), u0=0.1, nu=0.02; sha256 of every array
    after 1, 100 and 1000 steps (full arrays would be 10 MB per checkpoint)."""
    w, h = 512, 256
    o = Oracle(omega_from_viscosity(0.02), w, h)
    o.draw_points(disc_pairs(w, 128, 128, 16).astype(np.uint32))
    out, done = {}, 0
    for target in (1, 100, 1000):
        o.iterate(target - done)
        done = target
        for name, a in state_arrays(o).items():
            out[f"s{target}_{name}"] = np.array(digest(a))
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "small_cylinder_64x32.npz"), **case_small_cylinder())
    np.savez(os.path.join(HERE, "config1_512x256_digests.npz"), **case_config1_digests())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
