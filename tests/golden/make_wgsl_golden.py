"""Generates tests/golden/wgsl_golden.npz by INTERPRETING THE REFERENCE'S OWN WGSL SHADER TEXT
(/root/reference/lbm-wgpu/src/rewritten_shaders/*.wgsl, via oracle/wgsl_interp.py) on the scenarios of
tests/wgsl_cases.py.  Run in the build container, where /root/reference exists:

    python tests/golden/make_wgsl_golden.py

The GPU box has no /root/reference; the committed .npz travels instead.  A SHA-256 of every shader file that
was interpreted is stored alongside, so a test can tell whether the fixture still matches the reference tree."""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import wgsl_interp  # noqa: E402
from tests import wgsl_cases  # noqa: E402


def shader_digests(root=wgsl_interp.SHADER_ROOT):
    out = {}
    for d, _, files in sorted(os.walk(root)):
        for f in sorted(files):
            if f.endswith(".wgsl"):
                p = os.path.join(d, f)
                out[os.path.relpath(p, root)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    return out


def main():
    arrays = {}
    for name, (omega, w, h, u0, script) in wgsl_cases.cases().items():
        t = time.time()
        sim = wgsl_interp.WgslLBM(omega, w, h, inflow_ux=u0)
        shots = wgsl_cases.replay(script, sim, lambda s: s.state())
        for i, st in enumerate(shots):
            for k in wgsl_cases.STATE_KEYS:
                arrays[f"{name}/{i}/{k}"] = st[k]
        arrays[f"{name}/count"] = np.int64(len(shots))
        print(f"{name}: {len(shots)} snapshots, {time.time() - t:.1f} s", flush=True)
    # colour maps of one non-trivial output field (N3)
    sim = wgsl_interp.WgslLBM(1.25, 24, 13)
    sim.draw_points(wgsl_cases.disc_pairs(24, 8, 6, 2).astype(np.uint32))
    sim.iterate(30)
    for stat in range(5):
        sim.compute_summary(stat)
        for cmap in range(3):
            arrays[f"colors/{stat}/{cmap}"] = sim.colors_of(cmap)
    dig = shader_digests()
    arrays["shader_files"] = np.array(sorted(dig))
    arrays["shader_sha256"] = np.array([dig[k] for k in sorted(dig)])
    out = os.path.join(ROOT, "tests", "golden", "wgsl_golden.npz")
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
