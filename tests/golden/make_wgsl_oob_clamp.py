"""Generates tests/golden/wgsl_oob_clamp.npz: THE REFERENCE'S OWN WGSL SHADER TEXT EXECUTED WITH OUT-OF-RANGE READS
CLAMPED (oracle/wgsl_simt.py, OOB_POLICY = "clamp": min(u32(index), length - 1), what Tint's robustness transform and
naga's `Restrict` policy do) instead of returning 0 (the defined semantics of this repo: SURVEY.md section 8, what a
Vulkan device with robustBufferAccess2 does), on the scenarios of tests/wgsl_cases.py.  The oracle's clamp mode
(Oracle(..., oob="clamp")) must reproduce every buffer of every snapshot bit for bit.  There is no CUDA counterpart:
the vectors exist so that the size of the difference is known and a parity target that clamps
would find its oracle pinned.

    python tests/golden/make_wgsl_oob_clamp.py      # build container only (needs /root/reference); ~1 minute
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import wgsl_simt  # noqa: E402
from tests import wgsl_cases  # noqa: E402
from tests.golden.make_wgsl_golden import shader_digests  # noqa: E402


def main():
    arrays = {}
    wgsl_simt.OOB_POLICY = "clamp"
    try:
        for name, (omega, w, h, u0, script) in wgsl_cases.cases().items():
            t = time.time()
            sim = wgsl_simt.WgslLBMVec(omega, w, h, inflow_ux=u0)
            shots = wgsl_cases.replay(script, sim, lambda s: s.state())
            for i, st in enumerate(shots):
                for k in wgsl_cases.STATE_KEYS:
                    arrays[f"{name}/{i}/{k}"] = np.array(st[k]).copy()
            arrays[f"{name}/count"] = np.int64(len(shots))
            print(f"{name}: {len(shots)} snapshots, {time.time() - t:.1f} s", flush=True)
    finally:
        wgsl_simt.OOB_POLICY = "zero"
    dig = shader_digests()
    arrays["shader_files"] = np.array(sorted(dig))
    arrays["shader_sha256"] = np.array([dig[k] for k in sorted(dig)])
    out = os.path.join(ROOT, "tests", "golden", "wgsl_oob_clamp.npz")
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
