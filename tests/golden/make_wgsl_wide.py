"""Generates tests/golden/wgsl_wide.npz: the scenarios of tests/wgsl_cases.py: wide_cases() — lattices of up to
1024 x 512 cells — computed by EXECUTING THE REFERENCE'S OWN WGSL TEXT with the SIMT executor
(oracle/wgsl_simt.py), stored as one sha256 per buffer and snapshot.  The C oracle (CPU test) and the CUDA path
(GPU test) replay the same scripts and must reproduce every digest.

    python tests/golden/make_wgsl_wide.py [case ...]   # one process per case; named cases are merged into the file
                                                       # (box_4096x4096_1000steps: ~35 min on 8 threads, ~6 GB; the others: minutes)
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import wgsl_cases  # noqa: E402


THREADS = int(os.environ.get("WGSL_SIMT_THREADS", "8"))  # host threads per large case (chunks of one dispatch)


def run(name):
    from oracle.wgsl_simt import WgslLBMVec
    omega, w, h, u0, script = wgsl_cases.wide_cases()[name]
    t = time.time()
    sim = WgslLBMVec(omega, w, h, inflow_ux=u0, threads=THREADS if w * h >= (1 << 22) else 1)
    shots = wgsl_cases.replay(script, sim, lambda s: wgsl_cases.digest_snapshot(s.state()))
    extra = {}
    if name == wgsl_cases.COLOR_CASE:
        for stat in range(5):
            sim.compute_summary(stat)
            for cmap in range(3):
                extra[f"{name}/colors/{stat}/{cmap}"] = np.bytes_(wgsl_cases.digest_array(sim.colors_of(cmap)))
    print(f"{name}: {len(shots)} snapshots, {time.time() - t:.0f} s", flush=True)
    return name, shots, extra


def main():
    from tests.golden.make_wgsl_golden import shader_digests
    names = sys.argv[1:] or sorted(wgsl_cases.wide_cases())
    out = os.path.join(ROOT, "tests", "golden", "wgsl_wide.npz")
    arrays = dict(np.load(out)) if sys.argv[1:] and os.path.exists(out) else {}
    with ProcessPoolExecutor(min(len(names), os.cpu_count() or 1)) as ex:
        for name, shots, extra in ex.map(run, names):
            arrays.update(extra)
            arrays[f"{name}/count"] = np.int64(len(shots))
            for i, d in enumerate(shots):
                for k, v in d.items():
                    arrays[f"{name}/{i}/{k}"] = np.bytes_(v)
    dig = shader_digests()
    arrays["shader_files"] = np.array(sorted(dig))
    arrays["shader_sha256"] = np.array([dig[k] for k in sorted(dig)])
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
