"""Generates tests/golden/wgsl_contract.npz: THE REFERENCE'S OWN WGSL SHADER TEXT EXECUTED UNDER CONTRACTION
(oracle/wgsl_contract.py: the fusions are derived from the shader text by value numbering and use counting — a float
multiply with exactly one use is fused into the add / subtract that consumes it, left operand first — and the shaders
are then run by the SIMT executor with those additions evaluated as correctly rounded fma), on the scenarios of
tests/wgsl_cases.py.  The contraction-twin builds (oracle -DLBM_CONTRACT, CUDA -DBLBM_CONTRACT) must reproduce every
buffer of every snapshot bit for bit; the table of fusions travels along (wgsl_contract_fusions.json).

    python tests/golden/make_wgsl_contract.py      # build container only (needs /root/reference); ~1 minute
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import wgsl_contract  # noqa: E402
from tests import wgsl_cases  # noqa: E402
from tests.golden.make_wgsl_golden import shader_digests  # noqa: E402


def main():
    arrays = {}
    for name, (omega, w, h, u0, script) in wgsl_cases.cases().items():
        t = time.time()
        sim = wgsl_contract.WgslLBMContract(omega, w, h, inflow_ux=u0)
        shots = wgsl_cases.replay(script, sim, lambda s: s.state())
        for i, st in enumerate(shots):
            for k in wgsl_cases.STATE_KEYS:
                arrays[f"{name}/{i}/{k}"] = st[k]
        arrays[f"{name}/count"] = np.int64(len(shots))
        print(f"{name}: {len(shots)} snapshots, {time.time() - t:.1f} s", flush=True)
    sim = wgsl_contract.WgslLBMContract(1.25, 24, 13)
    sim.draw_points(wgsl_cases.disc_pairs(24, 8, 6, 2).astype(np.uint32))
    sim.iterate(30)
    for stat in range(5):
        sim.compute_summary(stat)
        for cmap in range(3):
            arrays[f"colors/{stat}/{cmap}"] = sim.colors_of(cmap)
    dig = shader_digests()
    arrays["shader_files"] = np.array(sorted(dig))
    arrays["shader_sha256"] = np.array([dig[k] for k in sorted(dig)])
    out = os.path.join(ROOT, "tests", "golden", "wgsl_contract.npz")
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out), "bytes")
    table = {k: [list(p) for p in v] if isinstance(v, list) else v for k, v in wgsl_contract.fusion_table().items()}
    with open(os.path.join(ROOT, "tests", "golden", "wgsl_contract_fusions.json"), "w") as f:
        json.dump(table, f, indent=1)
    print("fusions:", {k: len(v) for k, v in table.items() if isinstance(v, list) and v})


if __name__ == "__main__":
    main()
