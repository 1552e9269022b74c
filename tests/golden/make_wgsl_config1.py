"""Generates tests/golden/wgsl_config1.npz: BASELINE.json configs[0] — D2Q9 BGK channel flow past a cylinder,
512 x 256, 10,000 steps — computed by EXECUTING THE REFERENCE'S OWN WGSL TEXT (oracle/wgsl_simt.py: every dispatch
of /root/reference/lbm-wgpu/src/rewritten_shaders/*.wgsl evaluated for all 131,072 invocations, in lbm.rs's
dispatch order), for nu = 0.02 (Re 160, vortex shedding) and nu = 0.1 (Re 32).

Stored per viscosity: the sha256 of every buffer the reference keeps (2 x 9 populations, momentum x / y, density,
output = curl, barrier) after 1, 2, 100, 1000 and 10000 steps, and the full density / momentum / curl fields after
10000 steps (for a tolerance comparison should a bit ever differ).  The C oracle (CPU test) and the CUDA path (GPU
test) must reproduce these digests bit for bit; that pins the north star's "fields match the reference WGSL after
10k steps" to the shader source instead of to a restatement of it.  What it still cannot pin is what a particular
WebGPU backend does where WGSL leaves room (FMA contraction, out-of-range access): no wgpu run exists here.

    python tests/golden/make_wgsl_config1.py [nu ...]     # ~13 minutes per viscosity, one core each
"""
import hashlib
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.util import disc_pairs  # noqa: E402

W, H = 512, 256
CHECKPOINTS = (1, 2, 100, 1000, 10000)
VISCOSITIES = (0.02, 0.1)


def omega_from_viscosity(nu):
    f = np.float32
    return float(f(1.0) / (f(3.0) * f(nu) + f(0.5)))


def digest(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:  # canonicalise NaNs: their payload bits are not part of the contract
        a = a.copy()
        a[np.isnan(a)] = np.float32(np.nan)
    return hashlib.sha256(a.tobytes()).hexdigest()


def buffers(population, moments, output, barrier):
    """name -> array for everything the reference keeps in buffers; `population(buffer, k)` as oracle.population"""
    d = {f"f{b}_{k}": population(b, k) for b in (0, 1) for k in range(9)}
    d["mx"], d["my"], d["rho"] = moments
    d["out"] = output
    d["bar"] = np.asarray(barrier, dtype=np.uint32)
    return d


def tag(nu):
    return f"nu{nu:g}".replace(".", "p")


def run(nu):
    from oracle.wgsl_simt import WgslLBMVec
    sim = WgslLBMVec(omega_from_viscosity(nu), W, H)
    sim.draw_points(disc_pairs(W, 128, 128, 16).astype(np.uint32))
    out, done, t0 = {}, 0, time.time()
    for target in CHECKPOINTS:
        sim.iterate(target - done)  # n x compute_step, then calculate_summary (curl), lbm.rs:1065-1074
        done = target
        for name, a in buffers(sim.population, (sim.ux, sim.uy, sim.rho), sim.output, sim.barrier).items():
            out[f"{tag(nu)}_s{target}_{name}"] = np.bytes_(digest(a))
        print(f"nu={nu}: step {target} after {time.time() - t0:.0f} s", flush=True)
    for name, a in (("mx", sim.ux), ("my", sim.uy), ("rho", sim.rho), ("out", sim.output)):
        out[f"{tag(nu)}_final_{name}"] = a.reshape(H, W).copy()
    return out


def main():
    nus = [float(a) for a in sys.argv[1:]] or list(VISCOSITIES)
    path = os.path.join(HERE, "wgsl_config1.npz")
    merged = dict(np.load(path)) if os.path.exists(path) else {}
    with ProcessPoolExecutor(len(nus)) as ex:
        for res in ex.map(run, nus):
            merged.update(res)
    from tests.golden.make_wgsl_golden import shader_digests
    dig = shader_digests()
    merged["shader_files"] = np.array(sorted(dig))
    merged["shader_sha256"] = np.array([dig[k] for k in sorted(dig)])
    np.savez_compressed(path, **merged)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
