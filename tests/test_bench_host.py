"""Host-side logic of bench.py that needs no GPU: the workload description both arms print, the mask generator, the
slab decomposition the multi-GPU parity check relies on."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_config_names_the_workload_only_and_says_how_l2_is_handled():
    for name in bench.WORKLOADS:
        for n in (1, 2, 8):
            c = bench.workload_config(name, n)
            assert set(c) == {"workload", "W", "H", "H_per_gpu", "omega", "u0", "mask", "l2"}
            assert c["H_per_gpu"] * n == c["H"]
            assert "L2" in c["l2"]
    assert bench.workload_config("cylinder32768", 8)["H"] == 32768  # strong: the lattice is fixed
    assert bench.workload_config("porous16384", 8)["H"] == 8 * 16384  # weak: per-GPU work is fixed


def test_mask_rows_windows_tile_the_global_mask():
    """every rank generates only its own window; windows of any decomposition must agree with the undivided mask"""
    w, h = 640, 96
    discs = [(200, 48, 9), (400, 24, 5)]
    for kind in ("porous", "cylinder", "box", "none"):
        _, full = bench.mask_rows(kind, w, h, 0, h, discs)
        assert full[0].all() and full[-1].all()
        for r0, r1 in ((-2, 26), (22, 50), (46, 98)):
            a, part = bench.mask_rows(kind, w, h, r0, r1, discs)
            assert a == max(r0, 0)
            np.testing.assert_array_equal(part, full[a:min(r1, h)])
    _, m = bench.mask_rows("porous", w, h, 0, h)
    frac = m[1:-1, 2:-1].mean()
    assert 0.12 < frac < 0.18 and not m[1:-1, :2].any() and not m[1:-1, -1].any()


def test_reference_arm_prints_the_same_config_as_the_gpu_arm():
    """`--impl reference` on the smallest workload (runs the oracle for a few steps): one JSON line, impl marked, the
    workload-only config of the GPU arm, zero copy bytes"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cylinder512",
                        "--steps", "5", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["value"] > 0
    assert line["config"] == bench.workload_config("cylinder512", 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
