#!/usr/bin/env python
"""Where a small-lattice frame goes: device time per frame (timer_start/stop around 300 frames) of the pieces of the e2e
loop on the 512x256 cylinder (launch-bound: a step is ~2.5 us of kernel)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lbm_b200 import LBM  # noqa: E402

w, h, omega, u0, kind = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cylinder512"]
lbm = LBM(omega, w, h, inflow_ux=u0)
_, m = bench.mask_rows(kind, w, h, 0, h)
lbm.write_barrier_rows(0, m)
out = [torch.empty((h, w), dtype=torch.float32, pin_memory=True) for _ in range(2)]
stroke = np.zeros((64, 2), np.uint64)
stroke[:, 0] = [(h // 2 + j // 8) * w + (w // 2 + j % 8) for j in range(64)]
L = lbm._L
N = 300


def paint(fr):
    stroke[:, 1] = fr & 1 ^ 1
    assert L.blbm_draw_points64(lbm._h, stroke.ctypes.data, 64) == 0


def timed(name, body):
    for fr in range(20):
        body(fr)
    lbm.synchronize()
    lbm.timer_start()
    for fr in range(N):
        body(fr)
    ms = lbm.timer_stop()
    lbm.synchronize()
    res[name] = round(ms / N * 1e3, 2)


res = {}
timed("iterate15", lambda fr: lbm.iterate(15))
timed("iterate16", lambda fr: lbm.iterate(16))
timed("advance15_no_summary", lambda fr: lbm.advance(15))
timed("paint_only", paint)
timed("paint+iterate15", lambda fr: (paint(fr), lbm.iterate(15)))
timed("iterate15+readback", lambda fr: (lbm.iterate(15), lbm.read_output_async(out[fr & 1].data_ptr())))
timed("full_frame", lambda fr: (paint(fr), lbm.iterate(15), lbm.read_output_async(out[fr & 1].data_ptr())))
lbm.set_tuning(9, 0)  # steps launched plainly instead of as programmatic dependents of one another
timed("iterate15_nopdl", lambda fr: lbm.iterate(15))
timed("full_frame_nopdl", lambda fr: (paint(fr), lbm.iterate(15), lbm.read_output_async(out[fr & 1].data_ptr())))
lbm.set_tuning(9, -1)
lbm.set_tuning(5, 0)
timed("iterate15_nographs", lambda fr: lbm.iterate(15))
timed("full_frame_nographs", lambda fr: (paint(fr), lbm.iterate(15), lbm.read_output_async(out[fr & 1].data_ptr())))
print(json.dumps({"workload": sys.argv[1] if len(sys.argv) > 1 else "cylinder512", "us_per_frame": res}))
