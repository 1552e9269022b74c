#!/bin/bash
# GPU session V (round 1): 2-D grid (no per-thread division) + 32-bit plane offsets in the vec4 kernel —
# parity, then same-box A/B of the offset width on the porous and the empty-channel lattice.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_wgsl_pin.py -m gpu -q -x 2>&1 | tail -3
for wl in porous16384 channel16384; do
  for ix in 0 1 0 1; do
    timeout 300 python bench.py --workload $wl --index32 $ix --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
        > gpurun_out/v_${wl}_ix${ix}.json 2>>gpurun_out/v.err
    python - gpurun_out/v_${wl}_ix${ix}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
  done
done
