#!/bin/bash
# GPU session X (round 1): issue-order rework (chunk flag consumed last, warp-edge floats requested with the pulls)
# and the eager class-word flavour 3 again on top of it — parity, then same-box A/B.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_wgsl_pin.py -m gpu -q -x 2>&1 | tail -3
for dn in 2 3 2 3; do
  timeout 300 python bench.py --workload porous16384 --dense $dn --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/x_porous16384_dense${dn}.json 2>>gpurun_out/x.err
  python - gpurun_out/x_porous16384_dense${dn}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
done
for ix in 0 1; do
  timeout 300 python bench.py --workload porous16384 --dense 3 --index32 $ix --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/x_porous16384_dense3_ix${ix}.json 2>>gpurun_out/x.err
  python - gpurun_out/x_porous16384_dense3_ix${ix}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
done
for ix in 0 1; do
  timeout 300 python bench.py --workload channel16384 --index32 $ix --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/x_channel16384_ix${ix}.json 2>>gpurun_out/x.err
  python - gpurun_out/x_channel16384_ix${ix}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
done
