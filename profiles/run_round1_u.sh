#!/bin/bash
# GPU session U (round 1): dense flavour 2 (staged) vs 3 (staged + class words read without the chunk-flag test),
# porous and cylinder lattices; parity of flavour 3.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "staged or fuzz" 2>&1 | tail -3
for dn in 2 3 2 3; do
  timeout 300 python bench.py --workload porous16384 --dense $dn --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/u_porous16384_dense${dn}.json 2>>gpurun_out/u.err
  python - gpurun_out/u_porous16384_dense${dn}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
done
for dn in 0 2 3; do
  timeout 300 python bench.py --workload channel16384 --dense $dn --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/u_channel16384_dense${dn}.json 2>>gpurun_out/u.err
  python - gpurun_out/u_channel16384_dense${dn}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
done
