#!/bin/bash
# GPU session L (round 1): full -m gpu suite incl. colour maps and shapes; strong-scaling workload smoke at N=1.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 python bench.py --workload cylinder32768 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/l_cyl32768.json 2>gpurun_out/l.err; tail -2 gpurun_out/l.err
python - <<P
import json
d=json.load(open("gpurun_out/l_cyl32768.json"))
print("cylinder32768 N=1", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms", round(d["ms_per_step"],3), d["scaling"], d["clocks"])
P
