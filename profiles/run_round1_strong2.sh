#!/bin/bash
# Strong scaling of BASELINE.json configs[3] (32768^2 cylinder wake) with the current kernels: N = 1 and N = 8 on
# the same 8-GPU box.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 10 --workload cylinder32768 --no-e2e --no-cpu-baseline > gpurun_out/strong2_1.json 2>gpurun_out/strong2_1.err
bash profiles/run_round1_strong.sh 8
python - <<'P'
import json
a=json.load(open("gpurun_out/strong2_1.json")); b=json.load(open("gpurun_out/strong_8.json"))
print("N=1", round(a["value"]), "MLUPS", round(a["ms_per_step"],4), "ms/step frac", round(a["roofline"]["frac"],4), a["clocks"])
print("N=8", round(b["value"]), "MLUPS", round(b["ms_per_step"],4), "ms/step; parallel efficiency", round(a["ms_per_step"]/(8*b["ms_per_step"]),4))
P
