#!/bin/bash
# GPU session K (round 1): CUDA-graph replay of step chunks on small lattices.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
b() { lbl=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --steps 2000 --warmup 100 "$@" > gpurun_out/k_$lbl.json 2>gpurun_out/k.err || tail -3 gpurun_out/k.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/k_$lbl.json"))
    print("$lbl", round(d["value"]), "MLUPS us/step", round(d["ms_per_step"]*1e3,3), "e2e", d["e2e"] and round(d["e2e"]["value"]))
except Exception as e: print("$lbl FAILED", e)
P
}
b cyl512_graphs0 --workload cylinder512 --graphs 0
b cyl512_graphs1 --workload cylinder512 --graphs 1
b cyl512_graphs1_r1 --workload cylinder512 --graphs 1 --block-rows 1
b cyl512_graphs1_r2 --workload cylinder512 --graphs 1 --block-rows 2
