#!/bin/bash
# GPU session I (round 1): vec4 bounce-path variants.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "porous or random or cylinder" 2>&1 | tail -3
b() { lbl=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 150 "$@" > gpurun_out/i_$lbl.json 2>gpurun_out/i.err || tail -3 gpurun_out/i.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/i_$lbl.json"))
    print("$lbl", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$lbl FAILED", e)
P
}
for rep in 1 2; do
for wl in porous16384 channel16384; do
  b ${wl}_vec4_r4_$rep --workload $wl --kernel vec4 --block-rows 4
  b ${wl}_vec4_r2_$rep --workload $wl --kernel vec4 --block-rows 2
done; done
