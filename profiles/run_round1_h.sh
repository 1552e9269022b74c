#!/bin/bash
# GPU session H (round 1): TMA kernel after class-word software pipelining.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "Tma or tma or 3" 2>&1 | tail -4
b() { # label args...
  lbl=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 100 "$@" > gpurun_out/h_$lbl.json 2>gpurun_out/h.err || tail -3 gpurun_out/h.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/h_$lbl.json"))
    print("$lbl", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$lbl FAILED", e)
P
}
for wl in porous16384 channel16384; do
  b ${wl}_vec4 --workload $wl --kernel vec4
  b ${wl}_tma_r4_s4_c2 --workload $wl --kernel tma --tma-rows 4 --tma-stages 4 --tma-ctas 2
  b ${wl}_tma_r4_s2_c5 --workload $wl --kernel tma --tma-rows 4 --tma-stages 2 --tma-ctas 5
  b ${wl}_tma_r4_s3_c3 --workload $wl --kernel tma --tma-rows 4 --tma-stages 3 --tma-ctas 3
  b ${wl}_tma_r8_s2_c2 --workload $wl --kernel tma --tma-rows 8 --tma-stages 2 --tma-ctas 2
done
