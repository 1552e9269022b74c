#!/bin/bash
# GPU session G (round 1): TMA-staged kernel — parity, then launch-shape A/B against vec4.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
b() { # label args...
  lbl=$1; shift
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 100 "$@" > gpurun_out/g_$lbl.json 2>gpurun_out/g.err || tail -3 gpurun_out/g.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/g_$lbl.json"))
    print("$lbl", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$lbl FAILED", e)
P
}
for wl in porous16384 channel16384; do
  b ${wl}_vec4 --workload $wl --kernel vec4
  b ${wl}_tma_r4_s4_c2 --workload $wl --kernel tma --tma-rows 4 --tma-stages 4 --tma-ctas 2
  b ${wl}_tma_r4_s2_c4 --workload $wl --kernel tma --tma-rows 4 --tma-stages 2 --tma-ctas 4
  b ${wl}_tma_r4_s3_c3 --workload $wl --kernel tma --tma-rows 4 --tma-stages 3 --tma-ctas 3
  b ${wl}_tma_r8_s4_c1 --workload $wl --kernel tma --tma-rows 8 --tma-stages 4 --tma-ctas 1
  b ${wl}_tma_r8_s2_c2 --workload $wl --kernel tma --tma-rows 8 --tma-stages 2 --tma-ctas 2
done
