#!/bin/bash
# GPU session J (round 1): same-box A/B of the bounce-back fix-up (branchy vs branch-free build).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
b() { lbl=$1; lib=$2; shift; shift
  BLBM_LIBRARY=$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 150 "$@" > gpurun_out/j_$lbl.json 2>gpurun_out/j.err || tail -3 gpurun_out/j.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/j_$lbl.json"))
    print("$lbl", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$lbl FAILED", e)
P
}
A=$PWD/lbm_b200/libblbm.so; B=$PWD/lbm_b200/variants/libblbm_branchy64.so
for rep in 1 2; do
for wl in porous16384 channel16384; do
  b ${wl}_bf2batch64_$rep $A --workload $wl
  b ${wl}_branchy64_$rep $B --workload $wl
done; done
