#!/bin/bash
# GPU session E (round 1): block-shape A/B, small blocks.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for wl in porous16384 channel16384 cavity4096; do
  for rows in 4 2 1; do
    timeout 600 python bench.py --workload $wl --block-rows $rows --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/e_${wl}_r${rows}.json 2>gpurun_out/e.err
    python - <<P
import json
d=json.load(open("gpurun_out/e_${wl}_r${rows}.json"))
print("${wl} rows=${rows}", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
  done
done
