#!/bin/bash
# GPU session D (round 1): parity after the class re-encoding + row-chunk flags; block-shape A/B.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for wl in porous16384 channel16384; do
  for rows in 8 4 16; do
    timeout 600 python bench.py --workload $wl --block-rows $rows --no-cpu-baseline --no-e2e --steps 100 > gpurun_out/d_${wl}_r${rows}.json 2>gpurun_out/d.err
    python - <<P
import json
d=json.load(open("gpurun_out/d_${wl}_r${rows}.json"))
print("${wl} rows=${rows}", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
  done
done
timeout 600 python bench.py --workload porous16384 --no-cpu-baseline > gpurun_out/d_porous_e2e.json 2>gpurun_out/d.err; cut -c1-300 gpurun_out/d_porous_e2e.json; python -c "
import json; d=json.load(open('gpurun_out/d_porous_e2e.json')); print('e2e', d['e2e']['value'], 'value', d['value'])"
