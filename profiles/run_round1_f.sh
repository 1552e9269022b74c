#!/bin/bash
# GPU session F (round 1): parity after partial class rebuild; default bench lines.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/f_default.json 2>gpurun_out/f.err; tail -2 gpurun_out/f.err
python - <<P
import json
d=json.load(open("gpurun_out/f_default.json"))
print("default", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), d["clocks"], d["cpu_baseline"]["value"])
P
for wl in channel16384 cavity4096 cylinder512; do
timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/f_$wl.json 2>gpurun_out/f.err; tail -2 gpurun_out/f.err
python - <<P
import json
d=json.load(open("gpurun_out/f_$wl.json"))
print("$wl", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), d["clocks"])
P
done
