#!/bin/bash
# GPU session S (round 1): chain replay on a side stream (fork/join around the step launches) — full GPU suite,
# default bench (value, e2e), launch list of a short run to see the overlap.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/s_default.json 2>gpurun_out/s.err; tail -2 gpurun_out/s.err
python - <<'P'
import json
d=json.load(open("gpurun_out/s_default.json"))
print("default", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), d["clocks"], "launches", d["gpu_launches"])
P
timeout 600 python bench.py --frame-steps 30 --no-cpu-baseline > gpurun_out/s_default_fs30.json 2>>gpurun_out/s.err
python - <<'P'
import json
d=json.load(open("gpurun_out/s_default_fs30.json"))
print("fs30", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), d["clocks"])
P
