#!/bin/bash
# GPU session N (round 1): parity after the class-diff fix; default bench; e2e launch list.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/n_default.json 2>gpurun_out/n.err; tail -2 gpurun_out/n.err
python - <<P
import json
d=json.load(open("gpurun_out/n_default.json"))
print("default", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), d["clocks"], round(d["cpu_baseline"]["value"]))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_n_default.csv \
   python bench.py --steps 30 --warmup 15 --no-cpu-baseline > gpurun_out/ncu_launches_n.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 | cut -c1-330
