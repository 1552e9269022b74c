#!/bin/bash
# GPU session R (round 1): packed fp32 adds (FADD2) in the collision — parity subset, same-box A/B against the
# scalar-add kernels on the porous and the empty-channel lattice, executed-instruction counts from ncu.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_scripts or fuzz or cylinder" 2>&1 | tail -3
for wl in porous16384 channel16384; do
  for pk in 0 1 0 1; do
    timeout 300 python bench.py --workload $wl --packed $pk --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
        > gpurun_out/r_${wl}_pk${pk}.json 2>>gpurun_out/r.err
    python - gpurun_out/r_${wl}_pk${pk}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
  done
done
for pk in 0 1; do
  timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
     --clock-control none -k regex:step_vec4 -s 12 -c 3 --csv --log-file gpurun_out/r_inst_porous_pk${pk}.csv \
     python bench.py --workload porous16384 --packed $pk --steps 6 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  tail -16 gpurun_out/r_inst_porous_pk${pk}.csv | cut -d, -f5,13-
done
