#!/bin/bash
# GPU session C (round 1): ncu full captures of the step kernel on the porous and the empty channel.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for wl in porous16384 channel16384; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_vec4_kernel -s 20 -c 2 -o gpurun_out/prof_c_${wl} \
   python bench.py --workload $wl --steps 12 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c_${wl}.log 2>&1
tail -2 gpurun_out/ncu_c_${wl}.log | cut -c1-300
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c_porous.csv \
   python bench.py --steps 30 --warmup 15 --no-cpu-baseline > gpurun_out/ncu_launches_c.log 2>&1
ls -la gpurun_out | tail -8
