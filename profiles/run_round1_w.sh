#!/bin/bash
# GPU session W (round 1): full GPU suite, default bench + clock trace, ncu --set full of the staged dense kernel
# (porous) and of the sparse kernel with 32-bit offsets (channel), launch list of the default run.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_w.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/w_default.json 2>gpurun_out/w.err
kill $SMI
cut -c1-1700 gpurun_out/w_default.json
timeout 300 python bench.py --workload channel16384 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/w_channel16384.json 2>>gpurun_out/w.err
cut -c1-300 gpurun_out/w_channel16384.json
for wl in porous16384 channel16384; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_vec4_kernel -s 20 -c 2 -o gpurun_out/prof_w_${wl} \
   python bench.py --workload $wl --steps 12 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/ncu_w_${wl}.log 2>&1
tail -1 gpurun_out/ncu_w_${wl}.log | cut -c1-200
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_w_default.csv \
   python bench.py --steps 30 --warmup 15 --no-cpu-baseline > gpurun_out/ncu_launches_w.log 2>&1
tail -3 gpurun_out/launches_w_default.csv | cut -c1-200
