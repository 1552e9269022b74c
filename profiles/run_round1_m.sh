#!/bin/bash
# GPU session M (round 1): final ncu captures of the default step kernel (both flavours) + launch list.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for wl in porous16384 channel16384; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_vec4_kernel -s 20 -c 2 -o gpurun_out/prof_m_${wl} \
   python bench.py --workload $wl --steps 12 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/ncu_m_${wl}.log 2>&1
tail -1 gpurun_out/ncu_m_${wl}.log | cut -c1-200
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_m_default.csv \
   python bench.py --steps 30 --warmup 15 --no-cpu-baseline > gpurun_out/ncu_launches_m.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_m.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/m_default.json 2>gpurun_out/m.err
kill $SMI
cut -c1-400 gpurun_out/m_default.json
