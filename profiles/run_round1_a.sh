#!/bin/bash
# GPU session A (round 1): parity tests, first bench lines, ncu launch list + full capture of the step kernel.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for wl in porous16384 channel16384; do
  for k in vec4 scalar; do
    timeout 600 python bench.py --workload $wl --kernel $k --no-cpu-baseline > gpurun_out/bench_${wl}_${k}.json 2> gpurun_out/bench_${wl}_${k}.err
    tail -c 2500 gpurun_out/bench_${wl}_${k}.json; tail -3 gpurun_out/bench_${wl}_${k}.err
  done
done
timeout 600 python bench.py --workload cavity4096 --kernel vec4 --no-cpu-baseline > gpurun_out/bench_cavity4096_vec4.json 2>&1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_vec4 -s 4 -c 2 -o gpurun_out/prof_vec4_porous \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
