#!/bin/bash
# GPU session F of round 2 (1 GPU): whole GPU suite on the final kernels (64-register moment-storing launch, graph LRU),
# bench lines of every single-GPU configuration, launch list, and ncu --set full of the three kernels of a frame.
#   gpurun --timeout 1800 -- bash profiles/run_round2_f.sh
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -12 | tee gpurun_out/r2f_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2f_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench_default_steps20.json 2> gpurun_out/r2f_bench.err
tail -c 1200 gpurun_out/r2f_bench_default_steps20.json
for wl in cylinder512 cavity4096 channel16384 porous16384; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r2f_bench_$wl.json 2>> gpurun_out/r2f_bench.err
done
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_bench_reference.json 2>> gpurun_out/r2f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches_default.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"step_vec4_kernel<0" -s 20 -c 1 -o gpurun_out/r2f_step_vec4 \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"step_vec4_kernel<1" -s 1 -c 1 -o gpurun_out/r2f_step_vec4_mom \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/r2f_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_replay_kernel -s 2 -c 1 -o gpurun_out/r2f_chain_replay \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/r2f_ncu_full.log 2>&1
for k in step_vec4 step_vec4_mom chain_replay; do
  python profiles/ncu_summary.py gpurun_out/r2f_$k.ncu-rep > gpurun_out/r2f_ncu_$k.txt 2>&1 || true
  rm -f gpurun_out/r2f_$k.ncu-rep
done
tail -3 gpurun_out/r2f_bench.err
exit 0
