#!/bin/bash
# GPU session B of round 2 (1 GPU): the whole GPU suite on the group-handle / staging-ring / reset-handshake changes,
# smoke, then the bench line of every single-GPU configuration and the launch list of the default one.
#   gpurun --timeout 1500 -- bash profiles/run_round2_b.sh
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -25 | tee gpurun_out/r2b_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2b_smoke.log
python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
tail -c 3000 gpurun_out/r2b_bench_default.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_default_steps20.json 2>> gpurun_out/r2b_bench_default.err
for wl in cylinder512 cavity4096 channel16384; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r2b_bench_$wl.json 2> gpurun_out/r2b_bench_$wl.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_default.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_under_ncu.log 2>&1
tail -5 gpurun_out/*.err
