#!/bin/bash
# GPU session C of round 2 (1 GPU): the ordered barrier-chain table (coalesced replay, settled entries, moments merged
# by slot rank in the moment-storing launch, rank-based evict): whole GPU suite, bench + launch list.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -25 | tee gpurun_out/r2c_gpu_suite.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
tail -c 1500 gpurun_out/r2c_bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2c_launches_default.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2c_launches_cylinder512.csv \
    python bench.py --workload cylinder512 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c512_under_ncu.log 2>&1
tail -3 gpurun_out/*.err
exit 0
