#!/bin/bash
# GPU session O (round 1): ncu --set full of the TMA-staged kernel (why is it slower than vec4?).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_tma_kernel -s 20 -c 1 -o gpurun_out/prof_o_tma \
   python bench.py --workload channel16384 --kernel tma --steps 12 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/ncu_o.log 2>&1
tail -1 gpurun_out/ncu_o.log | cut -c1-200
