#!/bin/bash
# GPU session Z2 (1 GPU): summary made part of the epoch protocol of linked slabs (an epoch is published after the summary
# kernel).  Whole GPU suite with the deterministic regression test, then the soak command that exposed the race (400 group walks,
# 4 xdist workers time-slicing the GPU), a bench line for sanity.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 | tee gpurun_out/r2z2_gpu_suite.log
( time BLBM_FUZZ_SLAB_SEEDS=20001-20400 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -n 4 -rf -k "api_fuzz_slab_group" ) 2>&1 | tail -8 | tee gpurun_out/r2z2_fuzz_group_400_xdist.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2z2_bench_default_steps20.json 2> gpurun_out/r2z2.err
tail -2 gpurun_out/r2z2.err
exit 0
