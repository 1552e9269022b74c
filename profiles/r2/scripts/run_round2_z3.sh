#!/bin/bash
# GPU session Z3 (1 GPU): negative control - the regression test of the summary / next-push race against the library BEFORE the
# fix (built from commit 69abad7 into profiles/r2/ab_prev, not kept): it must fail there.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BLBM_LIBRARY=$PWD/profiles/r2/ab_prev/libblbm_prev.so timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --tb=line -k "running_ahead" > gpurun_out/r2z3_negative_control_prev_library.log 2>&1
tail -12 gpurun_out/r2z3_negative_control_prev_library.log
exit 0
