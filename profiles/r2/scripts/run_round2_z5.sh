#!/bin/bash
# GPU session Z5 (1 GPU): negative control for the stream half-step part of the skew test: the library of commit 2c45fb6 (summary
# epoch only; built into profiles/r2/ab_prev, not kept) must fail at "skewed stream + collide".
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BLBM_LIBRARY=$PWD/profiles/r2/ab_prev/libblbm_prev.so timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf --tb=line -k "running_ahead" > gpurun_out/r2z5_negative_control_stream.log 2>&1
tail -12 gpurun_out/r2z5_negative_control_stream.log
exit 0
