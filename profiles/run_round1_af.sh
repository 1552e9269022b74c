#!/bin/bash
# GPU session AF (round 1): the porous fuzz seeds 13-18, single process, with the failing operations logged.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "fuzz_against and (13 or 14 or 15 or 16 or 17 or 18)" > gpurun_out/af_fuzz.log 2>&1
tail -5 gpurun_out/af_fuzz.log
