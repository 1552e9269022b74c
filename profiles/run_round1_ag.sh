#!/bin/bash
# GPU session AG (round 1): after the fill_equilibrium/chain-table fix — fuzz seeds 1-18 and soak seeds 1000-1040, then 1041-1320,
# single process, failures with their operation logs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BLBM_FUZZ_SEEDS=1041-1320 timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "fuzz_against" > gpurun_out/ag_fuzz.log 2>&1
tail -4 gpurun_out/ag_fuzz.log
