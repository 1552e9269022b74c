#!/bin/bash
# GPU session AH (round 1): after the fill_equilibrium fix — the reset-heavy part of the suite and the WGSL pin.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_parity.py tests/test_wgsl_pin.py -m gpu -q --tb=short -x -k "fuzz_slab or random_scripts or porous_channel or single_cell or golden or wgsl or interpreted or create_state or fixed_point" > gpurun_out/ah.log 2>&1
tail -4 gpurun_out/ah.log
