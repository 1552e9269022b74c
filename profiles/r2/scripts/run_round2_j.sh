#!/bin/bash
# GPU session J of round 2 (2 GPUs): soak of the random API walks — 300 single-handle walks (half of them on a porous
# mask: ordered chain table, settled entries, evicts, omega changes, graph runs) and 150 walks over group handles of
# 2-4 slabs with the first two slabs on distinct devices — plus the preset tests on the GPU.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time BLBM_FUZZ_SEEDS=2001-2300 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "api_fuzz_against_oracle" ) 2>&1 | tail -8 | tee gpurun_out/r2j_fuzz_single.log
( time BLBM_FUZZ_SLAB_SEEDS=3001-3150 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "api_fuzz_slab_group" ) 2>&1 | tail -8 | tee gpurun_out/r2j_fuzz_group.log
( time timeout 600 python -m pytest tests/test_wasm_pin.py tests/test_shapes.py -m gpu -q -x ) 2>&1 | tail -6 | tee gpurun_out/r2j_wasm_pin.log
exit 0
