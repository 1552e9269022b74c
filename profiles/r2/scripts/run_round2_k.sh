#!/bin/bash
# GPU session K (1 GPU): where a small-lattice frame goes (profiles/frame_breakdown.py) + a bigger API-walk soak on the
# final library.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python profiles/frame_breakdown.py cylinder512 > gpurun_out/r2k_frame_breakdown_cylinder512.json 2> gpurun_out/r2k.err
cat gpurun_out/r2k_frame_breakdown_cylinder512.json
python profiles/frame_breakdown.py cavity4096 > gpurun_out/r2k_frame_breakdown_cavity4096.json 2>> gpurun_out/r2k.err
cat gpurun_out/r2k_frame_breakdown_cavity4096.json
( time BLBM_FUZZ_SEEDS=4001-5000 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "api_fuzz_against_oracle" ) 2>&1 | tail -6 | tee gpurun_out/r2k_fuzz_single_1000.log
tail -3 gpurun_out/r2k.err
exit 0
