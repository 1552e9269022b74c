#!/bin/bash
# GPU session O (1 GPU): graph runs that start with a pending class swap (the step after a paint): whole GPU suite, a 300-walk soak, the
# small-lattice bench lines and the frame breakdown again.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 | tee gpurun_out/r2o_gpu_suite.log
( time BLBM_FUZZ_SEEDS=6001-6300 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "api_fuzz_against_oracle" ) 2>&1 | tail -5 | tee gpurun_out/r2o_fuzz.log
python profiles/frame_breakdown.py cylinder512 > gpurun_out/r2o_frame_breakdown_cylinder512.json 2> gpurun_out/r2l.err
cat gpurun_out/r2o_frame_breakdown_cylinder512.json
for wl in cylinder512 cavity4096; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r2o_bench_$wl.json 2>> gpurun_out/r2l.err
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_default_steps20.json 2>> gpurun_out/r2l.err
tail -3 gpurun_out/r2l.err
exit 0
