#!/bin/bash
# GPU session V (1 GPU): final library (programmatic dependent launch on for stream-ordered steps, off inside graphs): whole GPU
# suite, smoke, bench lines of every single-GPU workload in the driver's shape, and the reference arm.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1100 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 | tee gpurun_out/r2v_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2v_smoke.log
for wl in cylinder512 cavity4096 channel16384; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r2v_bench_$wl.json 2>> gpurun_out/r2v.err
done
python bench.py --steps 20 --warmup 3 > gpurun_out/r2v_bench_default_steps20.json 2>> gpurun_out/r2v.err
python bench.py > gpurun_out/r2v_bench_default.json 2>> gpurun_out/r2v.err
tail -3 gpurun_out/r2v.err
exit 0
