#!/bin/bash
# GPU session H of round 2 (1 GPU): pre-primed CUDA-graph runs (16/8/4/2) — whole GPU suite, the small-lattice bench
# lines — and ncu --set full of the step kernel (both instantiations of a frame) for profiles/r2 and traffic.json.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 | tee gpurun_out/r2h_gpu_suite.log
for wl in cylinder512 cavity4096; do
  python bench.py --workload $wl --no-cpu-baseline > gpurun_out/r2h_bench_$wl.json 2>> gpurun_out/r2h_bench.err
  python bench.py --workload $wl --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r2h_bench_${wl}_steps20.json 2>> gpurun_out/r2h_bench.err
done
python bench.py --workload cylinder512 --no-cpu-baseline --graphs 0 > gpurun_out/r2h_bench_cylinder512_nographs.json 2>> gpurun_out/r2h_bench.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"step_vec4_kernel<\(bool\)0" -s 20 -c 1 -o gpurun_out/r2h_step_vec4 \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"step_vec4_kernel<\(bool\)1" -s 1 -c 1 -o gpurun_out/r2h_step_vec4_mom \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/r2h_ncu_full.log 2>&1
for k in step_vec4 step_vec4_mom; do
  python profiles/ncu_summary.py gpurun_out/r2h_$k.ncu-rep > gpurun_out/r2h_ncu_$k.txt 2>&1 || true
  ncu -i gpurun_out/r2h_$k.ncu-rep --page source --csv > gpurun_out/r2h_${k}_source.csv 2>/dev/null || true
  rm -f gpurun_out/r2h_$k.ncu-rep
done
tail -3 gpurun_out/r2h_bench.err gpurun_out/r2h_ncu_full.log
exit 0
