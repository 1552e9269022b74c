#!/bin/bash
# GPU session U (1 GPU): fused vec4 steps launched as programmatic dependents of one another (griddepcontrol.wait / launch_dependents;
# auto on small lattices).  Whole GPU suite, frame breakdown with the knob on / off, bench lines of the small and the default workload
# with the knob on / off, a fuzz soak and compute-sanitizer memcheck over the new launch paths.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1100 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 | tee gpurun_out/r2u_gpu_suite.log
python profiles/frame_breakdown.py cylinder512 > gpurun_out/r2u_frame_breakdown_cylinder512.json 2> gpurun_out/r2u.err
cat gpurun_out/r2u_frame_breakdown_cylinder512.json
for pdl in -1 0; do
  python bench.py --workload cylinder512 --no-cpu-baseline --pdl $pdl > gpurun_out/r2u_bench_cylinder512_pdl$pdl.json 2>> gpurun_out/r2u.err
done
python bench.py --workload cavity4096 --no-cpu-baseline --pdl 0 > gpurun_out/r2u_bench_cavity4096_pdl0.json 2>> gpurun_out/r2u.err
python bench.py --workload cavity4096 --no-cpu-baseline --pdl 1 > gpurun_out/r2u_bench_cavity4096_pdl1.json 2>> gpurun_out/r2u.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --pdl 1 > gpurun_out/r2u_bench_default_steps20_pdl1.json 2>> gpurun_out/r2u.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_default_steps20.json 2>> gpurun_out/r2u.err
( time BLBM_FUZZ_SEEDS=7001-7300 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "api_fuzz_against_oracle" ) 2>&1 | tail -5 | tee gpurun_out/r2u_fuzz.log
( time timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph_run_length and 150" ) 2>&1 | tail -8 | tee gpurun_out/r2u_compute_sanitizer_memcheck.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2u_smoke.log
tail -3 gpurun_out/r2u.err
exit 0
