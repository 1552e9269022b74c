#!/bin/bash
# GPU session T (1 GPU): one launch per small paint (mask + class rebuild), every even graph length, the moment-storing step
# inside the graph; contraction-twin library against the shader text executed under the derived contraction.  Whole GPU suite, the
# small-lattice frame breakdown and bench line A/B against the previous library (profiles/r2/ab_prev), the driver's default shape.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1100 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -8 | tee gpurun_out/r2t_gpu_suite.log
python profiles/frame_breakdown.py cylinder512 > gpurun_out/r2t_frame_breakdown_cylinder512.json 2> gpurun_out/r2t.err
cat gpurun_out/r2t_frame_breakdown_cylinder512.json
BLBM_LIBRARY=$PWD/profiles/r2/ab_prev/libblbm_prev.so python profiles/frame_breakdown.py cylinder512 > gpurun_out/r2t_frame_breakdown_cylinder512_prev.json 2>> gpurun_out/r2t.err
cat gpurun_out/r2t_frame_breakdown_cylinder512_prev.json
python bench.py --workload cylinder512 --no-cpu-baseline > gpurun_out/r2t_bench_cylinder512.json 2>> gpurun_out/r2t.err
BLBM_LIBRARY=$PWD/profiles/r2/ab_prev/libblbm_prev.so python bench.py --workload cylinder512 --no-cpu-baseline > gpurun_out/r2t_bench_cylinder512_prev.json 2>> gpurun_out/r2t.err
python bench.py --workload cavity4096 --no-cpu-baseline > gpurun_out/r2t_bench_cavity4096.json 2>> gpurun_out/r2t.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_default_steps20.json 2>> gpurun_out/r2t.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2t_smoke.log
tail -3 gpurun_out/r2t.err
exit 0
