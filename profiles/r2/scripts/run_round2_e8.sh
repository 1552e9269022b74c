#!/bin/bash
# GPU session E8 of round 2 (8 GPUs): exactly what the driver runs at round end — the default bench line at N = 8 with
# multi_gpu_parity before timing and the north_star block (65536^2 channel, 32768^2 strong) after it — then the same
# lattice through ONE group handle in one process (blbm_create_group over 8 distinct devices).
#   gpurun --gpus 8 --timeout 900 -- bash profiles/run_round2_e8.sh
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e8_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
( time $TR bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2e8_bench_n8.json 2> gpurun_out/r2e8_bench_n8.err
tail -c 4000 gpurun_out/r2e8_bench_n8.json
( time python bench.py --single-process --gpus 8 --steps 60 --warmup 6 --no-e2e ) > gpurun_out/r2e8_bench_n8_single.json 2> gpurun_out/r2e8_bench_n8_single.err
tail -c 1500 gpurun_out/r2e8_bench_n8_single.json
# NVLink evidence for the fused halo stores: the NVML per-link data counters around 2000 steps (expected per interior
# GPU and step: 2 faces x (3 W floats + 2 scalars) = 393 KiB sent, the same received; moment rows once per call)
nvidia-smi nvlink -gt d > gpurun_out/r2e8_nvlink_before.txt 2>&1
$TR bench.py --gpus 8 --steps 2000 --warmup 10 --no-e2e --no-parity --north-star 0 > gpurun_out/r2e8_bench_n8_2000.json 2>> gpurun_out/r2e8_bench_n8.err
nvidia-smi nvlink -gt d > gpurun_out/r2e8_nvlink_after.txt 2>&1
tail -5 gpurun_out/r2e8_bench_n8.err gpurun_out/r2e8_bench_n8_single.err
exit 0
