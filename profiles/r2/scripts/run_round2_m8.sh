#!/bin/bash
# GPU session M8 (8 GPUs): final validation of exactly what the driver runs at N = 8 — multi_gpu_parity (400 steps),
# the headline workload, e2e, north_star with its full-size parity check.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
( time $TR bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2m8_bench_n8.json 2> gpurun_out/r2m8_bench_n8.err
tail -c 4500 gpurun_out/r2m8_bench_n8.json
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 3 ) > gpurun_out/r2m8_bench_n4.json 2> gpurun_out/r2m8_bench_n4.err
tail -c 1500 gpurun_out/r2m8_bench_n4.json
tail -6 gpurun_out/r2m8_bench_n8.err gpurun_out/r2m8_bench_n4.err
exit 0
