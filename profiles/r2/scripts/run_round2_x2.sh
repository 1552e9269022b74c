#!/bin/bash
# GPU session X2 (2 GPUs): the library as committed on two devices: the group / linked-slab / multi-process tests with slabs on
# distinct GPUs, a 200-walk soak of the group fuzz, and the driver's N = 2 bench command (multi-GPU parity check inside) both as
# torchrun ranks and as one process over a group handle.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2x2_gpus.log
( time BLBM_FUZZ_SLAB_SEEDS=9001-9200 timeout 900 python -m pytest tests -m gpu -q -x -k "group or slab or multiproc or handshake or linked or peer" ) 2>&1 | tail -8 | tee gpurun_out/r2x2_gpu_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2x2_bench_n2.json 2> gpurun_out/r2x2.err
timeout 600 python bench.py --gpus 2 --single-process --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2x2_bench_n2_single_process.json 2>> gpurun_out/r2x2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload cavity4096 --strong --no-cpu-baseline > gpurun_out/r2x2_bench_n2_cavity4096_strong.json 2>> gpurun_out/r2x2.err
tail -3 gpurun_out/r2x2.err
exit 0
