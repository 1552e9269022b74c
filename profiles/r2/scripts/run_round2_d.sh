#!/bin/bash
# GPU session D of round 2 (2 GPUs): linked slabs on DISTINCT devices — IPC peers (one process per GPU), one group
# handle in one process (first hardware run of blbm_link_local across devices), multi_gpu_parity + timing in bench.py,
# handshake in-kernel vs wait/signal kernels, and the NVLink bytes of the fused halo stores (ncu on rank 0 only).
#   gpurun --gpus 2 --timeout 1500 -- bash profiles/run_round2_d.sh
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 python -m pytest tests/test_gpu_multiproc.py tests/test_cpp_host.py -m gpu -q -x ) 2>&1 | tail -8 | tee gpurun_out/r2d_multiproc.log
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "group_handle or halo_handshake or slab_group or group_of_one" ) 2>&1 | tail -8 | tee gpurun_out/r2d_group.log
$TR bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -c 2500 gpurun_out/r2d_bench_n2.json
$TR bench.py --gpus 2 --steps 100 --warmup 10 --link-in-kernel 0 --no-parity > gpurun_out/r2d_bench_n2_extern.json 2>> gpurun_out/r2d_bench_n2.err
python bench.py --single-process --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2d_bench_n2_single.json 2> gpurun_out/r2d_bench_n2_single.err
python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
for wl in cavity4096 cylinder512; do
  python bench.py --gpus 1 --workload $wl --steps 300 --warmup 30 --no-cpu-baseline --no-e2e --graphs 0 > gpurun_out/r2d_${wl}_n1.json 2>> gpurun_out/r2d_small.err
  $TR bench.py --gpus 2 --workload $wl --steps 300 --warmup 30 --no-e2e --no-parity > gpurun_out/r2d_${wl}_n2.json 2>> gpurun_out/r2d_small.err
  $TR bench.py --gpus 2 --workload $wl --steps 300 --warmup 30 --no-e2e --no-parity --link-in-kernel 0 > gpurun_out/r2d_${wl}_n2_extern.json 2>> gpurun_out/r2d_small.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 --no-python \
  profiles/ncu_rank0.sh r2d_nvlink --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum --clock-control none \
  -k regex:step_vec4 -s 8 -c 6 -- bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-parity > gpurun_out/r2d_nvlink.log 2>&1
tail -3 gpurun_out/*.err
exit 0
