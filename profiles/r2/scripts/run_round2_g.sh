#!/bin/bash
# GPU session G of round 2 (2 GPUs): the north_star block with its full-size parity check at N = 2, and another attempt
# at hardware evidence for the fused halo stores: L2 sectors addressed to the PEER aperture per step-kernel launch
# (ncu on rank 0 only).
#   gpurun --gpus 2 --timeout 900 -- bash profiles/run_round2_g.sh
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time $TR --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --north-star 1 ) > gpurun_out/r2g_bench_n2_ns.json 2> gpurun_out/r2g_bench_n2_ns.err
tail -c 3000 gpurun_out/r2g_bench_n2_ns.json
ncu --query-metrics 2>/dev/null | grep -i -E "peer|nvl" | head -60 > gpurun_out/r2g_metrics_peer.txt
$TR --master-port 29532 --no-python profiles/ncu_rank0.sh r2g_peer_sectors \
  --metrics lts__t_sectors_aperture_peer.sum,lts__t_sectors_aperture_peer_op_write.sum,lts__t_sectors_aperture_peer_op_read.sum,gpu__time_duration.sum \
  --clock-control none -k regex:step_vec4 -s 8 -c 6 -- bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-parity > gpurun_out/r2g_peer_sectors.log 2>&1
$TR --master-port 29533 --no-python profiles/ncu_rank0.sh r2g_time_only \
  --metrics gpu__time_duration.sum --clock-control none -k regex:step_vec4 -s 8 -c 6 -- bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-parity > gpurun_out/r2g_time_only.log 2>&1
tail -4 gpurun_out/r2g_peer_sectors.log gpurun_out/r2g_time_only.log gpurun_out/r2g_bench_n2_ns.err
exit 0
