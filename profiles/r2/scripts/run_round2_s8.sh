#!/bin/bash
# GPU session S8 (8 GPUs): strong scaling where it hurts — the 4096^2 closed box (2.1 M cells per GPU at N = 8) and the
# 512 x 256 cylinder (16 k cells per GPU) split over 8 GPUs, halo handshake inside the kernel vs wait/signal kernels,
# with the N = 1 time of the same lattice from the same box.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
port=29560
for wl in cavity4096 cylinder512; do
  python bench.py --gpus 1 --workload $wl --steps 400 --warmup 40 --no-cpu-baseline --no-e2e --graphs 0 > gpurun_out/r2s8_${wl}_n1.json 2>> gpurun_out/r2s8.err
  ns="2 4 8"; [ $wl = cylinder512 ] && ns="8"
  for n in $ns; do
    for mode in 1 0; do
      port=$((port+1))
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --strong \
        --steps 400 --warmup 40 --no-e2e --no-parity --link-in-kernel $mode > gpurun_out/r2s8_${wl}_n${n}_link${mode}.json 2>> gpurun_out/r2s8.err
    done
  done
done
tail -3 gpurun_out/r2s8.err
exit 0
