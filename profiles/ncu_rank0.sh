#!/bin/bash
# torchrun entry for a multi-rank run in which ONLY rank 0 runs under ncu (ncu replays kernels and must never wrap a
# whole multi-rank command): usage  torchrun ... --no-python profiles/ncu_rank0.sh OUT_BASENAME NCU_ARGS -- bench.py args
out=$1; shift
ncu_args=()
while [ "$1" != "--" ]; do ncu_args+=("$1"); shift; done
shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu "${ncu_args[@]}" --csv --log-file "gpurun_out/${out}.csv" python "$@"
else
  exec python "$@"
fi
