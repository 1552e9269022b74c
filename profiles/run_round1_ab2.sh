cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck2.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slab_group_on_one_device and 1-2-size0" 2>&1 | tail -40
tail -5 gpurun_out/memcheck2.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "slab_group_on_one_device and 1-2-size0" 2>&1 | tail -3
