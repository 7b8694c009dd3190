#!/bin/bash
# round 2, call X (1 GPU): compute-sanitizer memcheck over the single-buffer tests at HEAD (box tests, peer halos in one process)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=60000
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer_halo.py -m gpu -q -x -k "(inplace_engine or compact_body or slab_group_equals) and not ipc" > $O/r2x_memcheck.log 2>&1
echo "exit $?" >> $O/r2x_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" $O/r2x_memcheck.log; tail -n 8 $O/r2x_memcheck.log
