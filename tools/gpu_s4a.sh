#!/bin/bash
# session 4, call A: the single-buffer engine with obstacles -- GPU tests, memcheck of one small case, throughput probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/s4a_pytest.log
(timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "inplace_engine_with_obstacles and 19 and shape1" 2>&1 | tail -25) > $O/s4a_memcheck.log
{
python tools/perf_probe.py 512 30 inplace-obstacle
} > $O/s4a_probe.log 2>&1
tail -4 $O/s4a_pytest.log; tail -6 $O/s4a_memcheck.log; cat $O/s4a_probe.log
