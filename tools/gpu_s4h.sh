#!/bin/bash
# session 4, call H: special-word kernels instantiated with and without the force sum; one even launch when the force is off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inplace and (obstacle or refuses)" 2>&1 | tail -4) > $O/s4h_pytest.log
{
echo "== in-place obstacle, force accumulator on"; LBM_PROBE_FORCE=1 timeout 40 python tools/perf_one.py inplace_obst 512 30
echo "== in-place obstacle, force accumulator off"; timeout 40 python tools/perf_one.py inplace_obst 512 30
} > $O/s4h_probe.log 2>&1
tail -2 $O/s4h_pytest.log; cat $O/s4h_probe.log
