#!/bin/bash
# round 2, call X2 (1 GPU): compute-sanitizer racecheck over the single-buffer parity tests at HEAD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "compact_body or inplace_engine_with_obstacles or inplace_engine_equals" > $O/r2x2_racecheck.log 2>&1
echo "exit $?" >> $O/r2x2_racecheck.log
tail -n 6 $O/r2x2_racecheck.log
