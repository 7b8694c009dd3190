#!/bin/bash
# round 2, call A (1 GPU): sanity run of the GPU suite, config-1 error curve over 1000 steps, racecheck of a parity
# subset, and the 1024^3 single-GPU study (shapes + padded layouts + two ncu captures of the odd kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r2a_smi.log 2>&1
(timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/r2a_pytest.log
(timeout 400 python tools/config1_curve.py --steps 1000 --every 50 > $O/r2a_config1_curve.jsonl) 2> $O/r2a_config1_curve.err
(timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "(inplace_engine_with_obstacles and 19 and shape1) or bounce_back_obstacle or mass or unfused" 2>&1 | tail -15) > $O/r2a_racecheck.log
{
for shape in "896 896 896" "1024 1024 1024" "2048 1024 512" "512 1024 2048" "1024 512 2048" "128 1024 1024" "1024 1024 1008"; do
  timeout 120 python tools/box_probe.py $shape inplace 16
done
LBM_PAD_ROWS=6 timeout 120 python tools/box_probe.py 1024 1024 1024 inplace 16
LBM_PAD_LINES=3 timeout 120 python tools/box_probe.py 1024 1024 1024 inplace 16
LBM_PAD_ROWS=6 LBM_PAD_LINES=3 timeout 120 python tools/box_probe.py 1024 1024 1024 inplace 16
} > $O/r2a_box_probe.jsonl 2> $O/r2a_box_probe.err
for s in 896 1024; do
  timeout 200 ncu --set full --clock-control none -k regex:aa_odd_kernel --launch-skip 4 -c 1 -o $O/r2a_odd_$s -f \
     python tools/box_probe.py $s $s $s inplace 4 > $O/r2a_ncu_$s.log 2>&1
done
tail -3 $O/r2a_pytest.log; tail -4 $O/r2a_racecheck.log; cat $O/r2a_box_probe.jsonl; tail -3 $O/r2a_config1_curve.jsonl | cut -c1-600
