#!/bin/bash
# session 4, call G: momentum exchange fused into the special-word kernels of the single-buffer engine (no separate force kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inplace" 2>&1 | tail -5) > $O/s4g_pytest.log
{
echo "== in-place obstacle, force accumulator on"; LBM_PROBE_FORCE=1 timeout 100 python tools/perf_one.py inplace_obst 512 30
echo "== in-place obstacle, force accumulator off"; timeout 100 python tools/perf_one.py inplace_obst 512 30
} > $O/s4g_probe.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
LBM_PROBE_FORCE=1 timeout 120 ncu --metrics $M --clock-control none --launch-skip 12 -c 5 --csv --log-file $O/s4g_inplace_obst_launches.csv python tools/perf_one.py inplace_obst 512 4 > $O/s4g_ncu.log 2>&1
tail -2 $O/s4g_pytest.log; cat $O/s4g_probe.log; python tools/launch_table.py $O/s4g_inplace_obst_launches.csv
