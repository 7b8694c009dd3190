#!/bin/bash
# session 4, call E: lean z-edge path of the odd step (all-fluid and obstacle), row walkers at 10 / 12 blocks per SM, one row per warp
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inplace" 2>&1 | tail -5) > $O/s4e_pytest.log
{
echo "== default (row walkers, 10 blocks/SM)"; python tools/perf_one.py inplace_obst 512 30
echo "== variant a (12 blocks/SM)"; LBM_B200_LIB=$PWD/latticeboltzmanncuda_b200/liblbm_b200.a.so python tools/perf_one.py inplace_obst 512 30
echo "== variant b (one row per warp)"; LBM_B200_LIB=$PWD/latticeboltzmanncuda_b200/liblbm_b200.b.so python tools/perf_one.py inplace_obst 512 30
echo "== all-fluid"; python tools/perf_one.py inplace 512 30; python tools/perf_one.py inplace_walls 512 30
} > $O/s4e_probe.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --launch-skip 10 -c 4 --csv --log-file $O/s4e_inplace_obst_launches.csv python tools/perf_one.py inplace_obst 512 4 > $O/s4e_ncu.log 2>&1
tail -3 $O/s4e_pytest.log; cat $O/s4e_probe.log; python tools/launch_table.py $O/s4e_inplace_obst_launches.csv
