#!/bin/bash
# session 4, call C: row-walking obstacle kernels of the single-buffer engine -- tests, throughput at 8 / 10 / 12 blocks per SM, per-kernel list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inplace" 2>&1 | tail -5) > $O/s4c_pytest.log
{
echo "== default (8 blocks/SM)"; python tools/perf_one.py inplace_obst 512 30
for v in a b; do echo "== variant $v"; LBM_B200_LIB=$PWD/latticeboltzmanncuda_b200/liblbm_b200.$v.so python tools/perf_one.py inplace_obst 512 30; done
} > $O/s4c_probe.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --launch-skip 10 -c 8 --csv --log-file $O/s4c_inplace_obst_launches.csv python tools/perf_one.py inplace_obst 512 4 > $O/s4c_ncu.log 2>&1
tail -3 $O/s4c_pytest.log; cat $O/s4c_probe.log; python tools/launch_table.py $O/s4c_inplace_obst_launches.csv
