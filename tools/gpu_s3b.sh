#!/bin/bash
# session 3, call B: L2 prefetch of the next row (general kernel) and look-ahead over list entries (special kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
{
echo "== default (GEN_PREFETCH=1)"; python tools/perf_probe.py 512 40 headline
for v in a b c d e; do
  if [ -f $P/liblbm_b200.$v.so ]; then echo "== variant $v"; LBM_B200_LIB=$P/liblbm_b200.$v.so python tools/perf_probe.py 512 40 headline; fi
done
echo "== parts default"; python tools/perf_probe.py 512 40 parts
} > $O/s3b_probe.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3) > $O/s3b_pytest_default.log
(LBM_B200_LIB=$P/liblbm_b200.b.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3) > $O/s3b_pytest_b.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for v in so b.so; do
LBM_B200_LIB=$P/liblbm_b200.$v timeout 300 ncu --metrics $M --clock-control none -k regex:"step_special|step_general" --launch-skip 16 -c 3 --csv --log-file $O/s3b_ncu_$v.csv python tools/perf_one.py headline 512 3 > /dev/null 2>&1
echo "== ncu $v"; grep -v "^==" $O/s3b_ncu_$v.csv | awk -F'","' 'NR>1{print substr($5,1,60), $(NF-2), $NF}' | tr -d '"'
done >> $O/s3b_probe.log
cat $O/s3b_pytest_default.log $O/s3b_pytest_b.log; cat $O/s3b_probe.log
