#!/bin/bash
# session 2, call G: host-precomputed offsets; register / occupancy variants again
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/s2g_pytest.log
P=latticeboltzmanncuda_b200
{
echo "== default"; python tools/perf_probe.py 512 40 headline
for v in a b c d; do
  if [ -f $P/liblbm_b200.$v.so ]; then echo "== variant $v"; LBM_B200_LIB=$PWD/$P/liblbm_b200.$v.so python tools/perf_probe.py 512 40 headline; fi
done
echo "== parts"; python tools/perf_probe.py 512 40 parts
echo "== parts variant a"; LBM_B200_LIB=$PWD/$P/liblbm_b200.a.so python tools/perf_probe.py 512 40 parts
echo "== fast"; python tools/perf_probe.py 512 40 fast
echo "== inplace"; python tools/perf_probe.py 512 40 inplace
} > $O/s2g_probe.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for c in walls_y walls_yz; do
ncu --metrics $M --clock-control none -k regex:"step_general" --launch-skip 6 -c 1 --csv --log-file $O/s2g_ncu_$c.csv python tools/perf_one.py $c 512 3 > /dev/null 2>&1
echo "== ncu $c"; grep -v "^==" $O/s2g_ncu_$c.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
done >> $O/s2g_probe.log
tail -3 $O/s2g_pytest.log; cat $O/s2g_probe.log
