#!/bin/bash
# session 2, call B: tests + probes of the reworked special / in-place kernels and launch-shape variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/s2b_pytest.log
P=latticeboltzmanncuda_b200
{
echo "== default"; python tools/perf_probe.py 512 40 headline
echo "== cls not pinned"; LBM_PIN_CLS=0 python tools/perf_probe.py 512 40 headline
echo "== not concurrent"; LBM_SPECIAL_CONCURRENT=0 python tools/perf_probe.py 512 40 headline
for v in w8 w16 w16b3 m8 m6 r16 s8; do
  if [ -f $P/liblbm_b200.$v.so ]; then echo "== variant $v"; LBM_B200_LIB=$PWD/$P/liblbm_b200.$v.so python tools/perf_probe.py 512 40 headline; fi
done
echo "== parts"; python tools/perf_probe.py 512 40 parts
echo "== fast"; python tools/perf_probe.py 512 40 fast
echo "== inplace"; python tools/perf_probe.py 512 40 inplace
} > $O/s2b_probe.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"step_|aa_" --launch-skip 10 -c 4 --csv --log-file $O/s2b_launches_headline.csv python tools/perf_one.py headline 512 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"aa_" --launch-skip 8 -c 4 --csv --log-file $O/s2b_launches_inplace.csv python tools/perf_one.py inplace 512 4 > /dev/null 2>&1
tail -3 $O/s2b_pytest.log; cat $O/s2b_probe.log; grep -v "^==" $O/s2b_launches_headline.csv | cut -d, -f5,13- | tail -14; grep -v "^==" $O/s2b_launches_inplace.csv | cut -d, -f5,13- | tail -14
