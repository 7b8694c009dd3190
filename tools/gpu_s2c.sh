#!/bin/bash
# session 2, call C: tests, probes after the z-edge rework, variants, one full bench line + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/s2c_pytest.log
P=latticeboltzmanncuda_b200
{
echo "== default"; python tools/perf_probe.py 512 40 headline
for v in r4 w2 w8m4 m8; do
  if [ -f $P/liblbm_b200.$v.so ]; then echo "== variant $v"; LBM_B200_LIB=$PWD/$P/liblbm_b200.$v.so python tools/perf_probe.py 512 40 headline; fi
done
echo "== parts"; python tools/perf_probe.py 512 40 parts
echo "== inplace"; python tools/perf_probe.py 512 40 inplace
} > $O/s2c_probe.log 2>&1
python bench.py --steps 100 --warmup 5 > $O/s2c_bench.json 2> $O/s2c_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $O/s2c_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-extra > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"aa_" --launch-skip 8 -c 4 --csv --log-file $O/s2c_launches_inplace.csv python tools/perf_one.py inplace 512 4 > /dev/null 2>&1
tail -3 $O/s2c_pytest.log; cat $O/s2c_probe.log; cut -c1-1500 $O/s2c_bench.json; grep -v "^==" $O/s2c_launches_inplace.csv | cut -d, -f5,13- | tail -12
