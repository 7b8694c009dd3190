#!/bin/bash
# session 3, call A: validate HEAD on a fresh B200 -- GPU tests, both bench arms, launch list, one --set full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $O/s3a_smi.log 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/s3a_pytest.log
(timeout 400 python bench.py --impl reference --steps 10 --warmup 3 2>$O/s3a_bench_ref.err | tail -1) > $O/s3a_bench_ref.json
(timeout 600 python bench.py 2>$O/s3a_bench.err | tail -1) > $O/s3a_bench.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/s3a_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-extra > $O/s3a_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"step_special|step_general" --launch-skip 16 -c 3 -f -o $O/s3a_headline python tools/perf_one.py headline 512 3 > $O/s3a_ncu_full.log 2>&1
python tools/ncu_summary.py $O/s3a_headline.ncu-rep > $O/s3a_headline_summary.md 2>&1
{
echo "== probe headline/parts/fast/inplace"
python tools/perf_probe.py 512 40 headline
python tools/perf_probe.py 512 40 parts
python tools/perf_probe.py 512 40 fast
python tools/perf_probe.py 512 40 inplace
} > $O/s3a_probe.log 2>&1
tail -3 $O/s3a_pytest.log; cat $O/s3a_bench_ref.json; cat $O/s3a_bench.json; cat $O/s3a_probe.log
