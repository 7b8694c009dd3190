#!/bin/bash
# session 4, call F: HEAD on a fresh B200 -- all GPU tests, smoke, both bench arms, launch list of the bench, the single-buffer
# engine on the obstacle (10 / 12 blocks per SM, per-kernel list, memcheck of one case), the 896^3 obstacle on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > $O/s4f_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/s4f_smoke.log
(timeout 200 python bench.py --impl reference --steps 10 --warmup 3 2>$O/s4f_bench_ref.err | tail -1) > $O/s4f_bench_ref.json
(timeout 300 python bench.py 2>$O/s4f_bench.err | tail -1) > $O/s4f_bench.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/s4f_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-extra > $O/s4f_ncu_bench.log 2>&1
{
echo "== in-place obstacle, default (12 blocks/SM)"; timeout 100 python tools/perf_one.py inplace_obst 512 30
echo "== variant a (10 blocks/SM)"; LBM_B200_LIB=$PWD/latticeboltzmanncuda_b200/liblbm_b200.a.so timeout 100 python tools/perf_one.py inplace_obst 512 30
} > $O/s4f_probe.log 2>&1
timeout 150 ncu --metrics $M --clock-control none --launch-skip 10 -c 4 --csv --log-file $O/s4f_inplace_obst_launches.csv python tools/perf_one.py inplace_obst 512 4 > $O/s4f_ncu_inplace.log 2>&1
(timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "inplace_engine_with_obstacles and 19 and shape1 and bc2" 2>&1 | tail -4) > $O/s4f_memcheck.log
(timeout 150 python tools/run_config.py obstacle --size 896 --engine inplace --steps 40 --warmup 6 2>&1 | tail -2) > $O/s4f_obstacle896.log
tail -2 $O/s4f_pytest.log; cat $O/s4f_smoke.log; cut -c1-400 $O/s4f_bench_ref.json; echo; cat $O/s4f_bench.json; echo; cat $O/s4f_probe.log
python tools/launch_table.py $O/s4f_inplace_obst_launches.csv; tail -2 $O/s4f_memcheck.log; cat $O/s4f_obstacle896.log
