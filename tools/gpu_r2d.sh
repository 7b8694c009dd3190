#!/bin/bash
# round 2, call D (1 GPU): whole GPU suite at HEAD; single-buffer obstacle step with and without the live-block list
# (same box, A/B through LBM_AA_LIVE_LIST); bench; per-kernel list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=5000
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > $O/r2d_pytest.log
{
for rep in 1 2; do
echo "== live-block list ON, force on";  LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== live-block list OFF, force on"; LBM_AA_LIVE_LIST=0 LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
done
echo "== live-block list ON, force off";  timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== live-block list OFF, force off"; LBM_AA_LIVE_LIST=0 timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== two-buffer headline"; LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py headline 512 40
echo "== in-place all-fluid"; timeout 60 python tools/perf_one.py inplace 512 40
} > $O/r2d_probe.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
LBM_PROBE_FORCE=1 timeout 200 ncu --metrics $M --clock-control none --launch-skip 30 -c 12 --csv --log-file $O/r2d_inplace_obst_launches.csv python tools/perf_one.py inplace_obst 512 6 > $O/r2d_ncu.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5 2>$O/r2d_bench.err | tail -1) > $O/r2d_bench.json
tail -6 $O/r2d_pytest.log; cat $O/r2d_probe.log; python tools/launch_table.py $O/r2d_inplace_obst_launches.csv 2>/dev/null | tail -14; cut -c1-1200 $O/r2d_bench.json; tail -3 $O/r2d_bench.err
