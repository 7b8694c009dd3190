#!/bin/bash
# round 2, call C (1 GPU): peer-halo tests after the rounding fix (incl. reference mode on slabs), the single-buffer
# obstacle step with the force summed ahead of time / concurrent face launches (A/B through environment switches),
# 1024^3 as 2 and 4 slabs on one GPU, per-kernel list of the single-buffer obstacle step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=5000
(timeout 400 python -m pytest tests/test_gpu_peer_halo.py -m gpu -q 2>&1 | tail -25) > $O/r2c_pytest_peer.log
(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_io.py tests/test_gpu_compat.py -m gpu -q 2>&1 | tail -8) > $O/r2c_pytest.log
{
echo "== default (force ahead of time, faces + special on the side stream), force on";  LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== default, force off"; timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== LBM_AA_CONCURRENT=0, force on"; LBM_AA_CONCURRENT=0 LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== LBM_AA_FORCE_SPLIT=1, force on"; LBM_AA_FORCE_SPLIT=1 LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== LBM_AA_CONCURRENT=0 LBM_AA_FORCE_SPLIT=1 (round 1 behaviour), force on"; LBM_AA_CONCURRENT=0 LBM_AA_FORCE_SPLIT=1 LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== in-place all-fluid"; timeout 60 python tools/perf_one.py inplace 512 40
echo "== in-place walls"; timeout 60 python tools/perf_one.py inplace_walls 512 40
} > $O/r2c_probe.log 2>&1
{
timeout 120 python tools/box_probe.py 128 1024 1024 inplace 40
timeout 160 python tools/box_probe.py 1024 1024 1024 inplace 16 2
timeout 160 python tools/box_probe.py 1024 1024 1024 inplace 16 4
timeout 160 python tools/box_probe.py 1024 1024 1024 inplace 16 8
timeout 160 python tools/box_probe.py 512 1024 1024 inplace 16 1
timeout 160 python tools/box_probe.py 896 896 896 inplace 16 2
} > $O/r2c_slab_probe.jsonl 2> $O/r2c_slab_probe.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
LBM_PROBE_FORCE=1 timeout 200 ncu --metrics $M --clock-control none --launch-skip 30 -c 16 --csv --log-file $O/r2c_inplace_obst_launches.csv python tools/perf_one.py inplace_obst 512 6 > $O/r2c_ncu.log 2>&1
tail -12 $O/r2c_pytest_peer.log; tail -3 $O/r2c_pytest.log; cat $O/r2c_probe.log; cat $O/r2c_slab_probe.jsonl; tail -3 $O/r2c_slab_probe.err
python tools/launch_table.py $O/r2c_inplace_obst_launches.csv 2>/dev/null | tail -20
