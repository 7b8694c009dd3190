#!/bin/bash
# round 2, call B (1 GPU): the GPU suite incl. the peer-memory halo tests (slabs sharing the device, CUDA IPC between two
# processes on one GPU), the reworked bench (single-buffer headline), the 1024^3 look-ahead experiment, launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=5000
(timeout 300 python -m pytest tests/test_gpu_peer_halo.py -m gpu -x -q 2>&1 | tail -15) > $O/r2b_pytest_peer.log
(timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_peer_halo.py 2>&1 | tail -15) > $O/r2b_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > $O/r2b_smoke.log
(timeout 400 python tools/config1_curve.py --steps 1000 --every 100 > $O/r2b_config1_curve.jsonl) 2> $O/r2b_config1_curve.err
{
echo "== in-place obstacle, force on";  LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== in-place obstacle, force off"; timeout 60 python tools/perf_one.py inplace_obst 512 40
echo "== two-buffer headline (periodic x)"; LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py headline 512 40
echo "== in-place all-fluid"; timeout 60 python tools/perf_one.py inplace 512 40
} > $O/r2b_probe.log 2>&1
{
for a in 0 2 4 8 16; do LBM_TLB_AHEAD=$a timeout 120 python tools/box_probe.py 1024 1024 1024 inplace 16; done
LBM_TLB_AHEAD=4 LBM_TLB_MODE=2 timeout 120 python tools/box_probe.py 1024 1024 1024 inplace 16
LBM_TLB_AHEAD=8 LBM_TLB_MODE=2 timeout 120 python tools/box_probe.py 1024 1024 1024 inplace 16
LBM_TLB_AHEAD=4 timeout 120 python tools/box_probe.py 896 896 896 inplace 16
} > $O/r2b_tlb_probe.jsonl 2> $O/r2b_tlb_probe.err
(timeout 600 python bench.py --steps 20 --warmup 5 2>$O/r2b_bench.err | tail -1) > $O/r2b_bench.json
(timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>$O/r2b_bench_ref.err | tail -1) > $O/r2b_bench_ref.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 ncu --metrics $M --clock-control none -c 80 --csv --log-file $O/r2b_bench_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-extra > $O/r2b_ncu_bench.log 2>&1
tail -4 $O/r2b_pytest_peer.log; tail -4 $O/r2b_pytest.log; cat $O/r2b_smoke.log; cat $O/r2b_probe.log; cat $O/r2b_tlb_probe.jsonl; cut -c1-1500 $O/r2b_bench.json; echo; cut -c1-800 $O/r2b_bench_ref.json; tail -3 $O/r2b_bench.err $O/r2b_bench_ref.err
