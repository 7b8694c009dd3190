#!/bin/bash
# session 2, call A: tests, probes of the new special / in-place kernels, short ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/s2a_gpu.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/s2a_pytest.log
{
echo "== default"; python tools/perf_probe.py 512 40 headline
echo "== general MINB 12"; LBM_GENERAL_MINB=12 python tools/perf_probe.py 512 40 headline
echo "== cellinfo not pinned"; LBM_PIN_CELLINFO=0 python tools/perf_probe.py 512 40 headline
echo "== special MINB 8"; LBM_SPECIAL_MINB=8 python tools/perf_probe.py 512 40 headline
echo "== not concurrent"; LBM_SPECIAL_CONCURRENT=0 python tools/perf_probe.py 512 40 headline
echo "== parts"; python tools/perf_probe.py 512 40 parts
echo "== parts MINB 12"; LBM_GENERAL_MINB=12 python tools/perf_probe.py 512 40 parts
echo "== inplace"; python tools/perf_probe.py 512 40 inplace
} > $O/s2a_probe.log 2>&1
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats"
timeout 600 ncu $SECT --clock-control none --import-source on -k regex:"step_special|step_general" --launch-skip 10 -c 2 -f -o $O/prof_s2a_headline python tools/perf_one.py headline 512 3 > $O/s2a_ncu_headline.log 2>&1
timeout 600 ncu $SECT --clock-control none --import-source on -k regex:"aa_" --launch-skip 9 -c 3 -f -o $O/prof_s2a_inplace python tools/perf_one.py inplace 512 4 > $O/s2a_ncu_inplace.log 2>&1
tail -3 $O/s2a_pytest.log; cat $O/s2a_probe.log
