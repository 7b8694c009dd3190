#!/bin/bash
# session 2, call D: tests, final probe, full ncu capture of the two headline kernels, named configs on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/s2d_pytest.log
{
echo "== headline"; python tools/perf_probe.py 512 40 headline
echo "== cavity 256"; python tools/run_config.py cavity
echo "== taylor-green 256 two-buffer"; python tools/run_config.py taylor-green --size 256 --steps 400
echo "== taylor-green 256 in-place"; python tools/run_config.py taylor-green --size 256 --steps 400 --engine inplace
echo "== taylor-green 768 in-place"; python tools/run_config.py taylor-green --size 768 --steps 100 --engine inplace
echo "== channel 1024x256x256 one GPU"; python tools/run_config.py channel --size 256 --steps 100
} > $O/s2d_runs.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_special|step_general" --launch-skip 16 -c 3 -f -o $O/prof_r01c_headline python tools/perf_one.py headline 512 3 > $O/s2d_ncu.log 2>&1
tail -3 $O/s2d_pytest.log; cat $O/s2d_runs.log; tail -5 $O/s2d_ncu.log
