#!/bin/bash
# round 2, call O (1 GPU): number of all-solid boxes (LBM_AA_DEAD_BOXES = 0 / 2 / 4), A/B in one call, and the launch list
# of the single-buffer step with 4 boxes (obstacle) and of the all-fluid periodic box for comparison
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
R="timeout 90 python tools/perf_one.py inplace_obst 512 40"
export LBM_PROBE_FORCE=1 LBM_VERBOSE=1
{
for rep in 1 2; do
  for K in 0 2 3 4; do echo "== $K boxes (rep $rep)"; LBM_AA_DEAD_BOXES=$K $R; done
done
} > $O/r2o_probe.log 2>&1
grep -v "^lbm: all-solid box\|^    force" $O/r2o_probe.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:aa_ -c 24 --csv --log-file $O/r2o_obstacle_launches.csv python tools/perf_one.py inplace_obst 512 4 > $O/r2o_ncu1.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:aa_ -c 12 --csv --log-file $O/r2o_allfluid_launches.csv python tools/perf_one.py inplace 512 4 > $O/r2o_ncu2.log 2>&1
tail -n 2 $O/r2o_ncu1.log $O/r2o_ncu2.log
