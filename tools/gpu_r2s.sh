#!/bin/bash
# round 2, call S (1 GPU): the box tests in the two-buffer general kernel, A/B in one call on the two-buffer headline
# (512^3 obstacle, zero-inflow x, walls y/z, plane source, force on) and on the all-fluid lattice with a source
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_PROBE_FORCE=1
{
for rep in 1 2 3; do
  echo "== two-buffer obstacle, no boxes (rep $rep)"; LBM_AA_DEAD_BOXES=0 timeout 90 python tools/perf_one.py headline 512 40
  echo "== two-buffer obstacle, boxes = default (rep $rep)"; timeout 90 python tools/perf_one.py headline 512 40
done
echo "== two-buffer all-fluid + source (general kernel), no boxes"; LBM_AA_DEAD_BOXES=0 timeout 90 python tools/perf_one.py tasks_plain 512 40
echo "== two-buffer all-fluid + source (general kernel), boxes"; timeout 90 python tools/perf_one.py tasks_plain 512 40
echo "== two-buffer all-fluid walls y/z (general kernel), no boxes"; LBM_AA_DEAD_BOXES=0 timeout 90 python tools/perf_one.py walls_yz 512 40
echo "== two-buffer all-fluid walls y/z (general kernel), boxes"; timeout 90 python tools/perf_one.py walls_yz 512 40
} > $O/r2s_probe.log 2>&1
grep -v "^    force" $O/r2s_probe.log; grep force $O/r2s_probe.log | cut -c1-90 | sort | uniq -c
