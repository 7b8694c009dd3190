#!/bin/bash
# round 2, call P (1 GPU): no class byte outside the bounding box of the words that are not ordinary (LBM_AA_BUSY_BOX,
# default on), A/B in one call on the headline workload; the single-buffer GPU tests; one bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
R="timeout 90 python tools/perf_one.py inplace_obst 512 40"
export LBM_PROBE_FORCE=1
{
for rep in 1 2 3; do
  echo "== no boxes at all (rep $rep)"; LBM_AA_DEAD_BOXES=0 $R
  echo "== all-solid boxes only (rep $rep)"; LBM_AA_BUSY_BOX=0 $R
  echo "== all-solid boxes + bounding box = default (rep $rep)"; $R
done
} > $O/r2p_probe.log 2>&1
grep -v "^    force" $O/r2p_probe.log; grep force $O/r2p_probe.log | cut -c1-80 | sort | uniq -c
timeout 900 python -m pytest tests -m gpu -x -q -k "inplace or compact or slab or peer" 2>&1 | tail -n 3
timeout 600 python bench.py > $O/r2p_bench.json 2> $O/r2p_bench.err; cut -c1-900 $O/r2p_bench.json; tail -n 2 $O/r2p_bench.err
