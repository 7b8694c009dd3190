#!/bin/bash
# round 2, call W (1 GPU): byte offsets in the two big single-buffer kernels (2 instead of 4 address instructions per
# access), A/B in one call against the previous commit's build (liblbm_b200.elem.so): obstacle headline and all-fluid
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
export LBM_PROBE_FORCE=1
{
for rep in 1 2 3; do
  echo "== element offsets (previous commit) (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.elem.so timeout 90 python tools/perf_one.py inplace_obst 512 40
  echo "== byte offsets (rep $rep)"; timeout 90 python tools/perf_one.py inplace_obst 512 40
done
for rep in 1 2; do
  echo "== all-fluid periodic, element offsets (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.elem.so timeout 90 python tools/perf_one.py inplace 512 40
  echo "== all-fluid periodic, byte offsets (rep $rep)"; timeout 90 python tools/perf_one.py inplace 512 40
  echo "== all-fluid walls, element offsets (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.elem.so timeout 90 python tools/perf_one.py inplace_walls 512 40
  echo "== all-fluid walls, byte offsets (rep $rep)"; timeout 90 python tools/perf_one.py inplace_walls 512 40
done
} > $O/r2w_probe.log 2>&1
grep -v "^    force" $O/r2w_probe.log; grep force $O/r2w_probe.log | cut -c1-90 | sort | uniq -c
