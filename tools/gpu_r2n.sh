#!/bin/bash
# round 2, call N (1 GPU): all-solid boxes in the kernel parameters (LBM_AA_DEAD_BOXES, default on) and cache hints
# (-DLBM_AA_KEEP_L2=1, -DLBM_AA_STREAM=2) on the headline workload, A/B in one call
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
R="timeout 90 python tools/perf_one.py inplace_obst 512 40"
export LBM_PROBE_FORCE=1 LBM_VERBOSE=1
{
for rep in 1 2; do
  echo "== boxes off (rep $rep)"; LBM_AA_DEAD_BOXES=0 $R
  echo "== boxes on = default (rep $rep)"; $R
  echo "== boxes on, mask/class loads evict_last (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.keep.so $R
  echo "== boxes on, streaming stores (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.stst.so $R
  echo "== boxes on, both hints (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.keepstst.so $R
  echo "== boxes off, both hints (rep $rep)"; LBM_AA_DEAD_BOXES=0 LBM_B200_LIB=$P/liblbm_b200.keepstst.so $R
done
echo "== all-fluid periodic, default"; timeout 90 python tools/perf_one.py inplace 512 40
echo "== all-fluid periodic, streaming stores"; LBM_B200_LIB=$P/liblbm_b200.stst.so timeout 90 python tools/perf_one.py inplace 512 40
} > $O/r2n_probe.log 2>&1
cat $O/r2n_probe.log
