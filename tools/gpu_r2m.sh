#!/bin/bash
# round 2, call M (1 GPU): persistent blocks walking the live-block list (LBM_AA_LIVE_LIST=2) on the headline workload,
# A/B in one call against the default launch and the one-block-per-entry list (LBM_AA_LIVE_LIST=1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
R="timeout 90 python tools/perf_one.py inplace_obst 512 40"
export LBM_PROBE_FORCE=1
{
for rep in 1 2; do
  echo "== default (rep $rep)"; $R
  echo "== list=1 (rep $rep)"; LBM_AA_LIVE_LIST=1 $R
  echo "== list=2 persist 12/SM (rep $rep)"; LBM_AA_LIVE_LIST=2 $R
  echo "== list=2 persist 10/SM, 48 regs (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.pb10.so LBM_AA_LIVE_LIST=2 LBM_AA_PERSIST_BLOCKS=10 $R
  echo "== list=2 persist 6/SM (rep $rep)"; LBM_AA_LIVE_LIST=2 LBM_AA_PERSIST_BLOCKS=6 $R
  echo "== list=2 persist 24/SM (2 waves) (rep $rep)"; LBM_AA_LIVE_LIST=2 LBM_AA_PERSIST_BLOCKS=24 $R
  echo "== list=2 persist 48/SM (rep $rep)"; LBM_AA_LIVE_LIST=2 LBM_AA_PERSIST_BLOCKS=48 $R
done
} > $O/r2m_probe.log 2>&1
cat $O/r2m_probe.log
