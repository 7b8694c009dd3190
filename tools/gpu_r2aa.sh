#!/bin/bash
# round 2, call AA (1 GPU): special-word kernel at 9 blocks/SM (56 registers) against the shipped 8 (64 registers), A/B in one call
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
export LBM_PROBE_FORCE=1
{
for rep in 1 2 3; do
  echo "== 8 blocks/SM = default (rep $rep)"; timeout 90 python tools/perf_one.py inplace_obst 512 40
  echo "== 9 blocks/SM (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.spc9.so timeout 90 python tools/perf_one.py inplace_obst 512 40
done
} > $O/r2aa_probe.log 2>&1
grep -v "^    force" $O/r2aa_probe.log
