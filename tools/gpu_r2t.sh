#!/bin/bash
# round 2, call T (1 GPU): the multi-GPU bench's slab (128 x 1024 x 1024, four bodies) on one GPU: 0 / 4 / 8 all-solid boxes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
R="timeout 200 python bench.py --workload slab --no-e2e --no-cpu --no-extra --steps 40 --warmup 5"
export LBM_VERBOSE=1
{
for rep in 1 2; do
  echo "== no boxes (rep $rep)"; LBM_AA_DEAD_BOXES=0 $R 2>&1 | grep -v "all-solid box" | cut -c1-160
  echo "== 4 boxes = default (rep $rep)"; $R 2>&1 | grep -v "all-solid box" | cut -c1-160
  echo "== 8 boxes (rep $rep)"; LBM_B200_LIB=$P/liblbm_b200.db8.so $R 2>&1 | grep -v "all-solid box" | cut -c1-160
done
} > $O/r2t_probe.log 2>&1
cat $O/r2t_probe.log
