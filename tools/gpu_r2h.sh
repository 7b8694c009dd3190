#!/bin/bash
# round 2, call H (1 GPU): launch-shape variants of the single-buffer kernels on the headline workload, A/B in one call
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
{
for rep in 1 2; do
for v in default w2 w8 w16 mb10 mb16 spc10 spc12; do
  echo "== $v (rep $rep)"
  if [ $v = default ]; then LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40
  else LBM_B200_LIB=$P/liblbm_b200.$v.so LBM_PROBE_FORCE=1 timeout 60 python tools/perf_one.py inplace_obst 512 40; fi
done
done
for v in default w8 w16; do
  echo "== all-fluid $v"
  if [ $v = default ]; then timeout 60 python tools/perf_one.py inplace 512 40
  else LBM_B200_LIB=$P/liblbm_b200.$v.so timeout 60 python tools/perf_one.py inplace 512 40; fi
done
} > $O/r2h_probe.log 2>&1
cat $O/r2h_probe.log
