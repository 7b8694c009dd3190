#!/bin/bash
# round 2, call V (1 GPU): ncu --set full of the single-buffer headline step's kernels at HEAD (after the box tests)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
LBM_PROBE_FORCE=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"aa_odd_kernel|aa_odd_special|aa_even_kernel" --launch-skip 16 -c 5 -f -o $O/r2v_headline python tools/perf_one.py inplace_obst 512 4 > $O/r2v_ncu.log 2>&1
tail -n 3 $O/r2v_ncu.log; ls -la $O/r2v_headline.ncu-rep
