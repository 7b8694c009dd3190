#!/bin/bash
# session 4, call D: one --set full capture of the odd-step obstacle kernel of the single-buffer engine
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"aa_odd_rows" --launch-skip 2 -c 1 -f -o $O/s4d_odd_rows python tools/perf_one.py inplace_obst 512 2 > $O/s4d_ncu.log 2>&1
tail -3 $O/s4d_ncu.log; ls -la $O/s4d_odd_rows.ncu-rep
