#!/bin/bash
# round 2, call J (1 GPU): "more requests in flight per thread" experiments, A/B in one call:
#   er2 / er2b6 : even kernel of the single-buffer engine takes two rows per thread (64 / 80 registers)
#   gp8 / gp5 / gp4 : two-buffer general kernel pulls two ordinary rows before the first collision (64 / 96 / 128 registers)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
run() { if [ $1 = default ]; then shift; "$@"; else v=$1; shift; LBM_B200_LIB=$P/liblbm_b200.$v.so "$@"; fi; }
{
for rep in 1 2; do
for v in default er2 er2b6; do echo "== single-buffer obstacle, $v (rep $rep)"; LBM_PROBE_FORCE=1 run $v timeout 60 python tools/perf_one.py inplace_obst 512 40; done
for v in default gp8 gp5 gp4; do echo "== two-buffer headline, $v (rep $rep)"; LBM_PROBE_FORCE=1 run $v timeout 60 python tools/perf_one.py headline 512 40; done
done
for v in default er2; do echo "== single-buffer all-fluid, $v"; run $v timeout 60 python tools/perf_one.py inplace 512 40; done
for v in default gp8; do echo "== two-buffer periodic all-fluid, $v"; run $v timeout 60 python tools/perf_one.py periodic 512 40; done
} > $O/r2j_probe.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
LBM_B200_LIB=$P/liblbm_b200.gp8.so LBM_PROBE_FORCE=1 timeout 200 ncu --metrics $M --clock-control none --launch-skip 40 -c 8 --csv --log-file $O/r2j_headline_gp8_launches.csv python tools/perf_one.py headline 512 6 > $O/r2j_ncu.log 2>&1
cat $O/r2j_probe.log; python tools/launch_table.py $O/r2j_headline_gp8_launches.csv 2>/dev/null | tail -9
