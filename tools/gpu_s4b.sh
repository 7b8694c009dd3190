#!/bin/bash
# session 4, call B: per-kernel times and DRAM bytes of the single-buffer engine on the 512^3 obstacle; launch-shape variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --launch-skip 4 -c 24 --csv --log-file $O/s4b_inplace_obst_launches.csv python tools/perf_one.py inplace_obst 512 4 > $O/s4b_ncu.log 2>&1
{
echo "== default"; python tools/perf_one.py inplace_obst 512 30
for v in a b c; do echo "== variant $v"; LBM_B200_LIB=$PWD/latticeboltzmanncuda_b200/liblbm_b200.$v.so python tools/perf_one.py inplace_obst 512 30; done
} > $O/s4b_probe.log 2>&1
cat $O/s4b_probe.log; python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/s4b_inplace_obst_launches.csv')) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
idi = hdr.index('ID')
cur = {}
for r in rows[1:]:
    cur.setdefault((r[idi], r[ki][:70]), {})[r[mi]] = r[vi]
for (i, k), m in cur.items():
    print(i, k, m)
PY
