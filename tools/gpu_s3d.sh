#!/bin/bash
# session 3, call D: why is the in-place engine at 1024^3 slower per cell than at 512^3?  sizes around 1024, per-kernel
# times; the cavity on the in-place engine
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
{
echo "== cavity 256 in-place"; timeout 200 python tools/run_config.py cavity --engine inplace --steps 200 --warmup 20
echo "== cavity 512 in-place"; timeout 200 python tools/run_config.py cavity --size 512 --engine inplace --steps 100 --warmup 10
echo "== cavity 512 two-buffer"; timeout 200 python tools/run_config.py cavity --size 512 --steps 100 --warmup 10
for n in 768 896 1000 1024; do
echo "== taylor-green $n in-place"; timeout 300 python tools/run_config.py taylor-green --size $n --engine inplace --steps 40 --warmup 6
done
} > $O/s3d_configs.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:aa_ --launch-skip 6 -c 6 --csv --log-file $O/s3d_ncu_inplace1024.csv python tools/run_config.py taylor-green --size 1024 --engine inplace --steps 4 --warmup 6 > /dev/null 2>&1
grep -v "^\[W\|^W0" $O/s3d_configs.log | cut -c1-700
grep -v "^==" $O/s3d_ncu_inplace1024.csv | awk -F'","' 'NR>1{print substr($5,1,50), $(NF-2), $NF}' | tr -d '"'
