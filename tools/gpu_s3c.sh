#!/bin/bash
# session 3, call C: special-kernel look-ahead without prefetch, cache-streaming hints; BASELINE configs[1] and the
# 1024^3 Taylor-Green domain on ONE GPU (single buffer)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
P=$PWD/latticeboltzmanncuda_b200
{
echo "== default"; python tools/perf_probe.py 512 40 headline
for v in b c d e; do
  if [ -f $P/liblbm_b200.$v.so ]; then echo "== variant $v"; LBM_B200_LIB=$P/liblbm_b200.$v.so python tools/perf_probe.py 512 40 headline; fi
done
echo "== fast, variants d e"
python tools/perf_probe.py 512 40 fast
LBM_B200_LIB=$P/liblbm_b200.d.so python tools/perf_probe.py 512 40 fast
LBM_B200_LIB=$P/liblbm_b200.e.so python tools/perf_probe.py 512 40 fast
} > $O/s3c_probe.log 2>&1
{
echo "== cavity 256"; timeout 200 python tools/run_config.py cavity --steps 200 --warmup 20
echo "== taylor-green 768 two-buffer, 1 GPU"; timeout 300 python tools/run_config.py taylor-green --size 768 --steps 200 --warmup 20
echo "== taylor-green 1024 in-place, 1 GPU"; timeout 400 python tools/run_config.py taylor-green --size 1024 --engine inplace --steps 200 --warmup 20
} > $O/s3c_configs.log 2>&1
cat $O/s3c_probe.log; grep -v "^\[W\|^W0" $O/s3c_configs.log | cut -c1-1500
