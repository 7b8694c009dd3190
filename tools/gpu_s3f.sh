#!/bin/bash
# session 3, call F (8 GPUs): BASELINE configs[4] (Taylor-Green 1024^3, one acoustic period, decay check) and
# configs[3] (periodic channel 2048x512x512) on 8 x-slabs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
{
echo "== taylor-green 1024^3, 8 slabs, one acoustic period"; timeout 240 $TR tools/run_config.py taylor-green --steps 1774 --warmup 20
echo "== channel 2048x512x512, 8 slabs"; timeout 200 $TR tools/run_config.py channel --steps 200 --warmup 20
} > $O/s3f_runs.log 2>&1
grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM\|^$" $O/s3f_runs.log | cut -c1-1500
