#!/bin/bash
# round 2, call U4 (4 GPUs): the driver's bench command at N = 4 at HEAD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 400 $TR --nproc-per-node 4 --master-port 29581 bench.py --gpus 4 --steps 100 --warmup 5 2>$O/r2u4_bench.err | tail -1) > $O/r2u4_bench_n4.json
cut -c1-300 $O/r2u4_bench_n4.json; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $O/r2u4_bench.err | tail -n 3 | cut -c1-300
