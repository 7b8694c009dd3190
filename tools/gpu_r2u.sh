#!/bin/bash
# round 2, call U (8 GPUs): the driver's bench command at N = 8 (the 1024^3 domain) at HEAD
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 400 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 100 --warmup 5 2>$O/r2u_bench.err | tail -1) > $O/r2u_bench_n8.json
cut -c1-400 $O/r2u_bench_n8.json; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $O/r2u_bench.err | tail -n 3 | cut -c1-300
