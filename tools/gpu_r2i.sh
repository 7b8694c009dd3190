#!/bin/bash
# round 2, call I (8 GPUs): the driver's own command at N = 8 (single-buffer engine, peer-memory halos, e2e leg included)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=10000
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
(timeout 600 $TR8 bench.py --gpus 8 --steps 40 --warmup 5 2>$O/r2i_b8.err | tail -1) > $O/r2i_bench_8gpu_peer.json
cut -c1-1800 $O/r2i_bench_8gpu_peer.json; echo; grep -v "^\s*$" $O/r2i_b8.err | tail -n 8 | cut -c1-300
