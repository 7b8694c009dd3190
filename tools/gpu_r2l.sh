#!/bin/bash
# round 2, call L2 (8 GPUs): BASELINE.json configs[3], periodic channel 2048 x 512 x 512 on 2 / 4 / 8 slabs with the halo
# exchange inside the library (single-buffer engine; two-buffer at 8 for comparison)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
echo "== channel 2048x512x512, 8 slabs, single-buffer, peer"; timeout 300 $TR --nproc-per-node 8 --master-port 29541 tools/run_config.py channel --engine inplace --transport peer --steps 200 --warmup 20 2>$O/r2l_c8.err | tail -1
echo "== channel 2048x512x512, 8 slabs, two-buffer, peer"; timeout 300 $TR --nproc-per-node 8 --master-port 29542 tools/run_config.py channel --engine pull2 --transport peer --steps 200 --warmup 20 2>$O/r2l_c8t.err | tail -1
echo "== channel 2048x512x512, 4 slabs, single-buffer, peer"; timeout 300 $TR --nproc-per-node 4 --master-port 29543 tools/run_config.py channel --engine inplace --transport peer --steps 100 --warmup 10 2>$O/r2l_c4.err | tail -1
echo "== channel 2048x512x512, 2 slabs, single-buffer, peer"; timeout 300 $TR --nproc-per-node 2 --master-port 29544 tools/run_config.py channel --engine inplace --transport peer --steps 100 --warmup 10 2>$O/r2l_c2.err | tail -1
} > $O/r2l_named_configs.jsonl 2>&1
cut -c1-520 $O/r2l_named_configs.jsonl
for f in $O/r2l_c8.err $O/r2l_c8t.err $O/r2l_c4.err $O/r2l_c2.err; do grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $f | tail -n 3 | cut -c1-300; done
