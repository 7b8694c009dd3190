#!/bin/bash
# round 2, call L (8 GPUs): BASELINE.json's named configurations with the halo exchange inside the library --
# configs[3] periodic channel 2048 x 512 x 512 on 2 / 4 / 8 slabs (single-buffer engine; two-buffer at 8 for comparison),
# configs[1] lid-driven cavity 256^3 on one GPU on both engines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=10000
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
{
echo "== cavity 256 two-buffer"; timeout 200 python tools/run_config.py cavity 2>/dev/null | tail -1
echo "== cavity 256 single-buffer"; timeout 200 python tools/run_config.py cavity --engine inplace 2>/dev/null | tail -1
echo "== channel 2048x512x512, 8 slabs, single-buffer, peer"; timeout 300 TR 8 29541 tools/run_config.py channel --engine inplace --transport peer --steps 200 --warmup 20 2>$O/r2l_c8.err | tail -1
echo "== channel 2048x512x512, 8 slabs, two-buffer, peer"; timeout 300 TR 8 29542 tools/run_config.py channel --engine pull2 --transport peer --steps 200 --warmup 20 2>$O/r2l_c8t.err | tail -1
echo "== channel 2048x512x512, 4 slabs, single-buffer, peer"; timeout 300 TR 4 29543 tools/run_config.py channel --engine inplace --transport peer --steps 100 --warmup 10 2>$O/r2l_c4.err | tail -1
echo "== channel 2048x512x512, 2 slabs, single-buffer, peer"; timeout 300 TR 2 29544 tools/run_config.py channel --engine inplace --transport peer --steps 100 --warmup 10 2>$O/r2l_c2.err | tail -1
echo "== channel 2048x512x512, 1 GPU, single-buffer"; timeout 300 python tools/run_config.py channel --engine inplace --steps 60 --warmup 6 2>$O/r2l_c1.err | tail -1
} > $O/r2l_named_configs.jsonl 2>&1
cut -c1-520 $O/r2l_named_configs.jsonl
for f in $O/r2l_c8.err $O/r2l_c4.err $O/r2l_c2.err $O/r2l_c1.err; do grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $f | tail -n 3 | cut -c1-300; done
