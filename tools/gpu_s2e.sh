#!/bin/bash
# session 2, call E (2 GPUs): slab equality test + the named-config runner on two slabs at reduced size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -5) > $O/s2e_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
{
echo "== taylor-green 256, 2 slabs"; timeout 300 $TR tools/run_config.py taylor-green --size 256 --steps 443
echo "== channel 1024x256x256, 2 slabs"; timeout 300 $TR tools/run_config.py channel --size 256 --steps 100
echo "== bench 2 GPUs"; timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e
} > $O/s2e_runs.log 2>&1
cat $O/s2e_pytest.log; grep -v "^\[W\|^W0\|^\*\*\*" $O/s2e_runs.log | cut -c1-1200
