#!/bin/bash
# session 3, call E (2 GPUs): slab equality + global diagnostics test, bench.py under torchrun, channel and Taylor-Green on two slabs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -5) > $O/s3e_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
{
echo "== bench 2 GPUs"; timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 5
echo "== bench reference arm under torchrun"; timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1
echo "== channel 2048x512x512, 2 slabs"; timeout 300 $TR tools/run_config.py channel --steps 100 --warmup 10
echo "== taylor-green 768, 2 slabs"; timeout 300 $TR tools/run_config.py taylor-green --size 768 --steps 100 --warmup 10
} > $O/s3e_runs.log 2>&1
cat $O/s3e_pytest.log; grep -v "^\[W\|^W0\|^\*\*\*\|OMP_NUM\|^$" $O/s3e_runs.log | cut -c1-1800
