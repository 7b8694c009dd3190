#!/bin/bash
# round 2, call R (2 GPUs): HEAD after the box tests -- the 2-GPU tests and the driver's bench command at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests -m gpu -q -k "slab or peer or compat" 2>&1 | tail -8) > $O/r2r_pytest.log
(timeout 600 $TR --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 100 --warmup 5 2>$O/r2r_bench.err | tail -1) > $O/r2r_bench_n2.json
tail -n 6 $O/r2r_pytest.log; cut -c1-700 $O/r2r_bench_n2.json; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $O/r2r_bench.err | tail -n 3 | cut -c1-300
