#!/bin/bash
# round 2, call E (2 GPUs): slabs on two devices -- NCCL transport (both engines; the single-buffer slabs for the first time
# on hardware), peer-memory halos in one process and across processes (CUDA IPC), the C++ Lattice on two devices -- and the
# bench at N = 2 with both transports next to one slab on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=8000
nvidia-smi -L > $O/r2e_smi.log 2>&1
(LBM_TEST_INPLACE_SLABS=1 timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_peer_halo.py tests/test_gpu_compat.py -m gpu -q 2>&1 | tail -15) > $O/r2e_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
(timeout 300 python bench.py --workload slab --steps 40 --warmup 5 --no-e2e --no-cpu --no-extra 2>$O/r2e_b1.err | tail -1) > $O/r2e_bench_slab_1gpu.json
(timeout 400 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu 2>$O/r2e_b2p.err | tail -1) > $O/r2e_bench_2gpu_peer.json
(timeout 400 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --no-e2e --transport nccl 2>$O/r2e_b2n.err | tail -1) > $O/r2e_bench_2gpu_nccl.json
(timeout 400 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --no-e2e --engine pull2 2>$O/r2e_b2t.err | tail -1) > $O/r2e_bench_2gpu_peer_twobuffer.json
tail -6 $O/r2e_pytest.log
for f in $O/r2e_bench_slab_1gpu.json $O/r2e_bench_2gpu_peer.json $O/r2e_bench_2gpu_nccl.json $O/r2e_bench_2gpu_peer_twobuffer.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    b=json.load(open(sys.argv[1]))
    print({k:b.get(k) for k in ("value","ms_per_step","n_gpus","gpu_launches")}, "frac", b["roofline"]["frac"], "e2e", (b.get("e2e") or {}).get("value"))
except Exception as e:
    print("unreadable", e)
PY
done
tail -3 $O/r2e_b1.err $O/r2e_b2p.err $O/r2e_b2n.err $O/r2e_b2t.err
