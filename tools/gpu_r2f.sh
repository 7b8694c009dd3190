#!/bin/bash
# round 2, call F (8 GPUs): the bench at N = 8 and N = 4 (1024^3 domain at 8: one 128 x 1024 x 1024 slab per GPU) with
# peer-memory halos, NCCL transport and the two-buffer engine for comparison; Taylor-Green 1024^3 on 8 slabs of the
# single-buffer engine over one acoustic period (analytic decay); a subset of the peer-halo tests over 8 devices
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=10000
nvidia-smi -L > $O/r2f_smi.log 2>&1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
(timeout 500 $TR8 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu 2>$O/r2f_b8p.err | tail -1) > $O/r2f_bench_8gpu_peer.json
(timeout 400 $TR8 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu --no-e2e --transport nccl 2>$O/r2f_b8n.err | tail -1) > $O/r2f_bench_8gpu_nccl.json
(timeout 400 $TR8 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu --no-e2e --engine pull2 2>$O/r2f_b8t.err | tail -1) > $O/r2f_bench_8gpu_peer_twobuffer.json
(timeout 400 $TR4 bench.py --gpus 4 --steps 40 --warmup 5 --no-cpu --no-e2e 2>$O/r2f_b4p.err | tail -1) > $O/r2f_bench_4gpu_peer.json
(timeout 400 $TR8 tools/run_config.py taylor-green --size 1024 --engine inplace --transport peer --steps 1774 --warmup 20 2>$O/r2f_tg.err | tail -1) > $O/r2f_taylor_green_1024_8gpu_inplace_peer.json
(timeout 300 python -m pytest tests/test_gpu_peer_halo.py -m gpu -q -k "reference_mode or w3 or w4" 2>&1 | tail -6) > $O/r2f_pytest.log
for f in $O/r2f_bench_8gpu_peer.json $O/r2f_bench_8gpu_nccl.json $O/r2f_bench_8gpu_peer_twobuffer.json $O/r2f_bench_4gpu_peer.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    b=json.load(open(sys.argv[1]))
    print({k:b.get(k) for k in ("value","ms_per_step","n_gpus","gpu_launches")}, "frac", b["roofline"]["frac"], "e2e", (b.get("e2e") or {}).get("value"), b.get("clocks"))
except Exception as e:
    print("unreadable", e)
PY
done
cut -c1-900 $O/r2f_taylor_green_1024_8gpu_inplace_peer.json; echo; tail -4 $O/r2f_pytest.log
for f in $O/r2f_b8p.err $O/r2f_b8n.err $O/r2f_b8t.err $O/r2f_b4p.err $O/r2f_tg.err; do echo "-- $f"; tail -n 3 $f; done
