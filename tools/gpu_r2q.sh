#!/bin/bash
# round 2, call Q (1 GPU): HEAD after the box tests -- GPU suite, smoke, both bench arms, launch list of the bench step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=5000
(timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -22) > $O/r2q_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > $O/r2q_smoke.log
(timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>$O/r2q_bench_ref.err | tail -1) > $O/r2q_bench_ref.json
(timeout 600 python bench.py 2>$O/r2q_bench.err | tail -1) > $O/r2q_bench.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:aa_ -c 60 --csv --log-file $O/r2q_bench_launches.csv python bench.py --steps 2 --warmup 1 > $O/r2q_ncu.log 2>&1
tail -16 $O/r2q_pytest.log; cat $O/r2q_smoke.log; python - <<'PY'
import json
b=json.load(open("gpurun_out/r2q_bench.json"))
print({k:b[k] for k in ("value","ms_per_step","gpu_launches")}, b["roofline"]["frac"], b["roofline"]["traffic"])
for k in ("e2e","config1_64cubed","reference_mode","cpu_baseline","other_engines"):
    print(k, json.dumps(b.get(k))[:600])
r=json.load(open("gpurun_out/r2q_bench_ref.json"))
print("ref", r.get("value"), r.get("repetitions"), r.get("setup_seconds"), r.get("unavailable"))
PY
tail -n 3 $O/r2q_bench.err $O/r2q_bench_ref.err 2>/dev/null | cut -c1-300
