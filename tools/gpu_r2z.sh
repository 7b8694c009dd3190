#!/bin/bash
# round 2, call Z (1 GPU): final state -- GPU suite, smoke, both bench arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=5000
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6) > $O/r2z_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > $O/r2z_smoke.log
(timeout 600 python bench.py --impl reference 2>$O/r2z_bench_ref.err | tail -1) > $O/r2z_bench_ref.json
(timeout 600 python bench.py 2>$O/r2z_bench.err | tail -1) > $O/r2z_bench.json
tail -n 4 $O/r2z_pytest.log; cat $O/r2z_smoke.log; python - <<'PY'
import json
b=json.load(open("gpurun_out/r2z_bench.json"))
print({k:b[k] for k in ("value","ms_per_step","gpu_launches")}, b["roofline"]["frac"], b["e2e"]["value"], b["clocks"])
r=json.load(open("gpurun_out/r2z_bench_ref.json"))
print("ref", r.get("value"), r.get("steps"), r.get("unavailable"))
PY
tail -n 3 $O/r2z_bench.err $O/r2z_bench_ref.err 2>/dev/null | cut -c1-300
