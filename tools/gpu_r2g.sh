#!/bin/bash
# round 2, call G (1 GPU): the GPU suite at HEAD (CUDA-graph replay of long runs on small lattices is exercised by every
# test that takes >= 16 steps), config 1 with and without the graph, the bench, a --set full capture of the four kernels
# of one odd + one even step of the headline workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
export LBM_HALO_TIMEOUT_MS=5000
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12) > $O/r2g_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > $O/r2g_smoke.log
cat > $O/r2g_config1.py <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
import bench, latticeboltzmanncuda_b200 as lb
n, spacing, flow, obj = bench.config1_inputs()
l = lb.Lattice(n, n, n, None, None, lb.FluidData(1.0, 1.0), spacing)
start = np.zeros(n ** 3, dtype=lb.LATTICE_POINT); start["f"] = flow[None, :]
l.load_data(start); l.insertModel(obj); l.preProcessModel()
l.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow)); l.setReflectionData(lb.ReflectionData())
l.step(10); l.sync()
out = []
for rep in range(3):
    ms, nl = l.step_timed(1000)
    out.append(ms / 1000)
print(json.dumps({"LBM_STEP_GRAPH": os.environ.get("LBM_STEP_GRAPH", "default"), "ms_per_step": out, "mlups": [n ** 3 / (m * 1e-3) / 1e6 for m in out]}))
PY
{ LBM_STEP_GRAPH=0 timeout 100 python $O/r2g_config1.py; timeout 100 python $O/r2g_config1.py; } > $O/r2g_config1_graph.jsonl 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5 2>$O/r2g_bench.err | tail -1) > $O/r2g_bench.json
LBM_PROBE_FORCE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"aa_odd_kernel|aa_odd_special|aa_even_kernel" --launch-skip 16 -c 5 -f -o $O/r2g_headline python tools/perf_one.py inplace_obst 512 4 > $O/r2g_ncu.log 2>&1
tail -4 $O/r2g_pytest.log; cat $O/r2g_smoke.log; cat $O/r2g_config1_graph.jsonl; cut -c1-700 $O/r2g_bench.json; echo; tail -n 3 $O/r2g_bench.err; tail -n 2 $O/r2g_ncu.log; ls -la $O/r2g_headline.ncu-rep
