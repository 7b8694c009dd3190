#!/bin/bash
# round 2, call Y (1 GPU): diagnostics of the single-buffer engine with the warp-uniform short cuts; the single-buffer tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
timeout 300 python tools/diag_probe.py 512 > $O/r2y_diag.json 2> $O/r2y_diag.err; cat $O/r2y_diag.json; tail -n 3 $O/r2y_diag.err
timeout 900 python -m pytest tests -m gpu -x -q -k "inplace or compact or slab or peer or conservation or cavity" 2>&1 | tail -n 3
