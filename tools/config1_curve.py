#!/usr/bin/env python
"""BASELINE.json configs[0] at full length: the reference test geometry on a 64^3 grid, 1000 steps of the CLI loop
(LatticeBoltzmannCuda.cpp:29-93), compared every `--every` steps between

  * this library (EQ_FP64 = reference arithmetic, and EQ_FP32),
  * the CPU oracle (oracle/lbm_oracle.c), and
  * the reference's own kernels on the same GPU (oracle/_ref/libref_gpu.so), when that library was built.

Prints one JSON object per checkpoint: max |f - f_oracle| / max |f_oracle| for every pair, the minimum density of the
live cells (the open box drains, and u = m / rho amplifies rounding where rho -> 0) and the force accumulators.
Test infrastructure: the oracle is the checker here, nothing of it is on the product path.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--tau", type=float, default=1.0)
    ap.add_argument("--obj", default="cube_angled.obj")
    args = ap.parse_args()
    import _oracle as O
    import _ref_gpu as R
    import latticeboltzmanncuda_b200 as lb

    n, tau = args.n, args.tau
    obj = os.path.join(O.GOLDEN_DIR, "obj", args.obj)
    spacing = np.float32(1.0 / n)
    flow = np.full(27, 0.001, np.float32)
    flow[6] = 0.002
    cells = O.new_cells(n, n, n)
    cells["f"] = flow[None, :]
    mesh = O.Mesh(obj)
    assert mesh.insert(cells, n, n, n, spacing) == 1
    tables, tindex = mesh.preprocess(cells, n, n, n, spacing)
    live = cells["is_internal"] == 0
    force_o = np.zeros(3, np.float32)

    def make(eq, **kw):
        lat = lb.Lattice(n, n, n, (n // 8,) * 3, (8,) * 3, lb.FluidData(1.0, tau), float(spacing), eq_precision=eq, **kw)
        start = O.new_cells(n, n, n)
        start["f"] = flow[None, :]
        lat.load_data(start)
        lat.insertModel(obj)
        lat.preProcessModel()
        lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
        lat.setReflectionData(lb.ReflectionData())
        return lat

    arms = {"fp64": make(lb.EQ_FP64), "fp32": make(lb.EQ_FP32)}
    if hasattr(lb, "FORCE_ACCUM_REFERENCE_FP32"):
        arms["fp64_refforce"] = make(lb.EQ_FP64, force_accum=lb.FORCE_ACCUM_REFERENCE_FP32)
    ref = None
    if R.available():
        ref = R.RefLattice(n, n, n, tau=tau, spacing=float(spacing))
        start = O.new_cells(n, n, n)
        start["f"] = flow[None, :]
        ref.load(start)
        ref.insert_model(obj)
        ref.preprocess()
        ref.set_flow(2, flow)
        ref.set_reflection_data()
        ref.prime()

    cur = cells
    done = 0
    while done < args.steps:
        k = min(args.every, args.steps - done)
        cur = O.step_cli(cur, n, n, n, tau, tables, tindex, force_o, O.SRC_PLANE_X0, flow, steps=k)
        done += k
        want = cur["f"].astype(np.float64)
        scale = float(np.abs(want).max())
        rho = want.sum(axis=1)
        row = {"step": done, "scale": scale, "rho_min_live": float(rho[live].min()), "rho_max": float(rho.max()),
               "force_oracle": [float(v) for v in force_o]}
        for name, lat in arms.items():
            lat.step(k)
            got = lat.retrieve_data()["f"].astype(np.float64)
            r = lat.retrieve_reflection_data()
            row["err_" + name] = float(np.abs(got - want).max()) / scale
            row["force_" + name] = [r.x, r.y, r.z]
            row["force_relerr_" + name] = float(np.max(np.abs(np.array([r.x, r.y, r.z]) - force_o) / np.maximum(np.abs(force_o), 1e-30)))
        if ref is not None:
            ref.run_cli(k)
            got = ref.retrieve()["f"].astype(np.float64)
            row["err_reference_kernels_vs_oracle"] = float(np.abs(got - want).max()) / scale
            row["force_reference_kernels"] = [float(v) for v in ref.get_reflection_data()]
            got64 = arms["fp64"].retrieve_data()["f"].astype(np.float64)
            row["err_fp64_vs_reference_kernels"] = float(np.abs(got64 - got).max()) / float(np.abs(got).max())
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
