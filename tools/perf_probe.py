"""Quick throughput probe of the step kernel variants (device-resident, CUDA-event timed).
usage: python tools/perf_probe.py [size] [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmanncuda_b200 as lb
from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
try:
    import json
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0     # B200_PROFILING.md fallback


def run(name, q, eq, bc, boundary, obstacle, source, moment=lb.MOMENT_STANDARD, engine=lb.ENGINE_AUTO):
    lat = lb.Lattice(S, S, S, None, None, lb.FluidData(1.0, 0.6), 1.0 / S, q=q, bc=bc, moment_mode=moment,
                     eq_precision=eq, boundary_mode=boundary, engine=engine)
    w = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12, np.float32) if q == 19 else \
        np.array([8 / 27] + [2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8, np.float32)
    f = np.empty((q, S, S, S), np.float32)
    for i in range(q):
        f[i] = w[i]
    live = S ** 3
    if obstacle:
        solid = tapered_hexahedron_mask(S, S, S)
        f[:, solid.astype(bool)] = 0
        lat.set_solid_mask(solid)
        live -= int(solid.sum())
    lat.load_populations(f)
    del f
    if source:
        flow = np.zeros(27, np.float32); flow[:q] = 0.001; flow[6] = 0.002
        lat.setFlowData(lb.FlowData(((0, None, None), (0, None, None)), flow))
    if os.environ.get("LBM_PROBE_FORCE"):      # momentum exchange with the obstacle, like bench.py
        lat.setReflectionData(lb.ReflectionData())
    lat.step(5)
    lat.sync()
    ms, launches = lat.step_timed(K)
    mlups = S ** 3 * K / (ms * 1e-3) / 1e6
    gbs = live * K / (ms * 1e-3) * 2 * q * 4 / 1e9        # algorithmic bytes of the live (non-solid) cells
    st = lat.stats()
    name = "%s [dead %.0f%% special %.1f%%]" % (name, 100.0 * st["words_dead"] / st["words_total"], 100.0 * st["words_special"] / st["words_total"])
    print("%-44s %8.0f MLUPS %7.0f GB/s %5.1f%% of %.0f  (%.3f ms/step, %d launches)" %
          (name, mlups, gbs, 100 * gbs / PEAK, PEAK, ms / K, launches), flush=True)
    if os.environ.get("LBM_PROBE_FORCE"):      # a checksum to compare builds / launch modes with: must be identical
        print("    force %r mass %r" % (tuple(float(v) for v in lat.force_f64()), float(lat.mass())), flush=True)
    lat.close()


Z, P, W = lb.BC_ZERO_INFLOW, lb.BC_PERIODIC, lb.BC_WALL
BB, RF = lb.BOUNDARY_BOUNCE_BACK, lb.BOUNDARY_REFLECT
which = sys.argv[3] if len(sys.argv) > 3 else "all"
if which in ("all", "fast"):
    run("q19 fp32 fast (no geometry, zero-inflow)", 19, lb.EQ_FP32, (Z, Z, Z), BB, False, False)
if which in ("all", "headline"):
    run("q19 fp32 general obstacle+walls+source", 19, lb.EQ_FP32, (Z, W, W), BB, True, True)
if which in ("parts",):
    run("q19 general, source only (all ordinary)", 19, lb.EQ_FP32, (Z, Z, Z), BB, False, True)
    run("q19 general, walls y/z only", 19, lb.EQ_FP32, (Z, W, W), BB, False, False)
    run("q19 general, walls y only", 19, lb.EQ_FP32, (Z, W, Z), BB, False, False)
    run("q19 general, obstacle only", 19, lb.EQ_FP32, (Z, Z, Z), BB, True, False)
if which in ("all", "inplace"):
    run("q19 fp32 IN-PLACE periodic xyz (1 buffer)", 19, lb.EQ_FP32, (P, P, P), BB, False, False, engine=lb.ENGINE_INPLACE)
    run("q19 fp32 IN-PLACE walls xyz (1 buffer)", 19, lb.EQ_FP32, (W, W, W), BB, False, False, engine=lb.ENGINE_INPLACE)
    run("q27 fp32 IN-PLACE periodic xyz (1 buffer)", 27, lb.EQ_FP32, (P, P, P), BB, False, False, engine=lb.ENGINE_INPLACE)
if which in ("all", "inplace", "inplace-obstacle"):
    run("q19 fp32 two-buffer obstacle+walls+source, periodic x", 19, lb.EQ_FP32, (P, W, W), BB, True, True)
    run("q19 fp32 IN-PLACE obstacle+walls+source (1 buffer)", 19, lb.EQ_FP32, (P, W, W), BB, True, True, engine=lb.ENGINE_INPLACE)
    run("q19 fp32 IN-PLACE obstacle only, periodic xyz", 19, lb.EQ_FP32, (P, P, P), BB, True, False, engine=lb.ENGINE_INPLACE)
if which in ("all",):
    run("q19 fp32 periodic xyz (3 ghost fills)", 19, lb.EQ_FP32, (P, P, P), BB, False, False)
    run("q27 fp32 fast", 27, lb.EQ_FP32, (Z, Z, Z), BB, False, False)
    run("q27 fp32 general obstacle+walls+source", 27, lb.EQ_FP32, (Z, W, W), BB, True, True)
    run("q27 fp64-eq reference mode fast", 27, lb.EQ_FP64, (Z, Z, Z), RF, False, False, moment=lb.MOMENT_REFERENCE)
    run("q19 fp64-eq fast", 19, lb.EQ_FP64, (Z, Z, Z), BB, False, False)
