#!/usr/bin/env python
"""Time of the diagnostics of the single-buffer engine (mass / momentum reduction, moments) on the 512^3 obstacle lattice,
in both storage forms (after an even and after an odd number of steps).  Host clock around the call (each ends with a
device synchronisation); one JSON line.   python tools/diag_probe.py [size]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeboltzmanncuda_b200 as lb
from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
P, W = lb.BC_PERIODIC, lb.BC_WALL
lat = lb.Lattice(S, S, S, None, None, lb.FluidData(1.0, 0.6), 1.0 / S, q=19, bc=(P, W, W), moment_mode=lb.MOMENT_STANDARD,
                 eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK, engine=lb.ENGINE_INPLACE)
w = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12, np.float32)
f = np.empty((19, S, S, S), np.float32)
for i in range(19):
    f[i] = w[i]
solid = tapered_hexahedron_mask(S, S, S)
f[:, solid.astype(bool)] = 0
lat.set_solid_mask(solid)
lat.load_populations(f)
del f
out = {"size": S}
for steps, form in ((4, "parked (even number of steps)"), (1, "arrived (odd number of steps)")):
    lat.step(steps)
    lat.sync()
    lat.mass()
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); m = lat.mass(); t.append(time.perf_counter() - t0)
    out[form] = {"mass": m, "mass_ms": 1e3 * min(t)}
    import torch
    bufs = [torch.empty(S ** 3, dtype=torch.float32, device="cuda") for _ in range(4)]
    lat.moments(out=bufs)
    torch.cuda.synchronize()
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); lat.moments(out=bufs); lat.sync(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    out[form]["moments_ms_to_device_buffers"] = 1e3 * min(t)
    out[form]["rho_sum"] = float(bufs[0].double().sum())
    del bufs
print(json.dumps(out))
lat.close()
