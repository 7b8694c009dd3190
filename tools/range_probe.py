"""times one GPU stepping its lattice the way a slab does (edge planes, then interior on the side stream, commit),
without any exchange -- isolates the cost of split stepping from the cost of communication"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import latticeboltzmanncuda_b200 as lb
from latticeboltzmanncuda_b200 import capi
from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask
S, K = 512, 50
lat = lb.Lattice(S, S, S, None, None, lb.FluidData(1.0, 0.6), 1.0 / S, q=19, bc=(lb.BC_PERIODIC, lb.BC_WALL, lb.BC_WALL),
                 moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
solid = tapered_hexahedron_mask(S, S, S)
w = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12, np.float32)
f = np.empty((19, S, S, S), np.float32)
for i in range(19):
    f[i] = w[i]
f[:, solid.astype(bool)] = 0
lat.set_solid_mask(solid); lat.load_populations(f); del f
lat.setReflectionData(lb.ReflectionData())
lib, h = capi.load(), lat.handle
stream = torch.cuda.Stream()
capi.check(lib.lbm_set_stream(h, C.c_void_p(stream.cuda_stream)))
P = lb.PHASE_ALL
def whole(n):
    capi.check(lib.lbm_step(h, n))
def split(n, overlap=True):
    for _ in range(n):
        capi.check(lib.lbm_fill_ghosts(h))
        capi.check(lib.lbm_step_range(h, P, 0, 1, 0))
        capi.check(lib.lbm_step_range(h, P, S - 1, S, 0))
        capi.check(lib.lbm_step_range(h, P, 1, S - 1, 1 if overlap else 0))
        capi.check(lib.lbm_step_commit(h))
def one_range(n):
    for _ in range(n):
        capi.check(lib.lbm_fill_ghosts(h))
        capi.check(lib.lbm_step_range(h, P, 0, S, 0))
        capi.check(lib.lbm_step_commit(h))
for name, fn in (("lbm_step", whole), ("one range + commit", one_range), ("edges + interior(side) + commit", split),
                 ("edges + interior(main) + commit", lambda n: split(n, False))):
    with torch.cuda.stream(stream):
        fn(5); stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(K); e1.record(stream); e1.synchronize()
    print("%-36s %.3f ms/step" % (name, e0.elapsed_time(e1) / K), flush=True)
