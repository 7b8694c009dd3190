#!/usr/bin/env python
"""Throughput of an all-fluid, fully periodic NX x NY x NZ D3Q19 box on one GPU (CUDA events), one JSON line.

    python tools/box_probe.py NX NY NZ [inplace|pull2] [steps] [slabs]

`slabs` > 1 cuts the box into that many x slabs ON THE SAME GPU, joined by peer-memory halos (SlabGroup): the same
lattice and the same traffic, but several allocations instead of one.

Used for the 1024^3 study (profiles/r02_notes.md): the same kernel at 1024 x 1024 x 1024, at neighbouring shapes and
with the layout padded through LBM_PAD_ROWS / LBM_PAD_LINES (csrc/lbm_device.h).  The initial state is a uniform
equilibrium generated on the device.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import latticeboltzmanncuda_b200 as lb
    nx, ny, nz = (int(v) for v in sys.argv[1:4])
    engine = sys.argv[4] if len(sys.argv) > 4 else "inplace"
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    slabs = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    dev = torch.device("cuda", 0)
    if slabs > 1:
        return run_slabs(torch, lb, nx, ny, nz, engine, steps, slabs)
    w = [1 / 3] + [1 / 18] * 6 + [1 / 36] * 12
    f0 = torch.empty((19, nx, ny, nz), dtype=torch.float32, device=dev)
    for i in range(19):
        f0[i] = w[i]
    P = lb.BC_PERIODIC
    lat = lb.Lattice(nx, ny, nz, None, None, lb.FluidData(1.0, 0.6), 1.0, q=19, bc=(P, P, P), moment_mode=lb.MOMENT_STANDARD,
                     eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK, device=0,
                     engine=lb.ENGINE_INPLACE if engine == "inplace" else lb.ENGINE_PULL2)
    lat.load_populations(f0)
    del f0
    torch.cuda.empty_cache()
    lat.step(4)
    lat.sync()
    ms, launches = lat.step_timed(steps)
    lay = lat.layout()
    cells = nx * ny * nz
    gbs = cells * steps * 152 / (ms * 1e-3) / 1e9
    print(json.dumps({"box": [nx, ny, nz], "engine": engine, "pad_rows": os.environ.get("LBM_PAD_ROWS", "0"),
                      "pad_lines": os.environ.get("LBM_PAD_LINES", "0"), "yp": lay.yp, "zp": lay.zp,
                      "plane_bytes": lay.plane * 4, "ms_per_step": ms / steps, "glups": cells * steps / (ms * 1e-3) / 1e9,
                      "gbs": gbs, "state_gb": lay.state_bytes * lay.n_buffers / 1e9, "launches": launches}), flush=True)
    lat.close()


def run_slabs(torch, lb, nx, ny, nz, engine, steps, slabs):
    from latticeboltzmanncuda_b200.slab import SlabGroup
    P = lb.BC_PERIODIC
    grp = SlabGroup(nx, ny, nz, lb.FluidData(1.0, 0.6), [0] * slabs, bc_x=P, bc_yz=(P, P), q=19, moment_mode=lb.MOMENT_STANDARD,
                    eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK,
                    engine=lb.ENGINE_INPLACE if engine == "inplace" else lb.ENGINE_PULL2)
    w = [1 / 3] + [1 / 18] * 6 + [1 / 36] * 12
    for (x0, x1), s in zip(grp.ranges, grp.slabs):
        f0 = torch.empty((19, x1 - x0, ny, nz), dtype=torch.float32, device="cuda")
        for i in range(19):
            f0[i] = w[i]
        s.load_populations(f0)
        del f0
        torch.cuda.empty_cache()
    grp.step(4)
    grp.sync()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    grp.step(steps)
    grp.sync()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    cells = nx * ny * nz
    print(json.dumps({"box": [nx, ny, nz], "engine": engine, "slabs_on_one_gpu": slabs, "ms_per_step": ms / steps,
                      "glups": cells * steps / (ms * 1e-3) / 1e9, "gbs": cells * steps * 152 / (ms * 1e-3) / 1e9,
                      "timing": "host clock around %d steps and a device synchronisation" % steps,
                      "state_gb": sum(s.layout().state_bytes * s.layout().n_buffers for s in grp.slabs) / 1e9}), flush=True)
    grp.close()


if __name__ == "__main__":
    main()
