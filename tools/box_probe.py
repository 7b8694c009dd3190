#!/usr/bin/env python
"""Throughput of an all-fluid, fully periodic NX x NY x NZ D3Q19 box on one GPU (CUDA events), one JSON line.

    python tools/box_probe.py NX NY NZ [inplace|pull2] [steps]

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
    dev = torch.device("cuda", 0)
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


if __name__ == "__main__":
    main()
