#!/usr/bin/env python
"""Runs one of BASELINE.json's named configurations at full size through the public API and prints one JSON line.

    python tools/run_config.py cavity                      # configs[1]: lid-driven cavity 256^3 D3Q19, 1 GPU
    torchrun --nproc-per-node N tools/run_config.py channel   # configs[3]: periodic channel 2048x512x512, N = 2/4/8 x-slabs
    torchrun --nproc-per-node 8 tools/run_config.py taylor-green [--size 1024]   # configs[4]: Taylor-Green 1024^3, decay check
    python tools/run_config.py taylor-green --size 1024 --engine inplace         # the same domain on ONE GPU (single buffer)
    python tools/run_config.py obstacle [--size 896 --engine inplace]            # configs[2]: flow past the reference test geometry

`bench.py` keeps the weak-scaling contract (512^3 cells per GPU); this script is for the fixed-size domains the
configs name (strong scaling: the domain is split into N x-slabs).  Initial states are generated on the GPU with
torch and handed to the library as device pointers, so that set-up costs seconds, not minutes, of 8-GPU time.
Timing: CUDA events around K steps after W warm-up steps, maximum over ranks.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CX = [0, -1, 0, 0, 0, 0, 1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1]
CY = [0, 0, -1, 0, 0, 1, 0, -1, 0, 0, 1, -1, -1, 1, 1, -1, 0, 0, 1]
CZ = [0, 0, 0, -1, 1, 0, 0, 0, -1, 1, 0, -1, 1, -1, 1, 0, -1, 1, 0]
W19 = [1 / 3] + [1 / 18] * 6 + [1 / 36] * 12


def equilibrium(torch, rho, ux, uy, uz):
    """D3Q19 second-order equilibrium, float32 [19, ...] on the device of the inputs"""
    usq = ux * ux + uy * uy + uz * uz
    out = torch.empty((19,) + tuple(rho.shape), dtype=torch.float32, device=rho.device)
    for i in range(19):
        cu = CX[i] * ux + CY[i] * uy + CZ[i] * uz
        out[i] = (W19[i] * rho * (1.0 + 3.0 * cu + 4.5 * cu * cu - 1.5 * usq)).to(torch.float32)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cavity", "channel", "taylor-green", "obstacle"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--size", type=int, default=0, help="override the edge length (taylor-green, cavity) for small trials")
    ap.add_argument("--engine", default="pull2", choices=["pull2", "inplace"])
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="slabs: peer-memory halos below the C ABI, or NCCL send/recv")
    args = ap.parse_args()

    import torch
    import latticeboltzmanncuda_b200 as lb
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the JSON line only
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)

    tau = 0.6
    nu = (tau - 0.5) / 3.0
    P, W = lb.BC_PERIODIC, lb.BC_WALL
    common = dict(q=19, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    wall_u = None
    if args.config == "cavity":
        n = args.size or 256
        gx, gy, gz, bc = n, n, n, (W, W, W)
        wall_u = np.zeros((6, 3), np.float32); wall_u[3] = (0.05, 0.0, 0.0)      # lid: the y = Y-1 face moves along +x
        name = "lid-driven cavity %d^3 D3Q19 BGK fp32, bounce-back walls, lid u=(0.05,0,0)" % n
    elif args.config == "channel":
        gx, gy, gz, bc = (args.size * 4 if args.size else 2048), (args.size or 512), (args.size or 512), (P, W, P)
        name = "periodic channel %dx%dx%d D3Q19 BGK fp32, periodic x and z, bounce-back walls in y, initial u=(0.04,0,0)" % (gx, gy, gz)
    elif args.config == "obstacle":
        n = args.size or 512
        gx, gy, gz, bc = n, n, n, (P, W, W)
        name = ("flow past obstacle (reference test geometry, tapered hexahedron) %d^3 D3Q19 BGK fp32, half-way bounce-back, "
                "periodic x, walls y/z, initial u=(0.04,0,0), plane source at x=0, force accumulator on" % n)
        if world > 1:
            raise SystemExit("bench.py runs the obstacle configuration on slabs (one body per GPU)")
    else:
        n = args.size or 1024
        gx, gy, gz, bc = n, n, n, (P, P, P)
        name = "Taylor-Green vortex %d^3 D3Q19 BGK fp32, fully periodic, u0=0.04, tau=0.6" % n
    if world > 1 and (bc[0] != P or gx % world):
        raise SystemExit("slab runs need a periodic x axis divisible by the number of GPUs")
    nx = gx // world
    x0 = rank * nx

    # ---- initial state on the GPU
    t_setup = time.perf_counter()
    xs = torch.arange(x0, x0 + nx, device=dev, dtype=torch.float64)[:, None, None]
    ys = torch.arange(gy, device=dev, dtype=torch.float64)[None, :, None]
    one = torch.ones((nx, gy, gz), device=dev, dtype=torch.float32)
    if args.config == "taylor-green":
        u0, k = 0.04, 2.0 * math.pi / gx
        ux = (u0 * torch.sin(k * xs) * torch.cos(k * ys)).to(torch.float32) * one
        uy = (-u0 * torch.cos(k * xs) * torch.sin(k * ys)).to(torch.float32) * one
        rho = (1.0 - u0 * u0 * 0.75 * (torch.cos(2 * k * xs) + torch.cos(2 * k * ys))).to(torch.float32) * one
        f0 = equilibrium(torch, rho, ux, uy, torch.zeros_like(one))
        del ux, uy, rho
    elif args.config in ("channel", "obstacle"):
        f0 = equilibrium(torch, one, 0.04 * one, torch.zeros_like(one), torch.zeros_like(one))
    else:
        z = torch.zeros_like(one)
        f0 = equilibrium(torch, one, z, z, z)
        del z
    del one
    torch.cuda.empty_cache()          # hand the temporaries back before the library allocates its state

    engine = lb.ENGINE_INPLACE if args.engine == "inplace" else lb.ENGINE_PULL2
    if world == 1:
        lat = lb.Lattice(gx, gy, gz, None, None, lb.FluidData(1.0, tau), 1.0, bc=bc, wall_u=wall_u, device=local, engine=engine, **common)
        stepper = lat
    else:
        from latticeboltzmanncuda_b200.slab import SlabLattice
        stepper = SlabLattice(gx, gy, gz, lb.FluidData(1.0, tau), rank=rank, world=world, device=local, bc_x=P,
                              bc_yz=(bc[1], bc[2]), transport=args.transport, **common,
                              **({"engine": engine} if args.engine == "inplace" else {}))
        lat = stepper.lattice
    live = gx * gy * gz
    if args.config == "obstacle":
        from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask
        solid = tapered_hexahedron_mask(gx, gy, gz)
        live -= int(solid.sum())
        keep = torch.from_numpy(1 - solid).to(dev).to(torch.float32)
        for i in range(19):
            f0[i] *= keep          # solid cells hold zeros
        del keep
        lat.set_solid_mask(solid)
        flow = np.zeros(27, np.float32); flow[:19] = 1e-6; flow[6] = 2e-6
        lat.setFlowData(lb.FlowData(((0, None, None), (0, None, None)), flow))
        lat.setReflectionData(lb.ReflectionData())
        del solid
    stepper.load_populations(f0)
    del f0
    torch.cuda.empty_cache()
    n_local = nx * gy * gz

    def allsum(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def diagnostics():
        """global mass, kinetic energy sum rho (ux^2+uy^2+uz^2) and the mean x velocity of the lid layer, on the device"""
        m = torch.empty((4, n_local), dtype=torch.float32, device=dev)
        lat.moments(out=[m[0], m[1], m[2], m[3]])
        r = m[0].double()
        e = float((r * (m[1].double() ** 2 + m[2].double() ** 2 + m[3].double() ** 2)).sum().item())
        lid = float(m[1].view(nx, gy, gz)[:, -1, :].double().mean().item())
        mass = float(r.sum().item())
        del m, r
        return allsum(mass), allsum(e), lid

    def barrier():
        torch.cuda.synchronize()
        lat.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    mass0, _, _ = diagnostics()
    t_setup = time.perf_counter() - t_setup
    stepper.step(args.warmup)
    barrier()
    if world > 1:
        stepper.sync()
    # the energy baseline is taken after the warm-up steps: the state starts as a pure equilibrium and spends its
    # first steps growing the non-equilibrium part, which is not the decay being checked
    _, e0, _ = diagnostics()
    ms, launches = stepper.step_timed(args.steps)
    barrier()
    ms = allmax(ms)
    if world > 1:
        stepper.sync()
    mass1, e1, lid = diagnostics()
    cells = gx * gy * gz
    mlups = cells * args.steps / (ms * 1e-3) / 1e6
    out = {"config": name, "n_gpus": world, "engine": "in-place single buffer" if args.engine == "inplace" else "two-buffer fused pull",
           "transport": (args.transport if world > 1 else None), "slab": "%dx%dx%d per GPU" % (nx, gy, gz), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
           "mlups": mlups, "mlups_per_gpu": mlups / world, "hbm_gbs_per_gpu": mlups / world * 1e6 * 152 / 1e9 * live / cells,
           "gpu_launches": int(launches), "state_gb_per_gpu": lat.layout().state_bytes * lat.layout().n_buffers / 1e9,
           "mass_drift_rel": (mass1 - mass0) / mass0, "setup_seconds": t_setup}
    if args.config == "taylor-green":
        k = 2.0 * math.pi / gx
        expected = math.exp(-2.0 * nu * 2.0 * k * k * args.steps)
        out["energy_ratio"] = e1 / e0
        out["energy_ratio_analytic"] = expected
        out["decay_rate_rel_err"] = (math.log(e1 / e0) - math.log(expected)) / math.log(expected)
        # the initial pressure field is the incompressible one; what it misses sloshes through the box as a sound
        # wave with period L / c_s, so the ratio is cleanest over a whole number of periods
        out["acoustic_period_steps"] = gx * math.sqrt(3.0)
    if args.config == "cavity":
        out["lid_layer_mean_ux"] = lid
    if args.config == "obstacle":
        r = lat.retrieve_reflection_data()
        out["live_cells"] = live
        out["force_on_body"] = [r.x, r.y, r.z]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
