#!/usr/bin/env python
"""bench.py -- MLUPS and achieved HBM bandwidth of the fused collide-stream step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--q 19|27]

One "step" is one fused lattice update (pull-stream + bounce-back + BGK collide + source) of the
whole lattice.  Workload at N = 1: BASELINE.json configs[2], "flow past obstacle voxelised from
the reference test geometry, 512^3, bounce-back", D3Q19 fp32 -- the configuration BASELINE.json's
80%-of-HBM target is quoted on.  At N > 1 the lattice is slab-decomposed along x with the same
cells per GPU (weak scaling; N = 8 is the 1024^3 domain of configs[4]).

Prints ONE JSON line (see the task contract): value = whole-job MLUPS with the state resident in
HBM; e2e = the same job through the host API with pinned host buffers (state upload, per-step
force read-back like the reference CLI, moments download) inside the timed region; roofline =
algorithmic bytes (2*Q*4 per cell update) over the CUDA-event time of the step kernel against the
measured copy bandwidth in MEASURED_PEAKS.json; cpu_baseline = the CPU oracle port on this box's cores.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512, help="cells per axis of one GPU's share")
    ap.add_argument("--q", type=int, default=19, choices=[19, 27])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the all-fluid dense / in-place side measurements")
    ap.add_argument("--ref-size", type=int, default=int(os.environ.get("LBM_REF_SIZE", "256")))
    return ap.parse_args()


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clock and throttle reasons of one GPU while the timed region runs"""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def source_flow(q, scale=1.0):
    """the CLI's inflow (LatticeBoltzmannCuda.cpp:29-59, 66): every population gets the same amount, the +x
    population twice as much.  `scale` sets the amount relative to the CLI's 0.001."""
    flow = np.zeros(27, np.float32)
    flow[:q] = 0.001 * scale
    flow[6] = 0.002 * scale
    return flow


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------

def cpu_baseline(q, target_seconds=12.0):
    """the CPU oracle port (oracle/lbm_oracle.c, OpenMP) on a bounded sample of the same workload"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask
    n = 128
    solid = tapered_hexahedron_mask(n, n, n)
    w = O.weights(q).astype(np.float32)
    f = np.ascontiguousarray(np.broadcast_to(w[:, None, None, None], (q, n, n, n))).copy()
    f[:, solid.astype(bool)] = 0
    bc = (O.BC_ZERO, O.BC_WALL, O.BC_WALL)
    f = O.step_soa(f, n, n, n, q, 0.6, bc=bc, solid=solid, steps=1)          # warm-up
    steps, t0 = 0, time.perf_counter()
    while True:
        f = O.step_soa(f, n, n, n, q, 0.6, bc=bc, solid=solid, steps=2)
        steps += 2
        dt = time.perf_counter() - t0
        if dt >= target_seconds or steps >= 400:
            break
    return {"value": n ** 3 * steps / dt / 1e6, "unit": "MLUPS", "cores": int(O.lib().lbo_num_threads()),
            "kind": "port", "sample": "%d^3 D3Q%d bounce-back obstacle, %d steps in %.1f s (oracle/lbm_oracle.c, OpenMP)" % (n, q, steps, dt)}


def run_ours(args):
    import torch
    import latticeboltzmanncuda_b200 as lb
    from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask

    rank, world, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        # stdout carries ONE JSON line: NCCL writes its debug output (the version banner at WARN and above) to stdout
        # unless told otherwise, so whatever level the caller asked for goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")   # halo transfers must get SM slots next to the interior kernel
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    q, S = args.q, args.size
    tau = 0.6
    # global domain: a channel of N boxes of S^3 cells in a row along x, one box (x-slab) per GPU, each with its own
    # copy of the obstacle, so that every GPU has exactly the N = 1 work (weak scaling; cf. BASELINE.json configs[3])
    gx, gy, gz = world * S, S, S
    nx = gx // world
    x_off = rank * nx
    cells_total = gx * gy * gz
    # the CLI injects 0.001 per population per step into a fluid of density 0.028 with open (vacuum) edges;
    # here the fluid starts at density 1 with a uniform velocity, the x direction is periodic and the same
    # source shape at 0.1 % of the amplitude keeps the inflow phase in the step without driving it unstable
    flow = source_flow(q, 0.001)
    u0 = 0.04

    bc_x = lb.BC_PERIODIC if world == 1 else lb.BC_HALO
    kw = dict(q=q, bc=(bc_x, lb.BC_WALL, lb.BC_WALL), moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
              boundary_mode=lb.BOUNDARY_BOUNCE_BACK, device=local)
    if world == 1:
        lat = lb.Lattice(nx, gy, gz, None, None, lb.FluidData(1.0, tau), 1.0 / gx, **kw)
        slab = None
    else:
        from latticeboltzmanncuda_b200.slab import SlabLattice
        slab = SlabLattice(gx, gy, gz, lb.FluidData(1.0, tau), rank=rank, world=world, device=local,
                           bc_x=lb.BC_PERIODIC, **{k: v for k, v in kw.items() if k not in ("bc", "device")},
                           bc_yz=(lb.BC_WALL, lb.BC_WALL))
        lat = slab.lattice

    # synthetic inputs: obstacle mask + uniform initial state (host, pinned) + inflow plane source at global x == 0
    solid = tapered_hexahedron_mask(nx, gy, gz)
    if world > 1:
        # the neighbours' edge planes (periodic tiling of the same box): the obstacle keeps clear of the box faces
        ghost = np.zeros((1, gy, gz), np.uint8)
        solid = np.concatenate([ghost, solid, ghost], axis=0)
    w = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12, np.float32) if q == 19 else \
        np.array([8 / 27] + [2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8, np.float32)
    host_f = torch.empty((q, nx, gy, gz), dtype=torch.float32, pin_memory=True)
    hf = host_f.numpy()
    interior = solid[1:-1] if world > 1 else solid
    live_local = int(interior.size - int(interior.sum()))
    cx = np.array([0, -1, 0, 0, 0, 0, 1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1, 1, 1, 1, 1], np.float32)
    for i in range(q):
        hf[i] = w[i] * (1.0 + 3.0 * cx[i] * u0 + 4.5 * (cx[i] * u0) ** 2 - 1.5 * u0 * u0)   # equilibrium at rho = 1, u = (u0, 0, 0)
        hf[i][interior.astype(bool)] = 0.0

    stepper = slab if slab is not None else lat
    lat.set_solid_mask(solid, x_ghost_planes=(world > 1))
    stepper.load_populations(hf)
    if x_off == 0:
        lat.setFlowData(lb.FlowData(((0, None, None), (0, None, None)), flow))
    lat.setReflectionData(lb.ReflectionData())

    def barrier():
        torch.cuda.synchronize()
        lat.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K steps
    stepper.step(args.warmup)
    barrier()
    with ClockSampler(local) as clk:
        barrier()
        ms, launches = stepper.step_timed(args.steps)
        barrier()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    mlups = cells_total * args.steps / (ms * 1e-3) / 1e6
    b_alg = 2 * q * 4
    peak, peak_src = peak_hbm()
    live_total = live_local
    if world > 1:
        t = torch.tensor([live_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        live_total = int(t.item())
    # algorithmic bytes: every live (non-solid) cell reads Q and writes Q populations; solid cells move nothing
    per_gpu_gbs = (live_total / world) * args.steps * b_alg / (ms * 1e-3) / 1e9
    mass = lat.mass()
    finite = bool(np.isfinite(mass))
    state_gb = lat.layout().state_bytes / 1e9

    # ---- end to end through the host API: upload (pinned) + K x (step + force read-back) + moments download
    e2e = None
    if not args.no_e2e:
        n_local = nx * gy * gz
        host_m = torch.empty((4, n_local), dtype=torch.float32, pin_memory=True).numpy()
        lib, h = lb.capi.load(), lat.handle
        force = np.zeros(3, np.float32)
        barrier()
        t0 = time.perf_counter()
        stepper.load_populations(hf)
        for _ in range(args.steps):
            if x_off == 0:
                lat.setFlowData(lb.FlowData(((0, None, None), (0, None, None)), flow))   # this step's inflow, from the host
            stepper.step(1)
            lb.capi.check(lib.lbm_get_force(h, lb.capi.ptr(force)))                         # the CLI prints this every step
        lb.capi.check(lib.lbm_moments(h, lb.capi.ptr(host_m[0]), lb.capi.ptr(host_m[1]), lb.capi.ptr(host_m[2]),
                                      lb.capi.ptr(host_m[3])))
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": cells_total * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(hf.nbytes / args.steps + 27 * 4 + 24),
               "d2h_bytes_per_step": int(host_m.nbytes / args.steps + 12),
               "what": "lbm_upload_f_soa from pinned host + %d x (lbm_set_source, lbm_step, lbm_get_force) + lbm_moments to pinned host" % args.steps,
               "seconds": dt}

    # ---- the two dense engines on an all-fluid lattice of the same size (context for the roofline; N = 1 only)
    extra = None
    if world == 1 and not args.no_extra:
        extra = {}
        lat.close()
        for name, engine, bc in (("dense_two_buffer", lb.ENGINE_PULL2, (lb.BC_PERIODIC, lb.BC_PERIODIC, lb.BC_ZERO_INFLOW)),
                                 ("in_place_single_buffer", lb.ENGINE_INPLACE, (lb.BC_PERIODIC,) * 3)):
            l2 = lb.Lattice(nx, gy, gz, None, None, lb.FluidData(1.0, tau), 1.0 / gx, q=q, bc=bc, moment_mode=lb.MOMENT_STANDARD,
                            eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK, device=local, engine=engine)
            for i in range(q):
                hf[i] = w[i]
            l2.load_populations(hf)
            l2.step(4)
            l2.sync()
            ms2, _ = l2.step_timed(40)
            gbs2 = cells_total * 40 * b_alg / (ms2 * 1e-3) / 1e9
            extra[name] = {"mlups": cells_total * 40 / (ms2 * 1e-3) / 1e6, "gbs": gbs2, "frac": gbs2 / peak,
                           "state_gb": l2.layout().state_bytes * l2.layout().n_buffers / 1e9}
            l2.close()
        # the timed workload itself (obstacle, walls, plane source, force) on the single-buffer engine: half the memory
        l4 = lb.Lattice(nx, gy, gz, None, None, lb.FluidData(1.0, tau), 1.0 / gx, engine=lb.ENGINE_INPLACE,
                        **{**kw, "bc": (lb.BC_PERIODIC, lb.BC_WALL, lb.BC_WALL)})
        for i in range(q):
            hf[i] = w[i] * (1.0 + 3.0 * cx[i] * u0 + 4.5 * (cx[i] * u0) ** 2 - 1.5 * u0 * u0)
            hf[i][interior.astype(bool)] = 0.0
        l4.set_solid_mask(solid)
        l4.load_populations(hf)
        l4.setFlowData(lb.FlowData(((0, None, None), (0, None, None)), flow))
        l4.setReflectionData(lb.ReflectionData())
        l4.step(4)
        l4.sync()
        ms4, n4 = l4.step_timed(40)
        gbs4 = live_local * 40 * b_alg / (ms4 * 1e-3) / 1e9
        extra["obstacle_in_place_single_buffer"] = {
            "mlups": cells_total * 40 / (ms4 * 1e-3) / 1e6, "gbs": gbs4, "frac": gbs4 / peak, "ms_per_step": ms4 / 40,
            "launches_per_step": n4 / 40, "state_gb": l4.layout().state_bytes * l4.layout().n_buffers / 1e9,
            "mass_finite": bool(np.isfinite(l4.mass()))}
        l4.close()

    # ---- the reference arm's own workload in the reference's own mode (N = 1 only): D3Q27, SPLIT moments, fp64
    # equilibrium, zero-inflow edges, table reflection built from the OBJ fixture on the GPU, plane source, force
    # accumulator, driven through the four per-phase calls like the CLI -- the like-for-like number for --impl reference
    refmode = None
    if world == 1 and not args.no_extra:
        n = args.ref_size
        obj = os.path.join(ROOT, "tests", "golden", "obj", "cube_angled.obj")
        flow27 = source_flow(27)
        l3 = lb.Lattice(n, n, n, None, None, lb.FluidData(1.0, tau), float(np.float32(1.0 / n)), device=local)   # defaults = reference behaviour
        start = np.zeros(n * n * n, dtype=lb.LATTICE_POINT)
        start["f"] = flow27[None, :]
        l3.load_data(start)
        del start
        t_geo = time.perf_counter()
        l3.insertModel(obj)
        l3.preProcessModel()
        l3.sync()
        t_geo = time.perf_counter() - t_geo
        l3.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow27))
        l3.setReflectionData(lb.ReflectionData())
        l3.step(3)
        l3.sync()
        ms3, _ = l3.step_timed(20)
        st3 = l3.stats()
        refmode = {"mlups": n ** 3 * 20 / (ms3 * 1e-3) / 1e6, "ms_per_step": ms3 / 20,
                   "config": "%d^3 D3Q27, reference arithmetic (fp64 equilibrium, SPLIT moments), zero-inflow edges, table reflection "
                             "(%d links), plane source, force accumulator" % (n, st3["links"]),
                   "geometry_seconds": t_geo, "geometry": "lbm_insert_model + lbm_preprocess_model on the GPU (voxelise %d^3 cells, build tables)" % n}
        l3.close()

    out = None
    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("d3q%d_%d" % (q, S))
            except Exception:
                traffic = None
        out = {
            "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "flow past obstacle (reference test geometry, tapered hexahedron, one per 512-cell box) %dx%dx%d D3Q%d BGK, "
                                   "half-way bounce-back, periodic in x, walls in y/z, initial u=(0.04,0,0), plane source at x=0; "
                                   "%d x-slab(s) of %dx%dx%d" % (gx, gy, gz, q, world, nx, gy, gz),
                       "cells_per_gpu": nx * gy * gz, "tau": tau, "engine": "two-buffer fused pull, classified 32-cell words",
                       "l2": "state per GPU %.1f GB >> 126 MB L2, no flush needed" % state_gb},
            "gpu_launches": int(launches),
            "live_cells": live_total, "mflups": live_total * args.steps / (ms * 1e-3) / 1e6,
            "hbm_gbs_per_gpu": per_gpu_gbs,
            "roofline": {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_cell": b_alg,
                         "algorithmic_bytes_per_launch": live_total // world * b_alg,
                         "kernel": "step_general_kernel<%d> + step_special_kernel<%d> (one step)" % (q, q),
                         "note": "achieved = 2*Q*4 B x live (non-solid) cells per step / CUDA-event time per step; "
                                 "solid cells move no bytes and are not counted"},
            "clocks": clk.summary(),
            "state_finite": finite,
        }
        if e2e:
            out["e2e"] = e2e
        if extra:
            out["other_engines"] = extra
        if refmode:
            out["reference_mode"] = refmode
    if world > 1:
        dist.barrier()
    if rank == 0 and not args.no_cpu and world == 1:
        out["cpu_baseline"] = cpu_baseline(q)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# -------------------------------------------------------------------------------------------------
# reference arm: the reference's own Lattice class and CUDA kernels (oracle/_ref/libref_gpu.so)
# -------------------------------------------------------------------------------------------------

def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    base = {"impl": "reference", "metric": "MLUPS", "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic"}
    try:
        import _oracle as O
        import _ref_gpu as R
        if not R.available():
            raise RuntimeError("oracle/_ref/libref_gpu.so not built")
        n = args.ref_size
        spacing = float(np.float32(1.0 / n))
        flow = source_flow(27)
        ref = R.RefLattice(n, n, n, tau=0.6, spacing=spacing)
        cells = np.zeros(n * n * n, dtype=O.LATTICE_POINT)
        cells["f"] = flow[None, :]
        ref.load(cells)
        del cells
        ref.insert_model(os.path.join(O.GOLDEN_DIR, "obj", "cube_angled.obj"))
        ref.preprocess()
        ref.set_flow(2, flow)
        ref.set_reflection_data()
        ref.prime()
        ref.run_cli(args.warmup)
        R.lib().ref_sync()
        t0 = time.perf_counter()
        ref.run_cli(args.steps)          # the CLI loop incl. its per-step force read-back; ends with a device sync
        dt = time.perf_counter() - t0
        v = n ** 3 * args.steps / dt / 1e6
        base.update({
            "value": v, "ms_per_step": dt / args.steps * 1e3,
            "config": {"workload": "flow past obstacle (reference test geometry) D3Q27 BGK, table reflection, plane source at x=0, "
                                   "zero-inflow edges: the reference's own Lattice class + CUDA kernels compiled unmodified for "
                                   "sm_100a; bounded sample %d^3 of the 512^3 workload (the reference holds 2 x 144 B/cell on the "
                                   "device and leaks a 144 B/cell host array per retrieve_data call)" % n},
            "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": 0, "kind": "reference",
                             "sample": "%d^3, %d steps; the reference has no CPU path (SURVEY.md D2): this is its GPU path on the "
                                       "same B200, host loop of LatticeBoltzmannCuda.cpp:85-93" % (n, args.steps)},
            "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
    except Exception as exc:      # noqa: BLE001
        base = {"impl": "reference", "unavailable": "%s: %s" % (type(exc).__name__, exc)}
    print(json.dumps(base))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
