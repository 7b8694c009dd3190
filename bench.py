#!/usr/bin/env python
"""bench.py -- MLUPS and achieved HBM bandwidth of the fused collide-stream step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--q 19|27]

One "step" is one fused lattice update (stream + bounce-back + BGK collide + source) of the whole
lattice on ONE copy of the populations (single-buffer "AA" streaming, the north star's engine).
Workload at N = 1: BASELINE.json configs[2], "flow past obstacle voxelised from the reference test
geometry, 512^3, bounce-back", D3Q19 fp32 -- the configuration the 80%-of-HBM target is quoted on.
At N > 1 every GPU holds a 128 x 1024 x 1024 x-slab (the same 2^27 cells as 512^3) of a
128N x 1024 x 1024 domain -- N = 8 is the slab-decomposed 1024^3 domain of configs[4] -- with the
same obstacle squeezed into a 128 x 512 x 512 box and tiled 2 x 2 over the cross-section, so that
every GPU has the same work (weak scaling).  The slabs are joined by peer-memory halos below the
C ABI (lbm_halo_connect): no communication library runs in the step.

Prints ONE JSON line (see the task contract): value = whole-job MLUPS over the LIVE (non-solid)
cells -- solid cells move no bytes -- with the state resident in HBM; e2e = the same job through
the host API with pinned host buffers (state upload, per-step force read-back like the reference
CLI, moments download) inside the timed region; roofline = algorithmic bytes (2*Q*4 per live cell
update) over the CUDA-event time of the step against the measured copy bandwidth in
MEASURED_PEAKS.json; cpu_baseline = the CPU oracle port on this box's cores on BASELINE.json
configs[0] (64^3 reference geometry, the reference's own arithmetic).
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
    ap.add_argument("--size", type=int, default=512, help="N = 1: cells per axis")
    ap.add_argument("--slab", default="128,1024,1024", help="N > 1: cells of one GPU's x-slab (nx,ny,nz)")
    ap.add_argument("--workload", default="auto", choices=["auto", "slab"],
                    help="slab: run one slab of the N > 1 workload on ONE GPU (the per-GPU rate the scaling run is compared with)")
    ap.add_argument("--q", type=int, default=19, choices=[19, 27])
    ap.add_argument("--engine", default="inplace", choices=["inplace", "pull2"],
                    help="inplace = single-buffer AA streaming (headline), pull2 = two-buffer fused pull")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = peer-memory halos below the C ABI, nccl = torch.distributed P2P of the edge planes")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements (other engines, reference mode, config 1)")
    ap.add_argument("--ref-size", type=int, default=int(os.environ.get("LBM_REF_SIZE", "512")))
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

def config1_inputs(n=64):
    """BASELINE.json configs[0]: the reference test geometry on a 64^3 grid as the CLI sets it up
    (LatticeBoltzmannCuda.cpp:29-76): all f = 0.001, f[6] = 0.002, plane source at x == 0, force accumulator"""
    flow = source_flow(27)
    obj = os.path.join(ROOT, "tests", "golden", "obj", "cube_angled.obj")
    return n, float(np.float32(1.0 / n)), flow, obj


def cpu_baseline(target_seconds=10.0):
    """the CPU oracle port (oracle/lbm_oracle.c, OpenMP) on BASELINE.json configs[0]: 64^3 D3Q27, the reference's own
    arithmetic (table reflection, fp64 equilibrium, SPLIT moments, zero-inflow edges), CLI loop; all cores for
    ~target_seconds, then one thread for a fifth of that"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    n, spacing, flow, obj = config1_inputs()
    cells = O.new_cells(n, n, n)
    cells["f"] = flow[None, :]
    mesh = O.Mesh(obj)
    mesh.insert(cells, n, n, n, np.float32(spacing))
    tables, tindex = mesh.preprocess(cells, n, n, n, np.float32(spacing))
    force = np.zeros(3, np.float32)

    def run(seconds, cap):
        cur, steps, t0 = cells, 0, time.perf_counter()
        while True:
            cur = O.step_cli(cur, n, n, n, 1.0, tables, tindex, force, O.SRC_PLANE_X0, flow, steps=10)
            steps += 10
            dt = time.perf_counter() - t0
            if dt >= seconds or steps >= cap:
                return steps, dt

    cores = int(O.lib().lbo_num_threads())
    run(0.5, 10)                                                       # warm-up
    steps, dt = run(target_seconds, 1000)
    out = {"value": n ** 3 * steps / dt / 1e6, "unit": "MLUPS", "cores": cores, "kind": "port",
           "sample": "BASELINE configs[0]: %d^3 D3Q27 reference mode (table reflection, fp64 equilibrium), CLI loop, %d of 1000 steps in "
                     "%.1f s (oracle/lbm_oracle.c, OpenMP, %d threads)" % (n, steps, dt, cores)}
    if hasattr(O.lib(), "lbo_set_num_threads"):
        O.lib().lbo_set_num_threads(1)
        s1, d1 = run(target_seconds / 5.0, 200)
        O.lib().lbo_set_num_threads(cores)
        out["one_thread_mlups"] = n ** 3 * s1 / d1 / 1e6
    return out


def obstacle_mask(world, nx, gy, gz):
    """N = 1: the reference's tapered hexahedron in the box.  N > 1: the same solid squeezed into an nx x gy/2 x gz/2 box and
    tiled 2 x 2 over the slab's cross-section -- every slab holds the same four bodies, none touches a slab face"""
    from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask
    if world == 1:
        return tapered_hexahedron_mask(nx, gy, gz)
    return np.tile(tapered_hexahedron_mask(nx, gy // 2, gz // 2), (1, 2, 2))


def run_ours(args):
    import torch
    import latticeboltzmanncuda_b200 as lb

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
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    q, tau = args.q, 0.6
    slab_shape = world > 1 or args.workload == "slab"
    if not slab_shape:
        nx = gy = gz = args.size
    else:
        nx, gy, gz = (int(v) for v in args.slab.split(","))
    gx = nx * world
    x_off = rank * nx
    cells_total = gx * gy * gz
    engine = lb.ENGINE_INPLACE if args.engine == "inplace" else lb.ENGINE_PULL2
    # the CLI injects 0.001 per population per step into a fluid of density 0.028 with open (vacuum) edges;
    # here the fluid starts at density 1 with a uniform velocity, the x direction is periodic and the same
    # source shape at 0.1 % of the amplitude keeps the inflow phase in the step without driving it unstable
    flow = source_flow(q, 0.001)
    u0 = 0.04

    kw = dict(q=q, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK, engine=engine)
    if world == 1:
        lat = lb.Lattice(nx, gy, gz, None, None, lb.FluidData(1.0, tau), 1.0 / gx, bc=(lb.BC_PERIODIC, lb.BC_WALL, lb.BC_WALL),
                         device=local, **kw)
        slab = None
    else:
        from latticeboltzmanncuda_b200.slab import SlabLattice
        slab = SlabLattice(gx, gy, gz, lb.FluidData(1.0, tau), rank=rank, world=world, device=local, bc_x=lb.BC_PERIODIC,
                           bc_yz=(lb.BC_WALL, lb.BC_WALL), transport=args.transport, **kw)
        lat = slab.lattice

    # synthetic inputs: obstacle mask + uniform initial state (host, pinned) + inflow plane source at global x == 0
    interior = obstacle_mask(2 if slab_shape else 1, nx, gy, gz)
    solid = interior
    if world > 1:
        ghost = np.zeros((1, gy, gz), np.uint8)          # the bodies keep clear of the slab faces
        solid = np.concatenate([ghost, interior, ghost], axis=0)
    w = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12, np.float32) if q == 19 else \
        np.array([8 / 27] + [2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8, np.float32)
    host_f = torch.empty((q, nx, gy, gz), dtype=torch.float32, pin_memory=True)
    hf = host_f.numpy()
    live_local = int(interior.size - int(interior.sum()))
    cx = np.array([0, -1, 0, 0, 0, 0, 1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1, 1, 1, 1, 1], np.float32)
    fluid = interior == 0

    def fill_state(with_body=True):
        for i in range(q):
            v = w[i] * (1.0 + 3.0 * cx[i] * u0 + 4.5 * (cx[i] * u0) ** 2 - 1.5 * u0 * u0)   # equilibrium at rho = 1, u = (u0, 0, 0)
            if with_body:
                np.multiply(fluid, np.float32(v), out=hf[i], casting="unsafe")
            else:
                hf[i] = v

    fill_state()
    stepper = slab if slab is not None else lat
    lat.set_solid_mask(solid, x_ghost_planes=(world > 1))
    stepper.load_populations(hf)

    def set_inflow(l):
        if x_off == 0:
            l.setFlowData(lb.FlowData(((0, None, None), (0, None, None)), flow))

    set_inflow(lat)
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
    b_alg = 2 * q * 4
    peak, peak_src = peak_hbm()
    live_total = live_local
    if world > 1:
        t = torch.tensor([live_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        live_total = int(t.item())
    # every live (non-solid) cell reads Q and writes Q populations; solid cells move nothing and are not counted
    mflups = live_total * args.steps / (ms * 1e-3) / 1e6
    per_gpu_gbs = (live_total / world) * args.steps * b_alg / (ms * 1e-3) / 1e9
    mass = lat.mass() if world == 1 else slab.mass()
    finite = bool(np.isfinite(mass))
    lay = lat.layout()
    state_gb = lay.state_bytes * lay.n_buffers / 1e9

    # ---- end to end through the host API: upload (pinned) + K x (set source, step, force read-back) + moments download
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
            set_inflow(lat)                                                                    # this step's inflow, from the host
            stepper.step(1)
            lb.capi.check(lib.lbm_get_force(h, lb.capi.ptr(force)))                            # the CLI prints this every step
        if slab is not None:
            slab.sync()
        lb.capi.check(lib.lbm_moments(h, lb.capi.ptr(host_m[0]), lb.capi.ptr(host_m[1]), lb.capi.ptr(host_m[2]),
                                      lb.capi.ptr(host_m[3])))
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": live_total * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(hf.nbytes / args.steps + 27 * 4 + 24),
               "d2h_bytes_per_step": int(host_m.nbytes / args.steps + 12),
               "what": "lbm_upload_f_soa from pinned host + %d x (lbm_set_source, lbm_step, lbm_get_force) + lbm_moments to pinned host" % args.steps,
               "seconds": dt, "mlups_all_cells": cells_total * args.steps / dt / 1e6}

    extra, refmode, config1 = None, None, None
    if world == 1 and not args.no_extra:
        lat.close()
        extra = {}

        def probe(name, eng, bc, body):
            l = lb.Lattice(nx, gy, gz, None, None, lb.FluidData(1.0, tau), 1.0 / gx, q=q, bc=bc, moment_mode=lb.MOMENT_STANDARD,
                           eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK, device=local, engine=eng)
            fill_state(with_body=body)
            if body:
                l.set_solid_mask(interior)
            l.load_populations(hf)
            if body:
                set_inflow(l)
                l.setReflectionData(lb.ReflectionData())
            l.step(4)
            l.sync()
            ms2, n2 = l.step_timed(40)
            cells = live_local if body else cells_total
            gbs2 = cells * 40 * b_alg / (ms2 * 1e-3) / 1e9
            extra[name] = {"mflups": cells * 40 / (ms2 * 1e-3) / 1e6, "gbs": gbs2, "frac": gbs2 / peak, "ms_per_step": ms2 / 40,
                           "launches_per_step": n2 / 40, "state_gb": l.layout().state_bytes * l.layout().n_buffers / 1e9}
            l.close()

        # the same workload on the other engine, and both engines on an all-fluid lattice of the same size
        other = lb.ENGINE_PULL2 if engine == lb.ENGINE_INPLACE else lb.ENGINE_INPLACE
        probe("obstacle_two_buffer" if other == lb.ENGINE_PULL2 else "obstacle_in_place_single_buffer", other,
              (lb.BC_PERIODIC, lb.BC_WALL, lb.BC_WALL), True)
        probe("dense_two_buffer", lb.ENGINE_PULL2, (lb.BC_PERIODIC, lb.BC_PERIODIC, lb.BC_ZERO_INFLOW), False)
        probe("in_place_single_buffer", lb.ENGINE_INPLACE, (lb.BC_PERIODIC,) * 3, False)

        # ---- the reference arm's own workload in the reference's own mode: D3Q27, SPLIT moments, fp64 equilibrium,
        # zero-inflow edges, table reflection built from the OBJ fixture on the GPU, plane source, force accumulator --
        # the like-for-like number for --impl reference (same size, same arithmetic, same phases)
        n = args.ref_size
        obj = os.path.join(ROOT, "tests", "golden", "obj", "cube_angled.obj")
        flow27 = source_flow(27)
        try:
            l3 = lb.Lattice(n, n, n, None, None, lb.FluidData(1.0, tau), float(np.float32(1.0 / n)), device=local)   # defaults = reference behaviour
            f27 = torch.empty((27, n, n, n), dtype=torch.float32, device="cuda")
            for i in range(27):
                f27[i] = float(flow27[i])
            l3.load_populations(f27)
            del f27
            torch.cuda.empty_cache()
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
                                 "(%d links), plane source, force accumulator: the --impl reference workload on this library" % (n, st3["links"]),
                       "geometry_seconds": t_geo, "geometry": "lbm_insert_model + lbm_preprocess_model on the GPU (voxelise %d^3 cells, build tables)" % n}
            l3.close()
        except Exception as exc:      # noqa: BLE001
            refmode = {"error": "%s: %s" % (type(exc).__name__, exc)}

        # ---- BASELINE.json configs[0] on the GPU: 64^3, 1000 steps of the CLI loop in the reference's own mode.
        # 2.6e5 cells are 512 blocks of work: too little to fill 148 SMs (a CUDA graph of the step changes nothing,
        # profiles/r02_notes.md), so this line says nothing about bandwidth.
        n1, spacing1, flow1, obj1 = config1_inputs()
        l5 = lb.Lattice(n1, n1, n1, None, None, lb.FluidData(1.0, 1.0), spacing1, device=local)
        start = np.zeros(n1 ** 3, dtype=lb.LATTICE_POINT)
        start["f"] = flow1[None, :]
        l5.load_data(start)
        l5.insertModel(obj1)
        l5.preProcessModel()
        l5.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow1))
        l5.setReflectionData(lb.ReflectionData())
        l5.step(10)
        l5.sync()
        ms5, n5 = l5.step_timed(1000)
        config1 = {"workload": "BASELINE configs[0]: reference test geometry %d^3 D3Q27, reference mode, 1000 steps" % n1,
                   "mlups": n1 ** 3 * 1000 / (ms5 * 1e-3) / 1e6, "ms_per_step": ms5 / 1000, "launches_per_step": n5 / 1000,
                   "note": "too small to fill the GPU (512 blocks per launch); not a bandwidth figure"}
        l5.close()
        # the same configuration through the C++ drop-in: the reference's `main` rebuilt against the compatibility headers
        # (include_compat/lbm_cli.cpp), its step loop timed by the program itself, force trace printed every 100 steps
        cli = os.path.join(ROOT, "latticeboltzmanncuda_b200", "include_compat", "lbm_cli")
        if os.path.exists(cli):
            import re
            import subprocess
            try:
                res = subprocess.run([cli, obj1, "8", "8", "8", str(n1 // 8), str(n1 // 8), str(n1 // 8), "1.0", "1.0", repr(spacing1), "1000",
                                      "100", "--time", "1"], capture_output=True, text=True, timeout=300)
                m = re.search(r"in ([0-9.]+) s = ([0-9.]+) MLUPS", res.stderr)
                if m:
                    config1["cpp_cli"] = {"seconds": float(m.group(1)), "mlups": float(m.group(2)),
                                          "what": "lbm_cli (the reference's main on the C++ compatibility layer), 1000 steps, four "
                                                  "simulate* calls per step fused into one, force read back every 100 steps"}
            except Exception as exc:      # noqa: BLE001
                config1["cpp_cli"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    out = None
    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("%s_d3q%d_%d" % (args.engine, q, nx if world == 1 else 0))
            except Exception:
                traffic = None
        eng_name = ("single-buffer AA streaming (one copy of the populations), classified 32-cell words" if engine == lb.ENGINE_INPLACE
                    else "two-buffer fused pull, classified 32-cell words")
        kern = ("aa_odd_kernel<%d> + aa_odd_special_kernel<%d> / aa_even_kernel<%d> + aa_even_special_kernel<%d> (odd / even step)" % (q, q, q, q)
                if engine == lb.ENGINE_INPLACE else "step_general_kernel<%d> + step_special_kernel<%d> (one step)" % (q, q))
        if not slab_shape:
            wl = ("flow past obstacle (reference test geometry, tapered hexahedron) %dx%dx%d D3Q%d BGK, half-way bounce-back, periodic in x, "
                  "walls in y/z, initial u=(0.04,0,0), plane source at x=0, force accumulator on" % (gx, gy, gz, q))
        else:
            wl = ("flow past obstacles %dx%dx%d D3Q%d BGK (N=8: the 1024^3 domain), %d x-slabs of %dx%dx%d joined by %s; per slab four copies "
                  "of the reference test geometry (squeezed to %dx%dx%d, tiled 2x2 in y/z), half-way bounce-back, periodic in x, walls in "
                  "y/z, initial u=(0.04,0,0), plane source at x=0, force accumulator on"
                  % (gx, gy, gz, q, world, nx, gy, gz,
                     "peer-memory halos below the C ABI (no communication library in the step)" if args.transport == "peer"
                     else "NCCL send/recv of the edge planes", nx, gy // 2, gz // 2))
        out = {
            "metric": "MLUPS", "value": mflups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "cells_per_gpu": nx * gy * gz, "live_cells_per_gpu": live_total // world, "tau": tau,
                       "engine": eng_name, "transport": (args.transport if world > 1 else None),
                       "counting": "value counts live (non-solid) cell updates only; mlups_all_cells counts every lattice cell",
                       "parity": "extension of the reference (D3Q19, half-way bounce-back, periodic / wall edges: the reference has "
                                 "D3Q27, table reflection and zero-inflow edges only), checked against the CPU oracle; the reference's "
                                 "own mode on this library is `reference_mode`, checked against the reference's kernels",
                       "l2": "state per GPU %.1f GB >> 126 MB L2, no flush needed" % state_gb},
            "gpu_launches": int(launches),
            "live_cells": live_total, "mlups_all_cells": cells_total * args.steps / (ms * 1e-3) / 1e6,
            "hbm_gbs_per_gpu": per_gpu_gbs,
            "roofline": {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_cell": b_alg,
                         "algorithmic_bytes_per_launch": live_total // world * b_alg,
                         "kernel": kern,
                         "note": "achieved = 2*Q*4 B x live (non-solid) cells per step / CUDA-event time per step, per GPU; "
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
        if config1:
            out["config1_64cubed"] = config1
    if world > 1:
        dist.barrier()
    if rank == 0 and not args.no_cpu and world == 1:
        out["cpu_baseline"] = cpu_baseline()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# -------------------------------------------------------------------------------------------------
# reference arm: the reference's own Lattice class and CUDA kernels (oracle/_ref/libref_gpu.so)
# -------------------------------------------------------------------------------------------------

def run_reference(args):
    """The reference's own Lattice class and CUDA kernels, compiled unmodified for sm_100a (oracle/Makefile), driven through
    its CLI loop on the repo arm's lattice size.  The K timed steps are split into three repetitions; `value` is the median
    repetition (the reference allocates and frees two whole lattices per step, which makes single runs noisy)."""
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
        t_setup = time.perf_counter()
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
        t_setup = time.perf_counter() - t_setup
        ref.run_cli(args.warmup)
        R.lib().ref_sync()
        reps = [args.steps // 3 + (1 if r < args.steps % 3 else 0) for r in range(3)]
        rates, total = [], 0.0
        for k in reps:
            if k == 0:
                continue
            t0 = time.perf_counter()
            ref.run_cli(k)                   # the CLI loop incl. its per-step force read-back; ends with a device sync
            dt = time.perf_counter() - t0
            total += dt
            rates.append(n ** 3 * k / dt / 1e6)
        v = float(np.median(rates))
        base.update({
            "value": v, "ms_per_step": n ** 3 / v / 1e3,
            "repetitions": {"steps": [k for k in reps if k], "mlups": rates, "statistic": "median", "seconds_total": total},
            "setup_seconds": t_setup,
            "config": {"workload": "flow past obstacle (reference test geometry) %dx%dx%d D3Q27 BGK, table reflection, plane source at x=0, "
                                   "zero-inflow edges: the reference's own Lattice class + CUDA kernels compiled unmodified for sm_100a, "
                                   "driven through its CLI loop (LatticeBoltzmannCuda.cpp:85-93); the reference has one velocity set "
                                   "(D3Q27), one boundary model and one edge rule, so Q / boundary / edges are the reference's, the "
                                   "lattice size is the repo arm's.  The same workload on this library: `reference_mode` in the "
                                   "repo arm's line" % (n, n, n)},
            "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": 0, "kind": "reference",
                             "sample": "%d^3, %d steps in 3 repetitions (median); the reference has no CPU path (SURVEY.md D2): this is its "
                                       "GPU path on the same B200, host loop of LatticeBoltzmannCuda.cpp:85-93" % (n, args.steps)},
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
