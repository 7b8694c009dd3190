"""Slab decomposition of the lattice along x across the GPUs of one box (one process per GPU).

The reference is single-GPU (no cudaSetDevice, no NCCL/MPI anywhere); this is a build
extension with one requirement: the same answer as one GPU.  x is the slowest axis of the SoA layout
f[q][x][y][z], so a ghost plane of one population is one contiguous run of yp*zp floats and needs no
packing.  Per step every rank hands the populations that leave through its +x face (c_x = +1) to the low
ghost plane of its right neighbour and those that leave through its -x face (c_x = -1) to the high ghost
plane of its left neighbour: 5 planes per face for D3Q19, 9 for D3Q27.

`HaloExchange` holds the neighbour logic and runs on any torch.distributed backend (NCCL on the GPUs,
gloo in the CPU tests).  `SlabLattice` drives one `Lattice` handle per rank: edge planes first, then the
exchange of the new edge planes into the neighbours' ghost planes, overlapped with the interior planes.
"""
import ctypes as C
import os

import numpy as np

from . import capi
from .lattice import FlowData, Lattice

_DIRS_X = np.array([0, -1, 0, 0, 0, 0, 1, -1, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 1, -1, -1, -1, -1, 1, 1, 1, 1])


def face_populations(q, side):
    """populations that cross the +x face (side=1, c_x=+1) or the -x face (side=0, c_x=-1)"""
    want = 1 if side else -1
    return [i for i in range(q) if _DIRS_X[i] == want]


def slab_range(gx, rank, world):
    """[x0, x1) of the global x planes owned by `rank` (as even as possible, larger slabs first)"""
    base, extra = divmod(gx, world)
    x0 = rank * base + min(rank, extra)
    return x0, x0 + base + (1 if rank < extra else 0)


class HaloExchange:
    """who sends which plane to whom.  `periodic` closes the ring; otherwise the end slabs keep their
    outer ghost planes as the domain-edge rule left them (zero inflow)."""

    def __init__(self, rank, world, periodic, q):
        self.rank, self.world, self.periodic, self.q = rank, world, bool(periodic), q
        self.right = (rank + 1) % world if (periodic or rank + 1 < world) else None
        self.left = (rank - 1) % world if (periodic or rank > 0) else None
        self.up = face_populations(q, 1)      # leave through +x, enter the right neighbour's low ghost plane
        self.down = face_populations(q, 0)    # leave through -x, enter the left neighbour's high ghost plane

    def transfers(self, send_plane, recv_plane):
        """list of ("send"|"recv", tensor, peer, tag).  send_plane(q, which) gives my edge plane of population
        q (which = "lo" first interior plane, "hi" last interior plane); recv_plane(q, which) gives my ghost plane
        (which = "lo" ghost below x=0, "hi" ghost above x=nx-1).  Tags keep the two directions apart when
        left == right (two ranks, periodic)."""
        # NCCL matches the sends and receives of a rank pair in the order they are posted and ignores tags,
        # so every rank posts the rightward wave first and the leftward wave second; with two ranks on a
        # ring (left == right) the pair then sees "up, down" on both ends.
        ops = []
        if self.right is not None:
            for i in self.up:
                ops.append(("send", send_plane(i, "hi"), self.right, i))
        if self.left is not None:
            for i in self.up:
                ops.append(("recv", recv_plane(i, "lo"), self.left, i))
            for i in self.down:
                ops.append(("send", send_plane(i, "lo"), self.left, 100 + i))
        if self.right is not None:
            for i in self.down:
                ops.append(("recv", recv_plane(i, "hi"), self.right, 100 + i))
        return ops

    def run(self, send_plane, recv_plane, dist):
        """posts every transfer of one exchange and returns the work handles"""
        ops = self.transfers(send_plane, recv_plane)
        if not ops:
            return []
        p2p = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer, tag=tag) for kind, t, peer, tag in ops]
        return dist.batch_isend_irecv(p2p)


def allreduce_sum(values, dist, device=None):
    """sum of a small float64 vector over all ranks (mass, momentum[3], force[3] are reduced only
    on the steps that ask for them; the step path itself has no collective besides the halo exchange).
    `device` = the CUDA device for NCCL, None for a host tensor (gloo)."""
    import torch
    t = torch.as_tensor(np.asarray(values, np.float64).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


class _DevicePlane:
    """zero-copy view of a device plane for torch.as_tensor (CUDA array interface)"""

    def __init__(self, ptr, nfloat):
        self.__cuda_array_interface__ = {"shape": (nfloat,), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class SlabLattice:
    """one x-slab of a gx x gy x gz lattice, driven like a `Lattice`"""

    def __init__(self, gx, gy, gz, fluid, rank, world, device, bc_x=capi.BC_PERIODIC, bc_yz=(capi.BC_PERIODIC,) * 2,
                 overlap=True, transport="nccl", **kw):
        """transport: "nccl" = this module posts the edge planes with torch.distributed P2P every step (any backend, also
        gloo); "peer" = the library's own peer-memory halos (lbm_halo_connect: the edge cells address the neighbouring
        slab's memory over NVLink, nothing is copied and nothing runs between the steps; CUDA only, one box)."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.transport = transport
        if bc_x not in (capi.BC_PERIODIC, capi.BC_ZERO_INFLOW):
            raise ValueError("the global x edge must be periodic or zero-inflow in slab mode")
        self.rank, self.world = rank, world
        self.x0, self.x1 = slab_range(gx, rank, world)
        self.gx, self.nx = gx, self.x1 - self.x0
        if self.nx < 2:
            raise ValueError("every slab needs at least two x planes")
        self.lattice = Lattice(self.nx, gy, gz, None, None, fluid, 1.0, bc=(capi.BC_HALO, bc_yz[0], bc_yz[1]),
                               device=device, **kw)
        self.q = self.lattice.q
        # single-buffer engine: no ghost refresh per step; the edge planes travel around every odd step instead
        self.inplace = kw.get("engine") == capi.ENGINE_INPLACE
        self._aa_odd_next = True
        if self.inplace and bc_x != capi.BC_PERIODIC:
            raise ValueError("single-buffer slabs need a periodic global x axis")
        self.halo = HaloExchange(rank, world, bc_x == capi.BC_PERIODIC, self.q)
        self.overlap = bool(overlap) and not os.environ.get("LBM_SLAB_NOOVERLAP")
        self._lib = capi.load()
        self._views = {}
        # everything this handle launches is ordered on a torch stream, which NCCL transfers order against
        self.stream = torch.cuda.Stream(device=device)
        capi.check(self._lib.lbm_set_stream(self.lattice.handle, C.c_void_p(self.stream.cuda_stream)))
        self._ghosts_current = False
        if transport == "peer":
            self._connect_peers()
        elif transport != "nccl":
            raise ValueError("transport must be 'nccl' or 'peer'")

    def _connect_peers(self):
        """every rank exports its slab (CUDA IPC handles), the blobs travel once through the process group, and each rank
        connects to the slab below and above; from then on lbm_step does the whole step"""
        blob = (C.c_ubyte * capi.HALO_BLOB_BYTES)()
        capi.check(self._lib.lbm_halo_export(self.lattice.handle, blob, capi.HALO_BLOB_BYTES))
        blobs = [None] * self.world
        self.dist.all_gather_object(blobs, bytes(blob))
        lo = None if self.halo.left is None else (C.c_ubyte * capi.HALO_BLOB_BYTES).from_buffer_copy(blobs[self.halo.left])
        hi = None if self.halo.right is None else (C.c_ubyte * capi.HALO_BLOB_BYTES).from_buffer_copy(blobs[self.halo.right])
        capi.check(self._lib.lbm_halo_connect(self.lattice.handle, lo, hi))
        self.dist.barrier()              # nobody steps before everybody is connected

    # ---- plane views ----------------------------------------------------------------------------
    def _plane(self, q, xp, nxt):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        fn = self._lib.lbm_next_plane_ptr if nxt else self._lib.lbm_plane_ptr
        capi.check(fn(self.lattice.handle, q, xp, C.byref(ptr), C.byref(nbytes)))
        key = ptr.value
        v = self._views.get(key)
        if v is None:
            v = self.torch.as_tensor(_DevicePlane(ptr.value, nbytes.value // 4), device=self.torch.device("cuda", self.lattice.cfg.device))
            self._views[key] = v
        return v

    def _exchange(self, nxt):
        """edge planes of the (next or current) state -> the neighbours' ghost planes of the same state"""
        if os.environ.get("LBM_SLAB_NOEXCHANGE"):        # diagnosis only: step without communication (wrong results)
            return []
        nx = self.nx
        send = lambda q, which: self._plane(q, 1 if which == "lo" else nx, nxt)
        recv = lambda q, which: self._plane(q, 0 if which == "lo" else nx + 1, nxt)
        with self.torch.cuda.stream(self.stream):
            return self.halo.run(send, recv, self.dist)

    # ---- stepping --------------------------------------------------------------------------------
    def prime(self):
        """fill the ghost planes of the freshly loaded state (needed once before the first step)"""
        for w in self._exchange(nxt=False):
            w.wait()
        self._ghosts_current = True

    def _inplace_exchange(self, fetch):
        """single-buffer engine, around an odd step.  fetch=True: the neighbours' edge planes of the populations my
        edge cells pull (slots with c_x = -1 of the left neighbour's last plane, c_x = +1 of the right neighbour's
        first plane) -> my ghost planes.  fetch=False: the same slots, now holding what my edge cells pushed, go
        back from my ghost planes to the planes they belong to.  Same posting order as HaloExchange (rightward wave
        first), so NCCL pairs sends and receives correctly."""
        nx, h = self.nx, self.halo
        plane = lambda i, xp: self._plane(i, xp, False)
        ops = []
        if fetch:
            if h.right is not None:
                ops += [("send", plane(i, nx), h.right, i) for i in h.down]            # my last plane -> right's low ghost
            if h.left is not None:
                ops += [("recv", plane(i, 0), h.left, i) for i in h.down]
                ops += [("send", plane(i, 1), h.left, 100 + i) for i in h.up]          # my first plane -> left's high ghost
            if h.right is not None:
                ops += [("recv", plane(i, nx + 1), h.right, 100 + i) for i in h.up]
        else:
            # delivered into the slots of the receiver's ghost planes that it does not use itself (its low ghost plane
            # holds c_x = -1 slots only); lbm_inplace_merge_edges then moves what the neighbour really wrote
            if h.right is not None:
                ops += [("send", plane(i, nx + 1), h.right, i) for i in h.up]          # pushed across +x -> right's low ghost
            if h.left is not None:
                ops += [("recv", plane(i, 0), h.left, i) for i in h.up]
                ops += [("send", plane(i, 0), h.left, 100 + i) for i in h.down]        # pushed across -x -> left's high ghost
            if h.right is not None:
                ops += [("recv", plane(i, nx + 1), h.right, 100 + i) for i in h.down]
        if not ops:
            return
        dist = self.dist
        p2p = [dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer, tag=tag) for kind, t, peer, tag in ops]
        for w in dist.batch_isend_irecv(p2p):
            w.wait()                                                                   # stream-ordered for NCCL, blocking for gloo

    def _step_inplace(self, n, phases):
        lib, h = self._lib, self.lattice.handle
        with self.torch.cuda.stream(self.stream):
            for _ in range(int(n)):
                if self._aa_odd_next:
                    self._inplace_exchange(fetch=True)
                    capi.check(lib.lbm_step_phases(h, 1, int(phases)))
                    self._inplace_exchange(fetch=False)
                    capi.check(lib.lbm_inplace_merge_edges(h))
                else:
                    capi.check(lib.lbm_step_phases(h, 1, int(phases)))                 # even steps are cell-local
                self._aa_odd_next = not self._aa_odd_next

    def step(self, n=1, phases=capi.PHASE_ALL):
        lib, h = self._lib, self.lattice.handle
        if self.transport == "peer":
            capi.check(lib.lbm_step_phases(h, int(n), int(phases)))      # handshake, edge planes and interior: all below the C ABI
            return
        if self.inplace:
            return self._step_inplace(n, phases)
        if not self._ghosts_current:
            self.prime()
        with self.torch.cuda.stream(self.stream):
            for _ in range(int(n)):
                capi.check(lib.lbm_fill_ghosts(h))                 # y faces, including the x ghost planes
                if self.overlap and self.nx >= 4:
                    # edge planes first, their exchange overlaps the interior planes on the side stream
                    capi.check(lib.lbm_step_range(h, phases, 0, 1, 0))
                    capi.check(lib.lbm_step_range(h, phases, self.nx - 1, self.nx, 0))
                    capi.check(lib.lbm_step_range(h, phases, 1, self.nx - 1, 1))
                    works = self._exchange(nxt=True)
                else:
                    capi.check(lib.lbm_step_range(h, phases, 0, self.nx, 0))
                    works = self._exchange(nxt=True)
                for w in works:
                    w.wait()                                       # stream-ordered for NCCL, blocking for gloo
                capi.check(lib.lbm_step_commit(h))

    def step_timed(self, n, phases=capi.PHASE_ALL):
        torch = self.torch
        st0 = self.lattice.stats()["kernel_launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not self._ghosts_current and not self.inplace and self.transport != "peer":
            self.prime()
        self.stream.synchronize()
        e0.record(self.stream)
        self.step(n, phases)
        e1.record(self.stream)
        e1.synchronize()
        return e0.elapsed_time(e1), self.lattice.stats()["kernel_launches"] - st0

    def sync(self):
        """this slab's streams have drained and (peer transport) so have the neighbours': part of an edge cell's state may
        sit in a neighbouring slab's memory, so nothing is read back before every rank is here"""
        self.lattice.sync()
        if self.transport == "peer":
            self.dist.barrier()

    # ---- inputs in global coordinates ---------------------------------------------------------------
    def load_populations(self, f_local):
        self.lattice.load_populations(f_local)
        self._ghosts_current = False
        self._aa_odd_next = True               # a fresh state is in the parked form: the next step pulls and pushes

    def set_source_global(self, lo, hi, flow):
        """source box in global cell coordinates (None = whole axis); ranks the box misses get none"""
        gl = 0 if lo[0] is None else lo[0]
        gh = self.gx - 1 if hi[0] is None else hi[0]
        l, hgh = max(gl, self.x0) - self.x0, min(gh, self.x1 - 1) - self.x0
        if l > hgh:
            self.lattice.setFlowData(None)
        else:
            self.lattice.setFlowData(FlowData(((l, lo[1], lo[2]), (hgh, hi[1], hi[2])), flow))

    # ---- diagnostics over the whole domain (one small all-reduce, only when asked for) -----------------
    def _dev(self):
        return self.torch.device("cuda", self.lattice.cfg.device)

    def diagnostics(self):
        """{"mass": sum rho, "momentum": sum rho u [3], "force": obstacle force [3]} of the global lattice: every
        rank reduces its own slab on the device (lbm_mass / lbm_momentum / lbm_get_force), then one fp64
        all-reduce of 7 numbers"""
        self.sync()
        lat = self.lattice
        loc = np.zeros(7, np.float64)
        loc[0] = lat.mass()
        loc[1:4] = lat.momentum()
        if lat.m_reflectionData is not None:
            loc[4:7] = lat.force_f64()
        tot = allreduce_sum(loc, self.dist, self._dev() if self.dist.get_backend() == "nccl" else None)
        return {"mass": float(tot[0]), "momentum": tot[1:4].copy(), "force": tot[4:7].copy()}

    def mass(self):
        return self.diagnostics()["mass"]

    def momentum(self):
        return self.diagnostics()["momentum"]


class SlabGroup:
    """All x slabs of a gx x gy x gz lattice in ONE process, one per device in `devices` (repeating a device is fine:
    the slabs then share it) -- what a single-process caller such as the C++ `Lattice` drop-in does.  The slabs are
    connected through lbm_halo_connect and stepped in lockstep by lbm_group_step; inputs and outputs are global arrays."""

    def __init__(self, gx, gy, gz, fluid, devices, bc_x=capi.BC_PERIODIC, bc_yz=(capi.BC_PERIODIC,) * 2, **kw):
        if bc_x not in (capi.BC_PERIODIC, capi.BC_ZERO_INFLOW):
            raise ValueError("the global x edge must be periodic or zero-inflow in slab mode")
        self.gx, self.gy, self.gz = gx, gy, gz
        self.world = len(devices)
        self.periodic = bc_x == capi.BC_PERIODIC
        self.ranges = [slab_range(gx, r, self.world) for r in range(self.world)]
        self._lib = capi.load()
        self.slabs = [Lattice(x1 - x0, gy, gz, None, None, fluid, 1.0, bc=(capi.BC_HALO, bc_yz[0], bc_yz[1]), device=dev, **kw)
                      for (x0, x1), dev in zip(self.ranges, devices)]
        self.q = self.slabs[0].q
        blobs = []
        for s in self.slabs:
            b = (C.c_ubyte * capi.HALO_BLOB_BYTES)()
            capi.check(self._lib.lbm_halo_export(s.handle, b, capi.HALO_BLOB_BYTES))
            blobs.append(b)
        for r, s in enumerate(self.slabs):
            lo = blobs[(r - 1) % self.world] if (self.periodic or r > 0) else None
            hi = blobs[(r + 1) % self.world] if (self.periodic or r + 1 < self.world) else None
            capi.check(self._lib.lbm_halo_connect(s.handle, lo, hi))
        self._handles = (C.c_void_p * self.world)(*[s.handle for s in self.slabs])
        for (x0, _), s in zip(self.ranges, self.slabs):
            capi.check(self._lib.lbm_set_slab_origin(s.handle, x0, gx))
        self.m_simModel = None

    def close(self):
        for s in self.slabs:
            capi.check(self._lib.lbm_halo_disconnect(s.handle))
        for s in self.slabs:
            s.close()

    def step(self, n=1, phases=capi.PHASE_ALL):
        capi.check(self._lib.lbm_group_step(self._handles, self.world, int(n), int(phases)))

    def sync(self):
        for s in self.slabs:
            s.sync()

    def load_populations(self, f):
        f = np.ascontiguousarray(f, np.float32)
        for (x0, x1), s in zip(self.ranges, self.slabs):
            s.load_populations(np.ascontiguousarray(f[:, x0:x1]))

    def retrieve_populations(self):
        self.sync()
        return np.concatenate([s.retrieve_populations() for s in self.slabs], axis=1)

    def set_solid_mask(self, solid):
        """global uint8 [gx, gy, gz]: every slab gets its planes plus its neighbours' edge planes (periodic images at the
        ends of a periodic axis, fluid beyond a zero-inflow end)"""
        solid = np.ascontiguousarray(solid, np.uint8).reshape(self.gx, self.gy, self.gz)
        for (x0, x1), s in zip(self.ranges, self.slabs):
            lo = solid[(x0 - 1) % self.gx] if (self.periodic or x0 > 0) else np.zeros_like(solid[0])
            hi = solid[x1 % self.gx] if (self.periodic or x1 < self.gx) else np.zeros_like(solid[0])
            s.set_solid_mask(np.concatenate([lo[None], solid[x0:x1], hi[None]], axis=0), x_ghost_planes=True)

    def insertModel(self, filename, spacing):
        """Lattice::insertModel on every slab, in global coordinates (lbm_set_slab_origin): each slab voxelises its own
        planes and its neighbours' edge planes"""
        from .model import Model
        self.m_simModel = Model(filename)
        self._spacing = float(spacing)
        ok = True
        for s in self.slabs:
            accepted = C.c_int()
            capi.check(self._lib.lbm_insert_model(s.handle, self.m_simModel.handle, self._spacing, C.byref(accepted)))
            ok = ok and bool(accepted.value)
        return ok

    def preProcessModel(self):
        """Lattice::preProcessModel on every slab: the tables of its own boundary-solid cells and of those in the
        neighbours' edge planes"""
        if self.m_simModel is None:
            return
        for s in self.slabs:
            capi.check(self._lib.lbm_preprocess_model(s.handle, self.m_simModel.handle, self._spacing))

    def load_data(self, cells):
        """the reference's AoS records (144 bytes, host order: x slabs are contiguous runs)"""
        per = self.gy * self.gz
        cells = np.ascontiguousarray(cells, dtype=capi.LATTICE_POINT)
        for (x0, x1), s in zip(self.ranges, self.slabs):
            s.load_data(cells[x0 * per:x1 * per])
        solid = cells["is_internal"].reshape(self.gx, self.gy, self.gz)
        if solid.any():
            self.set_solid_mask(solid)          # load_data alone knows nothing of the neighbours' edge planes

    def retrieve_data(self):
        self.sync()
        parts = []
        for (x0, _), s in zip(self.ranges, self.slabs):
            p = s.retrieve_data()
            p["x"] += x0
            parts.append(p)
        return np.concatenate(parts)

    def set_source_global(self, lo, hi, flow):
        gl = 0 if lo[0] is None else lo[0]
        gh = self.gx - 1 if hi[0] is None else hi[0]
        for (x0, x1), s in zip(self.ranges, self.slabs):
            a, b = max(gl, x0) - x0, min(gh, x1 - 1) - x0
            s.setFlowData(None if a > b else FlowData(((a, lo[1], lo[2]), (b, hi[1], hi[2])), flow))

    def enable_force(self):
        from .lattice import ReflectionData
        for s in self.slabs:
            s.setReflectionData(ReflectionData())

    def force(self):
        self.sync()
        tot = np.zeros(3, np.float64)
        for s in self.slabs:
            tot += s.force_f64()
        return tot

    def mass(self):
        self.sync()
        return float(sum(s.mass() for s in self.slabs))

    def moments(self):
        self.sync()
        parts = [s.moments() for s in self.slabs]
        return tuple(np.concatenate([p[k] for p in parts], axis=0) for k in range(4))
