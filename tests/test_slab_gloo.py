"""Slab decomposition along x with world_size 2 and 3 over gloo on the CPU: the neighbour logic and plane
selection of latticeboltzmanncuda_b200.slab.HaloExchange, with the CPU oracle standing in for the step
kernel.  The decomposed run must reproduce the single-domain run bit for bit (no arithmetic crosses the
cut, every cell sees exactly the same operands)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle as O
from latticeboltzmanncuda_b200.slab import HaloExchange, face_populations, slab_range


def test_face_populations_match_velocity_set():
    for q, n in ((19, 5), (27, 9)):
        up, down = face_populations(q, 1), face_populations(q, 0)
        assert len(up) == n and len(down) == n
        assert all(O.DIRECTIONS[i][0] == 1 for i in up) and all(O.DIRECTIONS[i][0] == -1 for i in down)
        assert sorted(int(O.INVERSE[i]) for i in up) == down          # the two faces carry opposite populations


def test_slab_ranges_tile_the_axis():
    for gx, world in ((16, 2), (17, 3), (1024, 8), (10, 4)):
        edges = [slab_range(gx, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == gx
        assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
        assert max(b - a for a, b in edges) - min(b - a for a, b in edges) <= 1


def test_neighbours_and_tags():
    h = HaloExchange(0, 2, True, 19)
    assert (h.left, h.right) == (1, 1)
    ops = h.transfers(lambda q, w: ("s", q, w), lambda q, w: ("r", q, w))
    assert len(ops) == 20                                   # 5 + 5 sends, 5 + 5 receives
    # with left == right the two directions must not share a tag
    sends = {(peer, tag) for kind, t, peer, tag in ops if kind == "send"}
    recvs = {(peer, tag) for kind, t, peer, tag in ops if kind == "recv"}
    assert len(sends) == 10 and sends == recvs
    # NCCL ignores tags and matches a pair's sends and receives in posting order: the order in which rank 0
    # posts its sends to rank 1 must be the order in which rank 1 posts its receives from rank 0
    other = HaloExchange(1, 2, True, 19).transfers(lambda q, w: ("s", q, w), lambda q, w: ("r", q, w))
    assert [tag for kind, t, peer, tag in ops if kind == "send"] == [tag for kind, t, peer, tag in other if kind == "recv"]
    h = HaloExchange(0, 4, False, 27)
    assert h.left is None and h.right == 1
    assert len(h.transfers(lambda q, w: 0, lambda q, w: 0)) == 18
    assert HaloExchange(3, 4, False, 27).right is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, gx, Y, Z, periodic_x, bc_yz, steps, f_global, solid_global, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x0, x1 = slab_range(gx, rank, world)
        nx = x1 - x0
        # local state with one ghost plane on each side of x: [q, nx+2, Y, Z]
        loc = torch.zeros((q, nx + 2, Y, Z), dtype=torch.float32)
        loc[:, 1:nx + 1] = torch.from_numpy(f_global[:, x0:x1])
        xs = np.arange(x0 - 1, x1 + 1)
        if periodic_x:
            solid = solid_global[xs % gx]
        else:
            solid = np.zeros((nx + 2, Y, Z), np.uint8)
            ok = (xs >= 0) & (xs < gx)
            solid[ok] = solid_global[xs[ok]]
        halo = HaloExchange(rank, world, periodic_x, q)
        for _ in range(steps):
            loc[:, 0] = 0
            loc[:, nx + 1] = 0
            bufs = {}

            def send_plane(i, which):
                return loc[i, 1 if which == "lo" else nx].contiguous()

            def recv_plane(i, which):
                bufs[(i, which)] = torch.zeros((Y, Z), dtype=torch.float32)
                return bufs[(i, which)]

            for w in halo.run(send_plane, recv_plane, dist):
                w.wait()
            for (i, which), t in bufs.items():
                loc[i, 0 if which == "lo" else nx + 1] = t
            # the oracle steps the slab with its ghost planes as ordinary cells behind an open x edge
            new = O.step_soa(loc.numpy(), nx + 2, Y, Z, q, 0.7, bc=(O.BC_ZERO,) + tuple(bc_yz), solid=solid,
                             moment=O.MOMENT_STANDARD, eq64=False, steps=1)
            loc[:, 1:nx + 1] = torch.from_numpy(new[:, 1:nx + 1])
        np.save(out_path % rank, loc[:, 1:nx + 1].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,q,periodic_x,bc_yz", [(2, 19, True, (1, 1)), (2, 27, False, (2, 1)), (3, 19, True, (2, 2))])
def test_decomposed_run_equals_single_domain(tmp_path, world, q, periodic_x, bc_yz):
    gx, Y, Z, steps = 11, 5, 6, 6
    rng = np.random.default_rng(3)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1 + 0.05 * rng.standard_normal((q, gx, Y, Z)))).astype(np.float32)
    solid = (rng.random((gx, Y, Z)) < 0.1).astype(np.uint8)
    f[:, solid.astype(bool)] = 0
    bc = ((O.BC_PERIODIC if periodic_x else O.BC_ZERO),) + tuple(bc_yz)
    want = O.step_soa(f, gx, Y, Z, q, 0.7, bc=bc, solid=solid, moment=O.MOMENT_STANDARD, eq64=False, steps=steps)
    out = str(tmp_path / "slab_%d.npy")
    mp.spawn(_worker, args=(world, _free_port(), q, gx, Y, Z, periodic_x, bc_yz, steps, f, solid, out), nprocs=world, join=True)
    got = np.concatenate([np.load(out % r) for r in range(world)], axis=1)
    assert np.array_equal(got, want)


def _diag_worker(rank, world, port, gx, f_global, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from latticeboltzmanncuda_b200.slab import allreduce_sum
        x0, x1 = slab_range(gx, rank, world)
        loc = f_global[:, x0:x1].astype(np.float64)
        q = loc.shape[0]
        c = np.asarray(O.DIRECTIONS[:q], np.float64)
        part = np.concatenate([[loc.sum()], [(loc * c[:, a, None, None, None]).sum() for a in range(3)]])
        tot = allreduce_sum(part, dist)                 # host tensor over gloo; NCCL takes the same call with a device
        np.save(out_path % rank, tot)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_global_diagnostics_are_the_sum_over_slabs(tmp_path, world):
    """mass and momentum of the whole lattice = one fp64 all-reduce of the per-slab sums"""
    gx, Y, Z, q = 10, 4, 5, 19
    rng = np.random.default_rng(11)
    f = rng.random((q, gx, Y, Z)).astype(np.float32)
    out = str(tmp_path / "diag_%d.npy")
    mp.spawn(_diag_worker, args=(world, _free_port(), gx, f, out), nprocs=world, join=True)
    c = np.asarray(O.DIRECTIONS[:q], np.float64)
    f64 = f.astype(np.float64)
    want = np.array([f64.sum()] + [(f64 * c[:, a, None, None, None]).sum() for a in range(3)])
    for r in range(world):
        got = np.load(out % r)
        assert np.allclose(got, want, rtol=1e-13, atol=1e-12)     # every rank holds the global sums


# ---- the product's own SlabLattice (edge planes first, lbm_step_range / lbm_step_commit, plane pointers, halo masks)
# ---- over gloo, on the host build of the library (tests/_emu.py): what `bench.py --gpus N` runs, minus NCCL and the GPU

def _slab_lattice_worker(rank, world, port, q, gx, Y, Z, bc_yz, steps, f_global, solid_global, flow, out_path, inplace):
    import contextlib
    import ctypes as C
    import _emu
    import latticeboltzmanncuda_b200 as lb
    from latticeboltzmanncuda_b200.slab import SlabLattice
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class HostStream:                      # stands in for torch.cuda.Stream: the host build runs everything in order
        cuda_stream = 0

        def __init__(self, device=None):
            pass

        def synchronize(self):
            pass

    class HostSlab(SlabLattice):           # plane views and reductions on host memory; everything else is the product's
        def _plane(self, i, xp, nxt):
            ptr, nbytes = C.c_void_p(), C.c_size_t()
            fn = self._lib.lbm_next_plane_ptr if nxt else self._lib.lbm_plane_ptr
            lb.capi.check(fn(self.lattice.handle, i, xp, C.byref(ptr), C.byref(nbytes)))
            return torch.from_numpy(np.ctypeslib.as_array((C.c_float * (nbytes.value // 4)).from_address(ptr.value)))

        def _dev(self):
            return None

    try:
        with _emu.emulated_library():
            torch.cuda.Stream, torch.cuda.stream = HostStream, lambda s: contextlib.nullcontext()
            slab = HostSlab(gx, Y, Z, lb.FluidData(1.0, 0.7), rank=rank, world=world, device=0, bc_x=lb.BC_PERIODIC, bc_yz=bc_yz,
                            q=q, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK,
                            **({"engine": lb.ENGINE_INPLACE} if inplace else {}))
            x0, x1 = slab.x0, slab.x1
            xs = np.arange(x0 - 1, x1 + 1) % gx
            slab.lattice.set_solid_mask(np.ascontiguousarray(solid_global[xs]), x_ghost_planes=True)
            slab.load_populations(np.ascontiguousarray(f_global[:, x0:x1]))
            slab.set_source_global((0, None, None), (0, None, None), flow)
            slab.lattice.setReflectionData(lb.ReflectionData())
            slab.step(steps)
            d = slab.diagnostics()
            np.save(out_path % rank, slab.lattice.retrieve_populations())
            np.save((out_path % rank) + ".diag.npy", np.concatenate([[d["mass"]], d["momentum"], d["force"]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,q,bc_yz,gx,inplace", [(2, 19, (2, 2), 12, False), (2, 19, (2, 2), 12, True),
                                                      (3, 27, (1, 2), 13, True), (2, 19, (1, 1), 4, True)])
def test_product_slab_lattice_on_the_host_build(tmp_path, world, q, bc_yz, gx, inplace):
    """inplace: one copy of the populations per slab; the edge planes travel before and after every odd step"""
    Y, Z, steps = 5, 37, 5
    rng = np.random.default_rng(5)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1 + 0.05 * rng.standard_normal((q, gx, Y, Z)))).astype(np.float32)
    solid = (rng.random((gx, Y, Z)) < 0.08).astype(np.uint8)
    solid[0, 2, 3] = 1; solid[gx - 1, 2, 4] = 1                 # bodies on the slab edges: their images live in the ghost planes
    f[:, solid.astype(bool)] = 0
    flow = np.zeros(27, np.float32); flow[:q] = 1e-3
    bc = (O.BC_PERIODIC,) + tuple(bc_yz)
    force = np.zeros(3, np.float64)
    want = f
    for _ in range(steps):
        want = O.step_soa(want, gx, Y, Z, q, 0.7, bc=bc, solid=solid, moment=O.MOMENT_STANDARD, eq64=False, steps=1, force=force)
        add = np.zeros_like(want); add[:, 0] = flow[:q, None, None]; add[:, solid.astype(bool)] = 0
        want = want + add
    out = str(tmp_path / "slab_%d.npy")
    import _emu
    _emu.full_library()                                         # built once here, not by every worker
    mp.spawn(_slab_lattice_worker, args=(world, _free_port(), q, gx, Y, Z, bc_yz, steps, f, solid, flow, out, inplace), nprocs=world, join=True)
    got = np.concatenate([np.load(out % r) for r in range(world)], axis=1)
    scale = float(np.abs(want).max())
    assert float(np.abs(got.astype(np.float64) - want).max()) / scale <= 1e-5
    c = np.asarray(O.DIRECTIONS[:q], np.float64)
    w64 = want.astype(np.float64)
    for r in range(world):
        d = np.load((out % r) + ".diag.npy")
        assert d[0] == pytest.approx(w64.sum(), rel=1e-6)                                   # every rank holds the global mass ...
        assert np.allclose(d[1:4], [(w64 * c[:, a, None, None, None]).sum() for a in range(3)], atol=1e-4)
        assert np.allclose(d[4:7], force, rtol=1e-4, atol=1e-4)                            # ... and the force on all bodies
