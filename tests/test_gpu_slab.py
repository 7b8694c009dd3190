"""Two x-slabs on two GPUs (NCCL) against the same lattice on one GPU.  Needs >= 2 CUDA devices; the
single-GPU round-end run skips it (the host-side logic is covered on the CPU by test_slab_gloo.py)."""
import os
import socket

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs", allow_module_level=True)

import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q, gx, Y, Z, bc_x, bc_yz, steps, overlap, f, solid, out, engine=None):
    import torch.distributed as dist
    import latticeboltzmanncuda_b200 as lb
    from latticeboltzmanncuda_b200.slab import SlabLattice, slab_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        slab = SlabLattice(gx, Y, Z, lb.FluidData(1.0, 0.7), rank, world, rank, bc_x=bc_x, bc_yz=bc_yz, overlap=overlap,
                           q=q, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK,
                           **({"engine": engine} if engine is not None else {}))
        x0, x1 = slab_range(gx, rank, world)
        xs = np.arange(x0 - 1, x1 + 1)
        if bc_x == lb.BC_PERIODIC:
            sl = solid[xs % gx]
        else:
            sl = np.zeros((len(xs), Y, Z), np.uint8)
            ok = (xs >= 0) & (xs < gx)
            sl[ok] = solid[xs[ok]]
        slab.lattice.set_solid_mask(sl, x_ghost_planes=True)
        slab.load_populations(np.ascontiguousarray(f[:, x0:x1]))
        slab.step(steps)
        np.save(out % rank, slab.lattice.retrieve_populations())
        d = slab.diagnostics()                      # one fp64 all-reduce: the global sums, on every rank
        np.save((out % rank) + ".diag.npy", np.concatenate([[d["mass"]], d["momentum"]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("q,bc_x,bc_yz,overlap", [(19, 1, (2, 2), True), (27, 0, (1, 2), True), (19, 1, (1, 1), False)])
def test_two_slabs_equal_one_gpu(tmp_path, q, bc_x, bc_yz, overlap):
    import latticeboltzmanncuda_b200 as lb
    gx, Y, Z, steps = 24, 10, 40, 9
    rng = np.random.default_rng(5)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1 + 0.05 * rng.standard_normal((q, gx, Y, Z)))).astype(np.float32)
    solid = (rng.random((gx, Y, Z)) < 0.08).astype(np.uint8)
    solid[10:14, 3:7, 8:30] = 1
    f[:, solid.astype(bool)] = 0
    one = lb.Lattice(gx, Y, Z, None, None, lb.FluidData(1.0, 0.7), 1.0, q=q, bc=(bc_x,) + tuple(bc_yz),
                     moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    one.set_solid_mask(solid)
    one.load_populations(f)
    one.step(steps)
    want = one.retrieve_populations()
    want_diag = np.concatenate([[one.mass()], one.momentum()])
    out = str(tmp_path / "slab_%d.npy")
    mp.spawn(_worker, args=(2, _free_port(), q, gx, Y, Z, bc_x, bc_yz, steps, overlap, f, solid, out), nprocs=2, join=True)
    got = np.concatenate([np.load(out % r) for r in range(2)], axis=1)
    assert np.array_equal(got, want)          # same kernel, same operands: bit-identical
    for r in range(2):
        got_diag = np.load((out % r) + ".diag.npy")
        assert np.allclose(got_diag, want_diag, rtol=1e-10, atol=1e-9 * abs(want_diag[0]))   # fp64 sums, different order


@pytest.mark.parametrize("q,bc_yz,steps", [(19, (2, 2), 9), (27, (1, 2), 6)])
def test_two_single_buffer_slabs_equal_one_gpu(tmp_path, q, bc_yz, steps):
    """one copy of the populations per slab: the edge planes travel before and after every odd step (NCCL transport;
    first run on two B200s in round 2, profiles/r02e_pytest_2gpu.log)"""
    import latticeboltzmanncuda_b200 as lb
    gx, Y, Z = 24, 10, 40
    rng = np.random.default_rng(5)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1 + 0.05 * rng.standard_normal((q, gx, Y, Z)))).astype(np.float32)
    solid = (rng.random((gx, Y, Z)) < 0.08).astype(np.uint8)
    solid[10:14, 3:7, 8:30] = 1
    f[:, solid.astype(bool)] = 0
    one = lb.Lattice(gx, Y, Z, None, None, lb.FluidData(1.0, 0.7), 1.0, q=q, bc=(1,) + tuple(bc_yz),
                     moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    one.set_solid_mask(solid)
    one.load_populations(f)
    one.step(steps)
    want = one.retrieve_populations()
    out = str(tmp_path / "slab_%d.npy")
    mp.spawn(_worker, args=(2, _free_port(), q, gx, Y, Z, 1, bc_yz, steps, False, f, solid, out, lb.ENGINE_INPLACE), nprocs=2, join=True)
    got = np.concatenate([np.load(out % r) for r in range(2)], axis=1)
    scale = float(np.abs(want).max())
    assert float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale <= 2e-6
