"""Peer-memory halos below the C ABI (csrc/lbm_halo.cu) on the GPU.

1. Several slabs in ONE process (lbm_halo_connect + lbm_group_step) against one handle on the whole domain, bit for bit,
   and against the CPU oracle: the bodies of tests/_peer_halo_cases.py.  The slabs are spread over the GPUs of the box;
   on a one-GPU box they share the device (peer pointers are then plain pointers, the epoch handshake still runs).
2. One process per slab (the bench's launch mode): CUDA IPC handles exchanged through the process group, then lbm_step on
   every rank.  With two GPUs each rank takes its own; on a one-GPU box both ranks share device 0 and the process group
   is gloo (NCCL refuses two ranks on one device; it only carries the 320-byte blobs and barriers here).
"""
import os
import socket

import numpy as np
import pytest

import _oracle as O
import _peer_halo_cases as cases

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

import torch.multiprocessing as mp  # noqa: E402

os.environ.setdefault("LBM_HALO_TIMEOUT_MS", "5000")      # a broken handshake fails the test instead of stalling it


def _devices(world):
    n = torch.cuda.device_count()
    return [r % n for r in range(world)]


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_slab_group_equals_single_domain(name):
    cases.run_case(name, devices=_devices)


@pytest.mark.parametrize("world,fused", [(2, True), (3, True), (4, False)])
def test_reference_mode_on_slabs(world, fused):
    cases.run_reference_mode_case(world, devices=_devices, steps=12 if fused else 5, fused=fused)


def test_connect_refuses_mismatched_slabs():
    cases.check_refusals()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ngpu, q, gx, Y, Z, bc_yz, steps, engine, f, solid, out):
    import torch.distributed as dist
    import latticeboltzmanncuda_b200 as lb
    from latticeboltzmanncuda_b200.slab import SlabLattice, slab_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    if ngpu >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slab = SlabLattice(gx, Y, Z, lb.FluidData(1.0, 0.7), rank, world, dev, bc_x=lb.BC_PERIODIC, bc_yz=bc_yz, transport="peer",
                           q=q, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK,
                           engine=engine)
        x0, x1 = slab_range(gx, rank, world)
        xs = np.arange(x0 - 1, x1 + 1)
        slab.lattice.set_solid_mask(solid[xs % gx], x_ghost_planes=True)
        slab.load_populations(np.ascontiguousarray(f[:, x0:x1]))
        slab.lattice.setReflectionData(lb.ReflectionData())
        slab.step(steps)
        slab.sync()
        np.save(out % rank, slab.lattice.retrieve_populations())
        d = slab.diagnostics()
        np.save((out % rank) + ".diag.npy", np.concatenate([[d["mass"]], d["momentum"], d["force"]]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("engine,q,bc_yz,steps", [("inplace", 19, (2, 2), 9), ("pull2", 19, (2, 1), 6), ("inplace", 27, (1, 2), 5)])
def test_one_process_per_slab_over_cuda_ipc(tmp_path, engine, q, bc_yz, steps):
    import latticeboltzmanncuda_b200 as lb
    world = 2
    gx, Y, Z = 24, 10, 40
    rng = np.random.default_rng(5)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1 + 0.05 * rng.standard_normal((q, gx, Y, Z)))).astype(np.float32)
    solid = (rng.random((gx, Y, Z)) < 0.08).astype(np.uint8)
    solid[10:14, 3:7, 8:30] = 1                        # a body across the cut at x = 12
    solid[23, 2:5, 3:9] = 1                            # and one on the periodic seam
    f[:, solid.astype(bool)] = 0
    eng = lb.ENGINE_INPLACE if engine == "inplace" else lb.ENGINE_PULL2
    one = lb.Lattice(gx, Y, Z, None, None, lb.FluidData(1.0, 0.7), 1.0, q=q, bc=(1,) + tuple(bc_yz), engine=eng,
                     moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    one.set_solid_mask(solid)
    one.load_populations(f)
    one.setReflectionData(lb.ReflectionData())
    one.step(steps)
    want = one.retrieve_populations()
    want_diag = np.concatenate([[one.mass()], one.momentum(), one.force_f64()])
    one.close()
    out = str(tmp_path / "slab_%d.npy")
    mp.spawn(_worker, args=(world, _free_port(), torch.cuda.device_count(), q, gx, Y, Z, bc_yz, steps, eng, f, solid, out),
             nprocs=world, join=True)
    got = np.concatenate([np.load(out % r) for r in range(world)], axis=1)
    assert np.array_equal(got, want)                   # same kernels, same operands: bit-identical
    for r in range(world):
        got_diag = np.load((out % r) + ".diag.npy")
        assert np.allclose(got_diag, want_diag, rtol=1e-9, atol=1e-9 * abs(want_diag[0]))
