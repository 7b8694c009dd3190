"""field / state I/O: the VTK writer on the CPU, checkpoint round trip on the GPU"""
import os
import re
import struct

import numpy as np
import pytest

import _oracle as O
from latticeboltzmanncuda_b200 import io as lio


def test_write_vti_layout(tmp_path):
    nx, ny, nz = 3, 4, 5
    rng = np.random.default_rng(0)
    rho, ux, uy, uz = (rng.random((nx, ny, nz), dtype=np.float32) for _ in range(4))
    solid = (rng.random((nx, ny, nz)) < 0.3).astype(np.uint8)
    p = str(tmp_path / "out.vti")
    lio.write_vti(p, rho, ux, uy, uz, spacing=0.5, solid=solid)
    raw = open(p, "rb").read()
    head, tail = raw.split(b'<AppendedData encoding="raw">\n_', 1)
    assert b'WholeExtent="0 2 0 3 0 4"' in head and b'Spacing="0.5 0.5 0.5"' in head
    offs = [int(m) for m in re.findall(rb'offset="(\d+)"', head)]
    n = nx * ny * nz
    assert offs == [0, 8 + 4 * n, 8 + 4 * n + 8 + 12 * n]
    (count,) = struct.unpack("<Q", tail[:8])
    assert count == 4 * n
    dens = np.frombuffer(tail[8:8 + 4 * n], "<f4").reshape(nz, ny, nx)
    assert np.array_equal(dens, rho.transpose(2, 1, 0))             # x runs fastest in the file
    vel = np.frombuffer(tail[offs[1] + 8:offs[1] + 8 + 12 * n], "<f4").reshape(nz, ny, nx, 3)
    assert np.array_equal(vel[..., 1], uy.transpose(2, 1, 0))


@pytest.mark.gpu
def test_checkpoint_round_trip_continues_bit_for_bit(tmp_path):
    import latticeboltzmanncuda_b200 as lb
    n = 16
    obj = os.path.join(O.GOLDEN_DIR, "obj", "cube_angled.obj")
    flow = np.full(27, 0.001, np.float32); flow[6] = 0.002

    def fresh():
        return lb.Lattice(n, n, n, None, None, lb.FluidData(1.0, 0.9), float(np.float32(1.0 / n)))

    a = fresh()
    start = O.new_cells(n, n, n)
    start["f"] = flow[None, :]
    a.load_data(start)
    a.insertModel(obj)
    a.preProcessModel()
    a.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
    a.setReflectionData(lb.ReflectionData())
    a.step(7)
    ck = str(tmp_path / "state")                         # no suffix: both calls complete it the same way
    assert lio.save_checkpoint(a, ck) == ck + ".npz"
    a.step(9)
    want = a.retrieve_populations()
    wf = a.retrieve_reflection_data()
    wf64 = a.force_f64()
    b = fresh()
    # stale state on the restoring side must not survive: another body, a source and a force accumulator
    b.set_solid_mask(np.ones((n, n, n), np.uint8))
    b.setFlowData(lb.FlowData(lb.FlowCriterion.ptestCriteria, flow * 3))
    b.setReflectionData(lb.ReflectionData(5.0, 6.0, 7.0))
    lio.load_checkpoint(b, ck)
    b.step(9)
    assert np.array_equal(b.retrieve_populations(), want)
    bf = b.retrieve_reflection_data()
    assert np.allclose([bf.x, bf.y, bf.z], [wf.x, wf.y, wf.z], rtol=1e-6)
    assert np.allclose(b.force_f64(), wf64, rtol=1e-12)      # the accumulator is restored as the double it is (order of the warp sums aside)
    with pytest.raises(ValueError):
        b.retrieve_populations(out=np.zeros((27, n, n, n), np.float64))
    with pytest.raises(ValueError):
        b.moments(out=[np.zeros(n ** 3 - 1, np.float32)] * 4)
    rho, ux, uy, uz = b.moments()
    lio.write_vti(str(tmp_path / "field.vti"), rho, ux, uy, uz, 1.0 / n, b.get_solid_mask())
    assert os.path.getsize(str(tmp_path / "field.vti")) > 16 * n ** 3
