"""The product's host-side geometry helpers (latticeboltzmanncuda_b200/csrc/lbm_geometry.cu, host code paths)
through the C ABI: the reference's ModelTest / GeometryTest / TestDistributeVector / CompressionTest known
answers, and bit-equality with the CPU oracle.  No GPU needed: these entry points run on the host."""
import math
import os

import numpy as np
import pytest

import _oracle as O
import latticeboltzmanncuda_b200 as lb

OBJ = os.path.join(O.GOLDEN_DIR, "obj")


def obj(n):
    return os.path.join(OBJ, n)


def test_every_declared_symbol_is_exported():
    """the C-ABI library loads and exports every symbol include/lbm_b200.h declares"""
    import re
    lib = lb.capi.load()
    hdr = open(os.path.join(O.ROOT, "include", "lbm_b200.h")).read()
    declared = set(re.findall(r"LBM_API\s+[\w\s\*]+?\b(lbm_\w+)\s*\(", hdr))
    assert len(declared) >= 45
    bound = {name for name, _, _ in lb.capi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert getattr(lib, name) is not None


def test_model_known_answers():
    """ModelTest.*, test/ModelTest.cpp:5-74"""
    assert lb.Model(obj("triangle.obj")).number_of_vertices() == 3
    assert lb.Model(obj("cube.obj")).isModelClosed()
    assert not lb.Model(obj("triangle.obj")).isModelClosed()
    cube = lb.Model(obj("cube.obj"))
    assert cube.isPointInsideModel((0.0, 0.0, 0.0)) and not cube.isPointInsideModel((2.0, 0.0, 0.0))
    assert lb.Model(obj("quad.obj")).checkError()
    m = lb.Model(obj("compat_cube.obj"))
    assert np.allclose(m.reflectionVector((0.5, 0.5, -1.0), (0.5, 0.5, 1.0)), (0.0, 0.0, -1.0), atol=1e-6)
    assert tuple(m.reflectionVector((0.5, 0.5, -1.0), (0.5, 0.5, -2.0))) == (0.0, 0.0, 0.0)
    missing = lb.Model(obj("nope.obj"))
    assert missing.number_of_vertices() == 0 and not missing.isPointInsideModel((0.5, 0.5, 0.5))


def test_quaternion_rotation_known_answer():
    """GeometryTest.TestQuaternionRoatation, test/GeometryTest.cpp:11-32"""
    axis = (math.cos(math.pi * 0.25), math.sin(math.pi * 0.25), 0.0)
    assert np.allclose(lb.rotateVectorAroundAxis(axis, (1, 0, 0), math.pi), (0, 1, 0), atol=1e-6)
    assert np.allclose(lb.rotateVectorAroundAxis(axis, (1, 0, 0), math.pi / 2), (0.5, 0.5, -math.cos(math.pi * 0.25)), atol=1e-6)


def test_distribute_vector_known_answer():
    """CoordinatesTest.TestDistributeVector, test/CoordinatesTest.cpp:171-198"""
    ranked, s, d = lb.distributeVector((1.0, 0.1, 0.1))
    norm = s[0] + s[1] + s[2]
    assert list(d[:3]) == [6, 18, 17]
    assert s[0] / norm == pytest.approx(0.4134132615337746, abs=1e-6)
    assert s[1] / norm == pytest.approx(0.2932933692331128, abs=1e-6)


def test_compress_ints_known_answer():
    """CalculationsTest.CompressionTest, test/CalculationsTest.cpp:279-284"""
    assert lb.compressInts([1, 2, 3, 4, 5, 6], 6, 5) == 35754150


def test_host_helpers_bit_equal_to_oracle():
    rng = np.random.default_rng(11)
    for _ in range(300):
        axis, vec = rng.standard_normal(3).astype(np.float32), rng.standard_normal(3).astype(np.float32)
        ang = np.float32(rng.uniform(-7, 7))
        assert np.array_equal(lb.rotateVectorAroundAxis(axis, vec, ang), O.rotate(axis, vec, ang))
        v = rng.standard_normal(3).astype(np.float32)
        _, s, d = lb.distributeVector(v)
        so, do = O.distribute(v)
        assert np.array_equal(s, so) and np.array_equal(d, do)
    for name in ("cube.obj", "compat_cube.obj", "cube_angled.obj"):
        m, mo = lb.Model(obj(name)), O.Mesh(obj(name))
        lo, hi = m.bounding_box()
        pts = rng.uniform(lo - 0.2, hi + 0.2, (400, 3)).astype(np.float32)
        grid = (np.round(pts * 10) / 10).astype(np.float32)          # lattice-like points, many on facets and edges
        for p in np.concatenate([pts, grid]):
            assert m.isPointInsideModel(p) == mo.point_inside(*map(float, p))
        for _ in range(200):
            a, b = rng.uniform(lo - 0.5, hi + 0.5, 3).astype(np.float32), rng.uniform(lo, hi, 3).astype(np.float32)
            assert np.array_equal(m.reflectionVector(a, b), mo.reflection_vector(a, b))


def test_obj_with_out_of_range_face_indices_is_an_invalid_mesh(tmp_path):
    """`f 0 ...`, an index past the vertex list or a too-negative relative index must give an invalid (empty) model, as a
    reader that rejects the file would, not an out-of-range read in the mesh queries"""
    verts = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\n"
    good = "f 1 2 3\nf 1 2 4\nf 2 3 4\nf 1 3 4\n"
    for bad in ("f 0 1 2\n", "f 1 2 9\n", "f -7 1 2\n"):
        p = tmp_path / "bad.obj"
        p.write_text(verts + good + bad)
        m = lb.Model(str(p))
        assert m.checkError() and not m.isModelClosed()
        assert not m.isPointInsideModel((0.1, 0.1, 0.1))
        assert tuple(m.reflectionVector((0.1, 0.1, -1.0), (0.1, 0.1, 1.0))) == (0.0, 0.0, 0.0)
    p = tmp_path / "good.obj"
    p.write_text(verts + good)
    assert lb.Model(str(p)).isPointInsideModel((0.1, 0.1, 0.1))
