"""Pins the CPU oracle on every known-answer vector the reference's own gtest suite holds
for the hot path and its side inputs.  CPU only.

Each test names the reference test it re-expresses (paths relative to /root/reference).
"""
import math
import os

import numpy as np
import pytest

import _oracle as O

OBJ = os.path.join(O.GOLDEN_DIR, "obj")


def obj(name):
    return os.path.join(OBJ, name)


# ---- test/CalculationsTest.cpp ---------------------------------------------------------

def test_streaming_known_answer():
    """CalculationsTest.TestStreaming, test/CalculationsTest.cpp:19-45"""
    cells = O.new_cells(1, 2, 1)
    cells[0]["f"][5] = 2.0
    out = O.stream(cells, 1, 2, 1)
    assert out[1]["f"][5] == pytest.approx(2.0, abs=1e-6)
    assert out[0]["f"][5] == 0.0        # pulled from outside the box: zero inflow


def test_collision_known_answer():
    """CalculationsTest.TestCollision, test/CalculationsTest.cpp:47-75"""
    cells = O.new_cells(1, 1, 1)
    cells[0]["f"][0] = 2.0
    out = O.collide(cells, 1, 1, 1, 1.0)
    assert out[0]["f"][0] == pytest.approx(0.592592593, abs=1e-6)
    assert out[0]["f"][1] == pytest.approx(0.148148148, abs=1e-6)
    assert out[0]["f"][2] == pytest.approx(0.148148148, abs=1e-6)


def test_full_known_answer():
    """CalculationsTest.TestFull, test/CalculationsTest.cpp:77-105 (pins the SPLIT velocity)"""
    cells = O.new_cells(1, 2, 1)
    cells[0]["f"][5] = 2.0
    out = O.step_lattice(cells, 1, 2, 1, 1.0)
    assert out[1]["f"][0] == pytest.approx(-0.296296296, abs=1e-6)
    assert out[1]["f"][1] == pytest.approx(-0.074074074, abs=1e-6)
    assert out[1]["f"][2] == pytest.approx(0.148148148, abs=1e-6)


def _angled_3cube():
    X = Y = Z = 3
    cells = O.new_cells(X, Y, Z)
    mesh = O.Mesh(obj("cube_angled.obj"))
    assert mesh.insert(cells, X, Y, Z, 0.5) == 1
    tables, tindex = mesh.preprocess(cells, X, Y, Z, 0.5)
    return cells, tables, tindex


@pytest.mark.parametrize("with_force", [False, True])
def test_reflections_known_answer(with_force):
    """CalculationsTest.TestReflections / TestReflectionsData, test/CalculationsTest.cpp:107-197"""
    cells, tables, tindex = _angled_3cube()
    p = lambda x, y, z: O.idx(x, y, z, 3, 3)
    # only the centre cell is solid, and it reflects
    assert cells["is_internal"].sum() == 1 and cells[p(1, 1, 1)]["is_internal"] == 1
    assert cells[p(1, 1, 1)]["is_reflected"] == 1
    cells[p(1, 1, 1)]["f"][1] = 1.0
    force = np.zeros(3, np.float32) if with_force else None
    out = O.reflect(cells, 3, 3, 3, tables, tindex, force)
    assert out[p(2, 1, 2)]["f"][17] == pytest.approx(0.367402, abs=1e-6)
    assert out[p(1, 1, 2)]["f"][4] == pytest.approx(0.341121, abs=1e-6)
    assert out[p(2, 2, 2)]["f"][26] == pytest.approx(0.291477, abs=1e-6)
    assert out[p(1, 1, 1)]["f"][1] == pytest.approx(0.0, abs=1e-6)
    if with_force:
        assert force[0] == pytest.approx(0.28086, abs=1e-6)
        assert force[1] == pytest.approx(-0.194318, abs=1e-6)
        assert force[2] == pytest.approx(0.648261, abs=1e-6)


def test_flow_generation_known_answer():
    """CalculationsTest.TestFlowGeneration, test/CalculationsTest.cpp:199-234"""
    cells = O.new_cells(3, 3, 3)
    flow = np.zeros(27, np.float32); flow[0] = 1.0
    out = O.source(cells, 3, 3, 3, O.SRC_CELL_111, flow)
    assert out[O.idx(1, 1, 1, 3, 3)]["f"][0] == pytest.approx(1.0, abs=1e-6)
    assert out[O.idx(0, 0, 0, 3, 3)]["f"][0] == pytest.approx(0.0, abs=1e-6)


def test_plane_flow_generation_known_answer():
    """CalculationsTest.TestPlaneFlowGeneration, test/CalculationsTest.cpp:236-277"""
    cells = O.new_cells(3, 3, 3)
    flow = np.zeros(27, np.float32); flow[0] = 1.0
    out = O.source(cells, 3, 3, 3, O.SRC_PLANE_X0, flow)
    for y in range(3):
        for z in range(3):
            assert out[O.idx(0, y, z, 3, 3)]["f"][0] == pytest.approx(1.0, abs=1e-6)
            assert out[O.idx(1, y, z, 3, 3)]["f"][0] == pytest.approx(0.0, abs=1e-6)


def test_compress_ints_known_answer():
    """CalculationsTest.CompressionTest, test/CalculationsTest.cpp:279-284"""
    a = np.array([1, 2, 3, 4, 5, 6], np.int32)
    assert O.lib().lbo_compress_ints(O.ptr(a), 6, 5) == 35754150


# ---- test/CoordinatesTest.cpp ----------------------------------------------------------

def test_coordinates_with_model():
    """CoordinatesTest.TestCoordinatesWithModel, test/CoordinatesTest.cpp:137-169
    (pins ON_BOUNDARY => inside at index 1: 1*0.1f equals the parsed vertex coordinate)"""
    X = Y = Z = 10
    cells = O.new_cells(X, Y, Z)
    mesh = O.Mesh(obj("compat_cube.obj"))
    assert mesh.insert(cells, X, Y, Z, np.float32(0.1)) == 1
    for y in range(10):
        for x in range(10):
            assert cells[10 * y + x]["is_internal"] == 0
    for y in range(1, 9):
        for x in range(1, 9):
            assert cells[500 + 10 * y + x]["is_internal"] == 1
    # index 9: 9*0.1f rounds above 0.9f, so the far faces stay fluid
    assert cells[O.idx(5, 9, 5, Y, Z)]["is_internal"] == 0
    assert cells[O.idx(5, 5, 9, Y, Z)]["is_internal"] == 0


def test_distribute_vector():
    """CoordinatesTest.TestDistributeVector, test/CoordinatesTest.cpp:171-198"""
    s, d = O.distribute([1.0, 0.1, 0.1])
    norm = s[0] + s[1] + s[2]
    assert list(d[:3]) == [6, 18, 17]          # tie between 17 and 18: larger index first
    assert s[0] / norm == pytest.approx(0.4134132615337746, abs=1e-6)
    assert s[1] / norm == pytest.approx(0.2932933692331128, abs=1e-6)
    assert s[2] / norm == pytest.approx(0.2932933692331128, abs=1e-6)


def test_distribute_vector_ray_miss_order():
    """a (0,0,0) mirror vector ranks dir 0 first, then 26, 25, ... at equal score"""
    s, d = O.distribute([0.0, 0.0, 0.0])
    assert d[0] == 0 and s[0] == pytest.approx(2.0)
    assert list(d[1:]) == list(range(26, 0, -1))
    assert np.allclose(s[1:], 1.0, atol=1e-6)


def test_reflection_table():
    """CoordinatesTest.TestReflection, test/CoordinatesTest.cpp:200-229"""
    cells, tables, tindex = _angled_3cube()
    assert tindex[13] == 0 and len(tables) == 1
    t = tables[0]
    assert t["w"][3] == pytest.approx(0.367402, abs=1e-6)
    assert t["w"][4] == pytest.approx(0.341121, abs=1e-6)
    assert t["w"][5] == pytest.approx(0.291477, abs=1e-6)
    tt = tables[0:1].copy()
    get = lambda s: O.lib().lbo_get_reflection_direction(O.ptr(tt), s)
    assert (get(3), get(4), get(5)) == (17, 4, 26)
    # weights of every filled incoming direction sum to one
    for dinv in range(1, 27):
        w = t["w"][3 * dinv:3 * dinv + 3]
        dirs = [get(3 * dinv + k) for k in range(3)]
        if any(dirs):
            assert sum(w[k] for k in range(3) if dirs[k]) == pytest.approx(1.0, abs=1e-5)


def test_direction_packing_roundtrip():
    """Lattice::set/getReflectionDirection, simulation/source/Lattice.cpp:450-477"""
    t = np.zeros(1, O.REFLECTION_VALUES)
    rng = np.random.default_rng(1)
    vals = rng.integers(0, 27, 81)
    for s, v in enumerate(vals):
        O.lib().lbo_set_reflection_direction(O.ptr(t), s, int(v))
    for s, v in enumerate(vals):
        assert O.lib().lbo_get_reflection_direction(O.ptr(t), s) == v
    # slot s lives in int s/6 at bit 5*(s%6); bits 30-31 stay clear
    assert all((int(w) & 0xC0000000) == 0 for w in t[0]["dirs"].astype(np.uint32))
    assert (int(t[0]["dirs"][1]) >> 10) & 31 == vals[8]


# ---- test/ModelTest.cpp ----------------------------------------------------------------

def test_model_import():
    """ModelTest.importModel, test/ModelTest.cpp:5-12"""
    assert O.Mesh(obj("triangle.obj")).num_vertices == 3


def test_model_closed():
    """ModelTest.isModelClosed, test/ModelTest.cpp:14-25"""
    assert O.Mesh(obj("cube.obj")).is_closed
    assert not O.Mesh(obj("triangle.obj")).is_closed


def test_model_point_inside():
    """ModelTest.isPointInsideModel, test/ModelTest.cpp:27-40"""
    m = O.Mesh(obj("cube.obj"))
    assert m.point_inside(0.0, 0.0, 0.0)
    assert not m.point_inside(2.0, 0.0, 0.0)


def test_model_check_error():
    """ModelTest.checkErrorModel, test/ModelTest.cpp:42-49"""
    assert O.Mesh(obj("quad.obj")).check_error
    assert O.Mesh(obj("triangle.obj")).check_error      # open mesh => invalid (Model.cpp:40-43)
    assert not O.Mesh(obj("cube.obj")).check_error
    # an invalid mesh has no inside (Model.cpp:53-56)
    assert not O.Mesh(obj("quad.obj")).point_inside(0.5, 0.5, 0.0)


def test_model_intersection_normal():
    """ModelTest.intersectionNormal, test/ModelTest.cpp:51-74"""
    m = O.Mesh(obj("compat_cube.obj"))
    r = m.reflection_vector([0.5, 0.5, -1.0], [0.5, 0.5, 1.0])
    assert np.allclose(r, [0.0, 0.0, -1.0], atol=1e-6)
    r = m.reflection_vector([0.5, 0.5, -1.0], [0.5, 0.5, -2.0])
    assert tuple(r) == (0.0, 0.0, 0.0)


def test_missing_model_is_ignored():
    """Model::importModel returns silently on a missing file, simulation/source/Model.cpp:24-27"""
    m = O.Mesh(obj("does_not_exist.obj"))
    cells = O.new_cells(3, 3, 3)
    assert m.insert(cells, 3, 3, 3, 0.5) == 0
    assert cells["is_internal"].sum() == 0


def test_bounding_box_guard():
    """Lattice::insertModel leaves the lattice alone when the mesh leaves the box, Lattice.cpp:236-244"""
    cells = O.new_cells(3, 3, 3)
    assert O.Mesh(obj("cube.obj")).insert(cells, 3, 3, 3, 0.5) == 0       # xmin = -1 < 0
    assert O.Mesh(obj("compat_cube.obj")).insert(cells, 3, 3, 3, 0.25) == 0  # 0.9 > 3*0.25
    assert cells["is_internal"].sum() == 0


# ---- test/GeometryTest.cpp -------------------------------------------------------------

def test_quaternion_rotation():
    """GeometryTest.TestQuaternionRoatation, test/GeometryTest.cpp:11-32"""
    axis = [math.cos(math.pi * 0.25), math.sin(math.pi * 0.25), 0.0]
    r = O.rotate(axis, [1.0, 0.0, 0.0], np.float32(math.pi))
    assert np.allclose(r, [0.0, 1.0, 0.0], atol=1e-6)
    r = O.rotate(axis, [1.0, 0.0, 0.0], np.float32(math.pi / 2))
    assert np.allclose(r, [0.5, 0.5, -math.cos(math.pi * 0.25)], atol=1e-6)


# ---- self-consistency of the restatement -----------------------------------------------

def test_tables_are_consistent():
    assert np.array_equal(O.DIRECTIONS[O.INVERSE], -O.DIRECTIONS)
    assert np.array_equal(O.INVERSE[O.INVERSE], np.arange(27))
    assert set(O.INVERSE[:19]) == set(range(19))            # D3Q19 prefix is closed under inverse
    assert O.weights(27).sum() == pytest.approx(1.0) and O.weights(19).sum() == pytest.approx(1.0)


def test_stream_moves_every_population_one_cell():
    X, Y, Z = 4, 5, 6
    rng = np.random.default_rng(7)
    cells = O.new_cells(X, Y, Z)
    cells["f"] = rng.random((X * Y * Z, 27), dtype=np.float32)
    out = O.stream(cells, X, Y, Z)
    f_in = cells["f"].reshape(X, Y, Z, 27)
    f_out = out["f"].reshape(X, Y, Z, 27)
    for j in range(27):
        cx, cy, cz = O.DIRECTIONS[j]
        exp = np.zeros((X, Y, Z), np.float32)
        xs = slice(max(0, cx), X + min(0, cx)); xs_src = slice(max(0, -cx), X + min(0, -cx))
        ys = slice(max(0, cy), Y + min(0, cy)); ys_src = slice(max(0, -cy), Y + min(0, -cy))
        zs = slice(max(0, cz), Z + min(0, cz)); zs_src = slice(max(0, -cz), Z + min(0, -cz))
        exp[xs, ys, zs] = f_in[xs_src, ys_src, zs_src, j]
        assert np.array_equal(f_out[..., j], exp), j


def test_collide_conserves_mass_and_is_a_fixed_point_at_rest():
    X = Y = Z = 4
    cells = O.new_cells(X, Y, Z)
    w = O.weights(27).astype(np.float32)
    cells["f"] = w[None, :] * 1.3
    out = O.collide(cells, X, Y, Z, 0.8)
    assert np.allclose(out["f"], cells["f"], rtol=2e-7, atol=0)
    rng = np.random.default_rng(3)
    cells["f"] = (w[None, :] * (1 + 0.05 * rng.standard_normal((X * Y * Z, 27)))).astype(np.float32)
    out = O.collide(cells, X, Y, Z, 0.8)
    assert np.allclose(out["f"].sum(1), cells["f"].sum(1), rtol=1e-6)


def test_soa_stepper_matches_aos_phases():
    """the build-extension SoA stepper with zero-inflow edges and SPLIT moments is the same
    map as the reference's stream + collide"""
    X, Y, Z = 5, 4, 6
    rng = np.random.default_rng(11)
    w = O.weights(27).astype(np.float32)
    cells = O.new_cells(X, Y, Z)
    cells["f"] = (w[None, :] * (1 + 0.1 * rng.standard_normal((X * Y * Z, 27)))).astype(np.float32)
    ref = O.step_lattice(cells, X, Y, Z, 0.7, steps=3)
    f = np.ascontiguousarray(cells["f"].reshape(X, Y, Z, 27).transpose(3, 0, 1, 2))
    got = O.step_soa(f, X, Y, Z, 27, 0.7, bc=(O.BC_ZERO,) * 3, moment=O.MOMENT_REFERENCE, eq64=True, steps=3)
    assert np.array_equal(got.transpose(1, 2, 3, 0).reshape(-1, 27), ref["f"])


def test_soa_periodic_conserves_mass_and_momentum():
    X, Y, Z = 6, 5, 4
    rng = np.random.default_rng(5)
    for q in (19, 27):
        w = O.weights(q).astype(np.float32)
        f = (w[:, None, None, None] * (1 + 0.05 * rng.standard_normal((q, X, Y, Z)))).astype(np.float32)
        out = O.step_soa(f, X, Y, Z, q, 0.6, steps=20)
        assert out.sum(dtype=np.float64) == pytest.approx(f.sum(dtype=np.float64), rel=1e-6)
        c = O.DIRECTIONS[:q].astype(np.float64)
        m0 = np.tensordot(c.T, f.reshape(q, -1).sum(1, dtype=np.float64), 1)
        m1 = np.tensordot(c.T, out.reshape(q, -1).sum(1, dtype=np.float64), 1)
        assert np.allclose(m0, m1, atol=1e-4)


def test_soa_walls_conserve_mass():
    X = Y = Z = 6
    rng = np.random.default_rng(9)
    w = O.weights(19).astype(np.float32)
    f = (w[:, None, None, None] * (1 + 0.05 * rng.standard_normal((19, X, Y, Z)))).astype(np.float32)
    solid = np.zeros((X, Y, Z), np.uint8); solid[2:4, 2:4, 2:4] = 1
    f[:, solid.astype(bool)] = 0
    out = O.step_soa(f, X, Y, Z, 19, 0.7, bc=(O.BC_WALL,) * 3, solid=solid, steps=25)
    assert out.sum(dtype=np.float64) == pytest.approx(f.sum(dtype=np.float64), rel=1e-6)
    assert np.all(out[:, solid.astype(bool)] == 0)


# ---- the oracle against the reference's own host code (oracle/_ref/libref_host.so) ------

def _ref_host():
    import ctypes as C
    import subprocess
    so = os.path.join(O.ORACLE_DIR, "_ref", "libref_host.so")
    if not os.path.exists(so):
        if not os.path.isdir("/root/reference/simulation/source"):
            pytest.skip("reference sources absent and oracle/_ref/libref_host.so not prebuilt")
        subprocess.check_call(["make", "-s", "-C", O.ORACLE_DIR, "_ref/libref_host.so"])
    L = C.CDLL(so)
    L.ref_rotate_vector_around_axis.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.ref_compress_ints.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ref_compress_ints.restype = C.c_uint
    L.ref_constants.argtypes = [C.c_void_p] * 3
    return L


def test_oracle_tables_equal_reference_constants():
    """simulation/source/Constants.cpp:5-7 compiled unmodified"""
    L = _ref_host()
    d = np.zeros((27, 3), np.int32); inv = np.zeros(27, np.int32); tol = np.zeros(1, np.float32)
    L.ref_constants(O.ptr(d), O.ptr(inv), O.ptr(tol))
    assert np.array_equal(d, O.DIRECTIONS) and np.array_equal(inv, O.INVERSE)
    import ctypes as C
    od = (C.c_int * 81).in_dll(O.lib(), "LBO_DIRECTIONS")
    oi = (C.c_int * 27).in_dll(O.lib(), "LBO_INVERSE")
    assert list(od) == list(d.ravel()) and list(oi) == list(inv)
    assert tol[0] == np.float32(1e-6)


def test_oracle_rotation_bit_equal_to_reference():
    """rotateVectorAroundAxis, simulation/source/Geometry.cpp:22-42 compiled unmodified"""
    L = _ref_host()
    rng = np.random.default_rng(2026)
    for _ in range(500):
        axis = rng.standard_normal(3).astype(np.float32)
        vec = rng.standard_normal(3).astype(np.float32)
        ang = np.float32(rng.uniform(-7, 7))
        ref = np.zeros(3, np.float32)
        L.ref_rotate_vector_around_axis(O.ptr(axis), O.ptr(vec), ang, O.ptr(ref))
        got = O.rotate(axis, vec, ang)
        assert np.array_equal(got, ref)


def test_oracle_compress_ints_equal_to_reference():
    """Utils::compressInts, simulation/source/Utils.cpp:4-29 compiled unmodified"""
    L = _ref_host()
    rng = np.random.default_rng(5)
    for _ in range(200):
        n = int(rng.integers(1, 7))
        a = rng.integers(0, 40, n).astype(np.int32)      # values >= 32 exercise the skip branch
        assert O.lib().lbo_compress_ints(O.ptr(a), n, 5) == L.ref_compress_ints(O.ptr(a), n, 5)
