"""Parity of the sm_100a path (through the C ABI) against the CPU oracle and, where present,
against the reference's own CUDA kernels.  Needs a GPU: run with  pytest -m gpu.

Tolerances (BASELINE.json north_star): bit-exact for indexing / mask / streaming work;
<= 1e-5 relative (normalised by the largest population magnitude) for distributions and
moments in fp32; the reflection sums and the force are order-dependent fp32 sums in the
reference (atomics), so they get the same 1e-5.
"""
import os

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

lb = pytest.importorskip("latticeboltzmanncuda_b200")

OBJ = os.path.join(O.GOLDEN_DIR, "obj")
REL = 1e-5


def rel_err(got, want):
    scale = max(float(np.abs(want).max()), 1e-30)
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale


def random_cells(X, Y, Z, seed, amp=0.1):
    rng = np.random.default_rng(seed)
    cells = O.new_cells(X, Y, Z)
    w = O.weights(27).astype(np.float32)
    cells["f"] = (w[None, :] * (1.0 + amp * rng.standard_normal((X * Y * Z, 27)))).astype(np.float32)
    return cells


def make(X, Y, Z, tau=1.0, **kw):
    return lb.Lattice(X, Y, Z, (1, 1, 1), (X, Y, Z), lb.FluidData(1.0, tau), kw.pop("spacing", 0.1), **kw)


# ---- the reference's own known-answer tests, through the drop-in interface ---------------------

def test_load_and_retrieve():
    """CoordinatesTest.TestLoadAndRetrieve, test/CoordinatesTest.cpp:14-32"""
    lat = make(1, 2, 1)
    cells = O.new_cells(1, 2, 1)
    cells[0]["f"][5] = 2.0
    lat.load_data(cells)
    out = lat.retrieve_data()
    assert out[0]["f"][5] == 2.0


def test_coords_match_100_cubed():
    """CoordinatesTest.TestCoordsMatch / TestCoordinatesOnCopy, test/CoordinatesTest.cpp:34-135:
    host index z + 100 y + 10000 x <-> device cell (x, y, z), and a full state round trip"""
    X = Y = Z = 100
    lat = make(X, Y, Z)
    rng = np.random.default_rng(0)
    cells = O.new_cells(X, Y, Z)
    cells["f"] = rng.random((X * Y * Z, 27), dtype=np.float32)
    cells["is_internal"] = rng.random(X * Y * Z) < 0.1
    lat.load_data(cells)
    out = lat.retrieve_data()
    xs, ys, zs = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    assert np.array_equal(out["x"], xs.ravel()) and np.array_equal(out["y"], ys.ravel()) and np.array_equal(out["z"], zs.ravel())
    assert np.array_equal(out["f"], cells["f"])
    assert np.array_equal(out["is_internal"], cells["is_internal"])
    assert np.array_equal(lat.get_solid_mask().ravel(), cells["is_internal"])


def test_streaming_known_answer():
    """CalculationsTest.TestStreaming, test/CalculationsTest.cpp:19-45"""
    lat = make(1, 2, 1)
    cells = O.new_cells(1, 2, 1)
    cells[0]["f"][5] = 2.0
    lat.load_data(cells)
    lat.prime_points()
    lat.simulateStreaming()
    out = lat.retrieve_data()
    assert out[1]["f"][5] == pytest.approx(2.0, abs=1e-6)


@pytest.mark.parametrize("eq", [lb.EQ_FP64, lb.EQ_FP32])
def test_collision_known_answer(eq):
    """CalculationsTest.TestCollision, test/CalculationsTest.cpp:47-75"""
    lat = make(1, 1, 1, eq_precision=eq)
    cells = O.new_cells(1, 1, 1)
    cells[0]["f"][0] = 2.0
    lat.load_data(cells)
    lat.simulateCollision()
    out = lat.retrieve_data()
    assert out[0]["f"][0] == pytest.approx(0.592592593, abs=1e-6)
    assert out[0]["f"][1] == pytest.approx(0.148148148, abs=1e-6)
    assert out[0]["f"][2] == pytest.approx(0.148148148, abs=1e-6)


@pytest.mark.parametrize("eq", [lb.EQ_FP64, lb.EQ_FP32])
def test_full_known_answer(eq):
    """CalculationsTest.TestFull, test/CalculationsTest.cpp:77-105"""
    lat = make(1, 2, 1, eq_precision=eq)
    cells = O.new_cells(1, 2, 1)
    cells[0]["f"][5] = 2.0
    lat.load_data(cells)
    lat.simulateLattice()
    out = lat.retrieve_data()
    assert out[1]["f"][0] == pytest.approx(-0.296296296, abs=1e-6)
    assert out[1]["f"][1] == pytest.approx(-0.074074074, abs=1e-6)
    assert out[1]["f"][2] == pytest.approx(0.148148148, abs=1e-6)


def angled_3cube():
    cells = O.new_cells(3, 3, 3)
    mesh = O.Mesh(os.path.join(OBJ, "cube_angled.obj"))
    assert mesh.insert(cells, 3, 3, 3, 0.5) == 1
    tables, tindex = mesh.preprocess(cells, 3, 3, 3, 0.5)
    return cells, tables, tindex


@pytest.mark.parametrize("with_force", [False, True])
def test_reflections_known_answer(with_force):
    """CalculationsTest.TestReflections / TestReflectionsData, test/CalculationsTest.cpp:107-197"""
    cells, tables, tindex = angled_3cube()
    lat = make(3, 3, 3, spacing=0.5)
    p = lat.point_at_coords
    cells[p(1, 1, 1)]["f"][1] = 1.0
    lat.load_data(cells)
    lat.set_reflection_tables(tables, np.nonzero(tindex >= 0)[0])
    if with_force:
        lat.setReflectionData(lb.ReflectionData())
    lat.prime_points()
    lat.simulateReflections()
    out = lat.retrieve_data()
    assert out[p(2, 1, 2)]["f"][17] == pytest.approx(0.367402, abs=1e-6)
    assert out[p(1, 1, 2)]["f"][4] == pytest.approx(0.341121, abs=1e-6)
    assert out[p(2, 2, 2)]["f"][26] == pytest.approx(0.291477, abs=1e-6)
    assert out[p(1, 1, 1)]["f"][1] == pytest.approx(0.0, abs=1e-6)
    assert out[p(1, 1, 1)]["is_reflected"] == 1 and out[p(1, 1, 1)]["is_internal"] == 1
    if with_force:
        r = lat.retrieve_reflection_data()
        assert r.x == pytest.approx(0.28086, abs=1e-6)
        assert r.y == pytest.approx(-0.194318, abs=1e-6)
        assert r.z == pytest.approx(0.648261, abs=1e-6)


def test_flow_generation_known_answer():
    """CalculationsTest.TestFlowGeneration, test/CalculationsTest.cpp:199-234"""
    lat = make(3, 3, 3, spacing=0.5)
    flow = np.zeros(27, np.float32); flow[0] = 1.0
    lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestCriteria, flow))
    lat.simulateFlow()
    out = lat.retrieve_data()
    assert out[lat.point_at_coords(1, 1, 1)]["f"][0] == pytest.approx(1.0, abs=1e-6)
    assert out[lat.point_at_coords(0, 0, 0)]["f"][0] == pytest.approx(0.0, abs=1e-6)
    assert out["f"].sum() == pytest.approx(1.0)


def test_plane_flow_generation_known_answer():
    """CalculationsTest.TestPlaneFlowGeneration, test/CalculationsTest.cpp:236-277"""
    lat = make(3, 3, 3, spacing=0.5)
    flow = np.zeros(27, np.float32); flow[0] = 1.0
    lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
    lat.simulateFlow()
    out = lat.retrieve_data()
    for y in range(3):
        for z in range(3):
            assert out[lat.point_at_coords(0, y, z)]["f"][0] == pytest.approx(1.0, abs=1e-6)
            assert out[lat.point_at_coords(1, y, z)]["f"][0] == pytest.approx(0.0, abs=1e-6)


# ---- phase-by-phase parity on seeded random states ----------------------------------------------

SHAPES = [(1, 1, 1), (3, 3, 3), (5, 4, 7), (9, 6, 33), (4, 3, 70), (12, 10, 129)]


@pytest.mark.parametrize("shape", SHAPES)
def test_stream_bit_exact(shape):
    X, Y, Z = shape
    cells = random_cells(X, Y, Z, 1)
    lat = make(X, Y, Z)
    lat.load_data(cells)
    lat.simulateStreaming()
    got = lat.retrieve_data()
    want = O.stream(cells, X, Y, Z)
    assert np.array_equal(got["f"], want["f"])


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("eq", [lb.EQ_FP64, lb.EQ_FP32])
def test_collide_parity(shape, eq):
    X, Y, Z = shape
    cells = random_cells(X, Y, Z, 2, amp=0.3)
    lat = make(X, Y, Z, tau=0.8, eq_precision=eq)
    lat.load_data(cells)
    lat.simulateCollision()
    got = lat.retrieve_data()
    want = O.collide(cells, X, Y, Z, 0.8)
    assert rel_err(got["f"], want["f"]) <= (2e-7 if eq == lb.EQ_FP64 else REL)


def geometry_case(n, spacing=None, name="cube_angled.obj"):
    spacing = np.float32(1.0 / n) if spacing is None else spacing
    cells = O.new_cells(n, n, n)
    mesh = O.Mesh(os.path.join(OBJ, name))
    assert mesh.insert(cells, n, n, n, spacing) == 1
    tables, tindex = mesh.preprocess(cells, n, n, n, spacing)
    return cells, tables, tindex


def test_reflect_parity_16():
    n = 16
    geo, tables, tindex = geometry_case(n)
    cells = random_cells(n, n, n, 3)
    cells["is_internal"] = geo["is_internal"]
    cells["is_reflected"] = geo["is_reflected"]
    assert len(tables) > 50
    force = np.zeros(3, np.float32)
    want = O.reflect(cells, n, n, n, tables, tindex, force)
    lat = make(n, n, n)
    lat.load_data(cells)
    lat.set_reflection_tables(tables, np.nonzero(tindex >= 0)[0])
    lat.setReflectionData(lb.ReflectionData())
    lat.simulateReflections()
    got = lat.retrieve_data()
    assert rel_err(got["f"], want["f"]) <= REL
    assert np.array_equal(got["is_reflected"], geo["is_reflected"])
    r = lat.retrieve_reflection_data()
    assert np.allclose([r.x, r.y, r.z], force, rtol=1e-4, atol=1e-6)


def soa_equilibrium(q, n, solid):
    w = O.weights(q).astype(np.float32)
    f = np.ascontiguousarray(np.broadcast_to(w[:, None, None, None], (q, n, n, n))).copy()
    f[6] *= 1.05
    f[:, solid.astype(bool)] = 0
    return f


def lattice_pattern(x, y, z):
    """a flow criterion that is not an index box (restated in oracle/lbm_oracle.c, LBO_SRC_LATTICE_PATTERN, and compiled
    as a device predicate for the reference's own generate_flow in oracle/ref_extra_criterion.cu)"""
    return ((x + 2 * y + 3 * z) % 5 == 0) & (x >= 1)


def test_source_with_an_arbitrary_criterion():
    """FlowData takes ANY device predicate in the reference (FlowSurface.h:12, called per cell at SimCalcFuncs.cu:206);
    here it is evaluated once and kept as a bit per cell (lbm_set_source_mask): the source phase alone, the CLI loop with
    geometry, and the same mask on both throughput engines"""
    n = 16
    cells, tables, tindex = geometry_case(n)
    flow = (np.arange(27, dtype=np.float32) + 1.0) * 1e-4
    cells["f"] = np.float32(0.002)
    want1 = O.source(cells, n, n, n, O.SRC_LATTICE_PATTERN, flow)
    lat = make(n, n, n, tau=0.8)
    lat.load_data(cells)
    lat.setFlowData(lb.FlowData(lattice_pattern, flow))
    lat.simulateFlow()
    got1 = lat.retrieve_data()
    assert np.array_equal(got1["f"], want1["f"])                       # additions of the same floats: bit for bit
    hit = lattice_pattern(*np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")).ravel()
    assert 0 < hit.sum() < n ** 3 // 4 and np.all(got1["f"][~hit] == np.float32(0.002))
    # the CLI loop with this criterion
    force = np.zeros(3, np.float32)
    want = O.step_cli(cells, n, n, n, 0.8, tables, tindex, force, O.SRC_LATTICE_PATTERN, flow, steps=12)
    lat.load_data(cells)
    lat.set_reflection_tables(tables, np.nonzero(tindex >= 0)[0])
    lat.step(12)
    assert rel_err(lat.retrieve_data()["f"], want["f"]) <= REL
    # the same selection as an array, on the two throughput engines: identical states
    solid = cells["is_internal"].reshape(n, n, n).astype(np.uint8)
    sel = hit.reshape(n, n, n)
    f0 = soa_equilibrium(19, n, solid)
    got = []
    for engine in (lb.ENGINE_PULL2, lb.ENGINE_INPLACE):
        l2 = lb.Lattice(n, n, n, None, None, lb.FluidData(1.0, 0.7), 1.0, q=19, bc=(lb.BC_PERIODIC, lb.BC_WALL, lb.BC_PERIODIC),
                        moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK, engine=engine)
        l2.set_solid_mask(solid)
        l2.load_populations(f0)
        l2.setFlowData(lb.FlowData(sel, flow * 0.1))
        l2.step(7)
        got.append(l2.retrieve_populations())
        l2.close()
    assert rel_err(got[1], got[0]) <= 2e-6
    mass0 = float(f0.astype(np.float64).sum())
    fluid_hits = int((sel & (solid == 0)).sum())
    assert abs(float(got[0].astype(np.float64).sum()) - mass0 - 7 * fluid_hits * float(flow[:19].sum()) * 0.1) <= 1e-4 * mass0
    # a criterion that selects nothing switches the source off
    lat.setFlowData(lb.FlowData(lambda x, y, z: x < 0, flow))
    before = lat.retrieve_data()["f"]
    lat.simulateFlow()
    assert np.array_equal(lat.retrieve_data()["f"], before)


@pytest.mark.parametrize("tau,steps,fused", [(1.0, 40, True), (0.6, 25, True), (1.0, 12, False)])
def test_force_accumulator_in_the_reference_arithmetic(tau, steps, fused):
    """SURVEY.md 8c holds the force to 1e-5.  The reference adds every reflected slot into ONE float triple
    (SimCalcFuncs.cu:181-183), never reset (LatticeBoltzmannCuda.cpp:69-83): a large accumulator absorbs small
    increments, so the sum depends on the order.  LBM_FORCE_ACCUM_REFERENCE_FP32 fixes the oracle's order (solid
    cells in host order, slots ascending, one fused multiply-add per term) and must reproduce the oracle's floats;
    the default (double partial sums) is the exact sum and differs from that by the absorbed increments."""
    n = 16
    cells, tables, tindex = geometry_case(n)
    flow = np.full(27, 0.001, np.float32); flow[6] = 0.002
    cells["f"] = flow[None, :]
    force = np.zeros(3, np.float32)
    want = O.step_cli(cells, n, n, n, tau, tables, tindex, force, O.SRC_PLANE_X0, flow, steps=steps)
    lat = make(n, n, n, tau=tau, eq_precision=lb.EQ_FP64, force_accum=lb.FORCE_ACCUM_REFERENCE_FP32)
    lat.load_data(cells)
    lat.set_reflection_tables(tables, np.nonzero(tindex >= 0)[0])
    lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
    lat.setReflectionData(lb.ReflectionData())
    if fused:
        lat.step(steps)
    else:
        for _ in range(steps):
            lat.simulateStreaming(); lat.simulateReflections(); lat.simulateCollision(); lat.simulateFlow()
    got = lat.retrieve_data()
    assert rel_err(got["f"], want["f"]) <= REL
    r = lat.retrieve_reflection_data()
    assert np.allclose([r.x, r.y, r.z], force, rtol=1e-5, atol=0), ([r.x, r.y, r.z], force)
    # the slabs and the single-buffer engine have no such accumulator
    with pytest.raises(lb.LbmError):
        make(n, n, n, force_accum=lb.FORCE_ACCUM_REFERENCE_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)


@pytest.mark.parametrize("eq,tau,steps", [(lb.EQ_FP64, 1.0, 40), (lb.EQ_FP32, 1.0, 40), (lb.EQ_FP64, 0.6, 25), (lb.EQ_FP32, 0.6, 25)])
def test_cli_steps_parity_with_geometry(eq, tau, steps):
    """the CLI's loop (LatticeBoltzmannCuda.cpp:29-93) on the reference geometry at 16^3: all f = 0.001,
    f[6] = 0.002, plane source at x == 0, force accumulator on; fused step vs the oracle's four phases"""
    n = 16
    cells, tables, tindex = geometry_case(n)
    flow = np.full(27, 0.001, np.float32); flow[6] = 0.002
    cells["f"] = flow[None, :]
    force = np.zeros(3, np.float32)
    want = O.step_cli(cells, n, n, n, tau, tables, tindex, force, O.SRC_PLANE_X0, flow, steps=steps)
    lat = make(n, n, n, tau=tau, eq_precision=eq)
    lat.load_data(cells)
    lat.set_reflection_tables(tables, np.nonzero(tindex >= 0)[0])
    lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
    lat.setReflectionData(lb.ReflectionData())
    lat.step(steps)
    got = lat.retrieve_data()
    assert rel_err(got["f"], want["f"]) <= REL
    inside = cells["is_internal"] == 1          # the reflect phase empties solid cells; the source plane x == 0 holds none
    assert inside.any() and np.all(got["f"][inside] == 0) and np.all(want["f"][inside] == 0)
    r = lat.retrieve_reflection_data()
    assert np.allclose([r.x, r.y, r.z], force, rtol=2e-4, atol=1e-7)
    # moments (SPLIT-weighted, like the reference) agree too
    rho, ux, uy, uz = lat.moments()
    f64 = want["f"].astype(np.float64)
    assert np.allclose(rho.ravel(), f64.sum(1), rtol=1e-5, atol=1e-7)


def test_unfused_phases_equal_fused_step():
    n = 12
    cells, tables, tindex = geometry_case(n)
    flow = np.full(27, 0.001, np.float32); flow[6] = 0.002
    cells["f"] = random_cells(n, n, n, 5)["f"] * 0.01

    def run(fused):
        lat = make(n, n, n, tau=0.9)
        lat.load_data(cells)
        lat.set_reflection_tables(tables, np.nonzero(tindex >= 0)[0])
        lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
        lat.setReflectionData(lb.ReflectionData())
        for _ in range(6):
            if fused:
                lat.step(1)
            else:
                lat.simulateStreaming(); lat.simulateReflections(); lat.simulateCollision(); lat.simulateFlow()
        r = lat.retrieve_reflection_data()
        return lat.retrieve_data()["f"], np.array([r.x, r.y, r.z])

    a, fa = run(True)
    b, fb = run(False)
    assert np.array_equal(a, b)
    assert np.allclose(fa, fb, rtol=1e-6)


# ---- build-extension configurations (throughput modes) vs the oracle's SoA stepper ---------------

def soa_state(q, X, Y, Z, seed, solid=None):
    rng = np.random.default_rng(seed)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1.0 + 0.05 * rng.standard_normal((q, X, Y, Z)))).astype(np.float32)
    if solid is not None:
        f[:, solid.astype(bool)] = 0
    return f


@pytest.mark.parametrize("q", [19, 27])
@pytest.mark.parametrize("eq", [lb.EQ_FP32, lb.EQ_FP64])
@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (1, 0, 1), (2, 2, 2), (1, 2, 1), (2, 1, 0)])
def test_soa_engine_parity(q, eq, bc):
    X, Y, Z = 10, 7, 37
    f = soa_state(q, X, Y, Z, 17)
    # open (zero-inflow) edges drain the box until rho -> 0 and u = m / rho blows up: any rounding
    # difference is then amplified without bound, so those cases stop while the state is regular
    steps = 3 if 0 in bc else 12
    want = O.step_soa(f, X, Y, Z, q, 0.7, bc=bc, moment=O.MOMENT_STANDARD, eq64=(eq == lb.EQ_FP64), steps=steps)
    lat = make(X, Y, Z, tau=0.7, q=q, bc=bc, moment_mode=lb.MOMENT_STANDARD, eq_precision=eq,
               boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    lat.load_populations(f)
    lat.step(steps)
    got = lat.retrieve_populations()
    assert rel_err(got, want) <= REL


@pytest.mark.parametrize("q", [19, 27])
def test_bounce_back_obstacle_parity(q):
    X, Y, Z = 14, 12, 40
    rng = np.random.default_rng(23)
    solid = (rng.random((X, Y, Z)) < 0.08).astype(np.uint8)
    solid[4:8, 3:9, 30:36] = 1
    f = soa_state(q, X, Y, Z, 29, solid)
    bc = (1, 2, 1)
    force = np.zeros(3, np.float64)
    want = O.step_soa(f, X, Y, Z, q, 0.65, bc=bc, solid=solid, moment=O.MOMENT_STANDARD, eq64=False, steps=15, force=force)
    lat = make(X, Y, Z, tau=0.65, q=q, bc=bc, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
               boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    lat.set_solid_mask(solid)
    lat.load_populations(f)
    lat.setReflectionData(lb.ReflectionData())
    m0 = lat.mass()
    lat.step(15)
    got = lat.retrieve_populations()
    assert rel_err(got, want) <= REL
    r = lat.retrieve_reflection_data()          # momentum handed to the obstacle cells by bounce-back
    assert np.allclose([r.x, r.y, r.z], force, rtol=1e-4, atol=1e-4)
    assert np.all(got[:, solid.astype(bool)] == 0)
    assert lat.mass() == pytest.approx(m0, rel=1e-6)          # closed box: mass conserved to round-off
    assert np.array_equal(lat.get_solid_mask(), solid)


def test_lid_driven_cavity_parity():
    """BASELINE.json config 2 in small: walls on every face, moving lid at y = Y-1"""
    n = 24
    wall_u = np.zeros((6, 3), np.float32); wall_u[3] = (0.05, 0.0, 0.0)
    f = soa_state(19, n, n, n, 31) * 0 + O.weights(19).astype(np.float32)[:, None, None, None]
    want = O.step_soa(f, n, n, n, 19, 0.6, bc=(2, 2, 2), wall_u=wall_u, moment=O.MOMENT_STANDARD, eq64=False, steps=30)
    lat = make(n, n, n, tau=0.6, q=19, bc=(2, 2, 2), wall_u=wall_u, moment_mode=lb.MOMENT_STANDARD,
               eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    lat.load_populations(f)
    lat.step(30)
    got = lat.retrieve_populations()
    assert rel_err(got, want) <= REL
    rho, ux, uy, uz = lat.moments()
    assert ux[:, -1, :].mean() > 1e-3            # the lid drags the top layer along +x


# ---- size-independent properties at sizes the oracle cannot reach -----------------------------------

@pytest.mark.parametrize("q", [19, 27])
def test_large_periodic_conservation(q):
    X, Y, Z = 192, 160, 256
    lat = make(X, Y, Z, tau=0.6, q=q, bc=(1, 1, 1), moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
               boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    f = soa_state(q, X, Y, Z, 41)
    lat.load_populations(f)
    m0, p0 = lat.mass(), lat.momentum()
    assert m0 == pytest.approx(f.sum(dtype=np.float64), rel=1e-9)
    lat.step(50)
    m1, p1 = lat.mass(), lat.momentum()
    assert m1 == pytest.approx(m0, rel=1e-6)
    assert np.allclose(p1, p0, atol=1e-6 * m0)


def test_uniform_rest_state_is_a_fixed_point_256():
    n = 256
    lat = make(n, n, n, tau=0.6, q=19, bc=(1, 1, 1), moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
               boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    w = O.weights(19).astype(np.float32)
    f = np.broadcast_to(w[:, None, None, None], (19, n, n, n)).copy()
    lat.load_populations(f)
    lat.step(10)
    got = lat.retrieve_populations()
    assert np.allclose(got, f, rtol=1e-6, atol=0)


def test_taylor_green_decay_matches_viscosity():
    """BASELINE.json config 5 in small: 2-D Taylor-Green vortex extruded in z, fully periodic,
    standard moments; kinetic energy decays as exp(-2 nu (kx^2+ky^2) t), nu = (tau - 1/2)/3"""
    n, nzz, tau, u0 = 64, 32, 0.6, 0.04
    nu = (tau - 0.5) / 3.0
    k = 2 * np.pi / n
    x = np.arange(n)[:, None, None]; y = np.arange(n)[None, :, None]
    ux = (u0 * np.sin(k * x) * np.cos(k * y) * np.ones((1, 1, nzz))).astype(np.float64)
    uy = (-u0 * np.cos(k * x) * np.sin(k * y) * np.ones((1, 1, nzz))).astype(np.float64)
    rho = 1.0 - (u0 ** 2) * 3.0 / 4.0 * (np.cos(2 * k * x) + np.cos(2 * k * y)) * np.ones((1, 1, nzz))
    c = O.DIRECTIONS[:19].astype(np.float64); w = O.weights(19)
    cu = c[:, 0, None, None, None] * ux + c[:, 1, None, None, None] * uy
    usq = ux ** 2 + uy ** 2
    f = (w[:, None, None, None] * rho * (1 + 3 * cu + 4.5 * cu ** 2 - 1.5 * usq)).astype(np.float32)
    lat = make(n, n, nzz, tau=tau, q=19, bc=(1, 1, 1), moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
               boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    lat.load_populations(f)

    def energy():
        r, a, b, cz = lat.moments()
        return float((r.astype(np.float64) * (a.astype(np.float64) ** 2 + b.astype(np.float64) ** 2)).sum())

    e0 = energy()
    steps = 400
    lat.step(steps)
    e1 = energy()
    expected = np.exp(-2.0 * nu * 2 * k * k * steps)
    assert e1 / e0 == pytest.approx(expected, rel=0.02)


# ---- error behaviour of the boundary -------------------------------------------------------------------

def test_bad_arguments_are_reported():
    with pytest.raises(lb.LbmError):
        lb.Lattice(0, 4, 4)
    with pytest.raises(lb.LbmError):
        lb.Lattice(4, 4, 4, q=15)
    lat = make(4, 4, 4)
    with pytest.raises(lb.LbmError):
        lat.load_data(O.new_cells(2, 2, 2))
    with pytest.raises(lb.LbmError):
        lat.set_reflection_tables(np.zeros(1, O.REFLECTION_VALUES), [4 * 4 * 4 + 1])


# ---- single-buffer (AA pattern) engine ----------------------------------------------------------------------

@pytest.mark.parametrize("q", [19, 27])
@pytest.mark.parametrize("eq", [lb.EQ_FP32, lb.EQ_FP64])
@pytest.mark.parametrize("bc", [(1, 1, 1), (2, 2, 2), (1, 2, 1), (2, 1, 2)])
@pytest.mark.parametrize("steps", [1, 2, 7, 12])
def test_inplace_engine_equals_two_buffer_engine(q, eq, bc, steps):
    """the in-place engine holds one copy of the populations; after any number of steps (odd or even, i.e. in
    either of its two storage forms) the state it reports is the two-buffer engine's, and the oracle's"""
    X, Y, Z = 9, 6, 37
    f = soa_state(q, X, Y, Z, 53)
    wall_u = np.zeros((6, 3), np.float32)
    wall_u[3] = (0.03, 0.0, 0.01)          # moving y-high wall (used when y is a wall axis)
    wall_u[4] = (0.0, 0.02, 0.0)           # moving z-low wall
    kw = dict(tau=0.7, q=q, bc=bc, wall_u=wall_u, moment_mode=lb.MOMENT_STANDARD, eq_precision=eq,
              boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    ref = make(X, Y, Z, **kw)
    ref.load_populations(f)
    ref.step(steps)
    want = ref.retrieve_populations()
    lat = make(X, Y, Z, engine=lb.ENGINE_INPLACE, **kw)
    assert lat.layout().n_buffers == 1
    lat.load_populations(f)
    lat.step(steps)
    got = lat.retrieve_populations()
    assert rel_err(got, want) <= 2e-6
    orc = O.step_soa(f, X, Y, Z, q, 0.7, bc=bc, wall_u=wall_u, moment=O.MOMENT_STANDARD, eq64=(eq == lb.EQ_FP64), steps=steps)
    assert rel_err(got, orc) <= REL
    assert lat.mass() == pytest.approx(ref.mass(), rel=1e-6)
    assert np.allclose(lat.momentum(), ref.momentum(), atol=1e-4)
    for a, b in zip(lat.moments(), ref.moments()):
        assert np.allclose(a, b, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("bodies", [1, 2])
@pytest.mark.parametrize("bc", [(1, 1, 1), (1, 2, 2)])
def test_inplace_engine_with_a_compact_body(bc, bodies):
    """one compact body in a box: the single-buffer kernels skip the all-solid words inside it by a box test on the kernel
    parameters and read no class byte outside the bounding box of the body (csrc/lbm_capi.cu find_dead_boxes, busy_box);
    state and force stay the oracle's after every step"""
    X, Y, Z, q = 16, 14, 200, 19
    solid = np.zeros((X, Y, Z), np.uint8)
    solid[6:10, 5:9, 60:140] = 1                # words 2 and 3 of these rows are all solid; words 0, 5, 6 of every row are ordinary
    solid[7, 6, 59] = 1; solid[10, 8, 100] = 1  # (not a box)
    if bodies == 2:
        solid[12:15, 2:12, 96:192] = 1          # a second body: words 3, 4 and 5 of its rows, a second box of all-solid words
    f = soa_state(q, X, Y, Z, 31, solid)
    lat = make(X, Y, Z, engine=lb.ENGINE_INPLACE, tau=0.7, q=q, bc=bc, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
               boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    lat.set_solid_mask(solid)
    lat.load_populations(f)
    lat.setReflectionData(lb.ReflectionData())
    st = lat.stats()
    assert st["words_dead"] == 4 * 4 * 2 + (bodies - 1) * 3 * 10 * 3 and st["words_special"] > 0
    assert st["words_total"] > 2 * (st["words_dead"] + st["words_special"])
    force = np.zeros(3, np.float64)
    state = f
    c = O.DIRECTIONS[:q].astype(np.float64)
    for step in range(1, 7):
        lat.step(1)
        state = O.step_soa(state, X, Y, Z, q, 0.7, bc=bc, solid=solid, moment=O.MOMENT_STANDARD, eq64=False, steps=1, force=force)
        got = lat.retrieve_populations()
        assert rel_err(got, state) <= REL, "step %d" % step
        assert np.all(got[:, solid.astype(bool)] == 0)
        assert np.allclose(lat.force_f64(), force, rtol=1e-4, atol=1e-4), "step %d" % step
        # the diagnostics read either storage form (odd / even step count), with short cuts for ordinary words
        s64 = state.astype(np.float64)
        assert lat.mass() == pytest.approx(float(s64.sum()), rel=1e-6), "step %d" % step
        rho, ux, uy, uz = lat.moments()
        assert np.allclose(rho, s64.sum(axis=0), rtol=1e-5, atol=1e-6), "step %d" % step
        assert np.allclose(ux * rho, (s64 * c[:, 0, None, None, None]).sum(axis=0), rtol=1e-4, atol=1e-5), "step %d" % step
        assert np.allclose(uz * rho, (s64 * c[:, 2, None, None, None]).sum(axis=0), rtol=1e-4, atol=1e-5), "step %d" % step
    lat.close()


@pytest.mark.parametrize("q", [19, 27])
@pytest.mark.parametrize("bc", [(1, 1, 1), (2, 2, 2), (1, 2, 2), (2, 1, 2)])
@pytest.mark.parametrize("shape", [(14, 12, 40), (9, 8, 130)])
def test_inplace_engine_with_obstacles(q, bc, shape):
    """bounce-back obstacles, the plane source and the momentum exchange on ONE copy of the populations: after every
    step (either storage form) the state, the force and the moments are the two-buffer engine's and the oracle's"""
    X, Y, Z = shape
    rng = np.random.default_rng(23)
    solid = (rng.random((X, Y, Z)) < (0.08 if Z < 100 else 0.0)).astype(np.uint8)
    solid[3:6, 2:6, 30:(36 if Z < 100 else 100)] = 1       # the long block makes all-solid words, ordinary words remain
    solid[0, 4, Z // 2] = 1; solid[X - 1, Y - 1, Z - 1] = 1 # cells on faces and in a corner: images across every edge
    f = soa_state(q, X, Y, Z, 29, solid)
    wall_u = np.zeros((6, 3), np.float32)
    wall_u[3] = (0.03, 0.0, 0.01); wall_u[4] = (0.0, 0.02, 0.0)
    kw = dict(tau=0.65, q=q, bc=bc, wall_u=wall_u, moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
              boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    flow = np.zeros(27, np.float32); flow[:q] = 1e-3 * (1 + np.arange(q))
    src = lb.FlowData(((0, None, None), (0, None, None)), flow)          # the plane x == 0, like the CLI
    lats = []
    for engine in (lb.ENGINE_PULL2, lb.ENGINE_INPLACE):
        lat = make(X, Y, Z, engine=engine, **kw)
        lat.set_solid_mask(solid)
        lat.load_populations(f)
        lat.setReflectionData(lb.ReflectionData())
        lats.append(lat)
    two, one = lats
    assert one.layout().n_buffers == 1
    st1, st2 = one.stats(), two.stats()
    assert (st1["words_dead"], st1["words_special"]) == (st2["words_dead"], st2["words_special"])
    if Z > 100:
        assert st1["words_dead"] > 0 and st1["words_special"] > 0 and st1["words_total"] > 2 * st1["words_special"]
    force = np.zeros(3, np.float64)
    state = f
    for step in range(1, 6):
        two.step(1); one.step(1)
        state = O.step_soa(state, X, Y, Z, q, 0.65, bc=bc, wall_u=wall_u, solid=solid, moment=O.MOMENT_STANDARD, eq64=False,
                           steps=1, force=force)
        got = one.retrieve_populations()
        assert rel_err(got, two.retrieve_populations()) <= 2e-6, "step %d" % step
        assert rel_err(got, state) <= REL, "step %d" % step
        assert np.all(got[:, solid.astype(bool)] == 0)
        r1, r2 = one.retrieve_reflection_data(), two.retrieve_reflection_data()
        assert np.allclose([r1.x, r1.y, r1.z], force, rtol=1e-4, atol=1e-4)
        assert np.allclose([r1.x, r1.y, r1.z], [r2.x, r2.y, r2.z], rtol=1e-4, atol=1e-4)
    assert one.mass() == pytest.approx(two.mass(), rel=1e-6)
    for a, b in zip(one.moments(), two.moments()):
        assert np.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert np.array_equal(one.get_solid_mask(), solid)
    # the source: both engines add it after the collision to the live cells of the box
    for lat in lats:
        lat.setFlowData(src)
    two.step(4); one.step(4)
    assert rel_err(one.retrieve_populations(), two.retrieve_populations()) <= 2e-6
    two.step(3); one.step(3)
    assert rel_err(one.retrieve_populations(), two.retrieve_populations()) <= 4e-6


def test_inplace_obstacle_conserves_mass_256():
    """size-independent property at a size the oracle does not reach: closed box with a body, mass to round-off"""
    from latticeboltzmanncuda_b200.shapes import tapered_hexahedron_mask
    n = 256
    solid = tapered_hexahedron_mask(n, n, n)
    w = O.weights(19).astype(np.float32)
    f = np.broadcast_to(w[:, None, None, None], (19, n, n, n)).copy()
    f[6] *= 1.05; f[1] *= 0.95
    f[:, solid.astype(bool)] = 0
    lat = make(n, n, n, tau=0.6, q=19, bc=(1, 2, 2), moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32,
               boundary_mode=lb.BOUNDARY_BOUNCE_BACK, engine=lb.ENGINE_INPLACE)
    lat.set_solid_mask(solid)
    lat.load_populations(f)
    m0 = lat.mass()
    assert m0 == pytest.approx(f.sum(dtype=np.float64), rel=1e-9)
    lat.step(41)
    assert lat.mass() == pytest.approx(m0, rel=1e-6)
    rho = lat.moments()[0]
    assert np.all(rho[solid.astype(bool)] == 0) and rho[~solid.astype(bool)].min() > 0.5


def test_inplace_engine_refuses_what_it_cannot_do():
    with pytest.raises(lb.LbmError):
        make(8, 8, 8, engine=lb.ENGINE_INPLACE)                          # zero-inflow edges (the default)
    lat = make(8, 8, 8, engine=lb.ENGINE_INPLACE, bc=(1, 1, 1))
    with pytest.raises(lb.LbmError):
        lat.set_solid_mask(np.ones((8, 8, 8), np.uint8))                 # table-reflection mode (the default)
    with pytest.raises(lb.LbmError):
        lat.set_reflection_tables(np.zeros(1, O.REFLECTION_VALUES), [3])
    bb = make(8, 8, 8, engine=lb.ENGINE_INPLACE, bc=(1, 1, 1), boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    one_cell = np.zeros((8, 8, 8), np.uint8); one_cell[4, 4, 4] = 1
    bb.set_solid_mask(one_cell)
    bb.step(1)
    with pytest.raises(lb.LbmError):
        bb.set_solid_mask(one_cell)                                      # pushed form: the mask decides where populations sit
    bb.step(1)
    bb.set_solid_mask(one_cell)
    with pytest.raises(lb.LbmError):
        lat.simulateStreaming()                                          # fused steps only
    with pytest.raises(lb.LbmError):
        lat.load_data(O.new_cells(8, 8, 8))


def test_inplace_taylor_green_half_the_memory():
    n = 128
    tau, u0 = 0.6, 0.04
    k = 2 * np.pi / n
    x = np.arange(n)[:, None, None]; y = np.arange(n)[None, :, None]
    ux = (u0 * np.sin(k * x) * np.cos(k * y) * np.ones((1, 1, n))); uy = (-u0 * np.cos(k * x) * np.sin(k * y) * np.ones((1, 1, n)))
    c = O.DIRECTIONS[:19].astype(np.float64); w = O.weights(19)
    cu = c[:, 0, None, None, None] * ux + c[:, 1, None, None, None] * uy
    f = (w[:, None, None, None] * (1 + 3 * cu + 4.5 * cu ** 2 - 1.5 * (ux ** 2 + uy ** 2))).astype(np.float32)
    kw = dict(tau=tau, q=19, bc=(1, 1, 1), moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    two = make(n, n, n, **kw)
    one = make(n, n, n, engine=lb.ENGINE_INPLACE, **kw)
    assert one.layout().n_buffers * one.layout().state_bytes * 2 == two.layout().n_buffers * two.layout().state_bytes
    one.load_populations(f)
    m0 = one.mass()
    one.step(201)
    r, a, b, _ = one.moments()
    e1 = float((r.astype(np.float64) * (a.astype(np.float64) ** 2 + b.astype(np.float64) ** 2)).sum())
    e0 = float((ux ** 2 + uy ** 2).sum())
    assert e1 / e0 == pytest.approx(np.exp(-2.0 * (tau - 0.5) / 3.0 * 2 * k * k * 201), rel=0.02)
    assert one.mass() == pytest.approx(m0, rel=1e-6)
