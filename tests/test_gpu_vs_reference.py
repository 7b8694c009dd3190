"""The sm_100a path and the CPU oracle against the REFERENCE ITSELF: its unmodified Lattice.cpp,
SimCalcFuncs.cu and FlowCriterion.cu compiled into oracle/_ref/libref_gpu.so (oracle/Makefile;
only the CGAL-backed Model is substituted, see oracle/ref_model_stub.cpp).  Needs a GPU.

The reference's device indexing is only self-consistent for X == Z and its
launch must tile the lattice exactly, so all shapes here are X == Z.
"""
import os

import numpy as np
import pytest

import _oracle as O
import _ref_gpu as R

pytestmark = pytest.mark.gpu

lb = pytest.importorskip("latticeboltzmanncuda_b200")

if not R.available():
    pytest.skip("oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)", allow_module_level=True)

OBJ = os.path.join(O.GOLDEN_DIR, "obj")


def rel_err(got, want):
    scale = max(float(np.abs(want).max()), 1e-30)
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale


def random_cells(X, Y, Z, seed, amp=0.1):
    rng = np.random.default_rng(seed)
    cells = O.new_cells(X, Y, Z)
    w = O.weights(27).astype(np.float32)
    cells["f"] = (w[None, :] * (1.0 + amp * rng.standard_normal((X * Y * Z, 27)))).astype(np.float32)
    return cells


def ours(X, Y, Z, tau=1.0, spacing=0.1, **kw):
    return lb.Lattice(X, Y, Z, (1, 1, 1), (X, Y, Z), lb.FluidData(1.0, tau), spacing, **kw)


def test_reference_harness_reproduces_its_own_known_answer():
    """CalculationsTest.TestCollision through the real reference code: validates the harness"""
    ref = R.RefLattice(1, 1, 1, tau=1.0)
    cells = O.new_cells(1, 1, 1)
    cells[0]["f"][0] = 2.0
    ref.load(cells); ref.prime(); ref.collide()
    out = ref.retrieve()
    assert out[0]["f"][0] == pytest.approx(0.592592593, abs=1e-6)
    assert out[0]["f"][1] == pytest.approx(0.148148148, abs=1e-6)


@pytest.mark.parametrize("shape", [(4, 4, 4), (8, 6, 8), (16, 10, 16), (40, 8, 40)])
def test_stream_equals_reference_kernel(shape):
    X, Y, Z = shape
    cells = random_cells(X, Y, Z, 1)
    ref = R.RefLattice(X, Y, Z)
    ref.load(cells); ref.prime(); ref.stream()
    want = ref.retrieve()
    lat = ours(X, Y, Z)
    lat.load_data(cells); lat.simulateStreaming()
    got = lat.retrieve_data()
    assert np.array_equal(got["f"], want["f"])
    assert np.array_equal(O.stream(cells, X, Y, Z)["f"], want["f"])      # and so does the oracle


@pytest.mark.parametrize("shape", [(4, 4, 4), (16, 10, 16), (40, 8, 40)])
@pytest.mark.parametrize("tau", [1.0, 0.6])
def test_collide_equals_reference_kernel(shape, tau):
    X, Y, Z = shape
    cells = random_cells(X, Y, Z, 2, amp=0.3)
    ref = R.RefLattice(X, Y, Z, tau=tau)
    ref.load(cells); ref.prime(); ref.collide()
    want = ref.retrieve()
    lat = ours(X, Y, Z, tau=tau, eq_precision=lb.EQ_FP64)
    lat.load_data(cells); lat.simulateCollision()
    got = lat.retrieve_data()
    assert rel_err(got["f"], want["f"]) <= 2e-7        # reference arithmetic mode: last-bit agreement
    fast = ours(X, Y, Z, tau=tau, eq_precision=lb.EQ_FP32)
    fast.load_data(cells); fast.simulateCollision()
    assert rel_err(fast.retrieve_data()["f"], want["f"]) <= 1e-5
    assert rel_err(O.collide(cells, X, Y, Z, tau)["f"], want["f"]) <= 2e-7


def test_oracle_geometry_equals_reference_preprocessing():
    """Lattice::insertModel + preProcessModel (unmodified Lattice.cpp) vs the oracle's restatement:
    same solid flags, same reflecting cells, bit-identical tables"""
    n = 16
    spacing = float(np.float32(1.0 / n))
    path = os.path.join(OBJ, "cube_angled.obj")
    ref = R.RefLattice(n, n, n, spacing=spacing)
    ref.load(O.new_cells(n, n, n))
    ref.insert_model(path)
    ref.preprocess()
    rc = ref.retrieve()
    rtab, ridx = ref.tables(rc)
    cells = O.new_cells(n, n, n)
    mesh = O.Mesh(path)
    assert mesh.insert(cells, n, n, n, np.float32(spacing)) == 1
    tables, tindex = mesh.preprocess(cells, n, n, n, np.float32(spacing))
    assert np.array_equal(rc["is_internal"], cells["is_internal"])
    assert np.array_equal(rc["is_reflected"] != 0, cells["is_reflected"] != 0)
    assert np.array_equal(ridx, np.nonzero(tindex >= 0)[0])
    assert np.array_equal(rtab["dirs"], tables["dirs"])
    assert np.array_equal(rtab["w"], tables["w"])


def reference_cli_run(n, tau, steps, with_source=True):
    spacing = float(np.float32(1.0 / n))
    path = os.path.join(OBJ, "cube_angled.obj")
    flow = np.full(27, 0.001, np.float32); flow[6] = 0.002
    ref = R.RefLattice(n, n, n, tau=tau, spacing=spacing)
    cells = O.new_cells(n, n, n)
    cells["f"] = flow[None, :]
    ref.load(cells)
    ref.insert_model(path)
    ref.preprocess()
    if with_source:
        ref.set_flow(2, flow)
    # no force accumulator here: with it the reference kernel indexes DIRECTIONS[s]/SPLIT[s] with the slot
    # index s up to 80 (SimCalcFuncs.cu:181-183), past the end of its 27-entry tables -- undefined, and on
    # sm_100a an illegal address -- as soon as a table has a slot >= 27, i.e. on any real geometry
    ref.prime()
    start = ref.retrieve()
    rtab, ridx = ref.tables(start)
    trace = ref.run_cli(steps)
    return start, rtab, ridx, flow, ref.retrieve(), None, trace


@pytest.mark.parametrize("eq", [lb.EQ_FP64, lb.EQ_FP32])
@pytest.mark.parametrize("n,tau,steps", [(16, 1.0, 30), (32, 0.6, 20)])
def test_cli_loop_equals_reference(eq, n, tau, steps):
    """LatticeBoltzmannCuda.cpp:29-93 run by the reference's own Lattice class and kernels, against
    the fused sm_100a step fed with the reference's own mask and tables"""
    start, rtab, ridx, flow, want, want_force, trace = reference_cli_run(n, tau, steps)
    lat = ours(n, n, n, tau=tau, spacing=1.0 / n, eq_precision=eq)
    lat.load_data(start)
    lat.set_reflection_tables(rtab, ridx)
    lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
    lat.setReflectionData(lb.ReflectionData())
    lat.step(steps)
    got = lat.retrieve_data()
    assert rel_err(got["f"], want["f"]) <= 1e-5
    assert np.array_equal(got["is_internal"], want["is_internal"])


def test_oracle_cli_loop_equals_reference():
    """pins the CPU oracle's four phases on the reference's own multi-step output"""
    n, tau, steps = 16, 1.0, 30
    start, rtab, ridx, flow, want, want_force, trace = reference_cli_run(n, tau, steps)
    tindex = np.full(n * n * n, -1, np.int32)
    tindex[ridx] = np.arange(len(ridx))
    force = np.zeros(3, np.float32)
    got = O.step_cli(start, n, n, n, tau, rtab, tindex, force, O.SRC_PLANE_X0, flow, steps=steps)
    assert rel_err(got["f"], want["f"]) <= 1e-5


def test_reference_reflection_known_answer_and_ours():
    """CalculationsTest.TestReflectionsData (test/CalculationsTest.cpp:148-197) run by the reference's own
    insertModel / preProcessModel / simulateReflections, and the same state through the sm_100a path.
    Only slots 3..5 are filled here, so the reference's force indexing stays inside its tables."""
    path = os.path.join(OBJ, "cube_angled.obj")
    ref = R.RefLattice(3, 3, 3, spacing=0.5)
    ref.load(O.new_cells(3, 3, 3))
    ref.insert_model(path)
    ref.preprocess()
    cells = ref.retrieve()
    p = lambda x, y, z: O.idx(x, y, z, 3, 3)
    cells[p(1, 1, 1)]["f"][1] = 1.0
    ref.load(cells)
    ref.set_reflection_data()
    ref.prime()
    rtab, ridx = ref.tables(cells)
    ref.reflect()
    want = ref.retrieve()
    wf = ref.get_reflection_data()
    assert want[p(2, 1, 2)]["f"][17] == pytest.approx(0.367402, abs=1e-6)
    assert want[p(1, 1, 2)]["f"][4] == pytest.approx(0.341121, abs=1e-6)
    assert want[p(2, 2, 2)]["f"][26] == pytest.approx(0.291477, abs=1e-6)
    # The reference's force sum also visits the 75 filled slots >= 6 of this table; for slots >= 27 it
    # multiplies a zero population by DIRECTIONS[s]/SPLIT[s] read past the end of those tables
    # (SimCalcFuncs.cu:181-183).  Whatever lies there decides between the gtest's numbers and NaN
    # (0 * inf); both have been observed on sm_100a, so only a finite result is held to the golden value.
    golden = [0.28086, -0.194318, 0.648261]
    assert all(np.isnan(v) or abs(v - g) <= 1e-6 for v, g in zip(wf, golden))
    lat = ours(3, 3, 3, spacing=0.5)
    lat.load_data(cells)
    lat.set_reflection_tables(rtab, ridx)
    lat.setReflectionData(lb.ReflectionData())
    lat.simulateReflections()
    got = lat.retrieve_data()
    assert np.array_equal(got["f"], want["f"])
    r = lat.retrieve_reflection_data()
    assert np.allclose([r.x, r.y, r.z], golden, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("kind", [1, 2])
def test_reference_flow_generation_and_ours(kind):
    """CalculationsTest.TestFlowGeneration / TestPlaneFlowGeneration through the reference's own generate_flow
    (device function pointer and all) and through the sm_100a path"""
    n = 8
    cells = random_cells(n, n, n, 4)
    flow = np.linspace(0.001, 0.027, 27).astype(np.float32)
    ref = R.RefLattice(n, n, n)
    ref.load(cells); ref.prime(); ref.set_flow(kind, flow); ref.flow()
    want = ref.retrieve()
    lat = ours(n, n, n)
    lat.load_data(cells)
    crit = lb.FlowCriterion.ptestCriteria if kind == 1 else lb.FlowCriterion.ptestPlaneCriteria
    lat.setFlowData(lb.FlowData(crit, flow))
    lat.simulateFlow()
    got = lat.retrieve_data()
    assert np.array_equal(got["f"], want["f"])
    changed = np.nonzero((want["f"] != cells["f"]).any(1))[0]
    assert len(changed) == (1 if kind == 1 else n * n)


def test_source_with_a_criterion_that_is_not_a_box_equals_the_reference_kernel():
    """the reference calls a device predicate for every cell (SimCalcFuncs.cu:206); ours is a bit per cell evaluated on
    the host (lbm_set_source_mask).  The predicate is one of our own, compiled as a __device__ function-pointer symbol
    next to the reference's sources (oracle/ref_extra_criterion.cu) and handed to the reference's unmodified
    generate_flow through its own FlowData mechanism."""
    from test_gpu_parity import lattice_pattern
    X = Y = Z = 16
    cells = random_cells(X, Y, Z, 9)
    flow = (np.arange(27, dtype=np.float32) + 1.0) * 1e-3
    ref = R.RefLattice(X, Y, Z)
    ref.load(cells); ref.prime(); ref.set_flow(3, flow); ref.flow()
    want = ref.retrieve()
    lat = ours(X, Y, Z)
    lat.load_data(cells)
    lat.setFlowData(lb.FlowData(lattice_pattern, flow))
    lat.simulateFlow()
    got = lat.retrieve_data()
    assert np.array_equal(got["f"], want["f"])
    assert np.array_equal(O.source(cells, X, Y, Z, O.SRC_LATTICE_PATTERN, flow)["f"], want["f"])      # and the oracle's restatement
    changed = np.any(want["f"] != cells["f"], axis=1)
    assert 0 < changed.sum() < X * Y * Z // 4
