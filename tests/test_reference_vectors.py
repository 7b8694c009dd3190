"""The committed outputs of the reference itself (tests/golden/reference_vectors.npz, produced by
tests/golden/make_reference_vectors.py from the reference's unmodified Lattice class and kernels) against

  * the CPU oracle                       (always: this is what pins the oracle on a machine without /root/reference),
  * the sm_100a library through the C ABI (-m gpu).

Bit-exact for streaming, solid flags, reflecting cells and tables; collision within 1e-6 relative (the vectors come
from a g++ build of the reference's kernels: no FMA contraction, so a GPU run of the reference itself differs from them
in the last bits too); 20 steps of the CLI loop within the north star's 1e-5.
"""
import os

import numpy as np
import pytest

import _oracle as O

V = np.load(os.path.join(O.GOLDEN_DIR, "reference_vectors.npz"))
OBJ = os.path.join(O.GOLDEN_DIR, "obj", "cube_angled.obj")


def rel_err(got, want):
    scale = max(float(np.abs(want).max()), 1e-30)
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale


def small_cells():
    X, Y, Z = (int(v) for v in V["small_shape"])
    cells = O.new_cells(X, Y, Z)
    cells["f"] = V["small_f0"]
    return cells, X, Y, Z


def cli_case():
    n, tau, steps = V["cli_n_tau_steps"]
    return int(n), float(np.float32(tau)), int(steps), float(np.float32(1.0 / int(n)))


# ---- the oracle ---------------------------------------------------------------------------------------------------

def test_oracle_streaming_equals_the_reference():
    cells, X, Y, Z = small_cells()
    got = O.stream(cells, X, Y, Z)
    assert np.array_equal(got["f"], V["small_streamed"])


@pytest.mark.parametrize("tau,key", [(1.0, "small_collided_tau10"), (0.6, "small_collided_tau06")])
def test_oracle_collision_equals_the_reference(tau, key):
    cells, X, Y, Z = small_cells()
    got = O.collide(cells, X, Y, Z, np.float32(tau))
    assert rel_err(got["f"], V[key]) <= 1e-6


def test_oracle_geometry_equals_the_reference():
    n, tau, steps, spacing = cli_case()
    cells = O.new_cells(n, n, n)
    mesh = O.Mesh(OBJ)
    assert mesh.insert(cells, n, n, n, np.float32(spacing)) == 1
    tables, tindex = mesh.preprocess(cells, n, n, n, np.float32(spacing))
    assert np.array_equal(cells["is_internal"], V["cli_is_internal"])
    assert np.array_equal((cells["is_reflected"] != 0).astype(np.uint8), V["cli_is_reflected"])
    assert np.array_equal(np.nonzero(tindex >= 0)[0], V["cli_table_cells"])
    assert np.array_equal(tables["dirs"], V["cli_table_dirs"])
    assert np.array_equal(tables["w"], V["cli_table_w"])


def test_oracle_cli_loop_equals_the_reference():
    n, tau, steps, spacing = cli_case()
    flow = V["cli_flow"]
    cells = O.new_cells(n, n, n)
    cells["f"] = flow[None, :]
    mesh = O.Mesh(OBJ)
    mesh.insert(cells, n, n, n, np.float32(spacing))
    tables, tindex = mesh.preprocess(cells, n, n, n, np.float32(spacing))
    force = np.zeros(3, np.float32)
    got = O.step_cli(cells, n, n, n, np.float32(tau), tables, tindex, force, O.SRC_PLANE_X0, flow, steps=steps)
    assert rel_err(got["f"], V["cli_f_final"]) <= 1e-5


# ---- the CUDA library ---------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_library_streaming_equals_the_reference():
    import latticeboltzmanncuda_b200 as lb
    cells, X, Y, Z = small_cells()
    lat = lb.Lattice(X, Y, Z, (1, 1, 1), (X, Y, Z), lb.FluidData(1.0, 1.0), 0.1)
    lat.load_data(cells)
    lat.simulateStreaming()
    assert np.array_equal(lat.retrieve_data()["f"], V["small_streamed"])


@pytest.mark.gpu
@pytest.mark.parametrize("tau,key", [(1.0, "small_collided_tau10"), (0.6, "small_collided_tau06")])
def test_library_collision_equals_the_reference(tau, key):
    import latticeboltzmanncuda_b200 as lb
    cells, X, Y, Z = small_cells()
    lat = lb.Lattice(X, Y, Z, (1, 1, 1), (X, Y, Z), lb.FluidData(1.0, tau), 0.1)      # defaults = the reference's arithmetic
    lat.load_data(cells)
    lat.simulateCollision()
    assert rel_err(lat.retrieve_data()["f"], V[key]) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("eq", ["fp64", "fp32"])
def test_library_geometry_and_cli_loop_equal_the_reference(eq):
    import latticeboltzmanncuda_b200 as lb
    n, tau, steps, spacing = cli_case()
    flow = V["cli_flow"]
    lat = lb.Lattice(n, n, n, None, None, lb.FluidData(1.0, tau), spacing,
                     eq_precision=lb.EQ_FP64 if eq == "fp64" else lb.EQ_FP32)
    start = O.new_cells(n, n, n)
    start["f"] = flow[None, :]
    lat.load_data(start)
    assert lat.insertModel(OBJ)                          # GPU voxeliser
    lat.preProcessModel()                                # GPU table builder
    first = lat.retrieve_data()
    assert np.array_equal(first["is_internal"], V["cli_is_internal"])
    assert np.array_equal((first["is_reflected"] != 0).astype(np.uint8), V["cli_is_reflected"])
    tables, cells_with_table = lat.reflection_tables()
    order = np.argsort(cells_with_table)
    assert np.array_equal(cells_with_table[order], V["cli_table_cells"])
    assert np.array_equal(tables["dirs"][order], V["cli_table_dirs"])
    assert np.array_equal(tables["w"][order], V["cli_table_w"])
    lat.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
    lat.step(steps)
    assert rel_err(lat.retrieve_data()["f"], V["cli_f_final"]) <= 1e-5
