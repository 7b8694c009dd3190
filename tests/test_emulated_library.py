"""The GPU parity tests, run on the CPU.

tests/_emu.py compiles the CUDA sources of liblbm_b200 -- kernels, C ABI, geometry, in-place engine -- for the
host on a small SIMT executor (tests/emu/: every CUDA thread a fiber, warps meet at the shuffle / ballot
intrinsics, device memory is host memory) and, for the duration of each test here, points the Python binding at
that build.  The test bodies are the ones of tests/test_gpu_parity.py, tests/test_gpu_geometry.py and
tests/test_io.py, unchanged: same oracle, same tolerances, small lattices only.  This checks the kernels' logic
and the host logic of the C ABI without a GPU.  It is test infrastructure: nothing in the package can load the
host build, and `pytest -m gpu` runs the same bodies on the real kernels.
"""
import itertools

import pytest

import _emu
import test_gpu_compat as TC
import test_gpu_geometry as TG
import test_gpu_parity as TP
import test_io as TI
import test_reference_vectors as TV

# (module, test, keep every n-th parametrisation) -- sized so that the module runs in about a minute
SELECTION = [
    (TP, "test_load_and_retrieve", 1), (TP, "test_coords_match_100_cubed", 1), (TP, "test_streaming_known_answer", 1),
    (TP, "test_collision_known_answer", 1), (TP, "test_full_known_answer", 1), (TP, "test_reflections_known_answer", 1),
    (TP, "test_flow_generation_known_answer", 1), (TP, "test_plane_flow_generation_known_answer", 1),
    (TP, "test_stream_bit_exact", 1), (TP, "test_collide_parity", 1), (TP, "test_reflect_parity_16", 1),
    (TP, "test_cli_steps_parity_with_geometry", 2), (TP, "test_force_accumulator_in_the_reference_arithmetic", 1),
    (TP, "test_source_with_an_arbitrary_criterion", 1), (TP, "test_unfused_phases_equal_fused_step", 1),
    (TP, "test_soa_engine_parity", 1), (TP, "test_bounce_back_obstacle_parity", 1), (TP, "test_lid_driven_cavity_parity", 1),
    (TP, "test_bad_arguments_are_reported", 1), (TP, "test_inplace_engine_equals_two_buffer_engine", 9),
    (TP, "test_inplace_engine_with_obstacles", 5), (TP, "test_inplace_engine_with_a_compact_body", 1),
    (TP, "test_inplace_engine_refuses_what_it_cannot_do", 1),
    (TG, "test_coordinates_with_model", 1), (TG, "test_mask_and_tables_bit_exact", 1),
    (TG, "test_missing_file_and_bounding_box_guard", 1), (TG, "test_reflection_table_known_answer", 1),
    (TG, "test_reflections_known_answer_end_to_end", 1),
    (TV, "test_library_streaming_equals_the_reference", 1), (TV, "test_library_collision_equals_the_reference", 1),
    (TV, "test_library_geometry_and_cli_loop_equal_the_reference", 1),
]


def parametrisations(fn):
    """the keyword sets pytest would generate from the function's own @pytest.mark.parametrize marks"""
    names, values = [], []
    for m in getattr(fn, "pytestmark", []):
        if m.name != "parametrize":
            continue
        n = [s.strip() for s in m.args[0].split(",")]
        names.append(n)
        values.append([v if len(n) > 1 else (v,) for v in m.args[1]])
    for combo in itertools.product(*values):
        kw = {}
        for n, v in zip(names, combo):
            kw.update(dict(zip(n, v)))
        yield kw


def _cases():
    out = []
    for mod, name, every in SELECTION:
        fn = getattr(mod, name)
        for k, kw in enumerate(parametrisations(fn)):
            if k % every == 0:
                label = name[5:] + ("[%s]" % "-".join(str(v) for v in kw.values()) if kw else "")
                out.append(pytest.param(fn, kw, id=label.replace(" ", "")))
    return out


@pytest.fixture(scope="module")
def host_build():
    with _emu.emulated_library() as lib:
        yield lib


@pytest.mark.parametrize("fn,kw", _cases())
def test_on_the_host_build(host_build, fn, kw):
    fn(**kw)


def test_voxeliser_with_a_large_mesh_on_the_host_build(host_build, tmp_path, monkeypatch):
    """the binned voxeliser against the oracle's all-pairs loop: 1 280 triangles at 24^3 here, 20 480 at 128^3 on the GPU"""
    monkeypatch.setattr(TG, "ICO_CASE", (3, 24))
    TG.test_voxeliser_scales_with_the_mesh(tmp_path)


def test_table_builder_with_a_large_mesh_on_the_host_build(host_build, tmp_path, monkeypatch):
    """the binned first-hit search against the oracle's all-pairs loop: 1 280 triangles at 20^3 here, 20 480 at 64^3 on the GPU"""
    monkeypatch.setattr(TG, "ICO_TABLE_CASE", (3, 20))
    TG.test_table_builder_with_a_large_mesh(tmp_path)


def test_checkpoint_round_trip_on_the_host_build(host_build, tmp_path):
    TI.test_checkpoint_round_trip_continues_bit_for_bit(tmp_path)


@pytest.mark.parametrize("body", [TC.test_reference_gtest_cases_through_the_cpp_headers,
                                  lambda: TC.test_cli_force_trace_matches_oracle([]),
                                  lambda: TC.test_cli_force_trace_matches_oracle(["--gpus", "2"]),
                                  TC.test_cli_rejects_unknown_options],
                         ids=["reference_gtest_cases", "cli_force_trace", "cli_force_trace_two_slabs", "cli_usage"])
def test_cpp_layer_on_the_host_build(monkeypatch, body):
    """the C++ drop-in layer (the reference's gtest cases, the CLI's force trace -- also cut into two slabs) linked
    against the host build"""
    monkeypatch.setattr(TC, "_built", _emu.compat_executable)
    body()


# ---- lattice shapes the GPU tests do not sweep: one-cell axes, nz around the 32-cell word, every edge rule ---------
# ("device" buffers of the host build end at an inaccessible page, so an out-of-range access faults)

@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 2, 2), (3, 1, 5), (1, 8, 40), (8, 1, 40), (2, 9, 33), (5, 4, 31), (5, 4, 32), (2, 2, 65), (6, 3, 97)])
@pytest.mark.parametrize("bc", [(1, 1, 1), (2, 2, 2), (0, 2, 1)])
def test_degenerate_shapes_on_the_host_build(host_build, shape, bc):
    import numpy as np
    import _oracle as O
    import latticeboltzmanncuda_b200 as lb
    X, Y, Z = shape
    q = 19 if (X + Y + Z) % 2 else 27
    wall_u = np.zeros((6, 3), np.float32)
    wall_u[3] = (0.03, 0.0, 0.01)
    wall_u[4] = (0.0, 0.02, 0.0)
    solid = None
    if X * Y * Z >= 8:
        solid = np.zeros(shape, np.uint8)
        solid[X // 2, Y // 2, Z // 2] = 1
        solid[0, 0, 0] = 1
        solid[X - 1, Y - 1, Z - 1] = 1
    f = TP.soa_state(q, X, Y, Z, 3, solid)
    steps = 2 if 0 in bc else 4
    want = O.step_soa(f, X, Y, Z, q, 0.7, bc=bc, wall_u=wall_u, solid=solid, moment=O.MOMENT_STANDARD, eq64=False, steps=steps)
    for engine in [lb.ENGINE_PULL2] + ([lb.ENGINE_INPLACE] if 0 not in bc else []):
        lat = lb.Lattice(X, Y, Z, None, None, lb.FluidData(1.0, 0.7), 0.1, q=q, bc=bc, wall_u=wall_u, moment_mode=lb.MOMENT_STANDARD,
                         eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK, engine=engine)
        if solid is not None:
            lat.set_solid_mask(solid)
        lat.load_populations(f)
        lat.step(steps)
        # (open edges drain a one-cell box to almost nothing within two steps: what is left is rounding)
        tol = 1e-4 if (0 in bc and X * Y * Z == 1) else TP.REL
        assert TP.rel_err(lat.retrieve_populations(), want) <= tol, "engine %d" % engine
        assert lat.mass() == pytest.approx(float(want.astype(np.float64).sum()), rel=1e-5, abs=1e-7)
        lat.close()


# ---- the reference itself on the host executor ---------------------------------------------------------------------
# tests/_emu.py reference_library() compiles the reference's own Lattice class and CUDA kernels -- the unmodified
# sources, where they lie under /root/reference, exactly the set oracle/Makefile compiles for the GPU -- on the same
# SIMT executor (output: oracle/_ref/libref_simt.so, git-ignored).  The bodies of tests/test_gpu_vs_reference.py then
# compare the reference's kernels, the oracle and this library's kernels without a GPU.

import _ref_gpu

try:
    import test_gpu_vs_reference as TR          # skips itself at import when oracle/_ref/libref_gpu.so was never built
except pytest.skip.Exception:
    TR = None


@pytest.fixture(scope="module")
def reference_on_the_host():
    so = _emu.reference_library()
    if so is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libref_simt.so")
    saved = (_ref_gpu.SO, _ref_gpu._lib)
    _ref_gpu.SO, _ref_gpu._lib = so, None
    yield so
    _ref_gpu.SO, _ref_gpu._lib = saved


def _reference_cases():
    out = []
    for name in sorted(n for n in dir(TR) if n.startswith("test_")) if TR else []:
        fn = getattr(TR, name)
        for kw in parametrisations(fn):
            label = name[5:] + ("[%s]" % "-".join(str(v) for v in kw.values()) if kw else "")
            out.append(pytest.param(fn, kw, id=label.replace(" ", "")))
    return out


@pytest.mark.parametrize("fn,kw", _reference_cases())
def test_against_the_reference_on_the_host_build(host_build, reference_on_the_host, fn, kw):
    fn(**kw)
