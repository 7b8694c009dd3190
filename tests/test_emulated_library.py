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

# (module, test, keep every n-th parametrisation) -- sized so that the module runs in about a minute
SELECTION = [
    (TP, "test_load_and_retrieve", 1), (TP, "test_coords_match_100_cubed", 1), (TP, "test_streaming_known_answer", 1),
    (TP, "test_collision_known_answer", 1), (TP, "test_full_known_answer", 1), (TP, "test_reflections_known_answer", 1),
    (TP, "test_flow_generation_known_answer", 1), (TP, "test_plane_flow_generation_known_answer", 1),
    (TP, "test_stream_bit_exact", 1), (TP, "test_collide_parity", 1), (TP, "test_reflect_parity_16", 1),
    (TP, "test_cli_steps_parity_with_geometry", 2), (TP, "test_unfused_phases_equal_fused_step", 1),
    (TP, "test_soa_engine_parity", 1), (TP, "test_bounce_back_obstacle_parity", 1), (TP, "test_lid_driven_cavity_parity", 1),
    (TP, "test_bad_arguments_are_reported", 1), (TP, "test_inplace_engine_equals_two_buffer_engine", 9),
    (TP, "test_inplace_engine_with_obstacles", 5), (TP, "test_inplace_engine_refuses_what_it_cannot_do", 1),
    (TG, "test_coordinates_with_model", 1), (TG, "test_mask_and_tables_bit_exact", 1),
    (TG, "test_missing_file_and_bounding_box_guard", 1), (TG, "test_reflection_table_known_answer", 1),
    (TG, "test_reflections_known_answer_end_to_end", 1),
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


def test_checkpoint_round_trip_on_the_host_build(host_build, tmp_path):
    TI.test_checkpoint_round_trip_continues_bit_for_bit(tmp_path)


@pytest.mark.parametrize("body", [TC.test_reference_gtest_cases_through_the_cpp_headers, TC.test_cli_force_trace_matches_oracle],
                         ids=["reference_gtest_cases", "cli_force_trace"])
def test_cpp_layer_on_the_host_build(monkeypatch, body):
    """the C++ drop-in layer (the reference's gtest cases, the CLI's force trace) linked against the host build"""
    monkeypatch.setattr(TC, "_built", _emu.compat_executable)
    body()
