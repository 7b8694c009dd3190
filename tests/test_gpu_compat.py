"""The C++ drop-in layer (latticeboltzmanncuda_b200/include_compat): the reference's own gtest cases ported to a
plain executable and run against the reference-signature headers, and the CLI with the reference's positional
arguments, whose per-step force trace is compared with the CPU oracle.  Needs a GPU."""
import os
import subprocess

import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu

COMPAT = os.path.join(O.ROOT, "latticeboltzmanncuda_b200", "include_compat")
OBJ = os.path.join(O.GOLDEN_DIR, "obj")


def _built(name):
    path = os.path.join(COMPAT, name)
    if not os.path.exists(path):
        import latticeboltzmanncuda_b200.build as b
        b.build_native()
        subprocess.check_call(["make", "-s", "-C", COMPAT])
    return path


def test_reference_gtest_cases_through_the_cpp_headers():
    exe = _built("compat_selftest")
    res = subprocess.run([exe, OBJ], capture_output=True, text=True, timeout=600)
    print(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " 0 failed" in res.stdout


@pytest.mark.parametrize("extra", [[], ["--gpus", "2"], ["--gpus", "3", "--q", "27", "--bc", "zzz"]])
def test_cli_force_trace_matches_oracle(extra):
    """LatticeBoltzmannCuda.cpp:11-96: 11 positional arguments, one "Fx Fy Fz" line per step.  The named options are
    additions (the reference has no switches): --gpus N cuts the lattice into x slabs, one per device (devices are
    reused on a smaller box), joined by peer-memory halos -- the trace must not change."""
    exe = _built("lbm_cli")
    n, tau, steps = 16, 0.9, 25
    spacing = np.float32(1.0 / n)
    res = subprocess.run([exe, os.path.join(OBJ, "cube_angled.obj"), "8", "8", "8", "2", "2", "2", "1.0", str(tau),
                          repr(float(spacing)), str(steps)] + extra, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    got = np.array([[float(v) for v in line.split()] for line in res.stdout.strip().splitlines()])
    assert got.shape == (steps, 3)
    flow = np.full(27, 0.001, np.float32); flow[6] = 0.002
    cells = O.new_cells(n, n, n)
    cells["f"] = flow[None, :]
    mesh = O.Mesh(os.path.join(OBJ, "cube_angled.obj"))
    assert mesh.insert(cells, n, n, n, spacing) == 1
    tables, tindex = mesh.preprocess(cells, n, n, n, spacing)
    force = np.zeros(3, np.float32)
    want = []
    cur = cells
    for _ in range(steps):
        want.append(force.copy())               # the CLI prints the accumulator before each step
        cur = O.step_cli(cur, n, n, n, np.float32(tau), tables, tindex, force, O.SRC_PLANE_X0, flow, steps=1)
    want = np.array(want)
    assert np.allclose(got, want, rtol=5e-4, atol=1e-6)
    assert np.abs(got[-1]).max() > 0


def test_cli_rejects_unknown_options():
    exe = _built("lbm_cli")
    res = subprocess.run([exe, "x.obj", "8", "8", "8", "2", "2", "2", "1.0", "1.0", "0.1", "1", "--colour", "blue"],
                         capture_output=True, text=True, timeout=60)
    assert res.returncode != 0 and "usage" in res.stderr
