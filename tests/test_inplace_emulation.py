"""Host-side unit tests of the single-buffer engine's kernel source (tests/_emu.py: the CUDA source compiled for
the host, threads run one after the other) against the CPU oracle.  These check WHERE every population is read
from and written to -- both storage forms, obstacles, walls, periodic edges, source, momentum exchange -- without
a GPU; the GPU tests (tests/test_gpu_parity.py) run the real kernels through the C ABI."""
import numpy as np
import pytest

import _emu
import _oracle as O

REL = 1e-5


def rel_err(got, want):
    scale = max(float(np.abs(want).max()), 1e-30)
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale


def soa_state(q, X, Y, Z, seed, solid=None):
    rng = np.random.default_rng(seed)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1.0 + 0.1 * rng.standard_normal((q, X, Y, Z)))).astype(np.float32)
    if solid is not None:
        f[:, solid.astype(bool)] = 0
    return f


WALL_U = np.zeros((6, 3), np.float32)
WALL_U[3] = (0.03, 0.0, 0.01)
WALL_U[4] = (0.0, 0.02, 0.0)


@pytest.mark.parametrize("q", [19, 27])
@pytest.mark.parametrize("bc", [(1, 1, 1), (2, 2, 2), (1, 2, 1), (2, 1, 2)])
@pytest.mark.parametrize("steps", [1, 2, 5])
def test_all_fluid_lattice(q, bc, steps):
    """the harness itself, on the configuration the GPU tests already cover"""
    X, Y, Z = 9, 6, 37
    f = soa_state(q, X, Y, Z, 53)
    lat = _emu.Lattice(X, Y, Z, q, 0.7, bc, WALL_U)
    lat.load(f)
    lat.step(steps)
    want = O.step_soa(f, X, Y, Z, q, 0.7, bc=bc, wall_u=WALL_U, moment=O.MOMENT_STANDARD, eq64=False, steps=steps)
    assert rel_err(lat.retrieve(), want) <= REL


def obstacle(X, Y, Z, seed, frac=0.08):
    rng = np.random.default_rng(seed)
    solid = (rng.random((X, Y, Z)) < frac).astype(np.uint8)
    solid[4:8, 3:9, max(0, Z - 10):Z - 4] = 1          # a block that makes all-solid stretches
    solid[2, 1:5, :] = 1                               # a full z-row of solid words (dead words) ...
    solid[0, 0, 0] = 1; solid[X - 1, Y - 1, Z - 1] = 1 # ... and cells in the corners (images across every edge)
    return solid


@pytest.mark.parametrize("bc", [(1, 1, 1), (2, 2, 2), (1, 2, 1), (2, 1, 2), (1, 2, 2)])
@pytest.mark.parametrize("q,shape", [(19, (14, 12, 40)), (19, (6, 5, 70)), (19, (9, 7, 32)), (27, (6, 5, 70))])
def test_bounce_back_obstacles(q, bc, shape):
    X, Y, Z = shape
    solid = obstacle(X, Y, Z, 23 + X)
    f = soa_state(q, X, Y, Z, 29, solid)
    lat = _emu.Lattice(X, Y, Z, q, 0.65, bc, WALL_U)
    lat.set_solid_mask(solid)
    lat.load(f)
    lat.enable_force()
    force = np.zeros(3, np.float64)
    state = f
    for step in range(1, 8):
        lat.step(1)
        state = O.step_soa(state, X, Y, Z, q, 0.65, bc=bc, wall_u=WALL_U, solid=solid, moment=O.MOMENT_STANDARD, eq64=False,
                           steps=1, force=force)
        got = lat.retrieve()                           # after odd and after even step counts: both storage forms
        assert rel_err(got, state) <= REL, "step %d" % step
        assert np.all(got[:, solid.astype(bool)] == 0)
        assert np.allclose(lat.total_force(), force, rtol=1e-4, atol=1e-4), "step %d" % step
    rho, ux, uy, uz = lat.moments()
    c = O.DIRECTIONS[:q].astype(np.float64)
    r = state.astype(np.float64).sum(axis=0)
    assert np.allclose(rho, r, rtol=1e-5, atol=1e-6)
    mx = (state.astype(np.float64) * c[:, 0, None, None, None]).sum(axis=0)
    assert np.allclose(ux * rho, mx, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("q", [19, 27])
@pytest.mark.parametrize("bc", [(1, 1, 1), (2, 2, 2), (1, 2, 1)])
def test_sparse_obstacle_all_word_classes(q, bc):
    """a compact body in a mostly empty box: ordinary words (interior, z-edge and face launches), all-solid words
    and special words all occur, as in the 512^3 obstacle configuration"""
    X, Y, Z = 9, 8, 130
    solid = np.zeros((X, Y, Z), np.uint8)
    solid[3:6, 2:6, 30:100] = 1                        # covers words 1 and 2 of its rows completely
    solid[0, 4, 64] = 1                                # one cell on the x face
    solid[7, 7, 129] = 1                               # one on the y face at the z edge
    t = _emu.obstacle_tables(solid, list(bc))
    counts = [int((t["cls"][:, :, :Y] == c).sum()) for c in (0, 1, 2)]
    assert min(counts) > 0 and counts[0] > counts[2]
    f = soa_state(q, X, Y, Z, 11, solid)
    lat = _emu.Lattice(X, Y, Z, q, 0.7, bc, WALL_U)
    lat.set_solid_mask(solid)
    lat.enable_force()
    lat.load(f)
    force = np.zeros(3, np.float64)
    state = f
    for step in range(1, 6):
        lat.step(1)
        state = O.step_soa(state, X, Y, Z, q, 0.7, bc=bc, wall_u=WALL_U, solid=solid, moment=O.MOMENT_STANDARD, eq64=False,
                           steps=1, force=force)
        assert rel_err(lat.retrieve(), state) <= REL, "step %d" % step
        assert np.allclose(lat.total_force(), force, rtol=1e-4, atol=1e-4)
    assert lat.launches.value == 3 * 4 + 2 * 2      # odd: interior, y faces, x faces, special words; even: ordinary words, special words
    # without the force sum the even step is one launch and the solid lanes of the special kernel leave at once: same state
    quiet = _emu.Lattice(X, Y, Z, q, 0.7, bc, WALL_U)
    quiet.set_solid_mask(solid)
    quiet.load(f)
    quiet.step(5)
    assert np.array_equal(quiet.retrieve(), lat.retrieve())
    assert quiet.launches.value == 3 * 4 + 2 * 1


def test_obstacle_tables_layout():
    """the numpy rebuild of the mask / class / neighbourhood tables (what the GPU builders produce)"""
    X, Y, Z = 6, 5, 70
    solid = np.zeros((X, Y, Z), np.uint8)
    solid[3, 2, 40] = 1
    solid[1, :, 0:32] = 1                              # word 0 of every row of plane 1: all solid
    t = _emu.obstacle_tables(solid, [1, 2, 1])
    g = _emu.geom(X, Y, Z)
    assert t["mask"].shape == (g["xp"], g["yp"], g["wz"])
    assert (t["mask"][3 + 1, 2 + 1, (40 + 32) // 32] >> ((40 + 32) % 32)) & 1
    assert t["cls"][1, 0, :Y].tolist() == [1] * Y      # dead
    assert t["cls"][3, 1, 2] == 2 and t["cls"][3, 1, 1] == 2 and t["cls"][4, 1, 3] == 2     # special around the cell
    assert t["cls"][5, 2, 0] == 0
    k = int(np.nonzero(t["special"] == ((3 + 1) * g["yp"] + (2 + 1)) * g["wz"] + 2)[0][0])
    assert t["cellinfo"][k, 40 - 32] == 1 << _emu.nb_bit(0, 0, 0)
    assert t["cellinfo"][k, 41 - 32] == 1 << _emu.nb_bit(0, 0, -1)
    # periodic x: the image of plane 1's solid word is seen from plane 0 only (distance 1), not from plane X-1
    assert t["cls"][0, 0, 0] == 2 and t["cls"][X - 1, 0, 0] == 0


@pytest.mark.parametrize("q", [19, 27])
@pytest.mark.parametrize("with_solid", [False, True])
def test_box_source(q, with_solid):
    X, Y, Z = 10, 8, 40
    bc = (1, 2, 2)
    solid = obstacle(X, Y, Z, 5) if with_solid else None
    f = soa_state(q, X, Y, Z, 31, solid)
    src = (0.001 * (1 + np.arange(q))).astype(np.float32)
    lo, hi = (0, 0, 0), (0, Y - 1, Z - 1)             # the plane x == 0, like the CLI's source
    lat = _emu.Lattice(X, Y, Z, q, 0.8, bc)
    if with_solid:
        lat.set_solid_mask(solid)
    lat.set_source(lo, hi, src)
    lat.load(f)
    state = f
    for step in range(1, 6):
        lat.step(1)
        state = O.step_soa(state, X, Y, Z, q, 0.8, bc=bc, solid=solid, moment=O.MOMENT_STANDARD, eq64=False, steps=1)
        add = np.zeros_like(state)
        add[:, lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1] = src[:, None, None, None]
        if with_solid:
            add[:, solid.astype(bool)] = 0
        state = state + add
        assert rel_err(lat.retrieve(), state) <= REL, "step %d" % step


def test_reference_moments_and_fp64_equilibrium():
    X, Y, Z = 8, 6, 33
    solid = obstacle(X, Y, Z, 3)
    f = soa_state(27, X, Y, Z, 7, solid)
    lat = _emu.Lattice(X, Y, Z, 27, 0.9, (1, 1, 2), moment=0, eq64=True)
    lat.set_solid_mask(solid)
    lat.load(f)
    lat.step(4)
    want = O.step_soa(f, X, Y, Z, 27, 0.9, bc=(1, 1, 2), solid=solid, moment=O.MOMENT_REFERENCE, eq64=True, steps=4)
    assert rel_err(lat.retrieve(), want) <= REL
