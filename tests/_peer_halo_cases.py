"""Case bodies shared by tests/test_peer_halo_emulated.py (host build, CPU suite) and tests/test_gpu_peer_halo.py (GPU):
a SlabGroup -- several slabs in one process joined by peer-memory halos -- against ONE handle on the whole domain
(bit for bit) and against the CPU oracle (tolerance).  Test infrastructure."""
import numpy as np

import _oracle as O
import latticeboltzmanncuda_b200 as lb
from latticeboltzmanncuda_b200 import capi
from latticeboltzmanncuda_b200.slab import SlabGroup

P, W, Z0 = lb.BC_PERIODIC, lb.BC_WALL, lb.BC_ZERO_INFLOW

WALL_U = np.zeros((6, 3), np.float32)
WALL_U[3] = (0.03, 0.0, 0.01)      # the y = Y-1 wall moves
WALL_U[4] = (0.0, 0.02, 0.0)       # the z = 0 wall moves


def _state(q, X, Y, Z, seed, solid=None):
    rng = np.random.default_rng(seed)
    w = O.weights(q).astype(np.float32)
    f = (w[:, None, None, None] * (1.0 + 0.05 * rng.standard_normal((q, X, Y, Z)))).astype(np.float32)
    if solid is not None:
        f[:, solid.astype(bool)] = 0
    return f


def _body(X, Y, Z, kind):
    s = np.zeros((X, Y, Z), np.uint8)
    if kind == "edges":
        # blocks that sit ON the cuts of an even 2- or 3-way split and on the periodic seam
        s[X // 2 - 1:X // 2 + 1, 2:5, 3:9] = 1
        s[X // 3, 1:3, 40:44] = 1
        s[0, Y - 3:Y - 1, 10:14] = 1
        s[X - 1, 1:4, 20:23] = 1
        s[(2 * X) // 3 - 1:(2 * X) // 3 + 1, 4, 33:70] = 1
    if kind == "slab_of_solid":
        # whole 32-cell words and whole 4-word blocks of solid (dead words, dead blocks) next to partly solid ones
        s[1:X - 1, 1:Y - 1, 30:165] = 1
        s[X // 2, 2, 60:70] = 0
        s[0, 0, 128:160] = 1
    return s


# name -> (engine, q, (X, Y, Z), world, bc_x, (bc_y, bc_z), body, source, force, steps, eq64/moment)
CASES = {
    "inplace_q19_periodic_w2": dict(engine=lb.ENGINE_INPLACE, q=19, shape=(8, 5, 40), world=2, bc=(P, P, P), steps=5),
    "inplace_q19_body_walls_w3": dict(engine=lb.ENGINE_INPLACE, q=19, shape=(10, 7, 70), world=3, bc=(P, W, P), body="edges",
                                      source=True, force=True, steps=5, wall_u=WALL_U),
    "inplace_q27_walls_w2": dict(engine=lb.ENGINE_INPLACE, q=27, shape=(6, 4, 33), world=2, bc=(P, W, W), steps=4, wall_u=WALL_U),
    "inplace_q19_thin_slabs_w2": dict(engine=lb.ENGINE_INPLACE, q=19, shape=(4, 6, 64), world=2, bc=(P, P, W), body="edges",
                                      force=True, steps=3),
    # long rows with a slab of solid: all-solid words, so every slab finds boxes of them (csrc/lbm_capi.cu find_dead_boxes)
    "inplace_q19_all_solid_words_w2": dict(engine=lb.ENGINE_INPLACE, q=19, shape=(6, 5, 170), world=2, bc=(P, W, W), body="slab_of_solid",
                                           force=True, steps=5, wall_u=WALL_U),
    "inplace_q19_self_ring_w1": dict(engine=lb.ENGINE_INPLACE, q=19, shape=(5, 4, 36), world=1, bc=(P, W, P), steps=3),
    "pull2_q19_body_w3": dict(engine=lb.ENGINE_PULL2, q=19, shape=(10, 7, 70), world=3, bc=(P, W, P), body="edges", source=True,
                              force=True, steps=4, wall_u=WALL_U),
    "pull2_q27_zero_inflow_w2": dict(engine=lb.ENGINE_PULL2, q=27, shape=(7, 5, 34), world=2, bc=(Z0, Z0, Z0), steps=3),
    "pull2_q19_periodic_w4": dict(engine=lb.ENGINE_PULL2, q=19, shape=(9, 4, 40), world=4, bc=(P, P, P), steps=3),
    "pull2_q19_thin_w2_walls": dict(engine=lb.ENGINE_PULL2, q=19, shape=(4, 5, 32), world=2, bc=(P, W, W), body="edges", steps=3,
                                    wall_u=WALL_U),
}


def _make(kind, c, devices=None):
    X, Y, Z = c["shape"]
    kw = dict(q=c["q"], moment_mode=lb.MOMENT_STANDARD, eq_precision=lb.EQ_FP32, boundary_mode=lb.BOUNDARY_BOUNCE_BACK,
              engine=c["engine"], wall_u=c.get("wall_u"))
    if kind == "single":
        return lb.Lattice(X, Y, Z, None, None, lb.FluidData(1.0, 0.7), 1.0, bc=c["bc"], **kw)
    return SlabGroup(X, Y, Z, lb.FluidData(1.0, 0.7), devices, bc_x=c["bc"][0], bc_yz=c["bc"][1:], **kw)


def run_case(name, devices):
    c = CASES[name]
    X, Y, Z = c["shape"]
    q = c["q"]
    solid = _body(X, Y, Z, c["body"]) if c.get("body") else None
    f0 = _state(q, X, Y, Z, 5, solid)
    flow = np.zeros(27, np.float32)
    flow[:q] = 1e-3
    flow[6] = 2e-3
    one = _make("single", c)
    grp = _make("group", c, devices(c["world"]))
    try:
        if solid is not None:
            one.set_solid_mask(solid)
            grp.set_solid_mask(solid)
        one.load_populations(f0)
        grp.load_populations(f0)
        if c.get("source"):
            one.setFlowData(lb.FlowData(((0, None, 2), (X // 2, None, Z - 3)), flow))     # a box that spans several slabs
            grp.set_source_global((0, None, 2), (X // 2, None, Z - 3), flow)
        if c.get("force"):
            one.setReflectionData(lb.ReflectionData())
            grp.enable_force()
        done = 0
        for chunk in (1, c["steps"] - 1):          # read back after an odd and after the final number of steps
            one.step(chunk)
            grp.step(chunk)
            done += chunk
            a, b = one.retrieve_populations(), grp.retrieve_populations()
            assert np.array_equal(a, b), "%s: slabs differ from the single domain after %d steps (max %g)" % (
                name, done, float(np.abs(a - b).max()))
        if c.get("force"):
            r = one.retrieve_reflection_data()
            assert np.allclose(grp.force(), [r.x, r.y, r.z], rtol=1e-6, atol=1e-9)
        assert abs(grp.mass() - one.mass()) <= 1e-9 * abs(one.mass())
        rho_a, rho_b = one.moments()[0], grp.moments()[0]
        assert np.array_equal(rho_a, rho_b)
        # and the single domain is right: the oracle on the same inputs
        bc = tuple({P: O.BC_PERIODIC, W: O.BC_WALL, Z0: O.BC_ZERO}[v] for v in c["bc"])
        if not c.get("source"):
            want = O.step_soa(f0, X, Y, Z, q, 0.7, bc=bc, wall_u=c.get("wall_u"), solid=solid, moment=O.MOMENT_STANDARD,
                              eq64=False, steps=c["steps"])
            err = float(np.abs(b.astype(np.float64) - want).max() / np.abs(want).max())
            assert err <= 1e-5, "%s: %g against the oracle" % (name, err)
        import ctypes as C
        for s in grp.slabs:
            n = C.c_ulonglong()
            capi.check(capi.load().lbm_halo_timeouts(s.handle, C.byref(n)))
            assert n.value == 0
    finally:
        one.close()
        grp.close()


def run_reference_mode_case(world, devices, n=16, steps=12, fused=True):
    """The reference's own mode (D3Q27, table reflection, fp64 equilibrium, SPLIT moments, zero-inflow edges, plane source,
    force) on slabs: geometry and tables built per slab in global coordinates, a body that crosses the cuts.  Against one
    handle on the whole lattice (bit for bit) and the oracle's CLI loop."""
    import os
    obj = os.path.join(O.GOLDEN_DIR, "obj", "cube_angled.obj")
    spacing = float(np.float32(1.0 / n))
    flow = np.full(27, 0.001, np.float32)
    flow[6] = 0.002
    start = O.new_cells(n, n, n)
    start["f"] = flow[None, :]
    one = lb.Lattice(n, n, n, None, None, lb.FluidData(1.0, 0.9), spacing)
    grp = SlabGroup(n, n, n, lb.FluidData(1.0, 0.9), devices(world), bc_x=Z0, bc_yz=(Z0, Z0))
    try:
        one.load_data(start)
        grp.load_data(start)
        assert one.insertModel(obj) and grp.insertModel(obj, spacing)
        assert np.array_equal(np.concatenate([s.get_solid_mask() for s in grp.slabs], axis=0), one.get_solid_mask())
        one.preProcessModel()
        grp.preProcessModel()
        t1, k1 = one.reflection_tables()
        per = n * n
        tg = [s.reflection_tables() for s in grp.slabs]
        kg = np.concatenate([k + x0 * per for (t, k), (x0, _) in zip(tg, grp.ranges)])
        assert np.array_equal(kg, k1)                                   # the same boundary-solid cells, each owned by one slab
        assert np.array_equal(np.concatenate([t for t, _ in tg])["w"], t1["w"])
        one.setFlowData(lb.FlowData(lb.FlowCriterion.ptestPlaneCriteria, flow))
        grp.set_source_global((0, None, None), (0, None, None), flow)
        one.setReflectionData(lb.ReflectionData())
        grp.enable_force()
        if fused:
            one.step(steps)
            grp.step(steps)
        else:
            for _ in range(steps):
                for ph in (lb.PHASE_STREAM, lb.PHASE_REFLECT, lb.PHASE_COLLIDE, lb.PHASE_SOURCE):
                    one.step(1, ph)
                    grp.step(1, ph)
        a, b = one.retrieve_data(), grp.retrieve_data()
        assert np.array_equal(a["f"], b["f"])
        assert np.array_equal(a["is_internal"], b["is_internal"]) and np.array_equal(a["is_reflected"], b["is_reflected"])
        assert np.array_equal(a["x"], b["x"])
        assert np.allclose(grp.force(), one.force_f64(), rtol=1e-9, atol=1e-12)
        cells = O.new_cells(n, n, n)
        cells["f"] = flow[None, :]
        mesh = O.Mesh(obj)
        assert mesh.insert(cells, n, n, n, np.float32(spacing)) == 1
        tables, tindex = mesh.preprocess(cells, n, n, n, np.float32(spacing))
        force = np.zeros(3, np.float32)
        want = O.step_cli(cells, n, n, n, 0.9, tables, tindex, force, O.SRC_PLANE_X0, flow, steps=steps)
        err = float(np.abs(b["f"].astype(np.float64) - want["f"]).max() / np.abs(want["f"]).max())
        assert err <= 1e-5
        assert np.allclose(grp.force(), force, rtol=2e-4, atol=1e-7)
    finally:
        one.close()
        grp.close()


def check_refusals():
    import ctypes as C
    lib = capi.load()
    a = lb.Lattice(4, 4, 32, None, None, lb.FluidData(1.0, 0.7), 1.0, q=19, bc=(lb.BC_HALO, P, P), boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    b = lb.Lattice(4, 5, 32, None, None, lb.FluidData(1.0, 0.7), 1.0, q=19, bc=(lb.BC_HALO, P, P), boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    c = lb.Lattice(4, 4, 32, None, None, lb.FluidData(1.0, 0.7), 1.0, q=19, bc=(P, P, P), boundary_mode=lb.BOUNDARY_BOUNCE_BACK)
    try:
        ba, bb = (C.c_ubyte * capi.HALO_BLOB_BYTES)(), (C.c_ubyte * capi.HALO_BLOB_BYTES)()
        assert lib.lbm_halo_connect(a.handle, ba, ba) == capi.ERR_STATE            # connect before export
        capi.check(lib.lbm_halo_export(a.handle, ba, capi.HALO_BLOB_BYTES))
        capi.check(lib.lbm_halo_export(b.handle, bb, capi.HALO_BLOB_BYTES))
        assert lib.lbm_halo_export(c.handle, bb, capi.HALO_BLOB_BYTES) == capi.ERR_STATE   # x axis is not a halo axis
        assert lib.lbm_halo_export(a.handle, ba, 16) == capi.ERR_INVALID
        assert lib.lbm_halo_connect(a.handle, bb, bb) == capi.ERR_INVALID          # another cross-section
        junk = (C.c_ubyte * capi.HALO_BLOB_BYTES)()
        assert lib.lbm_halo_connect(a.handle, junk, junk) == capi.ERR_INVALID
        capi.check(lib.lbm_halo_connect(a.handle, ba, ba))                         # its own neighbour: a ring of one
        assert lib.lbm_halo_connect(a.handle, ba, ba) == capi.ERR_STATE            # twice
        capi.check(lib.lbm_halo_disconnect(a.handle))
    finally:
        a.close(); b.close(); c.close()
