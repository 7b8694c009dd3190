"""ctypes binding of oracle/liblbm_oracle.so -- the CPU checker.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# simulation/include/LatticePoint.h:9-17 (144 B) and :3-7 (380 B)
LATTICE_POINT = np.dtype({
    "names": ["x", "y", "z", "f", "is_reflected", "is_internal", "reflections", "d_reflections"],
    "formats": ["<i4", "<i4", "<i4", ("<f4", 27), "u1", "u1", "<u8", "<u8"],
    "offsets": [0, 4, 8, 12, 120, 121, 128, 136],
    "itemsize": 144,
})
REFLECTION_VALUES = np.dtype([("dirs", "<i4", 14), ("w", "<f4", 81)])
assert REFLECTION_VALUES.itemsize == 380

SRC_NONE, SRC_CELL_111, SRC_PLANE_X0, SRC_LATTICE_PATTERN = 0, 1, 2, 3
BC_ZERO, BC_PERIODIC, BC_WALL = 0, 1, 2
MOMENT_REFERENCE, MOMENT_STANDARD = 0, 1


def _build():
    so = os.path.join(ORACLE_DIR, "liblbm_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, n) for n in ("lbm_oracle.c", "geom_oracle.c", "lbm_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liblbm_oracle.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build())
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.lbo_prime_points.argtypes = [vp, i, i, i]
        L.lbo_stream.argtypes = [vp, vp, i, i, i, i]
        L.lbo_collide.argtypes = [vp, i, i, i, f, i, i]
        L.lbo_reflect.argtypes = [vp, vp, i, i, i, vp, vp, vp]
        L.lbo_source.argtypes = [vp, i, i, i, i, vp]
        L.lbo_step_cli.argtypes = [vp, vp, i, i, i, f, vp, vp, vp, i, vp]
        L.lbo_step_lattice.argtypes = [vp, vp, i, i, i, f]
        L.lbo_step_soa.argtypes = [vp, vp, i, i, i, i, f, vp, vp, vp, i, i, vp]
        L.lbo_set_reflection_direction.argtypes = [vp, i, C.c_uint]
        L.lbo_get_reflection_direction.argtypes = [vp, i]
        L.lbo_get_reflection_direction.restype = i
        L.lbo_compress_ints.argtypes = [vp, i, i]
        L.lbo_compress_ints.restype = C.c_uint
        L.lbo_num_threads.restype = i
        L.lbo_mesh_load.argtypes = [C.c_char_p]
        L.lbo_mesh_load.restype = vp
        L.lbo_mesh_free.argtypes = [vp]
        for n in ("lbo_mesh_num_vertices", "lbo_mesh_num_faces", "lbo_mesh_is_closed", "lbo_mesh_check_error"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = i
        L.lbo_mesh_bbox.argtypes = [vp, vp, vp]
        L.lbo_mesh_point_inside.argtypes = [vp, f, f, f]
        L.lbo_mesh_point_inside.restype = i
        L.lbo_mesh_reflection_vector.argtypes = [vp, vp, vp, vp]
        L.lbo_rotate_vector_around_axis.argtypes = [vp, vp, f, vp]
        L.lbo_distribute_vector.argtypes = [vp, vp, vp]
        L.lbo_insert_model.argtypes = [vp, vp, i, i, i, f]
        L.lbo_insert_model.restype = i
        L.lbo_preprocess_model.argtypes = [vp, vp, i, i, i, f, vp, i, vp]
        L.lbo_preprocess_model.restype = i
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def new_cells(X, Y, Z):
    """zeroed host lattice in the reference's host order z + y*Z + x*Z*Y, coordinates primed"""
    cells = np.zeros(X * Y * Z, dtype=LATTICE_POINT)
    lib().lbo_prime_points(ptr(cells), X, Y, Z)
    return cells


def idx(x, y, z, Y, Z):
    return z + y * Z + x * Z * Y


class Mesh:
    def __init__(self, path):
        self.h = lib().lbo_mesh_load(path.encode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().lbo_mesh_free(self.h)
            self.h = None

    num_vertices = property(lambda s: lib().lbo_mesh_num_vertices(s.h))
    num_faces = property(lambda s: lib().lbo_mesh_num_faces(s.h))
    is_closed = property(lambda s: bool(lib().lbo_mesh_is_closed(s.h)))
    check_error = property(lambda s: bool(lib().lbo_mesh_check_error(s.h)))

    def point_inside(self, x, y, z):
        return bool(lib().lbo_mesh_point_inside(self.h, x, y, z))

    def reflection_vector(self, p1, p2):
        a = np.asarray(p1, np.float32); b = np.asarray(p2, np.float32); o = np.zeros(3, np.float32)
        lib().lbo_mesh_reflection_vector(self.h, ptr(a), ptr(b), ptr(o))
        return o

    def insert(self, cells, X, Y, Z, spacing):
        return lib().lbo_insert_model(self.h, ptr(cells), X, Y, Z, spacing)

    def preprocess(self, cells, X, Y, Z, spacing):
        tables = np.zeros(max(1, int(cells["is_internal"].sum())), dtype=REFLECTION_VALUES)
        tindex = np.full(X * Y * Z, -1, dtype=np.int32)
        n = lib().lbo_preprocess_model(self.h, ptr(cells), X, Y, Z, spacing, ptr(tables), len(tables), ptr(tindex))
        assert n >= 0
        return tables[:n].copy(), tindex


def rotate(axis, vec, angle):
    a = np.asarray(axis, np.float32); v = np.asarray(vec, np.float32); o = np.zeros(3, np.float32)
    lib().lbo_rotate_vector_around_axis(ptr(a), ptr(v), angle, ptr(o))
    return o


def distribute(vec):
    v = np.asarray(vec, np.float32); s = np.zeros(27, np.float32); d = np.zeros(27, np.int32)
    lib().lbo_distribute_vector(ptr(v), ptr(s), ptr(d))
    return s, d


def stream(cells, X, Y, Z, bc=BC_ZERO):
    out = np.zeros_like(cells)
    lib().lbo_stream(ptr(cells), ptr(out), X, Y, Z, bc)
    return out


def collide(cells, X, Y, Z, tau, q=27, moment=MOMENT_REFERENCE):
    out = cells.copy()
    lib().lbo_collide(ptr(out), X, Y, Z, tau, q, moment)
    return out


def reflect(cells, X, Y, Z, tables, tindex, force=None):
    out = np.zeros_like(cells)
    lib().lbo_reflect(ptr(cells), ptr(out), X, Y, Z, ptr(tables), ptr(tindex), ptr(force))
    return out


def source(cells, X, Y, Z, kind, flow27):
    out = cells.copy()
    fl = np.ascontiguousarray(flow27, np.float32)
    lib().lbo_source(ptr(out), X, Y, Z, kind, ptr(fl))
    return out


def step_cli(cells, X, Y, Z, tau, tables=None, tindex=None, force=None, src_kind=SRC_NONE, flow27=None, steps=1):
    cur = cells.copy()
    scratch = np.zeros_like(cur)
    fl = None if flow27 is None else np.ascontiguousarray(flow27, np.float32)
    for _ in range(steps):
        lib().lbo_step_cli(ptr(cur), ptr(scratch), X, Y, Z, tau, ptr(tables), ptr(tindex), ptr(force), src_kind, ptr(fl))
    return cur


def step_lattice(cells, X, Y, Z, tau, steps=1):
    cur = cells.copy()
    scratch = np.zeros_like(cur)
    for _ in range(steps):
        lib().lbo_step_lattice(ptr(cur), ptr(scratch), X, Y, Z, tau)
    return cur


def step_soa(f, X, Y, Z, q, tau, bc=(BC_PERIODIC,) * 3, wall_u=None, solid=None,
             moment=MOMENT_STANDARD, eq64=False, steps=1, force=None):
    """f: float32 [q, X, Y, Z] dense.  Returns the state after `steps` steps."""
    a = np.ascontiguousarray(f, np.float32).copy()
    b = np.zeros_like(a)
    bcv = np.asarray(bc, np.int32)
    wu = None if wall_u is None else np.ascontiguousarray(wall_u, np.float32).reshape(6, 3)
    so = None if solid is None else np.ascontiguousarray(solid, np.uint8)
    for _ in range(steps):
        lib().lbo_step_soa(ptr(a), ptr(b), X, Y, Z, q, tau, ptr(bcv), ptr(wu), ptr(so), moment, int(eq64), ptr(force))
        a, b = b, a
    return a


DIRECTIONS = np.array([
    [0, 0, 0], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [0, 0, 1], [0, 1, 0], [1, 0, 0],
    [-1, -1, 0], [-1, 0, -1], [-1, 0, 1], [-1, 1, 0], [0, -1, -1], [0, -1, 1],
    [0, 1, -1], [0, 1, 1], [1, -1, 0], [1, 0, -1], [1, 0, 1], [1, 1, 0],
    [-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], np.int32)
INVERSE = np.array([0, 6, 5, 4, 3, 2, 1, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 26, 25, 24, 23, 22, 21, 20, 19], np.int32)


def weights(q):
    if q == 27:
        return np.array([8 / 27] + [2 / 27] * 6 + [1 / 54] * 12 + [1 / 216] * 8)
    return np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)
