"""Host-side mirror of the reference's `Model` (simulation/include/Model.h:23-47) and of the small geometry
helpers its API exposes, on top of the C ABI.  CGAL types do not appear: points are 3-sequences of floats."""
import ctypes as C

import numpy as np

from . import capi


def _f3(p):
    return np.ascontiguousarray(p, np.float32).reshape(3)


class Model:
    """Model::importModel / isModelClosed / isPointInsideModel / checkError / bounding_box / reflectionVector"""

    def __init__(self, filename=None):
        self._lib = capi.load()
        self._m = C.c_void_p()
        if filename is not None:
            self.importModel(filename)

    def importModel(self, filename):
        if self._m.value:
            self._lib.lbm_mesh_destroy(self._m)
            self._m = C.c_void_p()
        capi.check(self._lib.lbm_mesh_load(str(filename).encode(), C.byref(self._m)))

    def close(self):
        if getattr(self, "_m", None) is not None and self._m.value:
            self._lib.lbm_mesh_destroy(self._m)
            self._m = C.c_void_p()

    __del__ = close

    @property
    def handle(self):
        return self._m

    def _info(self):
        v = [C.c_int() for _ in range(4)]
        capi.check(self._lib.lbm_mesh_info(self._m, *[C.byref(x) for x in v]))
        return [x.value for x in v]

    def number_of_vertices(self):
        return self._info()[0]

    def number_of_faces(self):
        return self._info()[1]

    def isModelClosed(self):
        return bool(self._info()[2])

    def checkError(self):
        return not bool(self._info()[3])

    def bounding_box(self):
        lo, hi = np.zeros(3, np.float32), np.zeros(3, np.float32)
        capi.check(self._lib.lbm_mesh_bounding_box(self._m, capi.ptr(lo), capi.ptr(hi)))
        return lo, hi

    def isPointInsideModel(self, point):
        inside = C.c_int()
        p = _f3(point)
        capi.check(self._lib.lbm_mesh_point_inside(self._m, capi.ptr(p), C.byref(inside)))
        return bool(inside.value)

    def reflectionVector(self, point_1, point_2):
        a, b, out = _f3(point_1), _f3(point_2), np.zeros(3, np.float32)
        capi.check(self._lib.lbm_mesh_reflection_vector(self._m, capi.ptr(a), capi.ptr(b), capi.ptr(out)))
        return out


def rotateVectorAroundAxis(axis, vector, angle):
    """simulation/source/Geometry.cpp:22-42"""
    a, v, out = _f3(axis), _f3(vector), np.zeros(3, np.float32)
    capi.check(capi.load().lbm_rotate_vector_around_axis(capi.ptr(a), capi.ptr(v), float(np.float32(angle)), capi.ptr(out)))
    return out


def distributeVector(vector):
    """Lattice::distributeVector, simulation/source/Lattice.cpp:262-321 -> list of (score, direction)"""
    v, s, d = _f3(vector), np.zeros(27, np.float32), np.zeros(27, np.int32)
    capi.check(capi.load().lbm_distribute_vector(capi.ptr(v), capi.ptr(s), capi.ptr(d)))
    return list(zip(s.tolist(), d.tolist())), s, d


def compressInts(array, num, bitLength):
    """Utils::compressInts, simulation/source/Utils.cpp:4-29"""
    a = np.ascontiguousarray(array, np.int32)
    return int(capi.load().lbm_compress_ints(capi.ptr(a), int(num), int(bitLength)))


def getReflectionDirection(table, index):
    """Lattice::getReflectionDirection, Lattice.cpp:465-477: slot `index` of a ReflectionValues record"""
    word = int(np.asarray(table["dirs"]).astype(np.uint32)[index // 6])
    return (word >> (5 * (index % 6))) & 0x1F
