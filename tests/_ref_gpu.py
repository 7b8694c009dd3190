"""ctypes binding of oracle/_ref/libref_gpu.so: the reference's own Lattice class and CUDA
kernels compiled unmodified (oracle/Makefile, oracle/ref_harness.cpp).  Test infrastructure only.
Nothing here reads /root/reference at run time; the .so is prebuilt and travels with the repo."""
import ctypes as C
import os

import numpy as np

import _oracle as O

SO = os.path.join(O.ORACLE_DIR, "_ref", "libref_gpu.so")


def available():
    return os.path.exists(SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        O.lib()   # liblbm_oracle.so must be resolvable (rpath $ORIGIN/..)
        L = C.CDLL(SO)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_create.argtypes = [i] * 9 + [f, f, f]
        L.ref_create.restype = vp
        for n in ("ref_destroy", "ref_prime", "ref_stream", "ref_reflect", "ref_collide", "ref_flow",
                  "ref_lattice_step", "ref_preprocess"):
            getattr(L, n).argtypes = [vp]
        L.ref_load.argtypes = [vp, vp]
        L.ref_retrieve.argtypes = [vp, vp]
        L.ref_insert_model.argtypes = [vp, C.c_char_p]
        L.ref_set_flow.argtypes = [vp, i, vp]
        L.ref_set_reflection_data.argtypes = [vp, f, f, f]
        L.ref_get_reflection_data.argtypes = [vp, vp]
        L.ref_fetch_table.argtypes = [vp, vp]
        L.ref_run_cli.argtypes = [vp, i, vp]
        L.ref_last_cuda_error.restype = C.c_char_p
        _lib = L
    return _lib


def launch_shape(X, Y, Z):
    """blocks*threads must equal the lattice exactly (the reference has no bounds guard)"""
    def split(n):
        for t in (8, 4, 2, 1):
            if n % t == 0:
                return n // t, t
    (bx, tx), (by, ty), (bz, tz) = split(X), split(Y), split(Z)
    return (bx, by, bz), (tx, ty, tz)


class RefLattice:
    """the reference's `Lattice`, driven through oracle/ref_harness.cpp"""

    def __init__(self, X, Y, Z, tau=1.0, spacing=1.0):
        assert X == Z, "the reference's device indexing needs X == Z (SURVEY.md D7)"
        b, t = launch_shape(X, Y, Z)
        self.X, self.Y, self.Z = X, Y, Z
        self.h = lib().ref_create(X, Y, Z, b[0], b[1], b[2], t[0], t[1], t[2], 1.0, tau, spacing)

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    __del__ = close

    def load(self, cells):
        a = np.ascontiguousarray(cells, dtype=O.LATTICE_POINT)
        lib().ref_load(self.h, O.ptr(a))

    def retrieve(self):
        out = np.zeros(self.X * self.Y * self.Z, dtype=O.LATTICE_POINT)
        lib().ref_retrieve(self.h, O.ptr(out))
        return out

    def prime(self):
        lib().ref_prime(self.h)

    def stream(self):
        lib().ref_stream(self.h)

    def reflect(self):
        lib().ref_reflect(self.h)

    def collide(self):
        lib().ref_collide(self.h)

    def flow(self):
        lib().ref_flow(self.h)

    def lattice_step(self):
        lib().ref_lattice_step(self.h)

    def insert_model(self, path):
        lib().ref_insert_model(self.h, path.encode())

    def preprocess(self):
        lib().ref_preprocess(self.h)

    def set_flow(self, kind, f27):
        a = np.ascontiguousarray(f27, np.float32)
        lib().ref_set_flow(self.h, kind, O.ptr(a))

    def set_reflection_data(self, x=0.0, y=0.0, z=0.0):
        lib().ref_set_reflection_data(self.h, x, y, z)

    def get_reflection_data(self):
        out = np.zeros(3, np.float32)
        lib().ref_get_reflection_data(self.h, O.ptr(out))
        return out

    def run_cli(self, nsteps, trace=False):
        tr = np.zeros((nsteps, 3), np.float32) if trace else None
        lib().ref_run_cli(self.h, nsteps, O.ptr(tr))
        return tr

    def tables(self, cells):
        """device-side ReflectionValues of every reflecting cell, in host order"""
        idx = np.nonzero(cells["is_reflected"])[0]
        out = np.zeros(len(idx), O.REFLECTION_VALUES)
        for n, k in enumerate(idx):
            rec = cells[k:k + 1].copy()
            lib().ref_fetch_table(O.ptr(rec), O.ptr(out[n:n + 1]))
        return out, idx
