"""TEST INFRASTRUCTURE ONLY.  Host-side unit testing of the single-buffer engine's kernel source.

`latticeboltzmanncuda_b200/csrc/lbm_inplace.cu` is rewritten (kernel launches -> loops over blocks and threads,
launch bounds dropped), compiled with g++ against tests/emu/cuda_shim.h and driven through ctypes, one thread
after the other.  This checks the indexing logic of the kernels (which slot every population is read from and
written to, in both storage forms, with obstacles, walls and periodic edges) in the CPU test suite.  It is not a
fallback: nothing in the package imports it, and the GPU tests run the real kernels through the C ABI.

The obstacle data structures the kernels read (padded bit mask, word classes, special list, neighbourhood
words) are produced on the GPU by build_mask / build_cellinfo / fill_cellinfo kernels in lbm_kernels.cu; here
they are rebuilt in numpy from their documented layout (csrc/lbm_device.h, lbm_kernels.cu:811-872).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "latticeboltzmanncuda_b200", "csrc")
EMU = os.path.join(HERE, "emu")
BUILD = os.path.join(EMU, "_build")
PAD_Z = 32


def _balanced(s, i):
    """s[i] == '(' -> index just past the matching ')'"""
    depth = 0
    for k in range(i, len(s)):
        if s[k] == "(":
            depth += 1
        elif s[k] == ")":
            depth -= 1
            if depth == 0:
                return k + 1
    raise ValueError("unbalanced parentheses")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(src):
    """CUDA source -> C++: `name<T...><<<grid, block, shmem, stream>>>(args);` becomes a loop over the grid"""
    out = ""
    while True:
        k = src.find("__launch_bounds__")
        if k < 0:
            break
        out += src[:k]
        src = src[_balanced(src, src.index("(", k)):]
    src = (out + src).replace("__noinline__", "__attribute__((noinline))")     # (libstdc++ spells the attribute itself that way)
    pat = re.compile(r"((?:[A-Za-z_]\w*::)*[A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(")
    out = ""
    while True:
        m = pat.search(src)
        if not m:
            break
        end = _balanced(src, m.end() - 1)
        cfg = _split_top(m.group(2))
        args = src[m.end():end - 1]
        out += src[:m.start()] + "EMU_LAUNCH((%s), (%s), %s(%s))" % (cfg[0], cfg[1], m.group(1), args)
        src = src[end:]
    return out + src


_lib = None
_full = None
# _FORTIFY_SOURCE off: the executor longjmps between fiber stacks, which glibc's checked longjmp refuses
GXX = ["g++", "-std=c++17", "-O1", "-fPIC", "-Wno-attributes", "-Wno-unknown-pragmas", "-ffp-contract=off",
       "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0"]


def _includes():
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    return ["-I", cuda_inc, "-I", CSRC, "-I", BUILD, "-I", EMU, "-I", os.path.join(ROOT, "include")]


def _newer(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def _common_deps():
    return [os.path.join(EMU, "simt.h"), os.path.join(EMU, "simt_runtime.cpp"), os.path.abspath(__file__),
            os.path.join(ROOT, "include", "lbm_b200.h")] + \
           [os.path.join(CSRC, h) for h in ("lbm_device.h", "lbm_internal.h", "d3q.h", "lbm_collide.inl")]


def lib():
    """the in-place engine's launch functions only, driven directly (class Lattice below)"""
    global _lib
    if _lib is not None:
        return _lib
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(CSRC, "lbm_inplace.cu")
    inc = os.path.join(BUILD, "lbm_inplace_emu.inc")
    so = os.path.join(BUILD, "libinplace_emu.so")
    if _newer(so, [src, os.path.join(EMU, "emu_api.cpp")] + _common_deps()):
        with open(src) as fh:
            text = rewrite(fh.read())
        with open(inc, "w") as fh:
            fh.write(text)
        subprocess.check_call(GXX + ["-shared", "-Wl,-Bsymbolic"] + _includes() +
                              [os.path.join(EMU, "emu_api.cpp"), os.path.join(EMU, "simt_runtime.cpp"), "-o", so])
    _lib = C.CDLL(so)
    return _lib


def full_library():
    """all of liblbm_b200 (the four .cu files: kernels, C ABI, geometry, in-place engine) compiled for the host on the
    SIMT executor of tests/emu/: returns the path of tests/emu/_build/liblbm_b200_emu.so"""
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "liblbm_b200_emu.so")
    units = ["lbm_kernels.cu", "lbm_capi.cu", "lbm_geometry.cu", "lbm_inplace.cu", "lbm_halo.cu"]
    import fcntl
    with open(os.path.join(BUILD, ".lock"), "w") as lock:          # spawned test workers may all arrive here at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        _build_full(so, units)
    return so


def _build_full(so, units):
    if _newer(so, [os.path.join(CSRC, u) for u in units] + _common_deps()):
        procs, objs = [], []
        for u in units:
            cpp = os.path.join(BUILD, u + ".emu.cpp")
            with open(os.path.join(CSRC, u)) as fh:
                text = rewrite(fh.read())
            with open(cpp, "w") as fh:
                fh.write('#include "simt.h"\n#line 1 "%s"\n' % os.path.join(CSRC, u) + text)
            obj = cpp + ".o"
            objs.append(obj)
            procs.append(subprocess.Popen(GXX + ["-c"] + _includes() + [cpp, "-o", obj]))
        rt = os.path.join(BUILD, "simt_runtime.o")
        procs.append(subprocess.Popen(GXX + ["-c"] + _includes() + [os.path.join(EMU, "simt_runtime.cpp"), "-o", rt]))
        for pr in procs:
            if pr.wait() != 0:
                raise RuntimeError("host build of the library failed")
        # -Bsymbolic: the stand-ins for the CUDA runtime calls must win over a libcudart that torch may have loaded globally
        subprocess.check_call(["g++", "-shared", "-Wl,-Bsymbolic", "-o", so + ".tmp"] + objs + [rt])
        os.replace(so + ".tmp", so)


def compat_executable(name):
    """`compat_selftest` / `lbm_cli` of latticeboltzmanncuda_b200/include_compat, linked against the host build"""
    lib_so = full_library()
    compat = os.path.join(ROOT, "latticeboltzmanncuda_b200", "include_compat")
    exe = os.path.join(BUILD, name + "_emu")
    srcs = [os.path.join(compat, "lbm_compat.cpp"), os.path.join(compat, name + ".cpp")]
    if _newer(exe, srcs + [os.path.join(compat, "lbm_reference_api.hpp"), lib_so]):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", compat, "-I", cuda_inc] + srcs +
                              ["-o", exe, "-L", BUILD, "-llbm_b200_emu", "-Wl,-rpath," + BUILD])
    return exe


def reference_library():
    """the REFERENCE's own Lattice class and CUDA kernels (unmodified sources under /root/reference, compiled where they
    lie, like oracle/Makefile does for the GPU) on the host executor: oracle/_ref/libref_simt.so.  Rebuilt only when
    /root/reference is present; outputs go to oracle/_ref/ (git-ignored), never into the repository."""
    ref = os.environ.get("LBM_REFERENCE_ROOT", "/root/reference")
    oracle = os.path.join(ROOT, "oracle")
    out_dir = os.path.join(oracle, "_ref")
    so = os.path.join(out_dir, "libref_simt.so")
    src_dir = os.path.join(ref, "simulation", "source")
    if not os.path.isdir(src_dir):
        return so if os.path.exists(so) else None
    import fcntl
    os.makedirs(os.path.join(out_dir, "simt"), exist_ok=True)
    with open(os.path.join(out_dir, "simt", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        mine = [os.path.join(oracle, f) for f in ("ref_harness.cpp", "ref_model_stub.cpp")]
        host = [os.path.join(src_dir, f) for f in ("Lattice.cpp", "Geometry.cpp", "Constants.cpp", "Utils.cpp")]
        # (+ a flow criterion of our own that is not an index box, next to the reference's two: oracle/ref_extra_criterion.cu)
        dev = [os.path.join(src_dir, f) for f in ("SimCalcFuncs.cu", "FlowCriterion.cu")] + [os.path.join(oracle, "ref_extra_criterion.cu")]
        if _newer(so, mine + host + dev + [os.path.join(EMU, "simt.h"), os.path.join(EMU, "simt_runtime.cpp"), os.path.abspath(__file__)]):
            import _oracle
            _oracle.lib()                                              # liblbm_oracle.so answers the mesh queries of the model stub
            cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
            inc = ["-I", os.path.join(oracle, "cgal_stub"), "-I", os.path.join(ref, "simulation", "include"), "-I", oracle,
                   "-I", cuda_inc, "-I", EMU]
            objs = []
            for k, f in enumerate(mine + host):
                obj = os.path.join(out_dir, "simt", "%d_%s.o" % (k, os.path.basename(f)))
                subprocess.check_call(["g++", "-O1", "-fPIC", "-std=c++17", "-w"] + inc + ["-c", f, "-o", obj])
                objs.append(obj)
            for f in dev:
                # the rewritten text goes to the compiler through a pipe: no copy of a reference source is ever written
                with open(f) as fh:
                    text = '#include "simt.h"\n' + rewrite(fh.read())
                obj = os.path.join(out_dir, "simt", os.path.basename(f) + ".host.o")
                subprocess.run(GXX + ["-w"] + inc + ["-x", "c++", "-", "-c", "-o", obj], input=text.encode(), check=True)
                objs.append(obj)
            rt = os.path.join(out_dir, "simt", "simt_runtime.o")
            subprocess.check_call(GXX + ["-c", "-I", cuda_inc, "-I", EMU, os.path.join(EMU, "simt_runtime.cpp"), "-o", rt])
            subprocess.check_call(["g++", "-shared", "-Wl,-Bsymbolic", "-o", so + ".tmp"] + objs + [rt, "-L", oracle, "-llbm_oracle",
                                   "-Wl,-rpath," + oracle])
            os.replace(so + ".tmp", so)
    return so


class emulated_library:
    """context manager: the Python binding (latticeboltzmanncuda_b200.capi) talks to the host build for the duration.
    For the CPU test suite only -- the binding itself never looks for anything but the CUDA library."""

    def __enter__(self):
        from latticeboltzmanncuda_b200 import capi
        global _full
        if _full is None:
            _full = C.CDLL(full_library())
            for name, res, args in capi.SYMBOLS:
                fn = getattr(_full, name)
                fn.restype = res
                fn.argtypes = args
        self.capi = capi
        self.saved = capi._lib
        capi._lib = _full
        return _full

    def __exit__(self, *exc):
        self.capi._lib = self.saved


class EmuArgs(C.Structure):
    _fields_ = [("f", C.c_void_p),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("q", C.c_int), ("moment", C.c_int), ("eq64", C.c_int),
                ("tau", C.c_float), ("parity", C.c_int), ("bc", C.c_int * 3), ("wall_u", C.c_float * 18),
                ("solid", C.c_void_p), ("cls", C.c_void_p), ("cls_pitch", C.c_int),
                ("special_list", C.c_void_p), ("n_special", C.c_longlong), ("cellinfo", C.c_void_p),
                ("force", C.c_void_p), ("src_on", C.c_int), ("src_lo", C.c_int * 3), ("src_hi", C.c_int * 3),
                ("src_f", C.c_float * 27)]


def nb_bit(dx, dy, dz):
    return ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1)


def geom(nx, ny, nz):
    out = (C.c_longlong * 6)()
    lib().emu_geom(nx, ny, nz, out)
    return dict(zip(("xp", "yp", "zp", "wz", "plane", "np"), [int(v) for v in out]))


def obstacle_tables(solid, bc):
    """solid: uint8 [X, Y, Z]; bc per axis (1 periodic, 2 wall).  Returns the padded bit mask, the class bytes
    ([x][word][y], pitch), the sorted special list and the neighbourhood words, laid out as the GPU builds them."""
    X, Y, Z = solid.shape
    g = geom(X, Y, Z)
    xp, yp, zp, wz = g["xp"], g["yp"], g["zp"], g["wz"]
    s = solid.astype(bool)
    # ghost layers axis by axis: periodic -> the wrapped cell, wall -> clear
    def pad(a, axis, periodic):
        lo = np.take(a, [-1], axis=axis) if periodic else np.zeros_like(np.take(a, [0], axis=axis))
        hi = np.take(a, [0], axis=axis) if periodic else np.zeros_like(np.take(a, [0], axis=axis))
        return np.concatenate([lo, a, hi], axis=axis)
    p = pad(pad(pad(s, 0, bc[0] == 1), 1, bc[1] == 1), 2, bc[2] == 1)          # [X+2, Y+2, Z+2]
    P = np.zeros((xp, yp, zp + 32), bool)                                      # one spare word so z + 1 always exists
    P[:, :, PAD_Z - 1:PAD_Z + Z + 1] = p
    mask = np.packbits(P[:, :, :zp].reshape(xp, yp, wz, 32), axis=-1, bitorder="little").view("<u4").reshape(xp, yp, wz)
    wcount = (Z + 31) // 32
    zz = wcount * 32
    nb = np.zeros((X, Y, zz), np.uint32)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                sh = P[1 + dx:1 + dx + X, 1 + dy:1 + dy + Y, PAD_Z + dz:PAD_Z + dz + zz]
                nb |= sh.astype(np.uint32) << np.uint32(nb_bit(dx, dy, dz))
    live = (np.arange(zz) < Z)[None, None, :]
    own = P[1:1 + X, 1:1 + Y, PAD_Z:PAD_Z + zz]
    nbw = nb.reshape(X, Y, wcount, 32)
    livew = np.broadcast_to(live, nb.shape).reshape(X, Y, wcount, 32)
    dead = np.all(own.reshape(X, Y, wcount, 32) | ~livew, axis=-1)
    touches = np.any((nbw != 0) & livew, axis=-1)
    cls3 = np.where(dead, 1, np.where(touches, 2, 0)).astype(np.uint8)         # [X, Y, wcount]
    pitch = ((Y + 15) // 16) * 16
    cls = np.zeros((X, wcount, pitch), np.uint8)
    cls[:, :, :Y] = cls3.transpose(0, 2, 1)
    xs, ys, ws = np.nonzero(cls3 == 2)
    widx = (((xs + 1) * yp + (ys + 1)) * wz + (ws + PAD_Z // 32)).astype(np.uint32)
    order = np.argsort(widx, kind="stable")
    widx = np.ascontiguousarray(widx[order])
    cellinfo = np.ascontiguousarray(nbw[xs[order], ys[order], ws[order], :]).astype(np.uint32)
    return dict(mask=np.ascontiguousarray(mask), cls=cls, pitch=pitch, special=widx, cellinfo=cellinfo)


class Lattice:
    """the in-place engine's host logic (lbm_capi.cu: fill_inplace, one_step, upload / download) over the emulated launches"""

    def __init__(self, X, Y, Z, q, tau, bc, wall_u=None, moment=1, eq64=False):
        self.shape = (X, Y, Z)
        self.q = q
        g = geom(X, Y, Z)
        self.g = g
        self.f = np.zeros(q * g["np"], np.float32)
        self.force = np.zeros(3 * 128, np.float64)
        self.force_on = False
        self.dirty = False
        self.tables = None
        a = EmuArgs()
        a.f = self.f.ctypes.data
        a.nx, a.ny, a.nz, a.q, a.moment, a.eq64 = X, Y, Z, q, moment, int(eq64)
        a.tau = tau
        a.parity = 0
        a.bc[:] = list(bc)
        wu = np.zeros((6, 3), np.float32) if wall_u is None else np.asarray(wall_u, np.float32).reshape(6, 3)
        a.wall_u[:] = [float(v) for v in wu.ravel()]
        self.a = a
        self.launches = C.c_int(0)

    def set_solid_mask(self, solid):
        assert self.a.parity == 0
        t = obstacle_tables(np.asarray(solid, np.uint8), list(self.a.bc))
        self.tables = t
        a = self.a
        if solid.any():
            a.solid = t["mask"].ctypes.data
            a.cls = t["cls"].ctypes.data
            a.cls_pitch = t["pitch"]
            a.special_list = t["special"].ctypes.data
            a.n_special = len(t["special"])
            a.cellinfo = t["cellinfo"].ctypes.data
        self.dirty = True

    def set_source(self, lo, hi, f):
        a = self.a
        a.src_on = 1
        a.src_lo[:] = list(lo); a.src_hi[:] = list(hi)
        a.src_f[:] = [float(v) for v in f] + [0.0] * (27 - len(f))

    def enable_force(self):
        self.force_on = True

    def load(self, f):
        X, Y, Z = self.shape
        for i in range(self.q):
            d = np.ascontiguousarray(f[i], np.float32)
            assert lib().emu_scatter(C.byref(self.a), d.ctypes.data_as(C.c_void_p), i) == 0
        self.a.parity = 0
        self.dirty = True

    def step(self, n=1):
        a = self.a
        for _ in range(n):
            a.force = self.force.ctypes.data if (self.force_on and a.solid) else None
            if self.dirty:
                if a.solid:
                    assert lib().emu_zero_solid(C.byref(a)) == 0
                self.dirty = False
            assert lib().emu_step(C.byref(a), C.byref(self.launches)) == 0
            a.parity ^= 1

    def retrieve(self):
        X, Y, Z = self.shape
        out = np.zeros((self.q, X, Y, Z), np.float32)
        for i in range(self.q):
            assert lib().emu_gather(C.byref(self.a), out[i].ctypes.data_as(C.c_void_p), i) == 0
        return out

    def moments(self):
        X, Y, Z = self.shape
        r = [np.zeros((X, Y, Z), np.float32) for _ in range(4)]
        assert lib().emu_moments(C.byref(self.a), *[v.ctypes.data_as(C.c_void_p) for v in r]) == 0
        return r

    def total_force(self):
        return self.force.reshape(128, 3).sum(axis=0)
