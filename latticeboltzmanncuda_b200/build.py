"""Builds liblbm_b200.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblbm_b200.so")
SOURCES = ["lbm_kernels.cu", "lbm_capi.cu", "lbm_geometry.cu", "lbm_inplace.cu", "lbm_halo.cu"]
# per-file extra flags: the geometry predicates must evaluate identically on host and device
EXTRA = {"lbm_geometry.cu": ["-fmad=false"]}
HEADERS = ["d3q.h", "lbm_device.h", "lbm_internal.h", "lbm_collide.inl", os.path.join("..", "..", "include", "lbm_b200.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in sources() + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False, variant=None, defines=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a -> latticeboltzmanncuda_b200/liblbm_b200.so

    variant / defines: a measurement build `liblbm_b200.<variant>.so` with extra -D macros (the launch-shape
    knobs at the top of csrc/lbm_kernels.cu); load it with LBM_B200_LIB=<path> (capi.load)."""
    global_lib = LIB if not variant else os.path.join(HERE, "liblbm_b200.%s.so" % variant)
    if not variant and not force and not is_stale():
        return LIB
    # several ranks of one job may find the library stale at the same moment: one builds, the others wait and re-check
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not variant and not force and not is_stale():
            return LIB
        return _build_locked(global_lib, variant, defines, verbose)


def _build_locked(global_lib, variant, defines, verbose):
    nvcc = nvcc_path()
    common = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden"] + ["-D" + d for d in defines]
    if verbose:
        common.append("-Xptxas=-v")
    objdir = os.path.join(CSRC, "_obj" if not variant else "_obj_" + variant)
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([nvcc] + common + EXTRA.get(src, []) + ["-c", src, "-o", obj], cwd=CSRC))
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", global_lib + ".tmp"] + objs, cwd=CSRC)
    os.replace(global_lib + ".tmp", global_lib)          # never a half-written library under the final name
    return global_lib


if __name__ == "__main__":
    # python -m latticeboltzmanncuda_b200.build [--force] [-v] [--variant NAME -DMACRO=VALUE ...]
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=name,
                       defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
