"""ctypes binding of liblbm_b200.so (include/lbm_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at first use.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

# --- enums of include/lbm_b200.h ----------------------------------------------------------
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED = range(6)
BC_ZERO_INFLOW, BC_PERIODIC, BC_WALL, BC_HALO = range(4)
MOMENT_REFERENCE, MOMENT_STANDARD = 0, 1
EQ_FP32, EQ_FP64 = 0, 1
BOUNDARY_REFLECT, BOUNDARY_BOUNCE_BACK = 0, 1
FORCE_REFERENCE_SLOT_INDEX, FORCE_PHYSICAL = 0, 1
FORCE_ACCUM_F64_PARTIALS, FORCE_ACCUM_REFERENCE_FP32 = 0, 1
PHASE_STREAM, PHASE_REFLECT, PHASE_COLLIDE, PHASE_SOURCE, PHASE_ALL = 1, 2, 4, 8, 15
SOURCE_NONE, SOURCE_BOX = 0, 1
ENGINE_AUTO, ENGINE_PULL2, ENGINE_INPLACE = range(3)
HALO_BLOB_BYTES = 320

# simulation/include/LatticePoint.h:9-17 (144 B) and :3-7 (380 B) as numpy record types
LATTICE_POINT = np.dtype({
    "names": ["x", "y", "z", "f", "is_reflected", "is_internal", "reflections", "d_reflections"],
    "formats": ["<i4", "<i4", "<i4", ("<f4", 27), "u1", "u1", "<u8", "<u8"],
    "offsets": [0, 4, 8, 12, 120, 121, 128, 136],
    "itemsize": 144,
})
REFLECTION_VALUES = np.dtype([("dirs", "<i4", 14), ("w", "<f4", 81)])


class Config(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("q", C.c_int32),
        ("tau", C.c_float),
        ("bc", C.c_int32 * 3),
        ("moment_mode", C.c_int32),
        ("eq_precision", C.c_int32),
        ("boundary_mode", C.c_int32),
        ("force_mode", C.c_int32),
        ("engine", C.c_int32),
        ("device", C.c_int32),
        ("wall_u", (C.c_float * 3) * 6),
        ("force_accum", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("q", C.c_int32),
        ("xp", C.c_int32), ("yp", C.c_int32), ("zp", C.c_int32), ("pad_z", C.c_int32),
        ("plane", C.c_int64), ("pop_stride", C.c_int64), ("state_bytes", C.c_int64),
        ("n_buffers", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [("words_total", C.c_int64), ("words_dead", C.c_int64), ("words_special", C.c_int64),
                ("links", C.c_int64), ("kernel_launches", C.c_int64)]


class LbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lbm_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/lbm_b200.h declares: (name, restype, argtypes)
_vp, _i, _u = C.c_void_p, C.c_int, C.c_uint
SYMBOLS = [
    ("lbm_config_default", _i, [C.POINTER(Config)]),
    ("lbm_create", _i, [C.POINTER(Config), C.POINTER(_vp)]),
    ("lbm_destroy", _i, [_vp]),
    ("lbm_upload_aos", _i, [_vp, _vp, C.c_size_t]),
    ("lbm_download_aos", _i, [_vp, _vp, C.c_size_t]),
    ("lbm_upload_f_soa", _i, [_vp, _vp]),
    ("lbm_download_f_soa", _i, [_vp, _vp]),
    ("lbm_set_solid_mask", _i, [_vp, _vp, _i]),
    ("lbm_get_solid_mask", _i, [_vp, _vp]),
    ("lbm_set_reflection_table", _i, [_vp, _vp, _vp, _i]),
    ("lbm_set_reflection_table_slab", _i, [_vp, _vp, _vp, _i]),
    ("lbm_set_slab_origin", _i, [_vp, _i, _i]),
    ("lbm_set_source", _i, [_vp, _i, _vp, _vp, _vp]),
    ("lbm_set_source_mask", _i, [_vp, _vp, _vp]),
    ("lbm_set_force_accumulator", _i, [_vp, _i, _vp]),
    ("lbm_set_force_accumulator_f64", _i, [_vp, _i, _vp]),
    ("lbm_get_force", _i, [_vp, _vp]),
    ("lbm_get_force_f64", _i, [_vp, _vp]),
    ("lbm_prime_points", _i, [_vp]),
    ("lbm_stream", _i, [_vp]),
    ("lbm_reflect", _i, [_vp]),
    ("lbm_collide", _i, [_vp]),
    ("lbm_source", _i, [_vp]),
    ("lbm_collide_tau", _i, [_vp, C.c_float]),
    ("lbm_reflect_accumulate", _i, [_vp, _vp]),
    ("lbm_step", _i, [_vp, _i]),
    ("lbm_step_phases", _i, [_vp, _i, _u]),
    ("lbm_step_timed", _i, [_vp, _i, _u, C.POINTER(C.c_float), C.POINTER(_i)]),
    ("lbm_moments", _i, [_vp, _vp, _vp, _vp, _vp]),
    ("lbm_mass", _i, [_vp, C.POINTER(C.c_double)]),
    ("lbm_momentum", _i, [_vp, _vp]),
    ("lbm_sync", _i, [_vp]),
    ("lbm_last_error", C.c_char_p, []),
    ("lbm_version", C.c_char_p, []),
    ("lbm_plane_ptr", _i, [_vp, _i, _i, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    ("lbm_next_plane_ptr", _i, [_vp, _i, _i, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    ("lbm_face_populations", _i, [_vp, _i, _vp, C.POINTER(_i)]),
    ("lbm_get_stream", _i, [_vp, C.POINTER(_vp)]),
    ("lbm_set_stream", _i, [_vp, _vp]),
    ("lbm_step_range", _i, [_vp, _u, _i, _i, _i]),
    ("lbm_step_commit", _i, [_vp]),
    ("lbm_fill_ghosts", _i, [_vp]),
    ("lbm_inplace_merge_edges", _i, [_vp]),
    ("lbm_halo_export", _i, [_vp, _vp, C.c_size_t]),
    ("lbm_halo_connect", _i, [_vp, _vp, _vp]),
    ("lbm_halo_disconnect", _i, [_vp]),
    ("lbm_halo_timeouts", _i, [_vp, C.POINTER(C.c_ulonglong)]),
    ("lbm_group_step", _i, [C.POINTER(_vp), _i, _i, _u]),
    ("lbm_group_collide_tau", _i, [C.POINTER(_vp), _i, C.c_float]),
    ("lbm_group_reflect_accumulate", _i, [C.POINTER(_vp), _i, _vp]),
    ("lbm_device_count", _i, [C.POINTER(_i)]),
    ("lbm_get_layout", _i, [_vp, C.POINTER(Layout)]),
    ("lbm_get_stats", _i, [_vp, C.POINTER(Stats)]),
    ("lbm_mesh_load", _i, [C.c_char_p, C.POINTER(_vp)]),
    ("lbm_mesh_destroy", _i, [_vp]),
    ("lbm_mesh_info", _i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    ("lbm_mesh_bounding_box", _i, [_vp, _vp, _vp]),
    ("lbm_mesh_point_inside", _i, [_vp, _vp, C.POINTER(_i)]),
    ("lbm_mesh_reflection_vector", _i, [_vp, _vp, _vp, _vp]),
    ("lbm_rotate_vector_around_axis", _i, [_vp, _vp, C.c_float, _vp]),
    ("lbm_distribute_vector", _i, [_vp, _vp, _vp]),
    ("lbm_compress_ints", _u, [_vp, _i, _i]),
    ("lbm_insert_model", _i, [_vp, _vp, C.c_float, C.POINTER(_i)]),
    ("lbm_preprocess_model", _i, [_vp, _vp, C.c_float]),
    ("lbm_get_reflection_tables", _i, [_vp, _vp, _vp, _i, C.POINTER(_i)]),
]

_lib = None


def library_path():
    # LBM_B200_LIB: a measurement build of the same sources (build.py --variant); never a different implementation
    return os.environ.get("LBM_B200_LIB") or _build.LIB


def load(build_if_missing=True):
    """Loads liblbm_b200.so, building it with nvcc first if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if build_if_missing and path == _build.LIB and _build.is_stale():
        _build.build_native()
    if not os.path.exists(path):
        raise ImportError("liblbm_b200.so is missing (%s): run `python -m latticeboltzmanncuda_b200.build`; "
                          "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise LbmError(rc, load().lbm_last_error().decode(errors="replace"))


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def default_config():
    cfg = Config()
    check(load().lbm_config_default(C.byref(cfg)))
    return cfg
