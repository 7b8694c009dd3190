"""State and field I/O around the hot path.

The reference has no writer: `retrieve_data()` / `load_data()` (simulation/source/Lattice.cpp:73-113) are its only
dump / restore.  Here a checkpoint is the populations (SoA, what the kernels hold), the solid mask, the reflection
tables and the source / force settings in one `.npz`; `write_vti` dumps the density and velocity moments as VTK
ImageData for inspection (ParaView etc.).
"""
import struct

import numpy as np

from . import capi
from .lattice import FlowData, ReflectionData


def _npz_path(path):
    path = str(path)
    return path if path.endswith(".npz") else path + ".npz"


def save_checkpoint(lat, path):
    """everything needed to continue the run on a lattice created with the same options: populations bit for bit, the
    force accumulator as the double sum the kernels keep.  The file is `path` with `.npz` appended if it lacks it
    (numpy's rule); load_checkpoint applies the same rule.  Returns the file name written."""
    tables, cells = lat.reflection_tables()
    flow = lat.m_flowData
    force = lat.force_f64() if lat.m_reflectionData is not None else np.zeros(0, np.float64)
    path = _npz_path(path)
    with open(path, "wb") as fh:
        np.savez(fh,
                 dims=np.array([lat.nx, lat.ny, lat.nz, lat.q], np.int32),
                 f=lat.retrieve_populations(),
                 solid=lat.get_solid_mask(),
                 tables=tables.view(np.uint8).reshape(-1, 380) if len(tables) else np.zeros((0, 380), np.uint8),
                 table_cells=cells,
                 source_box=np.array([[-1 if v is None else v for v in c] for c in flow.pointCriterion], np.int64) if flow else np.zeros((0, 3), np.int64),
                 source_f=flow.inducedFlow if flow else np.zeros(0, np.float32),
                 force=force)
    return path


def load_checkpoint(lat, path):
    """restores EVERY part of the state the checkpoint describes, including the absence of a mask, of tables, of a source
    or of a force accumulator: whatever the lattice held before is replaced"""
    with np.load(_npz_path(path)) as z:
        dims = tuple(int(v) for v in z["dims"])
        if dims != (lat.nx, lat.ny, lat.nz, lat.q):
            raise ValueError("checkpoint is for a %r lattice, this one is %r" % (dims, (lat.nx, lat.ny, lat.nz, lat.q)))
        # populations first: on the single-buffer engine a fresh state is in the parked form, the only one in which
        # the mask may change (include/lbm_b200.h, enum lbm_engine)
        lat.load_populations(z["f"])
        lat.set_solid_mask(z["solid"] if z["solid"].any() else None)
        if lat.cfg.boundary_mode == capi.BOUNDARY_REFLECT and lat.cfg.engine != capi.ENGINE_INPLACE and lat.cfg.bc[0] != capi.BC_HALO:
            tables = z["tables"].reshape(-1).view(capi.REFLECTION_VALUES) if len(z["table_cells"]) else np.zeros(0, capi.REFLECTION_VALUES)
            lat.set_reflection_tables(tables, z["table_cells"])
        if len(z["source_f"]):
            lo, hi = ([None if v < 0 else int(v) for v in row] for row in z["source_box"])
            lat.setFlowData(FlowData((tuple(lo), tuple(hi)), z["source_f"]))
        else:
            lat.setFlowData(None)
        if len(z["force"]):
            lat.setReflectionData(ReflectionData(*[float(v) for v in z["force"]]), exact=z["force"])
        else:
            lat.setReflectionData(None)


def write_vti(path, rho, ux, uy, uz, spacing=1.0, solid=None):
    """density + velocity on the lattice points as VTK XML ImageData (appended raw binary, little endian).
    Arrays are [nx, ny, nz] as `Lattice.moments()` returns them; VTK wants x fastest, so they are transposed."""
    nx, ny, nz = rho.shape
    arrays = [("density", np.ascontiguousarray(rho.transpose(2, 1, 0), "<f4"), 1)]
    vel = np.stack([ux, uy, uz], axis=-1).transpose(2, 1, 0, 3)
    arrays.append(("velocity", np.ascontiguousarray(vel, "<f4"), 3))
    if solid is not None:
        arrays.append(("solid", np.ascontiguousarray(np.asarray(solid, np.uint8).transpose(2, 1, 0)), 1))
    head, blobs, offset = [], [], 0
    for name, a, comps in arrays:
        vt = "Float32" if a.dtype == np.dtype("<f4") else "UInt8"
        head.append('<DataArray type="%s" Name="%s" NumberOfComponents="%d" format="appended" offset="%d"/>' % (vt, name, comps, offset))
        raw = a.tobytes()
        blobs.append(struct.pack("<Q", len(raw)) + raw)
        offset += 8 + len(raw)
    ext = "0 %d 0 %d 0 %d" % (nx - 1, ny - 1, nz - 1)
    xml = ('<?xml version="1.0"?>\n<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n'
           '<ImageData WholeExtent="%s" Origin="0 0 0" Spacing="%g %g %g">\n<Piece Extent="%s">\n<PointData Scalars="density" Vectors="velocity">\n%s\n'
           '</PointData>\n</Piece>\n</ImageData>\n<AppendedData encoding="raw">\n_' % (ext, spacing, spacing, spacing, ext, "\n".join(head)))
    with open(path, "wb") as fh:
        fh.write(xml.encode())
        for b in blobs:
            fh.write(b)
        fh.write(b"\n</AppendedData>\n</VTKFile>\n")
