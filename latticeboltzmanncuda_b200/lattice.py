"""Host-side mirror of the reference's `Lattice` orchestrator for the per-timestep path
(simulation/include/Lattice.h:40-71, simulation/source/Lattice.cpp).  Method names and
argument meaning follow the reference so that its tests read the same here; the work is
done by the sm_100a kernels behind the C ABI (include/lbm_b200.h).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import LATTICE_POINT, REFLECTION_VALUES


class FluidData:
    """simulation/include/Lattice.h:24-31 (the density is stored and never used, like the reference)"""

    def __init__(self, fluidDensity=1.0, characteristicTimescale=1.0):
        self.m_fluidDensity = float(fluidDensity)
        self.m_characteristicTimescale = float(characteristicTimescale)


class ReflectionData:
    """simulation/include/Lattice.h:33-38"""

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class FlowCriterion:
    """simulation/source/FlowCriterion.cu:5-16 as data: an inclusive index box (None = whole axis)"""
    ptestCriteria = ((1, 1, 1), (1, 1, 1))           # x == 1 && y == 1 && z == 1
    ptestPlaneCriteria = ((0, None, None), (0, None, None))   # x == 0


class FlowData:
    """simulation/include/FlowSurface.h:9-47: which cells get the additive source, and how much"""

    def __init__(self, criterion, inducedFlow):
        self.pointCriterion = criterion
        self.inducedFlow = np.zeros(27, np.float32)
        flow = np.asarray(inducedFlow["f"] if isinstance(inducedFlow, np.void) else inducedFlow, np.float32)
        self.inducedFlow[:flow.size] = flow.ravel()[:27]

    def setCriterion(self, criterion):
        self.pointCriterion = criterion

    def setInducedFlow(self, flow):
        self.inducedFlow[:] = np.asarray(flow, np.float32).ravel()[:27]


class Lattice:
    """Lattice(x, y, z, blocks, threads, fluid, spacing) -- Lattice.cpp:22-39.

    `blocks`/`threads` are accepted for signature compatibility; the launch shape is chosen
    by the library (the reference needs lattice == blocks*threads exactly and has no bounds
    guard).  Keyword options select the build extensions; the defaults are the
    reference's behaviour.
    """

    def __init__(self, x, y, z, blocks=None, threads=None, fluid=None, spacing=1.0, *, q=27,
                 bc=(capi.BC_ZERO_INFLOW,) * 3, moment_mode=capi.MOMENT_REFERENCE, eq_precision=capi.EQ_FP64,
                 boundary_mode=capi.BOUNDARY_REFLECT, force_mode=capi.FORCE_REFERENCE_SLOT_INDEX,
                 wall_u=None, device=-1, engine=capi.ENGINE_AUTO, force_accum=capi.FORCE_ACCUM_F64_PARTIALS):
        self._lib = capi.load()
        self._h = C.c_void_p()
        fluid = fluid or FluidData()
        cfg = capi.default_config()
        cfg.nx, cfg.ny, cfg.nz, cfg.q = int(x), int(y), int(z), int(q)
        cfg.tau = fluid.m_characteristicTimescale
        for a in range(3):
            cfg.bc[a] = int(bc[a])
        cfg.moment_mode, cfg.eq_precision = int(moment_mode), int(eq_precision)
        cfg.boundary_mode, cfg.force_mode = int(boundary_mode), int(force_mode)
        cfg.device, cfg.engine = int(device), int(engine)
        cfg.force_accum = int(force_accum)
        if wall_u is not None:
            wu = np.asarray(wall_u, np.float32).reshape(6, 3)
            for i in range(6):
                for a in range(3):
                    cfg.wall_u[i][a] = float(wu[i, a])
        capi.check(self._lib.lbm_create(C.byref(cfg), C.byref(self._h)))
        self.cfg = cfg
        self.m_fluid = fluid
        self.m_latticeSpacing = float(spacing)
        self.m_blocks, self.m_threads = blocks, threads
        self.nx, self.ny, self.nz, self.q = int(x), int(y), int(z), int(q)
        self.m_reflectionData = None
        self.m_flowData = None
        self.m_reflections = np.zeros(0, REFLECTION_VALUES)
        self._table_cells = np.zeros(0, np.int64)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.lbm_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    @property
    def n_cells(self):
        return self.nx * self.ny * self.nz

    def getDimensions(self):
        return (self.nx, self.ny, self.nz)

    def point_at_coords(self, x, y, z):
        """host index of a cell: z + y*Z + x*Z*Y (Lattice.cpp:254)"""
        return z + y * self.nz + x * self.nz * self.ny

    # -- whole-lattice transfers (Lattice.cpp:73-113) ----------------------------------------
    def load_data(self, lattice_array):
        a = np.ascontiguousarray(lattice_array, dtype=LATTICE_POINT)
        capi.check(self._lib.lbm_upload_aos(self._h, capi.ptr(a), a.size))

    def retrieve_data(self):
        out = np.zeros(self.n_cells, dtype=LATTICE_POINT)
        capi.check(self._lib.lbm_download_aos(self._h, capi.ptr(out), out.size))
        return out

    def load_populations(self, f):
        """dense float32 [q, nx, ny, nz]: a numpy array (host) or a torch tensor (host or CUDA; the C ABI takes
        either kind of pointer)"""
        shape = (self.q, self.nx, self.ny, self.nz)
        if hasattr(f, "data_ptr"):
            if tuple(f.shape) != shape or str(f.dtype) != "torch.float32" or not f.is_contiguous():
                raise ValueError("expected a contiguous float32 tensor of shape %r" % (shape,))
            capi.check(self._lib.lbm_upload_f_soa(self._h, capi.ptr(f.data_ptr())))
            return
        a = np.ascontiguousarray(f, np.float32)
        if a.shape != shape:
            raise ValueError("expected shape %r" % (shape,))
        capi.check(self._lib.lbm_upload_f_soa(self._h, capi.ptr(a)))

    @staticmethod
    def _check_out(a, count, what):
        """a caller-supplied output buffer goes to C as a bare pointer: it must be float32, contiguous and large enough"""
        if isinstance(a, np.ndarray):
            ok = a.dtype == np.float32 and a.flags.c_contiguous and a.flags.writeable and a.size >= count
        elif hasattr(a, "data_ptr"):
            ok = str(a.dtype) == "torch.float32" and a.is_contiguous() and a.numel() >= count
        else:
            ok = False
        if not ok:
            raise ValueError("%s: expected a writable contiguous float32 buffer (numpy array or torch tensor) of at least %d elements" % (what, count))

    def retrieve_populations(self, out=None):
        if out is None:
            out = np.empty((self.q, self.nx, self.ny, self.nz), np.float32)
        else:
            self._check_out(out, self.q * self.n_cells, "retrieve_populations(out=)")
        capi.check(self._lib.lbm_download_f_soa(self._h, capi.ptr(out.data_ptr() if hasattr(out, "data_ptr") else out)))
        return out

    # -- geometry side inputs ----------------------------------------------------------------
    def set_solid_mask(self, solid, x_ghost_planes=False):
        if solid is None:
            capi.check(self._lib.lbm_set_solid_mask(self._h, None, 0))
            return
        a = np.ascontiguousarray(solid, np.uint8)
        planes = self.nx + (2 if x_ghost_planes else 0)
        if a.size != planes * self.ny * self.nz:
            raise ValueError("solid mask has %d cells, expected %d" % (a.size, planes * self.ny * self.nz))
        capi.check(self._lib.lbm_set_solid_mask(self._h, capi.ptr(a), int(bool(x_ghost_planes))))

    def get_solid_mask(self):
        out = np.zeros((self.nx, self.ny, self.nz), np.uint8)
        capi.check(self._lib.lbm_get_solid_mask(self._h, capi.ptr(out)))
        return out

    def set_reflection_tables(self, tables, cell_index):
        """what Lattice::preProcessModel leaves behind (Lattice.cpp:323-426): one
        ReflectionValues record per boundary-solid cell and that cell's host index"""
        t = np.ascontiguousarray(tables, dtype=REFLECTION_VALUES)
        k = np.ascontiguousarray(cell_index, dtype=np.int64)
        if t.size != k.size:
            raise ValueError("tables and cell_index differ in length")
        self.m_reflections, self._table_cells = t, k
        capi.check(self._lib.lbm_set_reflection_table(self._h, capi.ptr(t), capi.ptr(k), t.size))

    def insertModel(self, filename):
        """Lattice::insertModel (Lattice.cpp:228-260): voxelise the mesh on the GPU; a missing file or a mesh
        that leaves the box is silently ignored, like the reference"""
        from .model import Model
        self.m_simModel = Model(filename)
        accepted = C.c_int()
        capi.check(self._lib.lbm_insert_model(self._h, self.m_simModel.handle, self.m_latticeSpacing, C.byref(accepted)))
        return bool(accepted.value)

    def preProcessModel(self):
        """Lattice::preProcessModel (Lattice.cpp:323-426): build and install the reflection tables on the GPU"""
        if getattr(self, "m_simModel", None) is None:
            return
        capi.check(self._lib.lbm_preprocess_model(self._h, self.m_simModel.handle, self.m_latticeSpacing))
        self.m_reflections, self._table_cells = self.reflection_tables()

    def reflection_tables(self):
        n = C.c_int()
        capi.check(self._lib.lbm_get_reflection_tables(self._h, None, None, 0, C.byref(n)))
        t = np.zeros(n.value, REFLECTION_VALUES)
        k = np.zeros(n.value, np.int64)
        if n.value:
            capi.check(self._lib.lbm_get_reflection_tables(self._h, capi.ptr(t), capi.ptr(k), n.value, C.byref(n)))
        return t, k

    def distributeVector(self, vector):
        from .model import distributeVector
        return distributeVector(vector)[0]

    def setFlowData(self, flowData):
        """Lattice.cpp:428-436"""
        self.m_flowData = flowData
        if flowData is None:
            capi.check(self._lib.lbm_set_source(self._h, capi.SOURCE_NONE, None, None, None))
            return
        crit = flowData.pointCriterion
        f = np.ascontiguousarray(flowData.inducedFlow, np.float32)
        if callable(crit) or isinstance(crit, np.ndarray):
            # any predicate (the reference takes a device function pointer, FlowSurface.h:12): evaluated once, here,
            # on index grids -- crit(x, y, z) -> bool arrays -- or given as a [nx, ny, nz] array; kept as a bit per cell
            if callable(crit):
                x, y, z = np.meshgrid(np.arange(self.nx), np.arange(self.ny), np.arange(self.nz), indexing="ij", sparse=True)
                sel = np.broadcast_to(np.asarray(crit(x, y, z)), (self.nx, self.ny, self.nz))
            else:
                sel = crit
            sel = np.ascontiguousarray(sel, np.uint8).reshape(self.nx, self.ny, self.nz)
            capi.check(self._lib.lbm_set_source_mask(self._h, capi.ptr(sel), capi.ptr(f)))
            return
        lo_c, hi_c = crit
        dims = (self.nx, self.ny, self.nz)
        lo = np.array([0 if v is None else v for v in lo_c], np.int32)
        hi = np.array([dims[a] - 1 if v is None else v for a, v in enumerate(hi_c)], np.int32)
        capi.check(self._lib.lbm_set_source(self._h, capi.SOURCE_BOX, capi.ptr(lo), capi.ptr(hi), capi.ptr(f)))

    def setReflectionData(self, reflection, exact=None):
        """Lattice.cpp:61-71.  `exact`: three doubles to start the accumulator from instead of the float members of
        `reflection` (checkpoint restore: the kernels keep the sum in double)"""
        self.m_reflectionData = reflection
        if reflection is None:
            capi.check(self._lib.lbm_set_force_accumulator(self._h, 0, None))
            return
        if exact is not None:
            init64 = np.ascontiguousarray(exact, np.float64)
            if init64.shape != (3,):
                raise ValueError("exact: three doubles")
            capi.check(self._lib.lbm_set_force_accumulator_f64(self._h, 1, capi.ptr(init64)))
            return
        init = np.array([reflection.x, reflection.y, reflection.z], np.float32)
        capi.check(self._lib.lbm_set_force_accumulator(self._h, 1, capi.ptr(init)))

    def retrieve_reflection_data(self):
        """Lattice.cpp:438-443"""
        out = np.zeros(3, np.float32)
        capi.check(self._lib.lbm_get_force(self._h, capi.ptr(out)))
        if self.m_reflectionData is not None:
            self.m_reflectionData.x, self.m_reflectionData.y, self.m_reflectionData.z = map(float, out)
            return self.m_reflectionData
        return ReflectionData(*out)

    def force_f64(self):
        """the accumulated force before it is narrowed to the reference's float triple"""
        out = np.zeros(3, np.float64)
        capi.check(self._lib.lbm_get_force_f64(self._h, capi.ptr(out)))
        return out

    # -- the per-timestep path (Lattice.cpp:115-226) ------------------------------------------
    def prime_points(self):
        capi.check(self._lib.lbm_prime_points(self._h))

    def simulateStreaming(self):
        capi.check(self._lib.lbm_stream(self._h))

    def simulateReflections(self):
        capi.check(self._lib.lbm_reflect(self._h))

    def simulateCollision(self):
        capi.check(self._lib.lbm_collide(self._h))

    def simulateFlow(self):
        if self.m_flowData is not None:
            capi.check(self._lib.lbm_source(self._h))

    def simulateLattice(self):
        """stream + collide (Lattice.cpp:222-226), fused"""
        capi.check(self._lib.lbm_step_phases(self._h, 1, capi.PHASE_STREAM | capi.PHASE_COLLIDE))

    def step(self, n=1, phases=capi.PHASE_ALL):
        """n fused steps in the CLI's order: stream, reflect, collide, source
        (LatticeBoltzmannCuda.cpp:89-92)"""
        capi.check(self._lib.lbm_step_phases(self._h, int(n), int(phases)))

    def step_timed(self, n, phases=capi.PHASE_ALL):
        ms, launches = C.c_float(), C.c_int()
        capi.check(self._lib.lbm_step_timed(self._h, int(n), int(phases), C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    def sync(self):
        capi.check(self._lib.lbm_sync(self._h))

    # -- diagnostics ---------------------------------------------------------------------------
    def moments(self, out=None):
        """density and velocity fields [nx, ny, nz]; `out` = four float32 buffers of n_cells elements each (numpy
        arrays or torch tensors, host or CUDA) to fill instead of allocating host arrays"""
        n = self.n_cells
        if out is not None:
            if len(out) != 4:
                raise ValueError("moments(out=): four buffers (rho, ux, uy, uz)")
            for o in out:
                if o is not None:
                    self._check_out(o, n, "moments(out=)")
            ptrs = [None if o is None else capi.ptr(o.data_ptr() if hasattr(o, "data_ptr") else o) for o in out]
            capi.check(self._lib.lbm_moments(self._h, *ptrs))
            return out
        rho, ux, uy, uz = (np.empty(n, np.float32) for _ in range(4))
        capi.check(self._lib.lbm_moments(self._h, capi.ptr(rho), capi.ptr(ux), capi.ptr(uy), capi.ptr(uz)))
        shp = (self.nx, self.ny, self.nz)
        return rho.reshape(shp), ux.reshape(shp), uy.reshape(shp), uz.reshape(shp)

    def mass(self):
        m = C.c_double()
        capi.check(self._lib.lbm_mass(self._h, C.byref(m)))
        return m.value

    def momentum(self):
        out = np.zeros(3, np.float64)
        capi.check(self._lib.lbm_momentum(self._h, capi.ptr(out)))
        return out

    def stats(self):
        st = capi.Stats()
        capi.check(self._lib.lbm_get_stats(self._h, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in capi.Stats._fields_}

    def layout(self):
        lay = capi.Layout()
        capi.check(self._lib.lbm_get_layout(self._h, C.byref(lay)))
        return lay
