"""Analytic voxel masks of the reference's test geometry, for synthetic benchmark inputs.

The reference's flow-past-obstacle fixture is a tapered hexahedron (test/dataFiles/cube_angled.obj:
x in [0.1, 0.9] at z = 0.1 narrowing to [0.1, 0.4] at z = 0.9, y in [0.1, 0.9], in a unit box).
`tapered_hexahedron_mask` evaluates that solid on the lattice points (index * spacing, like
Lattice::insertModel, simulation/source/Lattice.cpp:248-257) without going through a mesh.
"""
import numpy as np


def tapered_hexahedron_mask(nx, ny, nz, x_planes=None):
    """uint8 [nx, ny, nz] (or the x-plane range `x_planes=(x0, x1)` of it), 1 = solid"""
    x0, x1 = (0, nx) if x_planes is None else x_planes
    px = (np.arange(x0, x1, dtype=np.float64) / nx)[:, None, None]
    py = (np.arange(ny, dtype=np.float64) / ny)[None, :, None]
    pz = (np.arange(nz, dtype=np.float64) / nz)[None, None, :]
    xmax = 0.9 - (pz - 0.1) * (0.5 / 0.8)
    inside = (py >= 0.1) & (py <= 0.9) & (pz >= 0.1) & (pz <= 0.9) & (px >= 0.1) & (px <= xmax)
    return inside.astype(np.uint8)
