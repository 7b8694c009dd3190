#!/usr/bin/env python
"""Writes an icosphere as a Wavefront OBJ: the closed, all-triangle mesh with many faces that the voxeliser tests need
(the reference's own fixtures have 12 triangles).  `level` subdivisions of the icosahedron give 20 * 4**level faces
(level 5: 20 480 faces, 10 242 vertices).  Centre and radius are chosen off the lattice so that no lattice point lies
exactly on a face.

    python tests/golden/make_icosphere.py out.obj [level] [cx cy cz radius]
"""
import sys

import numpy as np


def icosphere(level):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(level):
        cache, out = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            out += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = out
    return np.array(verts), np.array(f, np.int64)


def write_icosphere(path, level=5, centre=(0.5037, 0.4961, 0.5113), radius=0.3719):
    v, f = icosphere(level)
    v = v * radius + np.asarray(centre, np.float64)
    with open(path, "w") as fh:
        fh.write("# icosphere, %d subdivisions: %d vertices, %d faces\n" % (level, len(v), len(f)))
        for p in v:
            fh.write("v %.9g %.9g %.9g\n" % tuple(p))
        for t in f:
            fh.write("f %d %d %d\n" % tuple(t + 1))
    return len(v), len(f)


if __name__ == "__main__":
    a = sys.argv
    lvl = int(a[2]) if len(a) > 2 else 5
    if len(a) > 6:
        print(write_icosphere(a[1], lvl, tuple(float(x) for x in a[3:6]), float(a[6])))
    else:
        print(write_icosphere(a[1], lvl))
