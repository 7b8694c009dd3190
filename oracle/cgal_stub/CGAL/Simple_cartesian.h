// cgal_stub -- the smallest set of CGAL *type names* that lets the reference's
// simulation/include/Model.h and simulation/source/Lattice.cpp compile unmodified in a
// container that has no CGAL.  No CGAL algorithm is reproduced here: the
// Model member functions are provided by oracle/ref_model_stub.cpp on top of the CPU
// oracle's mesh queries (oracle/geom_oracle.c).  TEST INFRASTRUCTURE ONLY.
#ifndef LBM_CGAL_STUB_SIMPLE_CARTESIAN_H
#define LBM_CGAL_STUB_SIMPLE_CARTESIAN_H

namespace CGAL {

template <class FT>
struct Simple_cartesian {
    struct Point_3 {
        FT c[3];
        Point_3() : c{0, 0, 0} {}
        Point_3(FT x, FT y, FT z) : c{x, y, z} {}
        FT x() const { return c[0]; }
        FT y() const { return c[1]; }
        FT z() const { return c[2]; }
    };
    struct Iso_cuboid_3 {
        FT lo[3], hi[3];
        FT xmin() const { return lo[0]; }
        FT ymin() const { return lo[1]; }
        FT zmin() const { return lo[2]; }
        FT xmax() const { return hi[0]; }
        FT ymax() const { return hi[1]; }
        FT zmax() const { return hi[2]; }
    };
};

// handle types only; `impl` carries the oracle's mesh
template <class P> struct Surface_mesh { void* impl = nullptr; };
template <class M> struct AABB_face_graph_triangle_primitive {};
template <class K, class P> struct AABB_traits {};
template <class T> struct AABB_tree {};

}  // namespace CGAL
#endif
