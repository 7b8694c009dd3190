// ref_model_stub.cpp -- `Model` member functions (simulation/include/Model.h:23-47) for the
// reference-under-test build in oracle/_ref/.  The reference's own Model.cpp needs CGAL,
// which is absent here, so the mesh queries are answered by the CPU oracle's CGAL-free
// restatement (oracle/geom_oracle.c); every other reference source is compiled unmodified.
// TEST INFRASTRUCTURE ONLY.
#include "Model.h"
#include "lbm_oracle.h"

void Model::importModel(std::string filename)
{
    if (modelMesh.impl) lbo_mesh_free((lbo_mesh*)modelMesh.impl);
    lbo_mesh* m = lbo_mesh_load(filename.c_str());
    modelMesh.impl = m;
    meshValid = !lbo_mesh_check_error(m);
}

bool Model::isModelClosed() { return modelMesh.impl && lbo_mesh_is_closed((lbo_mesh*)modelMesh.impl); }

bool Model::isPointInsideModel(CGAL::Simple_cartesian<float>::Point_3 p)
{
    if (!meshValid || !modelMesh.impl) return false;
    return lbo_mesh_point_inside((lbo_mesh*)modelMesh.impl, p.x(), p.y(), p.z()) != 0;
}

void Model::lazyAabbTreeConstruction() {}

CGAL::Simple_cartesian<float>::Iso_cuboid_3 Model::bounding_box()
{
    CGAL::Simple_cartesian<float>::Iso_cuboid_3 b;
    if (modelMesh.impl && lbo_mesh_num_vertices((lbo_mesh*)modelMesh.impl) > 0) {
        lbo_mesh_bbox((lbo_mesh*)modelMesh.impl, b.lo, b.hi);
    } else {
        // an empty model must fail Lattice::insertModel's guard instead of dereferencing nothing
        for (int a = 0; a < 3; ++a) { b.lo[a] = -1.0f; b.hi[a] = -1.0f; }
    }
    return b;
}

Point_3 Model::reflectionVector(CoordsSystem::Point_3 a, CoordsSystem::Point_3 b)
{
    float p1[3] = {a.x(), a.y(), a.z()}, p2[3] = {b.x(), b.y(), b.z()}, out[3] = {0, 0, 0};
    if (modelMesh.impl) lbo_mesh_reflection_vector((lbo_mesh*)modelMesh.impl, p1, p2, out);
    return Point_3(out[0], out[1], out[2]);
}
