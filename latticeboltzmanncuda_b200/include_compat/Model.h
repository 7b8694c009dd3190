// Model.h -- the reference's mesh wrapper (simulation/include/Model.h:23-47) without CGAL: points are the
// reference's own Point_3, the bounding box is a small struct with the accessors Lattice::insertModel uses.
#pragma once

#include <string>

#include "Geometry.h"

struct lbm_mesh_s;

struct BoundingBox
{
    float lo[3], hi[3];
    float xmin() const { return lo[0]; }
    float ymin() const { return lo[1]; }
    float zmin() const { return lo[2]; }
    float xmax() const { return hi[0]; }
    float ymax() const { return hi[1]; }
    float zmax() const { return hi[2]; }
};

class Model
{
public:
    Model() = default;
    ~Model();
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    void importModel(std::string filename);

    bool isModelClosed();

    bool isPointInsideModel(Point_3 point);

    int number_of_vertices();      // getMesh().number_of_vertices() in the reference's tests

    bool checkError();

    BoundingBox bounding_box();

    Point_3 reflectionVector(Point_3 point_1, Point_3 point_2);

    lbm_mesh_s* handle() { return m_mesh; }

private:
    lbm_mesh_s* m_mesh = nullptr;
};
