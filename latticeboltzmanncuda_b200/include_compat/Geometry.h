// Geometry.h -- Point_3 and rotateVectorAroundAxis with the reference's signatures
// (simulation/include/Geometry.h:8-31), implemented by liblbm_b200 (lbm_rotate_vector_around_axis).
#ifndef LBM_COMPAT_GEOMETRY_H
#define LBM_COMPAT_GEOMETRY_H

#include <math.h>

struct Point_3
{
    float x, y, z;

    Point_3(float x_i, float y_i, float z_i) : x(x_i), y(y_i), z(z_i) {}

    bool operator==(const Point_3& a) const { return a.x == x && a.y == y && a.z == z; }

    void normaliseVector()
    {
        const float norm = sqrtf((x * x) + (y * y) + (z * z));
        x /= norm; y /= norm; z /= norm;
    }
};

Point_3 rotateVectorAroundAxis(Point_3 axis, Point_3 vector, float angle);
#endif
