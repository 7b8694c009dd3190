// ref_host_shim.cpp -- extern "C" doors into the reference's own host sources
// (simulation/source/Geometry.cpp, Constants.cpp, Utils.cpp), which are compiled unmodified
// from /root/reference into oracle/_ref/libref_host.so by oracle/Makefile.
// TEST INFRASTRUCTURE ONLY: used to pin the CPU oracle against the real reference code.
#include "Geometry.h"
#include "Constants.h"
#include "Utils.h"

extern "C" {

void ref_rotate_vector_around_axis(const float* axis, const float* vec, float angle, float* out)
{
    Point_3 r = rotateVectorAroundAxis(Point_3(axis[0], axis[1], axis[2]), Point_3(vec[0], vec[1], vec[2]), angle);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

unsigned ref_compress_ints(int* array, int num, int bit_length)
{
    return Utils::compressInts(array, num, bit_length);
}

void ref_constants(int* directions81, int* inverse27, float* tolerance)
{
    for (int i = 0; i < 27; ++i) {
        for (int a = 0; a < 3; ++a) directions81[i * 3 + a] = CONSTANTS::DIRECTIONS[i][a];
        inverse27[i] = CONSTANTS::INVERSE_DIRECTIONS[i];
    }
    *tolerance = CONSTANTS::GEOMETRIC_TOLERANCE;
}

}
