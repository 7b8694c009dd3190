// d3q.h -- the velocity set, shared by host and device code.
//
// Index order is API: it is the reference's DIRECTIONS / INVERSE_DIRECTIONS / SPLIT
// (simulation/source/SimCalcFuncs.cu:12-14, mirrored in simulation/source/Constants.cpp:6-7).
// Directions are sorted by shell (rest, 6 axis, 12 edge, 8 corner), so D3Q19 is the prefix
// 0..18 and is closed under inversion.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define LBM_HD __host__ __device__ __forceinline__
#else
#define LBM_HD inline
#endif

namespace lbm {

constexpr int QMAX = 27;

// component a of direction i
LBM_HD constexpr int dir_c(int i, int a)
{
    constexpr int8_t C[27][3] = {
        { 0, 0, 0},
        {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1}, { 0, 0, 1}, { 0, 1, 0}, { 1, 0, 0},
        {-1,-1, 0}, {-1, 0,-1}, {-1, 0, 1}, {-1, 1, 0}, { 0,-1,-1}, { 0,-1, 1},
        { 0, 1,-1}, { 0, 1, 1}, { 1,-1, 0}, { 1, 0,-1}, { 1, 0, 1}, { 1, 1, 0},
        {-1,-1,-1}, {-1,-1, 1}, {-1, 1,-1}, {-1, 1, 1}, { 1,-1,-1}, { 1,-1, 1}, { 1, 1,-1}, { 1, 1, 1}};
    return C[i][a];
}

// opposite direction: 0->0; 1..6 -> 7-i; 7..18 -> 25-i; 19..26 -> 45-i
LBM_HD constexpr int dir_inv(int i)
{
    return i == 0 ? 0 : (i < 7 ? 7 - i : (i < 19 ? 25 - i : 45 - i));
}

// number of non-zero components: 0 rest, 1 axis, 2 edge, 3 corner
LBM_HD constexpr int dir_shell(int i)
{
    return i == 0 ? 0 : (i < 7 ? 1 : (i < 19 ? 2 : 3));
}

// SPLIT[i] = c_i / |c_i|^2 scale of the reference's velocity moment
LBM_HD constexpr float dir_split(int i)
{
    return i == 0 ? 0.0f : (i < 7 ? 1.0f : (i < 19 ? 0.5f : (float)(1.0 / 3.0)));
}

// lattice weights: D3Q27 8/27, 2/27, 1/54, 1/216 (SimCalcFuncs.cu:125,134,143,152);
// D3Q19 1/3, 1/18, 1/36 (build extension)
template <int Q>
LBM_HD constexpr double dir_weight(int i)
{
    return Q == 27 ? (i == 0 ? 8.0 / 27.0 : (i < 7 ? 2.0 / 27.0 : (i < 19 ? 1.0 / 54.0 : 1.0 / 216.0)))
                   : (i == 0 ? 1.0 / 3.0 : (i < 7 ? 1.0 / 18.0 : 1.0 / 36.0));
}

// index of the (dx,dy,dz) neighbour in a 3x3x3 bit field
LBM_HD constexpr int nb_bit(int dx, int dy, int dz)
{
    return ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1);
}

}  // namespace lbm
