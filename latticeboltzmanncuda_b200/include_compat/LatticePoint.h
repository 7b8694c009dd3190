// LatticePoint.h -- the two POD records that cross the host boundary of the reference's API
// (simulation/include/LatticePoint.h:3-17).  Field names, order and sizes (380 / 144 bytes) are the interface;
// liblbm_b200 reads and writes host arrays of exactly these records (lbm_upload_aos / lbm_download_aos).
#pragma once

struct ReflectionValues
{
    int reflection_directions[14];   // 81 slots x 5 bits, six per int
    float reflection_weight[81];
};

struct LatticePoint
{
    int x, y, z;
    float particle_distribution[27];
    bool isReflected;
    bool isInternal;
    ReflectionValues* reflections = nullptr;     // host table of a reflecting cell (set by retrieve_data)
    ReflectionValues* d_reflections = nullptr;   // unused here: tables live inside the library
};

static_assert(sizeof(ReflectionValues) == 380, "ReflectionValues must stay 380 bytes");
static_assert(sizeof(LatticePoint) == 144, "LatticePoint must stay 144 bytes");
