// Constants.h -- host copies of the velocity set (simulation/include/Constants.h:4-9)
#ifndef LBM_COMPAT_CONSTANTS_H
#define LBM_COMPAT_CONSTANTS_H

namespace CONSTANTS
{
    extern const float GEOMETRIC_TOLERANCE;
    extern const int DIRECTIONS[27][3];
    extern const int INVERSE_DIRECTIONS[27];
}

#endif
