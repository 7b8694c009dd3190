// Utils.h -- simulation/include/Utils.h
#pragma once

namespace Utils
{
    unsigned int compressInts(int* array, const int num, const int bitLength);
}
