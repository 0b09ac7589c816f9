// Host-side PRN replica generation for the device tables (internal to libgnsscorr).
#pragma once
#include <stdint.h>

namespace gc {

// +-1 GPS C/A chips of one PRN (1..32), 1023 values.  GPS/GPS_L1CA/include/generateCAcode.m:42-90.
void ca_code(int prn, int8_t* out);

// C/A code resampled to the sampling frequency, N = samplesPerCode values.
// GPS/GPS_L1CA/include/makeCaTable.m:43-67 (index = ceil(ts*n/tc), last index forced to 1023).
void make_ca_table(int prn, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out);

}  // namespace gc
