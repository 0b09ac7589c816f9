// Host-side PRN replica generation for the device tables (internal to libgnsscorr).
#pragma once
#include <stdint.h>

namespace gc {

// +-1 GPS C/A chips of one PRN (1..32), 1023 values.  GPS/GPS_L1CA/include/generateCAcode.m:42-90.
void ca_code(int prn, int8_t* out);

// C/A code resampled to the sampling frequency, N = samplesPerCode values.
// GPS/GPS_L1CA/include/makeCaTable.m:43-67 (index = ceil(ts*n/tc), last index forced to 1023).
void make_ca_table(int prn, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out);

// +-1 BeiDou B3I chips of one PRN (1..63), 10230 values.  BDS/B3I/include/generateB3Icode.m:33-86.
void b3i_code(int prn, int8_t* out);

// A code resampled to the sampling frequency like makeCaTable.m / makeB3ITable.m: index = ceil(ts*n/tc),
// last index forced to codeLength.
void make_code_table(const int8_t* chips, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out);

// Galileo E1 BOC(1,1): primary chips c -> sub-chips [c -c] (GAL/GAL_E1C/include/generateE1Bcode.m:58-64),
// 2*codeLength values.
void boc11(const int8_t* primary, int codeLength, int8_t* out);

// BOC(1,1) sub-chip code resampled like makeE1BTable.m:42-55 / makeE1CTable.m: tc = 1/codeFreqBasis/2,
// index = ceil(ts*n/tc), last index forced to 2*codeLength, FIRST index forced to 1.
void make_boc_table(const int8_t* subchips, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out);

// codeValueIndex of the E1 fine search: floor((ts*(0:num-1)) / (1/codeFreqBasis/2)) rem (codeLength*2)
// (GAL_E1C/include/acquisition.m:211-214), 0-based.
void boc_fine_index(double fs, double codeFreqBasis, int codeLength, long long numSamples, int16_t* idx);

// +-1 GLONASS ST-code chips (511, 9-stage register, taps 5 and 9, output of stage 7).
// GLO/GLO_GL1/include/generateCAcode.m:95-108.
void glo_code(int8_t* out);

// MATLAB colon vector 0:step:(num*step)-step floored and wrapped to the code length: the sample ->
// chip index map of GLO/GLO_GL1/include/generateCAcode.m:110-116 (0-based chip indices).
void glo_sample_index(double codeRate, double fs, int codeLength, long long numSamples, int16_t* idx);

// codeValueIndex of the GPS fine search: floor((ts*(0:num-1)) / (1/codeFreqBasis)) mod codeLength
// (GPS/GPS_L1CA/include/acquisition.m:215-218), 0-based.  first = 1: the sample index runs 1..num
// (floor(ts*(1:num)/tc), GPS/GPS_L5C/include/acquisition.m:196).
void gps_fine_index(double fs, double codeFreqBasis, int codeLength, long long numSamples, int16_t* idx, int first = 0);

}  // namespace gc
