// The resident IF record and its sample format (internal to libgnsscorr).
// settings.fileType (1 = real samples, 2 = I,Q interleaved) and settings.dataType ('schar' / 'int16'):
// GPS/GPS_L1CA/initSettings.m:63-68, the dataAdaptCoeff / int16 branches of postProcessing.m:66-96 and tracking.m:141-153, 229-240.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gc {

struct Rec {
    const int8_t* p;          // raw file image, 16-byte aligned
    int fmt;                  // bit 0: int16 samples (else int8), bit 1: real samples (else I,Q pairs)
    __host__ __device__ __forceinline__ int bytes_per_sample() const { return ((fmt & 1) ? 2 : 1) * ((fmt & 2) ? 1 : 2); }
#ifdef __CUDACC__
    // sample n as (I, Q); real records have Q = 0 (rawSignal stays real, tracking.m:229-240)
    __device__ __forceinline__ short2 load(long long n) const
    {
        switch (fmt) {
            case 0: { const char2 v = reinterpret_cast<const char2*>(p)[n]; return make_short2(v.x, v.y); }
            case 1: return reinterpret_cast<const short2*>(p)[n];
            case 2: return make_short2(p[n], 0);
            default: return make_short2(reinterpret_cast<const short*>(p)[n], 0);
        }
    }
#endif
};

}  // namespace gc
