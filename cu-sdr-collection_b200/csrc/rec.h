// The resident IF record and its sample format (internal to libgnsscorr).
// settings.fileType (1 = real samples, 2 = I,Q interleaved) and settings.dataType ('schar' / 'int16'):
// GPS/GPS_L1CA/initSettings.m:63-68, the dataAdaptCoeff / int16 branches of postProcessing.m:66-96 and tracking.m:141-153, 229-240.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gc {

struct Rec {
    const int8_t* p;          // raw file image, 16-byte aligned
    int fmt;                  // bit 0: int16 samples (else int8), bit 1: real samples (else I,Q pairs); 4: 2-bit packed I,Q, two samples
                              // per byte (the input format of include/unpack_cplx.m, decoded on the fly: a quarter of the int8 traffic)
    // bytes of `n` samples
    __host__ __device__ __forceinline__ long long bytes_of(long long n) const
    {
        return fmt == 4 ? (n + 1) / 2 : n * (((fmt & 1) ? 2 : 1) * ((fmt & 2) ? 1 : 2));
    }
    __host__ __device__ __forceinline__ long long samples_in(long long bytes) const
    {
        return fmt == 4 ? bytes * 2 : bytes / (((fmt & 1) ? 2 : 1) * ((fmt & 2) ? 1 : 2));
    }
#ifdef __CUDACC__
    // sample n as (I, Q); real records have Q = 0 (rawSignal stays real, tracking.m:229-240)
    __device__ __forceinline__ short2 load(long long n) const
    {
        switch (fmt) {
            case 0: { const char2 v = reinterpret_cast<const char2*>(p)[n]; return make_short2(v.x, v.y); }
            case 1: return reinterpret_cast<const short2*>(p)[n];
            case 2: return make_short2(p[n], 0);
            case 3: return make_short2(reinterpret_cast<const short*>(p)[n], 0);
            default: {
                // unpack_cplx.m:17-20: byte -> I1 Q1 I2 Q2 through four 256-entry tables, i.e. per sample a sign bit and a
                // magnitude bit for I and for Q: first sample bits (0, 2) and (1, 3), second sample bits (4, 6) and (5, 7)
                const unsigned b = (unsigned)(unsigned char)p[n >> 1] >> ((n & 1) * 4);
                const int i = ((b & 4u) ? 3 : 1) * ((b & 1u) ? -1 : 1), q = ((b & 8u) ? 3 : 1) * ((b & 2u) ? -1 : 1);
                return make_short2((short)i, (short)q);
            }
        }
    }
#endif
};

}  // namespace gc
