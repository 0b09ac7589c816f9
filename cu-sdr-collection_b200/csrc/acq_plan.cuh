// Transform plans of the fused acquisition kernels: FFT length L = C x RA x RB (2*samplesPerCode),
// rows R = RA x RB by the Good-Thomas prime-factor mapping, columns C against R by the prime-factor
// mapping when gcd(C, R) = 1 and by one Cooley-Tukey twiddle otherwise (see acq_fused.cu).
#pragma once
#include <cuda_runtime.h>

namespace gc {

constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }

// C1 > 0: the column length C = C1 x C2 is too long for one register codelet and is done in two levels through shared
// memory (Cooley-Tukey inside the column pass: Galileo E1's 160000 = (20 x 10) x 800 and 144000 = (20 x 9) x 800, GPS L2C's
// 320000 = (20 x 20) x 800, BDS B1C's 360000 = (25 x 18) x 800, BDS B1I's 72000 = (10 x 9) x 800).
template <int C_, int RA_, int RB_, int C1_ = 0>
struct Plan {
    static constexpr int C = C_, RA = RA_, RB = RB_;
    static constexpr int C1 = C1_, C2 = C1_ ? C_ / C1_ : 0;
    static constexpr bool kBig = C1_ != 0;
    static_assert(C1_ == 0 || C_ % C1_ == 0, "C = C1 x C2");
    static constexpr int R = RA * RB, L = C * R;
    static constexpr bool kPfa = gcd_c(C, R) == 1;        // no twiddle between column and row pass
    static_assert(gcd_c(RA, RB) == 1 && RA == 32 && RB < 32, "row = 32 x RB with RB coprime to 32");
    // row position p = a*RB + b  ->  its part of the time / lag index
    __device__ static __forceinline__ int row_index(int p)
    {
        const int a = p / RB, b = p % RB;
        return kPfa ? ((L / RA) * a + (L / RB) * b) % L     // 3-D prime-factor map: contribution to n mod L
                    : (RB * a + RA * b) % R;                // 2-D map inside the row
    }
    // global time / lag index of (column index i1, row part)
    __device__ static __forceinline__ int index(int i1, int rowPart)
    {
        if (kPfa) { const int n = rowPart + R * i1; return n >= L ? n - L : n; }   // rowPart < L, R*i1 < L
        return rowPart + R * i1;
    }
};

using P32736 = Plan<33, 32, 31>;   // 16.368 Msps, 1 ms codes
using P36000 = Plan<45, 32, 25>;   // 18 Msps: the reference default of six signal folders
using P24000 = Plan<30, 32, 25>;   // 12 Msps: GLONASS default
using P32000 = Plan<40, 32, 25>;   // 16 Msps
using P40000 = Plan<50, 32, 25>;   // 20 Msps
using P160000 = Plan<200, 32, 25, 20>;   // Galileo E1 (4 ms codes) at 20 Msps: columns 200 = 20 x 10
using P144000 = Plan<180, 32, 25, 20>;   // Galileo E1 at 18 Msps (reference default): columns 180 = 20 x 9
using P320000 = Plan<400, 32, 25, 20>;   // GPS L2C at 8 Msps (40 ms block): columns 400 = 20 x 20
using P360000 = Plan<450, 32, 25, 25>;   // BDS B1C at 18 Msps (20 ms): columns 450 = 25 x 18
using P72000 = Plan<90, 32, 25, 10>;     // BDS B1I at 18 Msps (4 ms blocks): columns 90 = 10 x 9

#define GC_PLAN_DISPATCH(LEN, CALL)                             \
    switch (LEN) {                                              \
        case P32736::L: return Launch<P32736>::CALL;            \
        case P36000::L: return Launch<P36000>::CALL;            \
        case P24000::L: return Launch<P24000>::CALL;            \
        case P32000::L: return Launch<P32000>::CALL;            \
        case P40000::L: return Launch<P40000>::CALL;            \
        case P160000::L: return Launch<P160000>::CALL;          \
        case P144000::L: return Launch<P144000>::CALL;          \
        case P320000::L: return Launch<P320000>::CALL;          \
        case P360000::L: return Launch<P360000>::CALL;          \
        case P72000::L: return Launch<P72000>::CALL;            \
        default: return cudaErrorInvalidValue;                  \
    }

}  // namespace gc
