// Host-side PRN replica generation.  Integer LFSR work; the reference multiplies +-1 doubles,
// here the two 10-stage registers are bit registers and the chips come out as int8 +-1.
#include "codes.h"

#include <cmath>

namespace gc {

namespace {

// G2 delay in chips for PRN 1..32 (IS-GPS-200; the table generateCAcode.m:42-50 embeds)
const int kG2Delay[32] = {5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258,
                          469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862};

// maximal-length sequence of a 10-stage register started all-ones; output = stage 10;
// feedback = xor of the stages in `taps` (bit i = stage i+1)
void mseq(unsigned taps, uint8_t* out)
{
    unsigned reg = 0x3ff;
    for (int i = 0; i < 1023; ++i) {
        out[i] = (reg >> 9) & 1u;
        const unsigned fb = __builtin_parity(reg & taps);
        reg = ((reg << 1) | fb) & 0x3ff;
    }
}

}  // namespace

void ca_code(int prn, int8_t* out)
{
    uint8_t g1[1023], g2[1023];
    mseq((1u << 2) | (1u << 9), g1);                                                   // 1 + x^3 + x^10
    mseq((1u << 1) | (1u << 2) | (1u << 5) | (1u << 7) | (1u << 8) | (1u << 9), g2);   // 1+x^2+x^3+x^6+x^8+x^9+x^10
    const int d = kG2Delay[prn - 1];
    for (int i = 0; i < 1023; ++i) {
        const int j = (i - d + 1023) % 1023;                                           // G2 delayed by d chips
        out[i] = (g1[i] ^ g2[j]) ? 1 : -1;
    }
}

void make_ca_table(int prn, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out)
{
    int8_t chips[1023];
    ca_code(prn, chips);
    const double ts = 1 / fs, tc = 1 / codeFreqBasis;
    for (int n = 1; n <= N; ++n) {
        int idx = (int)std::ceil((ts * (double)n) / tc);                               // makeCaTable.m:59
        if (n == N) idx = codeLength;                                                  // :62
        out[n - 1] = chips[idx - 1];
    }
}

void glo_code(int8_t* out)
{
    // the reference multiplies +-1 registers loaded with -1; with bits (1 <-> -1) the product is an xor
    unsigned reg = 0x1ff;                                   // nine stages, all ones
    for (int i = 0; i < 511; ++i) {
        const unsigned o = (reg >> 6) & 1u;                 // stage 7
        out[i] = o ? -1 : 1;
        const unsigned fb = ((reg >> 4) ^ (reg >> 8)) & 1u; // stages 5 and 9
        reg = ((reg << 1) | fb) & 0x1ff;
    }
}

// MATLAB a:d:b with a = 0 (Cleve Moler's colonop): n+1 elements built from both ends
void glo_sample_index(double codeRate, double fs, int codeLength, long long numSamples, int16_t* idx)
{
    const double d = codeRate / fs;                         // stepSize
    const double b = ((double)numSamples * d) - d;
    const double tol = 2.0 * 2.220446049250313e-16 * std::fmax(0.0, std::fabs(b));
    long long n;
    if (d == 1.0) n = (long long)std::floor(b);
    else if (d == std::floor(d)) n = (long long)std::trunc(b / d);
    else {
        const double q = b / d;
        n = (long long)(q >= 0 ? std::floor(q + 0.5) : -std::floor(-q + 0.5));
        if ((0.0 + (double)n * d - b) > tol) n -= 1;
    }
    double c = 0.0 + (double)n * d;
    if ((c - b) > -tol) c = b;
    for (long long k = 0; k <= n && k < numSamples; ++k) {
        double v;
        if (2 * k < n) v = 0.0 + (double)k * d;
        else if (2 * k > n) v = c - (double)(n - k) * d;
        else v = (0.0 + c) / 2;
        idx[k] = (int16_t)std::fmod(std::floor(v), (double)codeLength);
    }
    for (long long k = n + 1; k < numSamples; ++k) idx[k] = 0;   // (never happens: n == numSamples-1)
}

void gps_fine_index(double fs, double codeFreqBasis, int codeLength, long long numSamples, int16_t* idx)
{
    const double ts = 1 / fs, tc = 1 / codeFreqBasis;
    for (long long k = 0; k < numSamples; ++k)
        idx[k] = (int16_t)((long long)std::floor((ts * (double)k) / tc) % codeLength);
}

}  // namespace gc
