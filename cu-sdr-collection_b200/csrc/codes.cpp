// Host-side PRN replica generation.  Integer LFSR work; the reference multiplies +-1 doubles,
// here the two 10-stage registers are bit registers and the chips come out as int8 +-1.
#include "codes.h"

#include <cmath>
#include <initializer_list>

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

void make_code_table(const int8_t* chips, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out)
{
    const double ts = 1 / fs, tc = 1 / codeFreqBasis;
    for (int n = 1; n <= N; ++n) {
        int idx = (int)std::ceil((ts * (double)n) / tc);                               // makeCaTable.m:59, makeB3ITable.m:47
        if (n == N) idx = codeLength;                                                  // :62 / :50
        out[n - 1] = chips[idx - 1];
    }
}

void boc11(const int8_t* primary, int codeLength, int8_t* out)
{
    for (int i = 0; i < codeLength; ++i) { out[2 * i] = primary[i]; out[2 * i + 1] = (int8_t)-primary[i]; }
}

void make_boc_table(const int8_t* subchips, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out)
{
    const double ts = 1 / fs, tc = 1 / codeFreqBasis / 2;                              // makeE1BTable.m:42-43
    for (int n = 1; n <= N; ++n) {
        int idx = (int)std::ceil((ts * (double)n) / tc);                               // :51
        if (n == N) idx = codeLength * 2;                                              // :54
        if (n == 1) idx = 1;                                                           // :55
        out[n - 1] = subchips[idx - 1];
    }
}

void boc_fine_index(double fs, double codeFreqBasis, int codeLength, long long numSamples, int16_t* idx)
{
    const double ts = 1 / fs, tc = 1 / codeFreqBasis / 2;
    for (long long k = 0; k < numSamples; ++k)
        idx[k] = (int16_t)((long long)std::floor((ts * (double)k) / tc) % (2LL * codeLength));
}

void make_ca_table(int prn, double fs, double codeFreqBasis, int codeLength, int N, int8_t* out)
{
    int8_t chips[1023];
    ca_code(prn, chips);
    make_code_table(chips, fs, codeFreqBasis, codeLength, N, out);
}

// BeiDou B3I: two 13-stage registers.  G1 (taps 1,3,4,13) is cut short: when it reaches the state
// 1111111111100 it is reloaded with all ones (period 8190); G2 (taps 1,5,6,7,9,10,12,13) is advanced
// by a PRN-specific number of steps before the first chip (generateB3Icode.m:36-84).  The reference
// works with +-1 values loaded with -1, so -1 <-> bit 1 and a product is an xor.
void b3i_code(int prn, int8_t* out)
{
    static const int kInit[63] = {4, 11, 13, 22, 30, 36, 44, 48, 88, 104, 116, 129, 376, 418, 458, 682, 696, 707, 1078, 2069,
                                  2248, 2574, 2596, 2731, 4294, 4436, 4647, 4978, 4986, 1, 5209, 5539, 6061, 6488, 7130, 7165,
                                  7403, 5879, 1681, 5080, 5938, 3983, 6208, 7223, 2996, 1814, 6906, 6144, 4713, 7406, 7264, 1766,
                                  5347, 3515, 7951, 7054, 3884, 6067, 4230, 3803, 869, 3683, 1205};
    auto tap = [](unsigned reg, std::initializer_list<int> stages) {
        unsigned f = 0;
        for (int st : stages) f ^= (reg >> (st - 1)) & 1u;     // bit i-1 = stage i
        return f;
    };
    const unsigned all = 0x1fff;
    // reset_state = [-1 x11, 1, 1] for stages 1..13  ->  bits 1 for stages 1..11, 0 for stages 12, 13
    const unsigned resetState = 0x07ff;
    unsigned a = all;
    uint8_t ca[10230];
    for (int i = 0; i < 10230; ++i) {
        ca[i] = (a >> 12) & 1u;                                // stage 13
        if (a == resetState) a = all;
        else a = ((a << 1) | tap(a, {1, 3, 4, 13})) & all;
    }
    unsigned b = all;
    for (int i = 0; i < kInit[prn - 1]; ++i) b = ((b << 1) | tap(b, {1, 5, 6, 7, 9, 10, 12, 13})) & all;
    for (int i = 0; i < 10230; ++i) {
        const unsigned cb = (b >> 12) & 1u;
        b = ((b << 1) | tap(b, {1, 5, 6, 7, 9, 10, 12, 13})) & all;
        out[i] = (cb ^ ca[i]) ? -1 : 1;                        // CB .* CA with -1 <-> bit 1
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

void gps_fine_index(double fs, double codeFreqBasis, int codeLength, long long numSamples, int16_t* idx, int first)
{
    const double ts = 1 / fs, tc = 1 / codeFreqBasis;
    for (long long k = 0; k < numSamples; ++k)
        idx[k] = (int16_t)((long long)std::floor((ts * (double)(k + first)) / tc) % codeLength);
}

}  // namespace gc
