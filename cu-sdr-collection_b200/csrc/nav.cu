// Bit and frame synchronisation front end of postNavigation for GPS L1 C/A:
// GPS/GPS_L1CA/include/NAVdecoding.m:69-170 (preamble cross-correlation over the prompt outputs, subframe-spacing test,
// parity of the TLM and HOW words on 20 ms bit sums, the 1500 navigation bits) with Common/navPartyChk.m.
// Integer / sign work on one row of trackResults.I_P per channel; everything after it (ephemeris decoding) is scalar
// MATLAB and stays there.
#include "nav.h"
#include "../../include/gnsscorr.h"

namespace gc {

namespace {

__constant__ int8_t kPreamble[8] = {1, -1, -1, -1, 1, -1, 1, 1};   // NAVdecoding.m:69

// Common/navPartyChk.m: ndat = [D29* D30* d1..d24 D25..D30] as +-1; the six parity products of IS-GPS-200 table 20-XIV
// as index sets; status != 0 <=> all six parities hold
__device__ bool nav_parity_ok(const int* ndat)
{
    // data-bit sets per parity bit (1-based d indices)
    const int8_t set[6][15] = {
        {1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23, 0},
        {2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24, 0},
        {1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22, 0},
        {2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23, 0},
        {1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24},
        {3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24, 0, 0}};
    const int8_t star[6] = {0, 1, 0, 1, 1, 0};                 // which of D29*, D30* enters (index into ndat)
    const int flip = (ndat[1] != 1) ? -1 : 1;                    // navPartyChk.m: data bits inverted when D30* is set
    int ok = 0;
    for (int p = 0; p < 6; ++p) {
        int v = ndat[star[p]];
        for (int q = 0; q < 15; ++q)
            if (set[p][q]) v *= flip * ndat[1 + set[p][q]];
        ok += (v == ndat[26 + p]);
    }
    return ok == 6;
}

// lag l (0-based) of xcorr(bits, preamble_ms) for non-negative lags, bits = sign of I_P(1 + offset : end) with <= 0 -> -1
// (NAVdecoding.m:79-86); cand[index - 1] = |corr| > 153 and 40 < index < msToProcess - 1199  (:95-101)
__global__ void nav_xcorr_kernel(const double* ip, int n, int offset, int msToProcess, uint8_t* cand)
{
    const int ch = blockIdx.y;
    const double* x = ip + (size_t)ch * n;
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int len = n - offset;
    if (l >= len) return;
    int acc = 0;
    for (int j = 0; j < 160 && l + j < len; ++j) acc += (x[offset + l + j] > 0 ? 1 : -1) * kPreamble[j / 20];
    const int index = l + 1 + offset;
    cand[(size_t)ch * n + index - 1] = (abs(acc) > 153 && index > 40 && index < msToProcess - (20 * 60 - 1)) ? 1 : 0;
}

// every candidate with another one 6000 ms later: 62 bits from 20 ms sums of I_P(index-40 : index+1199), parity of the TLM
// and HOW words (:104-141); the smallest passing index is the reference's first hit
__global__ void nav_verify_kernel(const double* ip, int n, const uint8_t* cand, int* first)
{
    const int ch = blockIdx.y;
    const int index = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (index > n) return;
    const uint8_t* c = cand + (size_t)ch * n;
    if (!c[index - 1] || index + 6000 > n || !c[index + 6000 - 1]) return;
    if (index - 40 < 1 || index + 20 * 60 - 1 > n) return;
    const double* x = ip + (size_t)ch * n + (index - 40 - 1);
    int bits[62];
    for (int b = 0; b < 62; ++b) {
        double s = 0;
        for (int t = 0; t < 20; ++t) s += x[20 * b + t];
        bits[b] = s > 0 ? 1 : -1;
    }
    if (nav_parity_ok(bits) && nav_parity_ok(bits + 30)) atomicMin(first + ch, index);
}

// navBits = sum over 20 ms of I_P(subFrameStart-20 : subFrameStart+1500*20-1) > 0   (:152-166): 1501 bits, the first one is
// the last bit of the previous subframe
__global__ void nav_bits_kernel(const double* ip, int n, const int* first, uint8_t* bits, int* valid)
{
    const int ch = blockIdx.y;
    const int sfs = first[ch];
    const bool ok = sfs != 0x7fffffff && sfs - 20 >= 1 && sfs + 1500 * 20 - 1 <= n;
    if (blockIdx.x == 0 && threadIdx.x == 0) valid[ch] = ok ? 1 : 0;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= GC_NAV_BITS) return;
    uint8_t v = 0;
    if (ok) {
        const double* x = ip + (size_t)ch * n + (sfs - 20 - 1) + 20 * b;
        double s = 0;
        for (int t = 0; t < 20; ++t) s += x[t];
        v = s > 0 ? 1 : 0;
    }
    bits[(size_t)ch * GC_NAV_BITS + b] = v;
}

__global__ void nav_init_kernel(int* first, int nCh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nCh) first[i] = 0x7fffffff;
}

}  // namespace

cudaError_t launch_nav_sync(const double* ip, int nCh, int n, int offset, int msToProcess, uint8_t* cand, int* first,
                            uint8_t* bits, int* valid, cudaStream_t st)
{
    nav_init_kernel<<<(nCh + 127) / 128, 128, 0, st>>>(first, nCh);
    nav_xcorr_kernel<<<dim3((n + 255) / 256, nCh), 256, 0, st>>>(ip, n, offset, msToProcess, cand);
    nav_verify_kernel<<<dim3((n + 127) / 128, nCh), 128, 0, st>>>(ip, n, cand, first);
    nav_bits_kernel<<<dim3((GC_NAV_BITS + 255) / 256, nCh), 256, 0, st>>>(ip, n, first, bits, valid);
    return cudaGetLastError();
}

}  // namespace gc
