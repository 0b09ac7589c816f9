// Acquisition variant B (BDS/B1I/include/acquisition.m:4-176, GPS/GPS_L2C/include/acquisition.m:4-118):
// one carrier wipe-off + forward FFT per sub-bin shift, the Doppler bins are circular shifts of that
// spectrum (IQfreqDomShift = circshift(IQfreqDom, frqBinIndex-1)), every (shift, bin, block) row is
// correlated with abs(ifft(. .* codeFreqDom)) and only the row holding the largest peak is kept; the
// metric is that peak over the second peak outside +-1 chip.  The transforms run on the generic
// mixed-radix passes (acq_generic.cu); this file adds the shifted spectrum multiply and the row reductions.
#include <algorithm>
#include "acq.h"
#include "common.cuh"

namespace gc {

namespace {

// out[r][j] = X[src(r)][(j - shift(r)) mod L] * Cc[rep(r)][j]      (acquisition.m:71-74 / B1I :107-113)
__global__ void __launch_bounds__(256)
mulshift_kernel(const float2* __restrict__ X, const float2* __restrict__ Cc, const VarbRow* __restrict__ rows,
                float2* __restrict__ out, int L)
{
    const VarbRow r = rows[blockIdx.y];
    const float2* x = X + (size_t)r.src * L;
    const float2* c = Cc + (size_t)r.rep * L;
    float2* o = out + (size_t)blockIdx.y * L;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
        int i = j - r.shift;
        if (i < 0) i += L;
        o[j] = cmul(x[i], __ldg(c + j));
    }
}

// per row: max(abs(W)) and its first index            (currmax = max(acqRes), acquisition.m:77; [maxPeak, codePhase] = max(corrVec), :90)
__global__ void __launch_bounds__(1024)
rowpeak_kernel(const float2* __restrict__ W, int L, float* peak, int* idx)
{
    const float2* w = W + (size_t)blockIdx.x * L;
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const float2 v = w[j];
        const float a = sqrtf(fmaf(v.x, v.x, v.y * v.y));
        if (a > best) { best = a; bi = j; }                 // j ascending per thread: first maximum kept
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __shared__ float sb[32];
    __shared__ int si[32];
    if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
            if (sb[q] > best || (sb[q] == best && si[q] < bi)) { best = sb[q]; bi = si[q]; }
        peak[blockIdx.x] = best;
        idx[blockIdx.x] = bi;
    }
}

// per row: max(abs(W)) over up to two 0-based index ranges [lo, hi]        (secondPeakSize = max(corrVec(codePhaseRange)), :108)
__global__ void __launch_bounds__(1024)
segmax_kernel(const float2* __restrict__ W, int L, const int4* __restrict__ seg, float* out)
{
    const float2* w = W + (size_t)blockIdx.x * L;
    const int4 s = seg[blockIdx.x];
    float best = -1.f;
    for (int j = s.x + threadIdx.x; j <= s.y; j += blockDim.x) { const float2 v = w[j]; best = fmaxf(best, sqrtf(fmaf(v.x, v.x, v.y * v.y))); }
    for (int j = s.z + threadIdx.x; j <= s.w; j += blockDim.x) { const float2 v = w[j]; best = fmaxf(best, sqrtf(fmaf(v.x, v.x, v.y * v.y))); }
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_down_sync(0xffffffffu, best, o));
    __shared__ float sb[32];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q) best = fmaxf(best, sb[q]);
        out[blockIdx.x] = best;
    }
}

// same on a row of magnitudes (the fused path writes corrVec itself)
__global__ void __launch_bounds__(1024)
segmax_mag_kernel(const float* __restrict__ mag, int L, const int4* __restrict__ seg, float* out)
{
    const float* w = mag + (size_t)blockIdx.x * L;
    const int4 s = seg[blockIdx.x];
    float best = -1.f;
    for (int j = s.x + threadIdx.x; j <= s.y; j += blockDim.x) best = fmaxf(best, w[j]);
    for (int j = s.z + threadIdx.x; j <= s.w; j += blockDim.x) best = fmaxf(best, w[j]);
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_down_sync(0xffffffffu, best, o));
    __shared__ float sb[32];
    if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q) best = fmaxf(best, sb[q]);
        out[blockIdx.x] = best;
    }
}

// variant C (BDS/B1C/include/acquisition.m:205-216): results(bin, :) = abs(ifft(X_shift .* Data)), with the pilot replica
// (results*sqrt(11) + abs(ifft(X_shift .* Pilot))*sqrt(29))/sqrt(40); per bin the maximum and its first index.
// W rows: bin*nRep + r.
__global__ void __launch_bounds__(1024)
varc_combine_kernel(const float2* __restrict__ W, int L, int nRep, float* partMax, int* partIdx, size_t outBase)
{
    const float2* a = W + (size_t)blockIdx.x * nRep * L;
    const float2* b = a + L;
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const float2 v = a[j];
        float r = sqrtf(fmaf(v.x, v.x, v.y * v.y));
        if (nRep == 2) {
            const float2 u = b[j];
            r = (r * 3.3166247903554f + sqrtf(fmaf(u.x, u.x, u.y * u.y)) * 5.385164807134504f) / 6.324555320336759f;
        }
        if (r > best) { best = r; bi = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __shared__ float sb[32];
    __shared__ int si[32];
    if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
            if (sb[q] > best || (sb[q] == best && si[q] < bi)) { best = sb[q]; bi = si[q]; }
        partMax[outBase + blockIdx.x] = best;
        partIdx[outBase + blockIdx.x] = bi;
    }
}

// variant C fine search (:236-250): FineResult(j) = abs(sum(signal0DC .* DataPriTable .* exp(-1i*f_j*t)))
// [*11 + the same with PilotPriTable *29, /40]; grid (nFine, nAcq), one 10 ms period per block.
__global__ void __launch_bounds__(1024)
varc_fine_kernel(Rec rec, long long winStart, int N, int nRep, const int8_t* tabs /*[slot][N]*/,
                 const int* tabSlot, const int* codePhase, const uint64_t* dphi, int nFine, double* fineResult)
{
    const int j = blockIdx.x, a = blockIdx.y;
    const long long x0 = winStart + (codePhase[a] - 1);
    const int8_t* d = tabs + (size_t)tabSlot[a] * N;
    const int8_t* pl = d + N;
    const uint64_t dp = dphi[a * nFine + j];
    double dr = 0, di = 0, pr = 0, pi = 0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const short2 v = rec.load(x0 + n);
        float sn, cs;
        fix_sincos(dp * (uint64_t)n, &sn, &cs);
        const float re = fmaf(cs, (float)v.x, sn * (float)v.y), im = fmaf(cs, (float)v.y, -sn * (float)v.x);
        const float cd = (float)d[n];
        dr += (double)(cd * re); di += (double)(cd * im);
        if (nRep == 2) { const float cp = (float)pl[n]; pr += (double)(cp * re); pi += (double)(cp * im); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        dr += __shfl_down_sync(0xffffffffu, dr, o); di += __shfl_down_sync(0xffffffffu, di, o);
        pr += __shfl_down_sync(0xffffffffu, pr, o); pi += __shfl_down_sync(0xffffffffu, pi, o);
    }
    __shared__ double sh[32][4];
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = dr; sh[threadIdx.x >> 5][1] = di; sh[threadIdx.x >> 5][2] = pr; sh[threadIdx.x >> 5][3] = pi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        dr = di = pr = pi = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { dr += sh[q][0]; di += sh[q][1]; pr += sh[q][2]; pi += sh[q][3]; }
        double r = hypot(dr, di);
        if (nRep == 2) r = (r * 11 + hypot(pr, pi) * 29) / 40;                           // :246-249
        fineResult[a * nFine + j] = r;
    }
}

// GPS L2C CL code phase (GPS/GPS_L2C/include/acquisition.m:100-137): powerArray(ind) = abs(sum(signal0DC .* CLCodeSample .* sigCarr))
// for the 75 CM-period segments of the CL sequence; signal0DC = x - mean(x) over one CM period starting at codePhase.
// sum((x - mu) .* c .* e) = sum(x .* c .* e) - mu * sum(c .* e); one block per segment.
__global__ void __launch_bounds__(1024)
l2c_clphase_kernel(Rec rec, long long start, int N, const int8_t* cl, int segLen, const int* codeIdx, uint64_t dphi, double* power)
{
    const int ind = blockIdx.x;
    const int8_t* c = cl + (size_t)ind * segLen;
    double s[6] = {0, 0, 0, 0, 0, 0};            // sum x.c.e (re, im), sum c.e (re, im), sum x (re, im)
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const short2 v = rec.load(start + n);
        float sn, cs;
        fix_sincos(dphi * (uint64_t)n, &sn, &cs);
        const float cv = (float)c[codeIdx[n] - 1];
        const float re = fmaf(cs, (float)v.x, sn * (float)v.y), im = fmaf(cs, (float)v.y, -sn * (float)v.x);
        s[0] += (double)(cv * re); s[1] += (double)(cv * im);
        s[2] += (double)(cv * cs); s[3] -= (double)(cv * sn);
        s[4] += (double)v.x; s[5] += (double)v.y;
    }
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 6; ++q) s[q] += __shfl_down_sync(0xffffffffu, s[q], o);
    __shared__ double sh[32][6];
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int q = 0; q < 6; ++q) sh[threadIdx.x >> 5][q] = s[q];
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 6; ++q) s[q] = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
#pragma unroll
            for (int q = 0; q < 6; ++q) s[q] += sh[w][q];
        const double mr = s[4] / N, mi = s[5] / N;           // mean(signal0DC), :103
        power[ind] = hypot(s[0] - (mr * s[2] - mi * s[3]), s[1] - (mr * s[3] + mi * s[2]));
    }
}

// replica tables of variant B into complex rows zero padded to L (code_kernel of acq_generic.cu, any table length)
__global__ void pad_kernel(const int8_t* tab, int n, float2* out, int L)
{
    const int r = blockIdx.y;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x)
        out[(size_t)r * L + j] = make_float2(j < n ? (float)tab[(size_t)r * n + j] : 0.f, 0.f);
}

}  // namespace

cudaError_t launch_varb_mulshift(const float2* X, const float2* Cc, const VarbRow* rows, int nRows, float2* out, int L, cudaStream_t st)
{
    dim3 grid(std::min((L + 255) / 256, 148 * 2), nRows);
    mulshift_kernel<<<grid, 256, 0, st>>>(X, Cc, rows, out, L);
    return cudaGetLastError();
}

cudaError_t launch_varb_rowpeak(const float2* W, int nRows, int L, float* peak, int* idx, cudaStream_t st)
{
    rowpeak_kernel<<<nRows, 1024, 0, st>>>(W, L, peak, idx);
    return cudaGetLastError();
}

cudaError_t launch_varb_segmax(const float2* W, int nRows, int L, const int4* seg, float* out, cudaStream_t st)
{
    segmax_kernel<<<nRows, 1024, 0, st>>>(W, L, seg, out);
    return cudaGetLastError();
}

cudaError_t launch_varb_segmax_mag(const float* mag, int nRows, int L, const int4* seg, float* out, cudaStream_t st)
{
    segmax_mag_kernel<<<nRows, 1024, 0, st>>>(mag, L, seg, out);
    return cudaGetLastError();
}

cudaError_t launch_varc_combine(const float2* W, int nBins, int L, int nRep, float* partMax, int* partIdx, size_t outBase, cudaStream_t st)
{
    varc_combine_kernel<<<nBins, 1024, 0, st>>>(W, L, nRep, partMax, partIdx, outBase);
    return cudaGetLastError();
}

cudaError_t launch_varc_fine(Rec rec, long long winStart, int N, int nRep, const int8_t* tabs, const int* tabSlot,
                             const int* codePhase, const uint64_t* dphi, int nFine, int nAcq, double* fineResult, cudaStream_t st)
{
    dim3 grid(nFine, nAcq);
    varc_fine_kernel<<<grid, 1024, 0, st>>>(rec, winStart, N, nRep, tabs, tabSlot, codePhase, dphi, nFine, fineResult);
    return cudaGetLastError();
}

cudaError_t launch_l2c_clphase(Rec rec, long long start, int N, const int8_t* cl, int segLen, const int* codeIdx,
                               uint64_t dphi, double* power, cudaStream_t st)
{
    l2c_clphase_kernel<<<75, 1024, 0, st>>>(rec, start, N, cl, segLen, codeIdx, dphi, power);
    return cudaGetLastError();
}

cudaError_t launch_varb_pad(const int8_t* tab, int n, int nRows, float2* out, int L, cudaStream_t st)
{
    dim3 grid(std::min((L + 255) / 256, 148 * 2), nRows);
    pad_kernel<<<grid, 256, 0, st>>>(tab, n, out, L);
    return cudaGetLastError();
}

}  // namespace gc
