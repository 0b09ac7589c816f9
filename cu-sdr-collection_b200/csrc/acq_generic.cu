// Acquisition, generic path: the same pipeline as acq_fused.cu for any FFT length 2*samplesPerCode
// whose prime factors are <= 64 (36000 at the reference's default 18 Msps, 24000 GLONASS, ...).
// Batched Stockham autosort passes through global memory, one pass per prime factor; correct for
// every configuration, not tuned — the length-specific fused plans are where the speed is.
// GPS/GPS_L1CA/include/acquisition.m:155-200.
#include <algorithm>
#include "acq.h"
#include "common.cuh"
#include "fft_codelets.cuh"

namespace gc {

namespace {

// out[(k*nonCoh+m)][n] = longSignal(m*N + n) * exp(-1i*f_k*phasePoints(n))   (acquisition.m:172-181)
__global__ void wipe_kernel(Rec rec, long long winStart, int N, int nonCoh, int swapIQ,
                            const uint64_t* dphi, float2* out, int L)
{
    const int km = blockIdx.y, k = km / nonCoh, m = km % nonCoh;
    const long long x0 = winStart + (long long)m * N;
    const uint64_t d = dphi[k];
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x) {
        const short2 s = rec.load(x0 + n);
        float sn, cs;
        fix_sincos(d * (uint64_t)n, &sn, &cs);
        const float I = swapIQ ? (float)s.y : (float)s.x, Q = swapIQ ? (float)s.x : (float)s.y;
        out[(size_t)km * L + n] = make_float2(fmaf(cs, I, sn * Q), fmaf(cs, Q, -sn * I));
    }
}

__global__ void code_kernel(const int8_t* codeTab, int N, float2* out, int L)
{
    const int prn = blockIdx.y;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x)
        out[(size_t)prn * L + n] = make_float2(n < N ? (float)codeTab[(size_t)prn * N + n] : 0.f, 0.f);
}

// One Stockham pass of radix RADIX (0 = run-time radix r): n = current sub-length, s = stride.
//   a_i = x[q + s*(p + i*m)],  y[q + s*(r*p + j)] = w_n^(p*j) * sum_i a_i w_r^(i*j)
template <int RADIX>
__global__ void __launch_bounds__(256)
stage_kernel(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ tw,
             int L, int n, int s, int rr, int inverse, long long batch)
{
    const int r = RADIX ? RADIX : rr;
    const int m = n / r;
    const long long perXform = (long long)m * s;       // butterflies per transform
    const long long total = perXform * batch;
    const int tstep = L / n, rstep = L / r;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / perXform;
        const int rem = (int)(t - b * perXform);
        const int p = rem / s, q = rem - p * s;
        const float2* x = src + (size_t)b * L;
        float2* y = dst + (size_t)b * L;
        float2 a[RADIX ? RADIX : 64];
#pragma unroll
        for (int i = 0; i < r; ++i) a[i] = x[q + s * (p + i * m)];
#pragma unroll
        for (int j = 0; j < r; ++j) {
            float2 acc = a[0];
#pragma unroll
            for (int i = 1; i < r; ++i) {
                const float2 w = __ldg(tw + (size_t)((i * j) % r) * rstep);
                acc = inverse ? make_float2(acc.x + fmaf(a[i].x, w.x, a[i].y * w.y), acc.y + fmaf(a[i].y, w.x, -a[i].x * w.y))
                              : make_float2(acc.x + fmaf(a[i].x, w.x, -a[i].y * w.y), acc.y + fmaf(a[i].x, w.y, a[i].y * w.x));
            }
            const float2 w = __ldg(tw + ((size_t)p * j * tstep) % L);
            y[q + s * (r * p + j)] = inverse ? cmul_conj(acc, w) : cmul(acc, w);
        }
    }
}

// Same pass with the radix-R butterfly done by a register codelet (fft_codelets.cuh: 8, 16, 25, 30..33, 40, 45, 50): four
// passes instead of nine for the 320000 / 360000-point transforms of GPS L2C / BDS B1C.
template <int R, bool INV>
__global__ void __launch_bounds__(128)
stage_codelet_kernel(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ tw,
                     int L, int n, int s, long long batch)
{
    const int m = n / R;
    const long long perXform = (long long)m * s;
    const long long total = perXform * batch;
    const int tstep = L / n;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / perXform;
        const int rem = (int)(t - b * perXform);
        const int p = rem / s, q = rem - p * s;
        const float2* x = src + (size_t)b * L;
        float2* y = dst + (size_t)b * L;
        float2 a[R];
#pragma unroll
        for (int i = 0; i < R; ++i) a[i] = x[q + s * (p + i * m)];
        const long long pt = (long long)p * tstep;
        codelet::dft<R, INV>(a, [&](int j, float re, float im) {
            const float2 w = __ldg(tw + (size_t)((pt * j) % L));
            const float2 v = make_float2(re, im);
            y[q + s * (R * p + j)] = INV ? cmul_conj(v, w) : cmul(v, w);
        });
    }
}

__global__ void mul_kernel(const float2* X, const float2* Cc, float2* out, int L, long long nKm)
{
    const long long total = nKm * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        out[i] = cmul(X[i], __ldg(Cc + (i % L)));       // IQfreqDom .* caCodeFreqDom (:186)
}

__global__ void conj_scale_kernel(float2* Cc, size_t n, float scale)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 v = Cc[i];
        Cc[i] = make_float2(v.x * scale, -v.y * scale);
    }
}

// results(k,:) = sum_m abs(W[k][m][:]) and the per-tile maximum / first arg-max (:188-198)
__global__ void __launch_bounds__(256)
absacc_kernel(const float2* W, const float2* W2, int L, int nonCoh, int parts, float* partMax, int* partIdx, size_t outBase)
{
    const int k = blockIdx.y, part = blockIdx.x;
    const int per = (L + parts - 1) / parts;
    const int lo = part * per, hi = min(L, lo + per);
    float best = -1.f; int bidx = 0x7fffffff;
    for (int n = lo + threadIdx.x; n < hi; n += 256) {
        float acc = 0.f;
        for (int m = 0; m < nonCoh; ++m) {
            const float2 v = W[((size_t)k * nonCoh + m) * L + n];
            float coh = sqrtf(fmaf(v.x, v.x, v.y * v.y));
            if (W2) {                                       // abs(ifft(convE1bIQ)) + abs(ifft(convE1cIQ)), GAL_E1C acquisition.m:192
                const float2 u = W2[((size_t)k * nonCoh + m) * L + n];
                coh += sqrtf(fmaf(u.x, u.x, u.y * u.y));
            }
            acc += coh;
        }
        if (acc > best) { best = acc; bidx = n; }      // n ascending per thread: first max kept
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    __shared__ float sb[8]; __shared__ int si[8];
    if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) if (sb[w] > best || (sb[w] == best && si[w] < bidx)) { best = sb[w]; bidx = si[w]; }
        partMax[outBase + (size_t)k * parts + part] = best;
        partIdx[outBase + (size_t)k * parts + part] = bidx;
    }
}

}  // namespace

cudaError_t launch_generic_stage(const GenericPlan& pl, int stage, int n, int s, bool inverse,
                                 const float2* src, float2* dst, long long batch, cudaStream_t st)
{
    const int r = pl.fac[stage];
    const long long total = (long long)(n / r) * s * batch;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 64);
    const int inv = inverse ? 1 : 0;
#define GC_STAGE(R) stage_kernel<R><<<grid, 256, 0, st>>>(src, dst, pl.tw, pl.L, n, s, r, inv, batch)
#define GC_CODELET(R)                                                                                             \
    case R: {                                                                                                     \
        const unsigned g = (unsigned)std::min<long long>((total + 127) / 128, 148LL * 64);                        \
        if (inverse) stage_codelet_kernel<R, true><<<g, 128, 0, st>>>(src, dst, pl.tw, pl.L, n, s, batch);         \
        else stage_codelet_kernel<R, false><<<g, 128, 0, st>>>(src, dst, pl.tw, pl.L, n, s, batch);                \
        break;                                                                                                    \
    }
    switch (r) {
        GC_CODELET(8) GC_CODELET(16) GC_CODELET(25) GC_CODELET(30) GC_CODELET(32) GC_CODELET(33) GC_CODELET(40) GC_CODELET(45) GC_CODELET(50)
        case 2: GC_STAGE(2); break;
        case 3: GC_STAGE(3); break;
        case 4: GC_STAGE(4); break;
        case 5: GC_STAGE(5); break;
        case 7: GC_STAGE(7); break;
        case 11: GC_STAGE(11); break;
        case 13: GC_STAGE(13); break;
        case 31: GC_STAGE(31); break;
        default: GC_STAGE(0); break;
    }
#undef GC_STAGE
#undef GC_CODELET
    return cudaGetLastError();
}

cudaError_t launch_generic_wipe(Rec rec, long long winStart, int N, int nonCoh, int nBins, int swapIQ,
                                const uint64_t* dphi, float2* out, int L, cudaStream_t st)
{
    dim3 grid((L + 255) / 256, nBins * nonCoh);
    wipe_kernel<<<grid, 256, 0, st>>>(rec, winStart, N, nonCoh, swapIQ, dphi, out, L);
    return cudaGetLastError();
}

cudaError_t launch_generic_code(const int8_t* codeTab, int N, int nPrn, float2* out, int L, cudaStream_t st)
{
    dim3 grid((L + 255) / 256, nPrn);
    code_kernel<<<grid, 256, 0, st>>>(codeTab, N, out, L);
    return cudaGetLastError();
}

cudaError_t launch_generic_mul(const float2* X, const float2* Cc, float2* out, int L, long long nKm, cudaStream_t st)
{
    mul_kernel<<<148 * 16, 256, 0, st>>>(X, Cc, out, L, nKm);
    return cudaGetLastError();
}

cudaError_t launch_generic_absacc(const float2* W, const float2* W2, int L, int nBins, int nonCoh, int parts,
                                  float* partMax, int* partIdx, size_t outBase, cudaStream_t st)
{
    dim3 grid(parts, nBins);
    absacc_kernel<<<grid, 256, 0, st>>>(W, W2, L, nonCoh, parts, partMax, partIdx, outBase);
    return cudaGetLastError();
}

cudaError_t launch_generic_conj_scale(float2* Cc, size_t n, float scale, cudaStream_t st)
{
    conj_scale_kernel<<<148 * 2, 256, 0, st>>>(Cc, n, scale);
    return cudaGetLastError();
}

}  // namespace gc
