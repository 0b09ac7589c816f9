// Acquisition, fused path for FFT length 32736 = 33 x 32 x 31 (2*samplesPerCode at 16.368 Msps).
//
// Replaces the PRN x Doppler x non-coherent-block loop of
// GPS/GPS_L1CA/include/acquisition.m:155-200.  The 2N-point transforms the reference does with
// MATLAB's fft/ifft are computed exactly at that length (no power-of-two padding, so the
// code-phase index space 1..2N is the reference's own) as a four-step transform
//
//      L = C x R,  C = 33 (= 3 x 11, prime-factor codelet),  R = 992 = 32 x 31
//
//   forward  (wipe-off + FFT, PRN independent, acquisition.m:169-183):
//      fwd_cols : one thread per column n2: 33 strided int8 I/Q samples, carrier wipe-off with a
//                 64-bit fixed-point phase, 33-point DFT in registers, twiddle w_L^(j1*n2)
//      fwd_rows : one warp per row j1: 992-point FFT as 32-point DFTs (31 lanes), twiddle,
//                 shared-memory transpose, 31-point DFTs (32 lanes)
//      spectrum layout X[j1][j2], frequency index j = j1 + 33*j2 (no transpose ever materialised)
//   inverse  (acquisition.m:186-190):
//      inv_rows : one warp per row: load X*conj(FFT(code))/L, inverse 992-point FFT, twiddle
//      inv_cols : one thread per column tau2: for each non-coherent block 33-point inverse DFT,
//                 |.|, accumulate in registers; after the last block the running maximum /
//                 first arg-max of the thread's 33 code phases tau = tau2 + 992*tau1.
//   `results(freqBin, :)` (acquisition.m:162,190) is therefore never written to memory.
#include "acq.h"
#include "common.cuh"
#include "fft_codelets.cuh"

namespace gc {

namespace {

constexpr int C = kFusedC;       // 33
constexpr int R = kFusedR;       // 992
constexpr int RA = 32, RB = 31;  // R = RA * RB
constexpr int L = C * R;         // 32736
constexpr int kPitch = RA + 1;   // smem row pitch (float2) for the 31 x 32 exchange
constexpr int kRowWarps = 8;     // warps (= rows in flight) per CTA in the row kernels

// ------------------------------------------------------------------ column pass (forward)
// grid (ceil(R/128), nRows), block 128.  MODE 0: IF samples with carrier wipe-off; MODE 1: code table.
template <int MODE>
__global__ void __launch_bounds__(128)
fwd_cols_kernel(FwdColsParams p)
{
    const int n2 = blockIdx.x * 128 + threadIdx.x;
    if (n2 >= R) return;
    const int row = blockIdx.y;               // MODE 0: km = k*nonCoh + m ; MODE 1: prn slot
    float2 x[C];
    if (MODE == 0) {
        const int k = row / p.nonCoh, m = row % p.nonCoh;
        const uint64_t dphi = p.dphi[k];
        const int8_t* src = p.rec + 2 * ((size_t)p.winStart + (size_t)m * p.N);   // window x((m-1)N+1 : (m+1)N)
#pragma unroll
        for (int n1 = 0; n1 < C; ++n1) {
            const int n = n1 * R + n2;
            const char2 s = *reinterpret_cast<const char2*>(src + 2 * (size_t)n);
            float sn, cs;
            fix_sincos(dphi * (uint64_t)n, &sn, &cs);          // exp(-1i*f*phasePoints(n)), :172
            const float I = (float)s.x, Q = (float)s.y;
            x[n1] = make_float2(fmaf(cs, I, sn * Q), fmaf(cs, Q, -sn * I));   // :180-181
        }
    } else {
        const int8_t* code = p.codeTab + (size_t)row * p.N;    // caCodesTable, zero padded to 2N (:160)
#pragma unroll
        for (int n1 = 0; n1 < C; ++n1) {
            const int n = n1 * R + n2;
            x[n1] = make_float2(n < p.N ? (float)code[n] : 0.f, 0.f);
        }
    }
    float2* dst = p.out + (size_t)row * L + n2;
    const float2* tw = p.twL + n2;
    codelet::dft33_fwd(x, [&](int j1, float re, float im) {
        const float2 w = __ldg(tw + (size_t)j1 * R);           // w_L^(j1*n2)
        dst[(size_t)j1 * R] = cmul(make_float2(re, im), w);
    });
}

// ------------------------------------------------------------------ row pass (both directions)
// One warp per 992-point row.  INV=false: plain forward FFT in place (spectrum rows).
// INV=true : load X*Cc, inverse FFT, multiply by conj(w_L^(j1*tau2)), store to the work buffer.
template <bool INV>
__global__ void __launch_bounds__(kRowWarps * 32)
rows_kernel(RowsParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);                 // [RA][RB] w_R^(b1*a2)
    float2* s_x = s_tw + RA * RB + (threadIdx.x >> 5) * (RB * kPitch);  // per-warp [RB][kPitch]
    for (int i = threadIdx.x; i < RA * RB; i += blockDim.x) s_tw[i] = p.twR[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // row decode
    const float2 *src, *mul = nullptr, *otw = nullptr;
    float2* dst;
    if (!INV) {
        const long long row = (long long)blockIdx.x * kRowWarps + warp;
        if (row >= p.nRows) return;
        src = p.X + row * R;
        dst = p.X + row * R;
    } else {
        // blockIdx.x = j1, blockIdx.y = k, blockIdx.z = (prn group, m group); warp -> (prn, m)
        const int j1 = blockIdx.x, k = blockIdx.y;
        const int mGroups = (p.nonCoh + p.mPerCta - 1) / p.mPerCta;
        const int pg = blockIdx.z / mGroups, mg = blockIdx.z % mGroups;
        const int pi = pg * p.prnPerCta + warp / p.mPerCta;     // prn slot within this launch's chunk
        const int m = mg * p.mPerCta + warp % p.mPerCta;
        if (pi >= p.nPrnChunk || m >= p.nonCoh) return;
        src = p.X + ((size_t)(k * p.nonCoh + m) * C + j1) * R;
        mul = p.Cc + ((size_t)p.prnList[p.prnSlot0 + pi] * C + j1) * R;
        otw = p.twL + (size_t)j1 * R;
        dst = p.W + (((size_t)(pi * p.nBins + k) * p.nonCoh + m) * C + j1) * R;
    }

    // stage 1: lane = a2 (< 31), elements a1*31 + a2, 32-point DFT over a1
    if (lane < RB) {
        float2 v[RA];
#pragma unroll
        for (int a1 = 0; a1 < RA; ++a1) {
            float2 t = src[a1 * RB + lane];
            if (INV) t = cmul(t, __ldg(mul + a1 * RB + lane));  // IQfreqDom .* caCodeFreqDom (:186)
            v[a1] = t;
        }
        auto put = [&](int b1, float re, float im) {
            const float2 w = s_tw[b1 * RB + lane];
            const float2 t = make_float2(re, im);
            s_x[lane * kPitch + b1] = INV ? cmul_conj(t, w) : cmul(t, w);
        };
        if (INV) codelet::dft32_inv(v, put); else codelet::dft32_fwd(v, put);
    }
    __syncwarp();
    // stage 2: lane = b1 (all 32), 31-point DFT over a2, output index b1 + 32*c2
    {
        float2 u[RB];
#pragma unroll
        for (int a2 = 0; a2 < RB; ++a2) u[a2] = s_x[a2 * kPitch + lane];
        auto put = [&](int c2, float re, float im) {
            const int o = lane + RA * c2;
            float2 t = make_float2(re, im);
            if (INV) t = cmul_conj(t, __ldg(otw + o));          // conj(w_L^(j1*tau2))
            dst[o] = t;
        };
        if (INV) codelet::dft31_inv(u, put); else codelet::dft31_fwd(u, put);
    }
}

// conj + 1/L scale of the replica spectra: caCodeFreqDom = conj(fft(...)) (:164) with the
// 1/(2N) of MATLAB's ifft (:188) folded in.
__global__ void finish_replica_kernel(float2* Cc, size_t n, float scale)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 v = Cc[i];
        Cc[i] = make_float2(v.x * scale, -v.y * scale);
    }
}

// ------------------------------------------------------------------ column pass (inverse) + |.| + sum + max
// grid (ceil(R/128), nBins, nPrnChunk), block 128: thread = code-phase column tau2 of one (PRN, bin).
__global__ void __launch_bounds__(128)
inv_cols_kernel(InvColsParams p)
{
    const int tau2 = blockIdx.x * 128 + threadIdx.x;
    const int k = blockIdx.y, pi = blockIdx.z;
    float acc[C];
#pragma unroll
    for (int i = 0; i < C; ++i) acc[i] = 0.f;
    if (tau2 < R) {
        const float2* base = p.W + ((size_t)(pi * p.nBins + k) * p.nonCoh) * L + tau2;
        for (int m = 0; m < p.nonCoh; ++m) {                    // acquisition.m:175
            float2 x[C];
#pragma unroll
            for (int j1 = 0; j1 < C; ++j1) x[j1] = __ldcs(base + (size_t)m * L + (size_t)j1 * R);
            codelet::dft33_inv(x, [&](int t1, float re, float im) {
                acc[t1] += sqrtf(fmaf(re, re, im * im));        // abs(ifft(.)) summed over blocks (:188-190)
            });
        }
    }
    // running maximum with MATLAB first-index tie breaking (smaller code phase wins)
    float best = -1.f;
    int bidx = 0x7fffffff;
    if (tau2 < R) {
#pragma unroll
        for (int t1 = 0; t1 < C; ++t1) {
            const int idx = tau2 + R * t1;
            if (acc[t1] > best || (acc[t1] == best && idx < bidx)) { best = acc[t1]; bidx = idx; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    __shared__ float s_b[4];
    __shared__ int s_i[4];
    if ((threadIdx.x & 31) == 0) { s_b[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w)
            if (s_b[w] > best || (s_b[w] == best && s_i[w] < bidx)) { best = s_b[w]; bidx = s_i[w]; }
        const size_t o = ((size_t)(p.prnSlot0 + pi) * p.nBins + k) * gridDim.x + blockIdx.x;
        p.partMax[o] = best;
        p.partIdx[o] = bidx;
    }
}

}  // namespace

int fused_row_smem_bytes() { return (int)(sizeof(float2) * (RA * RB + kRowWarps * RB * kPitch)); }
int fused_col_parts() { return (R + 127) / 128; }

cudaError_t launch_fwd_cols(const FwdColsParams& p, int nRows, bool codeMode, cudaStream_t s)
{
    dim3 grid((R + 127) / 128, nRows);
    if (codeMode) fwd_cols_kernel<1><<<grid, 128, 0, s>>>(p);
    else fwd_cols_kernel<0><<<grid, 128, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_rows(const RowsParams& p, cudaStream_t s)
{
    const int smem = fused_row_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((p.nRows + kRowWarps - 1) / kRowWarps);
    rows_kernel<false><<<grid, kRowWarps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_inv_rows(const RowsParams& p, cudaStream_t s)
{
    const int smem = fused_row_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int mGroups = (p.nonCoh + p.mPerCta - 1) / p.mPerCta;
    const int pGroups = (p.nPrnChunk + p.prnPerCta - 1) / p.prnPerCta;
    dim3 grid(C, p.nBins, pGroups * mGroups);
    rows_kernel<true><<<grid, kRowWarps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finish_replica(float2* Cc, size_t n, cudaStream_t s)
{
    finish_replica_kernel<<<148 * 2, 256, 0, s>>>(Cc, n, 1.0f / (float)L);
    return cudaGetLastError();
}

cudaError_t launch_inv_cols(const InvColsParams& p, cudaStream_t s)
{
    dim3 grid((R + 127) / 128, p.nBins, p.nPrnChunk);
    inv_cols_kernel<<<grid, 128, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gc
