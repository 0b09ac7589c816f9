// Acquisition, fused path for FFT length 32736 = 33 x 32 x 31 (2*samplesPerCode at 16.368 Msps).
//
// Replaces the PRN x Doppler x non-coherent-block loop of
// GPS/GPS_L1CA/include/acquisition.m:155-200.  The 2N-point transforms the reference does with
// MATLAB's fft/ifft are computed exactly at that length (no power-of-two padding, so the
// code-phase index space 1..2N is the reference's own) as a four-step transform
//
//      L = 33 x 32 x 31, pairwise coprime  ->  Good-Thomas prime-factor algorithm: with the time /
//      lag index n = (992*n1 + 1023*n2 + 1056*n3) mod L and the frequency index addressed by its
//      residues (j mod 33, j mod 32, j mod 31) the transform is a plain 3-D DFT, no twiddles.
//
//   forward  (wipe-off + FFT, PRN independent, acquisition.m:169-183):
//      fwd_cols : one thread per (n2, n3): 33 gathered int8 I/Q samples, carrier wipe-off with a
//                 64-bit fixed-point phase, 33-point DFT in registers (3 x 11 codelet)
//      fwd_rows : one warp per row k1: 32-point DFTs (31 lanes), shared-memory transpose,
//                 31-point DFTs (32 lanes)
//      spectrum layout X[k1][k3][k2]; the replica spectra use the same layout, so no index map
//      is ever applied in the frequency domain.
//   inverse  (acquisition.m:186-190):
//      inv_rows : one warp per row: load X*conj(FFT(code))/L, 31-point then 32-point inverse DFTs
//      inv_cols : one thread per (t2, t3): for each non-coherent block 33-point inverse DFT,
//                 |.|, accumulate in registers; after the last block the running maximum /
//                 first arg-max over the thread's 33 code phases (992*t1 + 1023*t2 + 1056*t3) mod L.
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
constexpr int kPitchF = RA + 1;  // smem pitch (float2) of the forward 31 x 32 exchange
constexpr int kPitchI = RB;      // smem pitch of the inverse 32 x 31 exchange
constexpr int kRowWarps = 8;     // warps (= rows in flight) per CTA in the row kernels
// Good-Thomas index maps (33, 32, 31 pairwise coprime): time / lag index
//   n = (992*n1 + 1023*n2 + 1056*n3) mod 32736,   992 = L/33, 1023 = L/32, 1056 = L/31
// while the frequency index j is addressed by its residues (j mod 33, j mod 32, j mod 31).
// With these maps the 32736-point DFT is a plain 33 x 32 x 31 three-dimensional DFT: no twiddle
// factors between the passes.
constexpr int kM1 = R, kM2 = L / RA, kM3 = L / RB;

// ------------------------------------------------------------------ column pass (forward)
// grid (ceil(R/128), nRows), block 128; thread = (n2, n3) = p / 31, p % 31.
// MODE 0: IF samples with carrier wipe-off; MODE 1: code table.
template <int MODE>
__global__ void __launch_bounds__(128)
fwd_cols_kernel(FwdColsParams p)
{
    const int pp = blockIdx.x * 128 + threadIdx.x;
    if (pp >= R) return;
    const int n2 = pp / RB, n3 = pp % RB;
    const int row = blockIdx.y;               // MODE 0: km = k*nonCoh + m ; MODE 1: prn slot
    const int base = (kM2 * n2 + kM3 * n3) % L;
    float2 x[C];
    if (MODE == 0) {
        const int k = row / p.nonCoh, m = row % p.nonCoh;
        const uint64_t dphi = p.dphi[k];
        const int8_t* src = p.rec + 2 * ((size_t)p.winStart + (size_t)m * p.N);   // window x((m-1)N+1 : (m+1)N)
#pragma unroll
        for (int n1 = 0; n1 < C; ++n1) {
            int n = base + kM1 * n1;
            n -= (n >= L) ? L : 0;
            const char2 s = *reinterpret_cast<const char2*>(src + 2 * (size_t)n);
            float sn, cs;
            fix_sincos(dphi * (uint64_t)n, &sn, &cs);          // exp(-1i*f*phasePoints(n)), :172
            const float I = p.swapIQ ? (float)s.y : (float)s.x, Q = p.swapIQ ? (float)s.x : (float)s.y;
            x[n1] = make_float2(fmaf(cs, I, sn * Q), fmaf(cs, Q, -sn * I));   // :180-181
        }
    } else {
        const int8_t* code = p.codeTab + (size_t)row * p.N;    // caCodesTable, zero padded to 2N (:160)
#pragma unroll
        for (int n1 = 0; n1 < C; ++n1) {
            int n = base + kM1 * n1;
            n -= (n >= L) ? L : 0;
            x[n1] = make_float2(n < p.N ? (float)code[n] : 0.f, 0.f);
        }
    }
    float2* dst = p.out + (size_t)row * L + pp;
    codelet::dft33_fwd(x, [&](int k1, float re, float im) { dst[(size_t)k1 * R] = make_float2(re, im); });
}

// ------------------------------------------------------------------ row pass, forward
// One warp per row (fixed k1): 32 x 31 two-dimensional DFT over (n2, n3) in place.
// in : element (n2, n3) at n2*31 + n3       out: element (k2, k3) at k3*32 + k2
__global__ void __launch_bounds__(kRowWarps * 32)
fwd_rows_kernel(RowsParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* s_x = reinterpret_cast<float2*>(smem_raw) + warp * (RB * kPitchF);
    const long long row = (long long)blockIdx.x * kRowWarps + warp;
    if (row >= p.nRows) return;
    float2* io = p.X + row * R;
    if (lane < RB) {                                             // lane = n3: DFT-32 over n2
        float2 v[RA];
#pragma unroll
        for (int n2 = 0; n2 < RA; ++n2) v[n2] = io[n2 * RB + lane];
        codelet::dft32_fwd(v, [&](int k2, float re, float im) { s_x[lane * kPitchF + k2] = make_float2(re, im); });
    }
    __syncwarp();
    {                                                            // lane = k2: DFT-31 over n3
        float2 u[RB];
#pragma unroll
        for (int n3 = 0; n3 < RB; ++n3) u[n3] = s_x[n3 * kPitchF + lane];
        codelet::dft31_fwd(u, [&](int k3, float re, float im) { io[k3 * RA + lane] = make_float2(re, im); });
    }
}

// ------------------------------------------------------------------ row pass, inverse (dominant kernel)
// One warp per row of one (PRN, bin, block): X .* Cc, inverse 31 x 32 DFT over (k3, k2) -> (t3, t2).
// in : element (k2, k3) at k3*32 + k2       out: element (t2, t3) at t2*31 + t3
// grid: x = k1 (33 rows), y = PRN group, z = bin * mGroups + block group, so that CTAs scheduled
// together share one X slice (L1/L2 hits) while the replica spectra stay L2 resident.
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
inv_rows_kernel(RowsParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* s_x = reinterpret_cast<float2*>(smem_raw) + warp * (RA * kPitchI);
    const int k1 = blockIdx.x, pg = blockIdx.y;
    const int mGroups = (p.nonCoh + p.mPerCta - 1) / p.mPerCta;
    const int k = blockIdx.z / mGroups, mg = blockIdx.z % mGroups;
    const int pi = pg * p.prnPerCta + warp / p.mPerCta;         // list slot within this launch's chunk
    const int m = mg * p.mPerCta + warp % p.mPerCta;
    if (pi >= p.nPrnChunk || m >= p.nonCoh) return;
    const float2* src = p.X + ((size_t)(k * p.nonCoh + m) * C + k1) * R;
    const float2* mul = p.Cc + ((size_t)p.prnList[p.prnSlot0 + pi] * C + k1) * R;
    float2* dst = p.W + (((size_t)(pi * p.nBins + k) * p.nonCoh + m) * C + k1) * R;
    {                                                            // lane = k2: DFT-31 over k3
        float2 u[RB];
#pragma unroll
        for (int k3 = 0; k3 < RB; ++k3)                          // IQfreqDom .* caCodeFreqDom (:186)
            u[k3] = cmul(src[k3 * RA + lane], __ldg(mul + k3 * RA + lane));
        codelet::dft31_inv(u, [&](int t3, float re, float im) { s_x[lane * kPitchI + t3] = make_float2(re, im); });
    }
    __syncwarp();
    if (lane < RB) {                                             // lane = t3: DFT-32 over k2
        float2 v[RA];
#pragma unroll
        for (int k2 = 0; k2 < RA; ++k2) v[k2] = s_x[k2 * kPitchI + lane];
        codelet::dft32_inv(v, [&](int t2, float re, float im) { __stcs(dst + t2 * RB + lane, make_float2(re, im)); });
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
// grid (ceil(R/128), nBins, nPrnChunk), block 128: thread = (t2, t3) column of one (PRN, bin).
__global__ void __launch_bounds__(128)
inv_cols_kernel(InvColsParams p)
{
    const int pp = blockIdx.x * 128 + threadIdx.x;
    const int k = blockIdx.y, pi = blockIdx.z;
    float acc[C];
#pragma unroll
    for (int i = 0; i < C; ++i) acc[i] = 0.f;
    if (pp < R) {
        const float2* base = p.W + ((size_t)(pi * p.nBins + k) * p.nonCoh) * L + pp;
        for (int m = 0; m < p.nonCoh; ++m) {                    // acquisition.m:175
            float2 x[C];
#pragma unroll
            for (int k1 = 0; k1 < C; ++k1) x[k1] = __ldcs(base + (size_t)m * L + (size_t)k1 * R);
            codelet::dft33_inv(x, [&](int t1, float re, float im) {
                acc[t1] += sqrtf(fmaf(re, re, im * im));        // abs(ifft(.)) summed over blocks (:188-190)
            });
        }
    }
    // running maximum with MATLAB first-index tie breaking (smaller code phase wins)
    float best = -1.f;
    int bidx = 0x7fffffff;
    if (pp < R) {
        const int rest = (kM2 * (pp / RB) + kM3 * (pp % RB)) % L;
#pragma unroll
        for (int t1 = 0; t1 < C; ++t1) {
            int idx = rest + kM1 * t1;                          // lag = (992*t1 + 1023*t2 + 1056*t3) mod L
            idx -= (idx >= L) ? L : 0;
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
    const int smem = (int)(sizeof(float2) * kRowWarps * RB * kPitchF);
    cudaError_t e = cudaFuncSetAttribute(fwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((p.nRows + kRowWarps - 1) / kRowWarps);
    fwd_rows_kernel<<<grid, kRowWarps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

template <int WARPS, int MINB>
static cudaError_t launch_inv_rows_t(const RowsParams& p, cudaStream_t s)
{
    const int smem = (int)(sizeof(float2) * WARPS * RA * kPitchI);
    cudaError_t e = cudaFuncSetAttribute(inv_rows_kernel<WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int mGroups = (p.nonCoh + p.mPerCta - 1) / p.mPerCta;
    const int pGroups = (p.nPrnChunk + p.prnPerCta - 1) / p.prnPerCta;
    dim3 grid(C, pGroups, p.nBins * mGroups);
    inv_rows_kernel<WARPS, MINB><<<grid, WARPS * 32, smem, s>>>(p);
    return cudaGetLastError();
}

// p.prnPerCta * p.mPerCta warps per CTA: 8 (2 x 4, 128 registers, 16 warps/SM), 6 (1 x 6 or 2 x 3,
// 96 registers, 18 warps/SM) or 5 (1 x 5, 96 registers, 20 warps/SM)
cudaError_t launch_inv_rows(const RowsParams& p, cudaStream_t s)
{
    switch (p.prnPerCta * p.mPerCta) {
        case 5: return launch_inv_rows_t<5, 4>(p, s);
        case 6: return launch_inv_rows_t<6, 3>(p, s);
        case 10: return launch_inv_rows_t<10, 2>(p, s);
        default: return launch_inv_rows_t<8, 2>(p, s);
    }
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
