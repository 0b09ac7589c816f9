// Tracking: correlate-and-dump with the DLL/PLL closed on the device.
//
// Replaces the epoch loop of GPS/GPS_L1CA/include/tracking.m:184-360.  One persistent CTA per
// channel walks the resident IF record; the 1 ms sample block of epoch e+1 is staged into shared
// memory by the TMA unit (cp.async.bulk + mbarrier, double buffered) while epoch e is correlated.
// Its start is known as soon as epoch e begins (start + blksize); only its length changes by a
// sample or so, so a fixed-size window is fetched.
//
// Numerics (targets: I/Q sums within 1e-6 relative of the float64 reference, every recorded
// state variable computed in float64 exactly as the reference does):
//   * code phase: t(k) is evaluated per sample in float64 with the reference's own operation
//     order (MATLAB colon vector a:d:b built from both ends, then ceil) because one chip flip
//     changes a sum by ~1e-3 relative;
//   * carrier: phase kept as a 64-bit fixed-point fraction of a turn; one sincospif per 8-sample
//     chunk, the 8 in-chunk rotations e^{-i*j*dphi} are per-epoch constants;
//   * sums: fp32 inside a thread (<= 8 samples per chunk, a few chunks), float64 across threads.
#include "common.cuh"
#include "track.h"

namespace gc {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kStage = 16;   // epochs of results staged in smem before a coalesced flush

struct EpochParams {          // written by lane 0 of warp 0, read by everybody
    double aE, cE, aP, cP, aL, cL, d;
    int nE, nP, nL;
    int blk;
    long long pos;            // first sample of this epoch (absolute, in complex samples)
    uint64_t phase0, dphi;    // carrier phase at sample 0 and per-sample increment (turns, 0.64)
    int stop;
};

struct LoopState {
    double codeFreq, remCodePhase, carrFreq, carrFreqBasis, remCarrPhase;
    double oldCodeNco, oldCodeError, oldCarrNco, oldCarrError;
    long long pos;
};

// MATLAB colon vector a:d:b for non-integer a (Cleve Moler's colonop): element count n+1 and
// last element c.  tracking.m:252-268.
__device__ __forceinline__ void colon_setup(double a, double d, double b, int* n_out, double* c_out)
{
    const double tol = 2.0 * 2.220446049250313e-16 * fmax(fabs(a), fabs(b));
    int n;
    if (a == floor(a) && d == 1.0) n = (int)(floor(b) - a);
    else if (a == floor(a) && d == floor(d)) n = (int)trunc(__ddiv_rn(__dsub_rn(b, a), d));
    else {
        const double q = __ddiv_rn(__dsub_rn(b, a), d);
        n = (int)(q >= 0 ? floor(q + 0.5) : -floor(-q + 0.5));
        if (__dsub_rn(__dadd_rn(a, __dmul_rn((double)n, d)), b) > tol) n -= 1;
    }
    double c = __dadd_rn(a, __dmul_rn((double)n, d));
    if (__dsub_rn(c, b) > -tol) c = b;
    *n_out = n;
    *c_out = c;
}
// element idx of that vector: a + idx*d from the left half, c - (n-idx)*d from the right half
__device__ __forceinline__ double colon_elem(double a, double d, double c, int n, int idx)
{
    const int two = 2 * idx;
    if (two == n) return __dmul_rn(__dadd_rn(a, c), 0.5);
    const bool left = two < n;
    const double base = left ? a : c;
    const double step = __dmul_rn((double)(left ? idx : (n - idx)), d);
    return left ? __dadd_rn(base, step) : __dsub_rn(base, step);
}

// Geometry of the next epoch from the loop state (tracking.m:219-222, 252-268, 273, 280-283).
__device__ void plan_epoch(const TrackParams& p, LoopState& st, EpochParams& ep)
{
    const double step = __ddiv_rn(st.codeFreq, p.fs);                                   // :219
    const int blk = (int)ceil(__ddiv_rn(__dsub_rn(p.codeLength, st.remCodePhase), step));   // :222
    ep.d = step;
    ep.blk = blk;
    ep.pos = st.pos;
    ep.stop = (st.pos + blk > p.recSamples) || blk <= 0;                                // :241
    const double rem = st.remCodePhase;
    const double span = __dmul_rn((double)(blk - 1), step);
    ep.aE = __dsub_rn(rem, p.spc);
    ep.aL = __dadd_rn(rem, p.spc);
    ep.aP = rem;
    // b = ((blksize-1)*codePhaseStep + remCodePhase) -/+ earlyLateSpc, left to right
    colon_setup(ep.aE, step, __dsub_rn(__dadd_rn(span, rem), p.spc), &ep.nE, &ep.cE);
    colon_setup(ep.aL, step, __dadd_rn(__dadd_rn(span, rem), p.spc), &ep.nL, &ep.cL);
    colon_setup(ep.aP, step, __dadd_rn(span, rem), &ep.nP, &ep.cP);
    ep.phase0 = turns_to_fix(st.remCarrPhase / kTwoPi);
    ep.dphi = turns_to_fix(st.carrFreq / p.fs);
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
track_kernel(TrackParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [buf0 | buf1 | code table (float) | warp partials | staging | params | mbarriers]
    int8_t* buf[2] = {reinterpret_cast<int8_t*>(smem_raw), reinterpret_cast<int8_t*>(smem_raw) + p.bufBytes};
    float* s_code = reinterpret_cast<float*>(smem_raw + 2 * (size_t)p.bufBytes);
    double* s_part = reinterpret_cast<double*>(s_code + ((p.codeLen + 2 + 3) & ~3));
    double* s_stage = s_part + kWarps * 6;                       // [15][kStage]
    EpochParams* s_ep = reinterpret_cast<EpochParams*>(s_stage + GC_TRACK_ROWS * kStage);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_ep + 1);     // 2 mbarriers
    int* s_issued = reinterpret_cast<int*>(s_bar + 2);           // copy in flight per stage

    const int ch = blockIdx.x;
    const TrackChan cinfo = p.chans[ch];
    if (cinfo.prn == 0) {                                        // tracking.m:136
        if (threadIdx.x == 0) p.epochsDone[ch] = 0;
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* out = p.out + (size_t)ch * GC_TRACK_ROWS * p.nEpochs;

    // wrapped code table [c(L) c(1..L) c(1)]  (tracking.m:156-158)
    for (int i = tid; i < p.codeLen + 2; i += kThreads)
        s_code[i] = (float)p.codeTables[(size_t)ch * p.codeStride + i];

    LoopState st;   // live in lane 0 / warp 0 only
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        s_issued[0] = s_issued[1] = 0;
        st.codeFreq = p.codeFreqBasis;                           // :163
        st.remCodePhase = 0.0;                                   // :165
        st.carrFreq = cinfo.acqFreq;                             // :167
        st.carrFreqBasis = cinfo.acqFreq;                        // :168
        st.remCarrPhase = 0.0;                                   // :170
        st.oldCodeNco = st.oldCodeError = st.oldCarrNco = st.oldCarrError = 0.0;   // :173-178
        st.pos = cinfo.startSample;                              // :150 (fseek)
        plan_epoch(p, st, *s_ep);
    }
    __syncthreads();

    const long long recBytesUp = (p.recSamples * 2 + 15) & ~15LL;
    auto prefetch = [&](long long startSample, int stage) {      // thread 0 only
        const long long b0 = (startSample * 2) & ~15LL;
        long long n = p.bufBytes;
        if (b0 + n > recBytesUp) n = recBytesUp - b0;
        if (b0 < 0 || n <= 0) return false;
        mbar_expect_tx(&s_bar[stage], (uint32_t)n);
        bulk_g2s(buf[stage], p.rec + b0, (uint32_t)n, &s_bar[stage]);
        s_issued[stage] = 1;
        return true;
    };
    if (tid == 0 && !s_ep->stop) prefetch(s_ep->pos, 0);
    __syncthreads();

    int e = 0;
    for (; e < p.nEpochs; ++e) {
        const EpochParams ep = *s_ep;                            // broadcast read
        const int stage = e & 1;
        if (ep.stop) {
            // never leave a bulk copy in flight into this CTA's shared memory
            if (s_issued[stage]) mbar_wait(&s_bar[stage], (e >> 1) & 1);
            break;
        }
        // window of epoch e+1 starts where this one ends; fetch it while we correlate
        if (tid == 0 && e + 1 < p.nEpochs) prefetch(ep.pos + ep.blk, stage ^ 1);
        mbar_wait(&s_bar[stage], (e >> 1) & 1);
        if (tid == 0) s_issued[stage] = 0;

        // in-chunk rotations e^{-i*2*pi*j*dphi}, j = 0..7 (per-epoch constants)
        float wc[8], ws[8];
        {
            float s, c;
            fix_sincos(ep.dphi * (uint64_t)(lane & 7), &s, &c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                wc[j] = __shfl_sync(0xffffffffu, c, j);
                ws[j] = __shfl_sync(0xffffffffu, s, j);
            }
        }
        const int off = (int)((ep.pos * 2) & 15) >> 1;           // samples skipped in the first 16-byte chunk
        const int nChunks = (off + ep.blk + 7) >> 3;
        const bool inBuf = ((long long)nChunks * 16 <= p.bufBytes);
        const int8_t* src = buf[stage];
        const int8_t* gsrc = p.rec + ((ep.pos * 2) & ~15LL);

        float aIE = 0, aQE = 0, aIP = 0, aQP = 0, aIL = 0, aQL = 0;
        for (int c = tid; c < nChunks; c += kThreads) {
            int4 raw;
            if (inBuf) raw = *reinterpret_cast<const int4*>(src + (size_t)c * 16);
            else raw = __ldg(reinterpret_cast<const int4*>(gsrc + (size_t)c * 16));   // oversize block: straight from L2
            const int k0 = c * 8 - off;
            const uint32_t wds[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
            float pIE = 0, pQE = 0, pIP = 0, pQP = 0, pIL = 0, pQL = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j;
                const uint32_t wd = wds[j >> 1] >> ((j & 1) * 16);
                const bool valid = (k >= 0) && (k < ep.blk);
                const float xi = valid ? (float)(int)(int8_t)(wd & 0xff) : 0.f;          // tracking.m:233-235
                const float xq = valid ? (float)(int)(int8_t)((wd >> 8) & 0xff) : 0.f;
                const int kk = valid ? k : 0;
                // code replicas (tracking.m:252-270): ceil(tcode) indexes [c(L) c c(1)] 0-based
                const int iE = __double2int_ru(colon_elem(ep.aE, ep.d, ep.cE, ep.nE, kk));
                const int iP = __double2int_ru(colon_elem(ep.aP, ep.d, ep.cP, ep.nP, kk));
                const int iL = __double2int_ru(colon_elem(ep.aL, ep.d, ep.cL, ep.nL, kk));
                const float cE = s_code[iE], cP = s_code[iP], cL = s_code[iL];
                // x * e^{-i*j*dphi}   (tracking.m:287-292 with the chunk phase factored out)
                const float ur = fmaf(wc[j], xi, ws[j] * xq);
                const float ui = fmaf(wc[j], xq, -ws[j] * xi);
                pIE = fmaf(cE, ur, pIE); pQE = fmaf(cE, ui, pQE);                          // :295-300
                pIP = fmaf(cP, ur, pIP); pQP = fmaf(cP, ui, pQP);
                pIL = fmaf(cL, ur, pIL); pQL = fmaf(cL, ui, pQL);
            }
            // rotate the chunk sums by e^{-i*phase(k0)}
            float s0, c0;
            fix_sincos(ep.phase0 + ep.dphi * (uint64_t)(long long)k0, &s0, &c0);
            aIE += fmaf(c0, pIE, s0 * pQE); aQE += fmaf(c0, pQE, -s0 * pIE);
            aIP += fmaf(c0, pIP, s0 * pQP); aQP += fmaf(c0, pQP, -s0 * pIP);
            aIL += fmaf(c0, pIL, s0 * pQL); aQL += fmaf(c0, pQL, -s0 * pIL);
        }
        // cross-thread reduction in float64: warp shuffle, then warp 0 over the warp partials
        double v[6] = {(double)aIE, (double)aQE, (double)aIP, (double)aQP, (double)aIL, (double)aQL};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < 6; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
        if (lane == 0)
#pragma unroll
            for (int q = 0; q < 6; ++q) s_part[warp * 6 + q] = v[q];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 6; ++q) v[q] = (lane < kWarps) ? s_part[lane * 6 + q] : 0.0;
#pragma unroll
            for (int o = kWarps / 2; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < 6; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
            if (lane == 0) {
                const double I_E = v[0], Q_E = v[1], I_P = v[2], Q_P = v[3], I_L = v[4], Q_L = v[5];
                double* sg = s_stage + (e % kStage);
                sg[GC_F_ABSOLUTE_SAMPLE * kStage] = (double)ep.pos;                 // :215 ftell/2
                sg[GC_F_REM_CODE_PHASE * kStage] = st.remCodePhase;                // :249
                sg[GC_F_REM_CARR_PHASE * kStage] = st.remCarrPhase;                // :277
                // :273 remCodePhase = (tcode(blksize) + codePhaseStep) - codeLength
                st.remCodePhase = __dsub_rn(__dadd_rn(colon_elem(ep.aP, ep.d, ep.cP, ep.nP, ep.blk - 1), ep.d), p.codeLength);
                // :280-283 trigarg(blksize+1), rem(.,2*pi)
                const double w = __dmul_rn(__dmul_rn(st.carrFreq, 2.0), 3.141592653589793);
                const double trigEnd = __dadd_rn(__dmul_rn(w, __ddiv_rn((double)ep.blk, p.fs)), st.remCarrPhase);
                st.remCarrPhase = fmod(trigEnd, kTwoPi);
                // PLL (:305-317)
                const double carrError = atan(__ddiv_rn(Q_P, I_P)) / kTwoPi;
                const double carrNco = __dadd_rn(__dadd_rn(st.oldCarrNco, __dmul_rn(p.pA, __dsub_rn(carrError, st.oldCarrError))),
                                                 __dmul_rn(carrError, p.pB));
                st.oldCarrNco = carrNco; st.oldCarrError = carrError;
                sg[GC_F_CARR_FREQ * kStage] = st.carrFreq;                         // :314
                st.carrFreq = __dadd_rn(st.carrFreqBasis, carrNco);                // :317
                // DLL (:322-335)
                const double sE = sqrt(__dadd_rn(__dmul_rn(I_E, I_E), __dmul_rn(Q_E, Q_E)));
                const double sL = sqrt(__dadd_rn(__dmul_rn(I_L, I_L), __dmul_rn(Q_L, Q_L)));
                const double codeError = __ddiv_rn(__dsub_rn(sE, sL), __dadd_rn(sE, sL));
                const double codeNco = __dadd_rn(__dadd_rn(st.oldCodeNco, __dmul_rn(p.cA, __dsub_rn(codeError, st.oldCodeError))),
                                                 __dmul_rn(codeError, p.cB));
                st.oldCodeNco = codeNco; st.oldCodeError = codeError;
                sg[GC_F_CODE_FREQ * kStage] = st.codeFreq;                         // :332
                st.codeFreq = __dsub_rn(p.codeFreqBasis, codeNco);                 // :335
                sg[GC_F_DLL_DISCR * kStage] = codeError;                           // :338-341
                sg[GC_F_DLL_DISCR_FILT * kStage] = codeNco;
                sg[GC_F_PLL_DISCR * kStage] = carrError;
                sg[GC_F_PLL_DISCR_FILT * kStage] = carrNco;
                sg[GC_F_I_E * kStage] = I_E; sg[GC_F_I_P * kStage] = I_P; sg[GC_F_I_L * kStage] = I_L;   // :343-348
                sg[GC_F_Q_E * kStage] = Q_E; sg[GC_F_Q_P * kStage] = Q_P; sg[GC_F_Q_L * kStage] = Q_L;
                st.pos = ep.pos + ep.blk;
                plan_epoch(p, st, *s_ep);
            }
        }
        __syncthreads();
        // coalesced flush of the staged rows every kStage epochs
        if ((e % kStage) == kStage - 1) {
            const int e0 = e - (kStage - 1);
            for (int i = tid; i < GC_TRACK_ROWS * kStage; i += kThreads) {
                const int f = i / kStage, q = i % kStage;
                out[(size_t)f * p.nEpochs + e0 + q] = s_stage[f * kStage + q];
            }
            __syncthreads();
        }
    }
    // tail flush (e = number of completed epochs)
    const int rem = e % kStage;
    if (rem) {
        const int e0 = e - rem;
        for (int i = tid; i < GC_TRACK_ROWS * kStage; i += kThreads) {
            const int f = i / kStage, q = i % kStage;
            if (q < rem) out[(size_t)f * p.nEpochs + e0 + q] = s_stage[f * kStage + q];
        }
    }
    if (tid == 0) p.epochsDone[ch] = e;
}

size_t track_smem_bytes(int bufBytes, int codeLen)
{
    size_t s = 2 * (size_t)bufBytes;
    s += sizeof(float) * ((codeLen + 2 + 3) & ~3);
    s += sizeof(double) * (kWarps * 6 + GC_TRACK_ROWS * kStage);
    s += sizeof(EpochParams) + 2 * sizeof(uint64_t) + 2 * sizeof(int) + 64;
    return s;
}

cudaError_t launch_track(const TrackParams& p, int nCh, cudaStream_t stream)
{
    const size_t smem = track_smem_bytes(p.bufBytes, p.codeLen);
    cudaError_t err = cudaFuncSetAttribute(track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    track_kernel<<<nCh, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// out rows pre-fill (tracking.m:51-77): zeros for absoluteSample and the six I/Q rows, +inf elsewhere
__global__ void track_fill_kernel(double* out, int nCh, int nEpochs)
{
    const size_t total = (size_t)nCh * GC_TRACK_ROWS * nEpochs;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)((i / nEpochs) % GC_TRACK_ROWS);
        out[i] = (f == GC_F_ABSOLUTE_SAMPLE || (f >= GC_F_I_P && f <= GC_F_Q_L)) ? 0.0 : inf;
    }
}

cudaError_t launch_track_fill(double* out, int nCh, int nEpochs, cudaStream_t stream)
{
    track_fill_kernel<<<148 * 4, 256, 0, stream>>>(out, nCh, nEpochs);
    return cudaGetLastError();
}

}  // namespace gc
