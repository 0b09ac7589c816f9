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
//     changes a sum by ~1e-3 relative.  ceil() is a round-up add of 2^52 (FP64 pipe) instead of a
//     float->int conversion (quarter-rate XU pipe);
//   * carrier: phase kept as a 64-bit fixed-point fraction of a turn; one sincospif per 8-sample
//     chunk, the 8 in-chunk rotations e^{-i*j*dphi} are per-epoch constants;
//   * sums: fp32 inside a thread (8 samples per chunk, 4 chunks), float64 across threads.
//
// The per-epoch scalar work is split so that little of it sits on the critical path:
//   warp 0 lane 0 : PLL (atan) and the 15 recorded values          after the sums are reduced
//   warp 1 lane 0 : DLL and the geometry of the next block         concurrently with warp 0
//   warp 2 lane 0 : NCO phases at the end of the block just started (fmod etc.) — needs only that
//                   block's parameters, so it runs in the shadow of the sample loop.
#include "common.cuh"
#include "track.h"

namespace gc {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kStage = 16;   // epochs of results staged in smem before a coalesced flush
constexpr double kCeilMagic = 6755399441055744.0;   // 1.5 * 2^52: t + magic stays in [2^52, 2^53) for |t| < 2^51

struct EpochParams {          // parameters of the block being correlated (read by every thread)
    double aE, aP, aL;        // first elements of the early/prompt/late tcode vectors (tracking.m:252-266)
    double cE, cP, cL;        // last elements
    double mE, mP, mL;        // middle elements (a+c)/2, used when the vector length is odd
    double d;                 // codePhaseStep (:219)
    int n;                    // blksize - 1
    int blk;                  // blksize (:222)
    int generic;              // the three colon vectors disagree on their length: slow path
    int nE, nP, nL;
    long long pos;            // first sample of this block (absolute, in complex samples)
    uint64_t phase0, dphi;    // carrier phase at sample 0 and per-sample increment (turns, 0.64)
    int stop;
    int pad;
    double codeFreq, carrFreq, remCodePhase, remCarrPhase;   // values used in this block (recorded)
};

struct NextPhases {           // NCO phases at the end of the block being correlated (warp 2)
    double remCodePhase, remCarrPhase;
    uint64_t phase0;
};

struct LoopMem {              // loop-filter memories
    double oldCodeNco, oldCodeError, oldCarrNco, oldCarrError, carrFreqBasis;
};

// MATLAB colon vector a:d:b for non-integer a (Cleve Moler's colonop): element count n+1 and
// last element c.  tracking.m:252-268.
__device__ __noinline__ void colon_setup(double a, double d, double b, int* n_out, double* c_out)
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
    const bool left = two < n;
    const double base = left ? a : c;
    const double step = __dmul_rn((double)(left ? idx : (n - idx)), left ? d : -d);
    const double v = __dadd_rn(base, step);
    return (two == n) ? __dmul_rn(__dadd_rn(a, c), 0.5) : v;
}
// ceil(t) for |t| < 2^31 as an integer: a round-up add of 1.5*2^52 leaves ceil(t) (two's complement)
// in the low word of the sum
__device__ __forceinline__ int ceil_idx(double t) { return __double2loint(__dadd_ru(t, kCeilMagic)); }

// Geometry of a block from the NCO state (tracking.m:219-222, 252-268).
__device__ void plan_epoch(const TrackParams& p, double codeFreq, double remCodePhase,
                           double remCarrPhase, uint64_t phase0, long long pos, EpochParams& ep)
{
    const double step = __ddiv_rn(codeFreq, p.fs);                                      // :219
    const int blk = (int)ceil(__ddiv_rn(__dsub_rn(p.codeLength, remCodePhase), step));  // :222
    ep.d = step;
    ep.blk = blk;
    ep.n = blk - 1;
    ep.pos = pos;
    ep.stop = (pos + blk > p.recSamples) || blk <= 0;                                   // :241
    const double rem = remCodePhase;
    const double span = __dmul_rn((double)(blk - 1), step);
    ep.aE = __dsub_rn(rem, p.spc);
    ep.aL = __dadd_rn(rem, p.spc);
    ep.aP = rem;
    // b = ((blksize-1)*codePhaseStep + remCodePhase) -/+ earlyLateSpc, evaluated left to right
    const double bE = __dsub_rn(__dadd_rn(span, rem), p.spc);
    const double bL = __dadd_rn(__dadd_rn(span, rem), p.spc);
    const double bP = __dadd_rn(span, rem);
    // colonop: n = round((b-a)/d), c = a + n*d snapped to b when within tolerance.  For these
    // vectors n is blksize-1 and c snaps to b; verify that cheaply and fall back to the full
    // algorithm otherwise.
    bool easy = true;
    {
        const double tolE = 4.440892098500626e-16 * fmax(fabs(ep.aE), fabs(bE));
        const double tolL = 4.440892098500626e-16 * fmax(fabs(ep.aL), fabs(bL));
        const double tolP = 4.440892098500626e-16 * fmax(fabs(ep.aP), fabs(bP));
        const double eE = __dsub_rn(__dadd_rn(ep.aE, span), bE);
        const double eL = __dsub_rn(__dadd_rn(ep.aL, span), bL);
        const double eP = __dsub_rn(__dadd_rn(ep.aP, span), bP);
        // (step is never an integer here, so colonop's integer special cases cannot apply)
        easy = (fabs(eE) <= tolE) && (fabs(eL) <= tolL) && (fabs(eP) <= tolP) && blk > 2 && step != floor(step);
    }
    if (easy) {
        ep.nE = ep.nP = ep.nL = blk - 1;
        ep.cE = bE; ep.cL = bL; ep.cP = bP;
    } else {
        colon_setup(ep.aE, step, bE, &ep.nE, &ep.cE);
        colon_setup(ep.aL, step, bL, &ep.nL, &ep.cL);
        colon_setup(ep.aP, step, bP, &ep.nP, &ep.cP);
    }
    ep.generic = !(ep.nE == blk - 1 && ep.nP == blk - 1 && ep.nL == blk - 1);
    ep.mE = __dmul_rn(__dadd_rn(ep.aE, ep.cE), 0.5);
    ep.mP = __dmul_rn(__dadd_rn(ep.aP, ep.cP), 0.5);
    ep.mL = __dmul_rn(__dadd_rn(ep.aL, ep.cL), 0.5);
    ep.phase0 = phase0;
    ep.codeFreq = codeFreq;
    ep.remCodePhase = remCodePhase; ep.remCarrPhase = remCarrPhase;
}
// carrier part of the block parameters (written by the PLL thread)
__device__ __forceinline__ void plan_carrier(const TrackParams& p, double carrFreq, EpochParams& ep)
{
    ep.carrFreq = carrFreq;
    ep.dphi = turns_to_fix(carrFreq / p.fs);
}

// NCO phases after the block described by ep (tracking.m:273, :280-283)
__device__ void end_phases(const TrackParams& p, const EpochParams& ep, NextPhases& nx)
{
    const double lastP = ep.generic ? colon_elem(ep.aP, ep.d, ep.cP, ep.nP, ep.blk - 1) : ep.cP;
    nx.remCodePhase = __dsub_rn(__dadd_rn(lastP, ep.d), p.codeLength);                        // :273
    const double w = __dmul_rn(__dmul_rn(ep.carrFreq, 2.0), 3.141592653589793);               // :281
    const double trigEnd = __dadd_rn(__dmul_rn(w, __ddiv_rn((double)ep.blk, p.fs)), ep.remCarrPhase);
    nx.remCarrPhase = fmod(trigEnd, kTwoPi);                                                  // :283
    nx.phase0 = turns_to_fix(nx.remCarrPhase / kTwoPi);
}

// exact float of a signed byte already xor-ed with 0x80: 0x4B0000bb = 2^23 + (b+128)
template <int B>
__device__ __forceinline__ float byte_to_float(uint32_t wx)
{
    return __uint_as_float(__byte_perm(wx, 0x4B000000u, 0x7650 | B)) - 8388736.0f;
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
track_kernel(TrackParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [buf0 | buf1 | code table (float) | warp partials | staging | params | mbarriers]
    int8_t* const buf0 = reinterpret_cast<int8_t*>(smem_raw);
    float* s_code = reinterpret_cast<float*>(smem_raw + 2 * (size_t)p.bufBytes);
    double* s_part = reinterpret_cast<double*>(s_code + ((p.codeLen + 2 + 3) & ~3));
    double* s_stage = s_part + kWarps * 6;                       // [15][kStage]
    EpochParams* s_ep = reinterpret_cast<EpochParams*>(s_stage + GC_TRACK_ROWS * kStage);   // [2]
    NextPhases* s_nx = reinterpret_cast<NextPhases*>(s_ep + 2);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_nx + 1);     // 2 mbarriers
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

    LoopMem lm;   // warp 0 lane 0: carrier memories; warp 1 lane 0: code memories
    lm.oldCodeNco = lm.oldCodeError = lm.oldCarrNco = lm.oldCarrError = 0.0;   // :173-178
    lm.carrFreqBasis = cinfo.acqFreq;                                          // :168
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        s_issued[0] = s_issued[1] = 0;
        // :163-170 codeFreq = codeFreqBasis, remCodePhase = 0, carrFreq = acquiredFreq, remCarrPhase = 0; :150 fseek
        plan_epoch(p, p.codeFreqBasis, 0.0, 0.0, 0ull, cinfo.startSample, s_ep[0]);
        plan_carrier(p, cinfo.acqFreq, s_ep[0]);
    }
    __syncthreads();

    const long long recBytesUp = (p.recSamples * 2 + 15) & ~15LL;
    auto prefetch = [&](long long startSample, int stage) {      // thread 0 only
        const long long b0 = (startSample * 2) & ~15LL;
        long long n = p.bufBytes;
        if (b0 + n > recBytesUp) n = recBytesUp - b0;
        if (b0 < 0 || n <= 0) return false;
        mbar_expect_tx(&s_bar[stage], (uint32_t)n);
        bulk_g2s(buf0 + (size_t)stage * p.bufBytes, p.rec + b0, (uint32_t)n, &s_bar[stage]);
        s_issued[stage] = 1;
        return true;
    };
    if (tid == 0 && !s_ep[0].stop) prefetch(s_ep[0].pos, 0);
    __syncthreads();

    int e = 0;
    for (; e < p.nEpochs; ++e) {
        const int stage = e & 1;
        const EpochParams& ep = s_ep[stage];
        if (ep.stop) {
            // never leave a bulk copy in flight into this CTA's shared memory
            if (s_issued[stage]) mbar_wait(&s_bar[stage], (e >> 1) & 1);
            break;
        }
        const int blk = ep.blk, n = ep.n;
        const long long pos = ep.pos;
        const uint64_t dphi = ep.dphi, phase0 = ep.phase0;
        // window of epoch e+1 starts where this one ends; fetch it while we correlate
        if (tid == 0 && e + 1 < p.nEpochs) prefetch(pos + blk, stage ^ 1);
        // NCO phases at the end of this block: off the critical path, in the shadow of the sample loop
        if (tid == 64) end_phases(p, ep, *s_nx);
        mbar_wait(&s_bar[stage], (e >> 1) & 1);
        if (tid == 0) s_issued[stage] = 0;

        // in-chunk rotations e^{-i*2*pi*j*dphi}, j = 0..7 (per-epoch constants)
        float wc[8], ws[8];
        {
            float s, c;
            fix_sincos(dphi * (uint64_t)(lane & 7), &s, &c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                wc[j] = __shfl_sync(0xffffffffu, c, j);
                ws[j] = __shfl_sync(0xffffffffu, s, j);
            }
        }
        const int off = (int)((pos * 2) & 15) >> 1;              // samples skipped in the first 16-byte chunk
        const int nChunks = (off + blk + 7) >> 3;
        const bool inBuf = ((long long)nChunks * 16 <= p.bufBytes);
        const int8_t* src = buf0 + (size_t)stage * p.bufBytes;
        const int8_t* gsrc = p.rec + ((pos * 2) & ~15LL);
        const double d = ep.d;
        const double aE = ep.aE, aP = ep.aP, aL = ep.aL, cE = ep.cE, cP = ep.cP, cL = ep.cL;
        const bool generic = ep.generic != 0;

        float aIE = 0, aQE = 0, aIP = 0, aQP = 0, aIL = 0, aQL = 0;
        for (int c = tid; c < nChunks; c += kThreads) {
            int4 raw;
            if (inBuf) raw = *reinterpret_cast<const int4*>(src + (size_t)c * 16);
            else raw = __ldg(reinterpret_cast<const int4*>(gsrc + (size_t)c * 16));   // oversize block: straight from L2
            const int k0 = c * 8 - off;
            const uint32_t w0 = (uint32_t)raw.x ^ 0x80808080u, w1 = (uint32_t)raw.y ^ 0x80808080u,
                           w2 = (uint32_t)raw.z ^ 0x80808080u, w3 = (uint32_t)raw.w ^ 0x80808080u;
            float xi[8], xq[8];                                   // tracking.m:233-235
            xi[0] = byte_to_float<0>(w0); xq[0] = byte_to_float<1>(w0); xi[1] = byte_to_float<2>(w0); xq[1] = byte_to_float<3>(w0);
            xi[2] = byte_to_float<0>(w1); xq[2] = byte_to_float<1>(w1); xi[3] = byte_to_float<2>(w1); xq[3] = byte_to_float<3>(w1);
            xi[4] = byte_to_float<0>(w2); xq[4] = byte_to_float<1>(w2); xi[5] = byte_to_float<2>(w2); xq[5] = byte_to_float<3>(w2);
            xi[6] = byte_to_float<0>(w3); xq[6] = byte_to_float<1>(w3); xi[7] = byte_to_float<2>(w3); xq[7] = byte_to_float<3>(w3);
            float pIE = 0, pQE = 0, pIP = 0, pQP = 0, pIL = 0, pQL = 0;
            // A chunk lies in the left half (t = a + k*d), the right half (t = c - (n-k)*d), or it is
            // one of the <= 3 special chunks per block (first, last, the one holding the middle).
            const bool allLeft = (2 * (k0 + 7) < n), allRight = (2 * k0 > n);
            if (k0 >= 0 && k0 + 7 < blk && (allLeft || allRight) && !generic) {
                const double sg = allLeft ? 1.0 : -1.0;
                const double f0 = (double)(allLeft ? k0 : (n - k0));
                const double ds = allLeft ? d : -d;
                const double bE = allLeft ? aE : cE, bP = allLeft ? aP : cP, bL = allLeft ? aL : cL;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const double st = __dmul_rn(__fma_rn(sg, (double)j, f0), ds);   // (+-)(k or n-k)*d, exact integer factor
                    // code replicas (tracking.m:252-270): ceil(tcode) indexes [c(L) c c(1)] 0-based
                    const float vE = s_code[ceil_idx(__dadd_rn(bE, st))];
                    const float vP = s_code[ceil_idx(__dadd_rn(bP, st))];
                    const float vL = s_code[ceil_idx(__dadd_rn(bL, st))];
                    // x * e^{-i*j*dphi}   (tracking.m:287-292 with the chunk phase factored out)
                    const float ur = fmaf(wc[j], xi[j], ws[j] * xq[j]);
                    const float ui = fmaf(wc[j], xq[j], -ws[j] * xi[j]);
                    pIE = fmaf(vE, ur, pIE); pQE = fmaf(vE, ui, pQE);                 // :295-300
                    pIP = fmaf(vP, ur, pIP); pQP = fmaf(vP, ui, pQP);
                    pIL = fmaf(vL, ur, pIL); pQL = fmaf(vL, ui, pQL);
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const int k = k0 + j;
                    if (k < 0 || k >= blk) continue;
                    const uint32_t wsel = (j < 2) ? w0 : (j < 4) ? w1 : (j < 6) ? w2 : w3;
                    const uint32_t hw = wsel >> ((j & 1) * 16);
                    const float sxi = (float)((int)(hw & 0xff) - 128), sxq = (float)((int)((hw >> 8) & 0xff) - 128);
                    const float vE = s_code[ceil_idx(colon_elem(aE, d, cE, ep.nE, k))];
                    const float vP = s_code[ceil_idx(colon_elem(aP, d, cP, ep.nP, k))];
                    const float vL = s_code[ceil_idx(colon_elem(aL, d, cL, ep.nL, k))];
                    float sj, cj;
                    fix_sincos(dphi * (uint64_t)j, &sj, &cj);
                    const float ur = fmaf(cj, sxi, sj * sxq);
                    const float ui = fmaf(cj, sxq, -sj * sxi);
                    pIE = fmaf(vE, ur, pIE); pQE = fmaf(vE, ui, pQE);
                    pIP = fmaf(vP, ur, pIP); pQP = fmaf(vP, ui, pQP);
                    pIL = fmaf(vL, ur, pIL); pQL = fmaf(vL, ui, pQL);
                }
            }
            // rotate the chunk sums by e^{-i*phase(k0)}
            float s0, c0;
            fix_sincos(phase0 + dphi * (uint64_t)(long long)k0, &s0, &c0);
            aIE += fmaf(c0, pIE, s0 * pQE); aQE += fmaf(c0, pQE, -s0 * pIE);
            aIP += fmaf(c0, pIP, s0 * pQP); aQP += fmaf(c0, pQP, -s0 * pIP);
            aIL += fmaf(c0, pIL, s0 * pQL); aQL += fmaf(c0, pQL, -s0 * pIL);
        }
        // cross-thread reduction in float64: warp shuffle, then warps 0 and 1 over the warp partials
        double v[6] = {(double)aIE, (double)aQE, (double)aIP, (double)aQP, (double)aIL, (double)aQL};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < 6; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
        if (lane == 0)
#pragma unroll
            for (int q = 0; q < 6; ++q) s_part[warp * 6 + q] = v[q];
        __syncthreads();
        if (warp < 2) {
#pragma unroll
            for (int q = 0; q < 6; ++q) v[q] = (lane < kWarps) ? s_part[lane * 6 + q] : 0.0;
#pragma unroll
            for (int o = kWarps / 2; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < 6; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
            const double I_E = v[0], Q_E = v[1], I_P = v[2], Q_P = v[3], I_L = v[4], Q_L = v[5];
            double* sg = s_stage + (e % kStage);
            if (tid == 0) {
                // PLL (tracking.m:305-317)
                const double carrError = atan(__ddiv_rn(Q_P, I_P)) / kTwoPi;
                const double carrNco = __dadd_rn(__dadd_rn(lm.oldCarrNco, __dmul_rn(p.pA, __dsub_rn(carrError, lm.oldCarrError))),
                                                 __dmul_rn(carrError, p.pB));
                lm.oldCarrNco = carrNco; lm.oldCarrError = carrError;
                plan_carrier(p, __dadd_rn(lm.carrFreqBasis, carrNco), s_ep[stage ^ 1]);   // :317 carrFreq of the next block
                sg[GC_F_ABSOLUTE_SAMPLE * kStage] = (double)pos;                    // :215 ftell/2
                sg[GC_F_REM_CODE_PHASE * kStage] = ep.remCodePhase;                 // :249
                sg[GC_F_REM_CARR_PHASE * kStage] = ep.remCarrPhase;                 // :277
                sg[GC_F_CARR_FREQ * kStage] = ep.carrFreq;                          // :314
                sg[GC_F_CODE_FREQ * kStage] = ep.codeFreq;                          // :332
                sg[GC_F_PLL_DISCR * kStage] = carrError;                            // :340-341
                sg[GC_F_PLL_DISCR_FILT * kStage] = carrNco;
                sg[GC_F_I_E * kStage] = I_E; sg[GC_F_I_P * kStage] = I_P; sg[GC_F_I_L * kStage] = I_L;   // :343-348
                sg[GC_F_Q_E * kStage] = Q_E; sg[GC_F_Q_P * kStage] = Q_P; sg[GC_F_Q_L * kStage] = Q_L;
            } else if (tid == 32) {
                // DLL (tracking.m:322-335)
                const double sE = sqrt(__dadd_rn(__dmul_rn(I_E, I_E), __dmul_rn(Q_E, Q_E)));
                const double sL = sqrt(__dadd_rn(__dmul_rn(I_L, I_L), __dmul_rn(Q_L, Q_L)));
                const double codeError = __ddiv_rn(__dsub_rn(sE, sL), __dadd_rn(sE, sL));
                const double codeNco = __dadd_rn(__dadd_rn(lm.oldCodeNco, __dmul_rn(p.cA, __dsub_rn(codeError, lm.oldCodeError))),
                                                 __dmul_rn(codeError, p.cB));
                lm.oldCodeNco = codeNco; lm.oldCodeError = codeError;
                sg[GC_F_DLL_DISCR * kStage] = codeError;                            // :338-339
                sg[GC_F_DLL_DISCR_FILT * kStage] = codeNco;
                // :335 codeFreq of the next block, then its geometry (phases come from end_phases())
                const NextPhases nx = *s_nx;
                plan_epoch(p, __dsub_rn(p.codeFreqBasis, codeNco), nx.remCodePhase, nx.remCarrPhase, nx.phase0,
                           pos + blk, s_ep[stage ^ 1]);
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
        }
    }
    __syncthreads();
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
    s += 2 * sizeof(EpochParams) + sizeof(NextPhases) + 2 * sizeof(double) + 2 * sizeof(uint64_t) + 2 * sizeof(int) + 64;
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
