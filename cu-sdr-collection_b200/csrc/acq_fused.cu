// Acquisition, fused transform plans for FFT lengths L = C x RA x RB (2*samplesPerCode):
//
//      32736 = 33 x 32 x 31   (16.368 Msps, 1 ms codes)      40000 = 50 x 32 x 25   (20 Msps)
//      36000 = 45 x 32 x 25   (18 Msps: the reference default of six signal folders)
//      24000 = 30 x 32 x 25   (12 Msps: GLONASS default)     32000 = 40 x 32 x 25   (16 Msps)
//
// Replaces the PRN x Doppler x non-coherent-block loop of
// GPS/GPS_L1CA/include/acquisition.m:155-200 (and the same loop of the other variant-A folders).
// The 2N-point transforms the reference does with MATLAB's fft/ifft are computed exactly at that
// length (no power-of-two padding, so the code-phase index space 1..2N is the reference's own).
//
//   rows   R = RA x RB with gcd(RA, RB) = 1: Good-Thomas prime-factor mapping, no twiddles:
//          row time index m = (RB*a + RA*b) mod R, row frequency index addressed by residues.
//   cols   C against R: prime-factor mapping again when gcd(C, R) = 1 (32736: a plain 3-D DFT, no
//          twiddle anywhere); otherwise one Cooley-Tukey twiddle w_L^(j1*m) between the passes,
//          read from a [C][R] table laid out like the data.
//
//   forward  (wipe-off + FFT, PRN independent, acquisition.m:169-183):
//      fwd_cols : one thread per row position: C gathered int8 I/Q samples, carrier wipe-off with a
//                 64-bit fixed-point phase, C-point DFT in registers
//      fwd_rows : one warp per row: RA-point DFTs on RB lanes, shared-memory transpose, RB-point
//                 DFTs on RA lanes
//      spectrum layout X[j1][kb][ka]; the replica spectra use the same layout, so no index map is
//      ever applied in the frequency domain.
//   inverse  (acquisition.m:186-190):
//      inv_rows : one warp per row: load X*conj(FFT(code))/L, RB-point then RA-point inverse DFTs
//      inv_cols : one thread per row position: for each non-coherent block C-point inverse DFT,
//                 |.|, accumulate in registers; after the last block the running maximum /
//                 first arg-max over the thread's C code phases.
//   `results(freqBin, :)` (acquisition.m:162,190) is therefore never written to memory.
#include "acq.h"
#include "common.cuh"
#include "fft_codelets.cuh"
#include "acq_plan.cuh"

namespace gc {

namespace {

constexpr int kRowWarps = 8;     // warps (= rows in flight) per CTA in the forward row kernel

// ------------------------------------------------------------------ column pass (forward)
// grid (ceil(R/128), nRows), block 128; thread = row position p.
// MODE 0: IF samples with carrier wipe-off; MODE 1: code table.
template <class P, int MODE>
__global__ void __launch_bounds__(128)
fwd_cols_kernel(FwdColsParams p)
{
    constexpr int C = P::C, R = P::R, L = P::L;
    const int pp = blockIdx.x * 128 + threadIdx.x;
    if (pp >= R) return;
    const int row = blockIdx.y;               // MODE 0: km = k*nonCoh + m ; MODE 1: replica slot
    const int base = P::row_index(pp);
    float2 x[C];
    if (MODE == 0) {
        const int k = row / p.nonCoh, m = row % p.nonCoh;
        const uint64_t dphi = p.dphi[k];
        const long long x0 = p.winStart + (long long)m * p.N;                       // window x((m-1)N+1 : (m+1)N)
#pragma unroll
        for (int n1 = 0; n1 < C; ++n1) {
            const int n = P::index(n1, base);
            const short2 s = p.rec.load(x0 + n);
            float sn, cs;
            fix_sincos(dphi * (uint64_t)n, &sn, &cs);          // exp(-1i*f*phasePoints(n)), :172
            const float I = p.swapIQ ? (float)s.y : (float)s.x, Q = p.swapIQ ? (float)s.x : (float)s.y;
            x[n1] = make_float2(fmaf(cs, I, sn * Q), fmaf(cs, Q, -sn * I));   // :180-181
        }
    } else {
        const int8_t* code = p.codeTab + (size_t)row * p.N;    // caCodesTable, zero padded to 2N (:160)
#pragma unroll
        for (int n1 = 0; n1 < C; ++n1) {
            const int n = P::index(n1, base);
            x[n1] = make_float2(n < p.N ? (float)code[n] : 0.f, 0.f);
        }
    }
    float2* dst = p.out + (size_t)row * L + pp;
    const float2* tw = p.tw + pp;
    codelet::dft<C, false>(x, [&](int k1, float re, float im) {
        float2 v = make_float2(re, im);
        if (!P::kPfa) v = cmul(v, __ldg(tw + (size_t)k1 * R));   // w_L^(j1*m)
        dst[(size_t)k1 * R] = v;
    });
}

// ------------------------------------------------------------------ row pass, forward
// One warp per row (fixed j1): RA x RB two-dimensional DFT over (a, b) in place.
// in : element (a, b) at a*RB + b       out: element (ka, kb) at kb*RA + ka
template <class P>
__global__ void __launch_bounds__(kRowWarps * 32)
fwd_rows_kernel(RowsParams p)
{
    constexpr int RA = P::RA, RB = P::RB, R = P::R;
    constexpr int kPitchF = RA + 1;                              // conflict-free pitch of the RB x RA exchange
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* s_x = reinterpret_cast<float2*>(smem_raw) + warp * (RB * kPitchF);
    const long long row = (long long)blockIdx.x * kRowWarps + warp;
    if (row >= p.nRows) return;
    float2* io = p.X + row * R;
    if (lane < RB) {                                             // lane = b: DFT-RA over a
        float2 v[RA];
#pragma unroll
        for (int a = 0; a < RA; ++a) v[a] = io[a * RB + lane];
        codelet::dft<RA, false>(v, [&](int ka, float re, float im) { s_x[lane * kPitchF + ka] = make_float2(re, im); });
    }
    __syncwarp();
    {                                                            // lane = ka: DFT-RB over b
        float2 u[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) u[b] = s_x[b * kPitchF + lane];
        codelet::dft<RB, false>(u, [&](int kb, float re, float im) { io[kb * RA + lane] = make_float2(re, im); });
    }
}

// ------------------------------------------------------------------ row pass, inverse (dominant kernel)
// One warp per row of one (PRN, bin, block): X .* Cc, inverse RB x RA DFT over (kb, ka) -> (tb, ta).
// in : element (ka, kb) at kb*RA + ka       out: element (ta, tb) at ta*RB + tb
// grid: x = j1 (C rows), y = PRN group, z = bin * mGroups + block group, so that CTAs scheduled
// together share one X slice (L1/L2 hits) while the replica spectra stay L2 resident.
// LOOP: a warp walks zLoop consecutive z positions (block groups, then bins) of its (SV, row): the CTA launch and the
// first-touch latencies of a 3 us row are paid once per zLoop rows and the replica row stays in L1 between them.  Nothing
// thread-dependent is carried from one row to the next (the thread index is re-read behind an opaque barrier, so the index
// arithmetic is redone per row and its registers are free during the codelets).
template <class P, int WARPS, int MINB, bool LOOP>
__global__ void __launch_bounds__(WARPS * 32, MINB)
inv_rows_kernel(RowsParams p)
{
    constexpr int C = P::C, RA = P::RA, RB = P::RB, R = P::R;
    constexpr int kPitchI = RB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
#pragma unroll 1
    for (int it = 0; it < (LOOP ? p.zLoop : 1); ++it) {
    int tid = threadIdx.x;
    if (LOOP) asm volatile("" : "+r"(tid));
    const int k1 = blockIdx.x, pg = blockIdx.y;
    const int vz = LOOP ? blockIdx.z * p.zLoop + it : blockIdx.z;
    if (LOOP && vz >= p.zTotal) break;
    const int M = p.nonCoh * p.nRep;                             // transforms per (SV, bin): blocks x replicas
    const int mGroups = p.mGroups;                               // ceil(M / mPerCta), from the launcher
    const int kz = vz / mGroups;
    const int mg = vz - kz * mGroups;
    const int lane = tid & 31, warp = tid >> 5;
    float2* s_x = reinterpret_cast<float2*>(smem_raw) + warp * (RA * kPitchI);
    // (a warp does this index arithmetic once per row: the run-time divisions are kept to the cases that need them)
    int k = kz;
    int wl = warp;
    if (p.binPerCta > 1) {                                       // several bins per CTA (variants B / C)
        const int wpb = p.prnPerCta * p.mPerCta;                 // warps per bin
        const int kb = warp / wpb;
        k = k * p.binPerCta + kb;
        wl = warp - kb * wpb;
    }
    const int pq = (p.mPerCta == 1) ? wl : wl / p.mPerCta;
    const int pi = pg * p.prnPerCta + pq;                       // list slot within this launch's chunk
    const int mv = mg * p.mPerCta + (wl - pq * p.mPerCta);
    if (pi >= p.nPrnChunk || mv >= M || k >= p.nBins) continue;
    const int m = (p.nRep == 1) ? mv : mv / p.nRep, r = mv - m * p.nRep;   // block, replica (data / pilot, GPS_L5C acquisition.m:171-175)
    // circshift(IQfreqDom, s) (BDS/B1I acquisition.m:88, GPS_L2C :73, B1C :203): product element j takes spectrum element j - s.
    // With j = j1 + C*j2 (row j1, in-row frequency j2 held by its residues mod RA and mod RB) that is row (j1 - s) mod C
    // and j2 - floor-part, i.e. a fixed source row and a circular shift of both residues.
    // Variant A whose Doppler step is a whole number of FFT bins (acqSearchStep * 2N / fs integer: 500 Hz at 16.368, 18 and 12 Msps)
    // uses the same addressing: bin k of the grid is the spectrum of bin 0 shifted by k steps, so only nonCoh forward transforms
    // exist (identity 2).  Prime-factor plans hold j by its three residues: each is rotated by the shift.
    int xrow = (p.bin0 + k) * p.nonCoh + m, srow = k1, sa = 0, sb = 0;
    const int grow = (p.slotGroup != nullptr) ? p.slotGroup[p.prnSlot0 + pi] * p.groupRows : 0;     // this SV's carrier grid
    if (p.binMap != nullptr) {
        const int2 bm = p.binMap[(p.binMapSlotStride ? (p.prnSlot0 + pi) * p.binMapSlotStride : 0) + p.bin0 + k];
        xrow = bm.x * p.nonCoh + m;
        if constexpr (!P::kPfa) {
            const int s1 = bm.y % C;
            int s2 = bm.y / C;
            srow = k1 - s1;
            if (srow < 0) { srow += C; s2 += 1; }
            sa = s2 % RA; sb = s2 % RB;
        } else {
            srow = k1 - bm.y % C;
            if (srow < 0) srow += C;
            sa = bm.y % RA; sb = bm.y % RB;
        }
    }
    xrow += grow;
    const float2* src = p.X + ((size_t)xrow * C + srow) * R;
    const float2* mul = p.Cc + ((size_t)(p.prnList[p.prnSlot0 + pi] + r * p.repStride) * C + k1) * R;
    float2* dst = p.W + (((size_t)(pi * p.nBins + k) * M + mv) * C + k1) * R;
    {                                                            // lane = ka: DFT-RB over kb
        float2 u[RB];
#pragma unroll
        for (int kb = 0; kb < RB; ++kb) {                        // IQfreqDom .* caCodeFreqDom (:186)
            int ks = kb - sb;
            if (ks < 0) ks += RB;
            u[kb] = cmul(src[ks * RA + ((lane - sa) & (RA - 1))], __ldg(mul + kb * RA + lane));
        }
        codelet::dft<RB, true>(u, [&](int tb, float re, float im) { s_x[lane * kPitchI + tb] = make_float2(re, im); });
    }
    __syncwarp();
    if (lane < RB) {                                             // lane = tb: DFT-RA over ka
        float2 v[RA];
#pragma unroll
        for (int ka = 0; ka < RA; ++ka) v[ka] = s_x[ka * kPitchI + lane];
        const float2* tw = p.tw + (size_t)k1 * R + lane;
        codelet::dft<RA, true>(v, [&](int ta, float re, float im) {
            float2 t = make_float2(re, im);
            if (!P::kPfa) t = cmul_conj(t, __ldg(tw + ta * RB));  // conj(w_L^(j1*tau2))
            __stcs(dst + ta * RB + lane, t);
        });
    }
    if (LOOP) __syncwarp();                                      // the next row overwrites s_x
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
// grid (ceil(R/128), nBins, nPrnChunk), block 128: thread = row position (ta, tb) of one (PRN, bin).
// p.persist: a fixed number of CTAs walks the (tile, bin, SV) items - used when the pass runs on a high-priority stream next to
// the row pass of the following chunk (GC_ACQ_OVERLAP), so that it takes one CTA slot per SM instead of every slot that frees up.
template <class P>
__global__ void __launch_bounds__(128)
inv_cols_kernel(InvColsParams p)
{
    constexpr int C = P::C, R = P::R, L = P::L;
    constexpr int kTiles = (R + 127) / 128;
    const long long nItems = p.persist ? (long long)kTiles * p.nBins * p.nPrnChunk : 1;
    for (long long item = p.persist ? blockIdx.x : 0; item < nItems; item += gridDim.x) {
    const int bx = p.persist ? (int)(item % kTiles) : blockIdx.x;
    const int pp = bx * 128 + threadIdx.x;
    const int k = p.persist ? (int)((item / kTiles) % p.nBins) : blockIdx.y, pi = p.persist ? (int)(item / ((long long)kTiles * p.nBins)) : blockIdx.z;
    float acc[C];
#pragma unroll
    for (int i = 0; i < C; ++i) acc[i] = 0.f;
    if (pp < R) {
        const float2* base = p.W + ((size_t)(pi * p.nBins + k) * p.nonCoh) * L + pp;
        for (int m = 0; m < p.nonCoh; ++m) {                    // acquisition.m:175
            float2 x[C];
#pragma unroll
            for (int k1 = 0; k1 < C; ++k1) x[k1] = __ldcs(base + (size_t)m * L + (size_t)k1 * R);
            const float wm = p.weighted ? ((m & 1) ? p.w1 : p.w0) : 1.f;
            codelet::dft<C, true>(x, [&](int t1, float re, float im) {
                acc[t1] = fmaf(wm, cabs_fast(re, im), acc[t1]);        // abs(ifft(.)) summed over blocks (:188-190)
            });
        }
        if (p.weighted)
#pragma unroll
            for (int i = 0; i < C; ++i) acc[i] *= p.wScale;
    }
    // running maximum with MATLAB first-index tie breaking (smaller code phase wins)
    float best = -1.f;
    int bidx = 0x7fffffff;
    if (pp < R) {
        const int rest = P::row_index(pp);
#pragma unroll
        for (int t1 = 0; t1 < C; ++t1) {
            const int idx = P::index(t1, rest);                 // code phase (lag) of this output
            if (p.magOut) p.magOut[(size_t)(pi * p.nBins + k) * L + idx] = acc[t1];
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
        const size_t o = ((size_t)(p.prnSlot0 + pi) * (p.nBinsTotal ? p.nBinsTotal : p.nBins) + p.bin0 + k) * kTiles + bx;
        p.partMax[o] = best;
        p.partIdx[o] = bidx;
    }
    if (p.persist) __syncthreads();                              // s_b / s_i are reused by the next item
    }
}

// ------------------------------------------------------------------ column passes for long columns (C = C1 x C2)
// Cooley-Tukey inside the pass: input index n1 = C2*a + b, output index k1 = ka + C1*kb.
//   phase 1: thread (b, col): C1-point DFT over a, times w_C^(ka*b), into shared memory Y[ka][b][col]
//   phase 2: thread (ka, col): C2-point DFT over b -> outputs k1 = ka + C1*kb
// A CTA handles TC adjacent row positions (columns), so global accesses are TC*8-byte runs.
template <class P> struct BigGeo {
    static constexpr int C1 = P::C1, C2 = P::C2;
    // 16 columns per CTA (128-byte runs) and several CTAs per SM: one CTA's loads overlap the other's DFT phases
    // (one 640-thread CTA per SM measured 10-25 % slower: its load, DFT and barrier phases serialise)
    static constexpr int TC = 16;
    // three CTAs per SM (64 registers) where both phases are at most 20 points (E1 36 x 81 grid 6.0 -> 5.4 ms, L2C 43 -> 41 ms);
    // a 25-point phase spills at that budget (B1C 46 -> 58 ms) and stays at two
    static constexpr int kMinCtas = (C1 > C2 ? C1 : C2) <= 20 ? 3 : 2;
    static constexpr int NT = (C1 > C2 ? C1 : C2) * TC;                        // threads: max of the two phases
    static constexpr size_t kSmem = sizeof(float2) * P::C * TC;
    static constexpr size_t kSmemInv = kSmem + sizeof(float2) * P::C;          // + the w_C^(ta*beta) table of the inverse pass
};

// w_C^(+-ka*b): computed once per thread (one sincospif)
__device__ __forceinline__ float2 unit_root(int num, int den, bool inverse)
{
    float s, c;
    sincospif(2.0f * (float)num / (float)den, &s, &c);
    return make_float2(c, inverse ? s : -s);
}

template <class P, int MODE>
__global__ void __launch_bounds__(BigGeo<P>::NT, BigGeo<P>::kMinCtas)
fwd_cols_big_kernel(FwdColsParams p)
{
    using G = BigGeo<P>;
    constexpr int C = P::C, C1 = P::C1, C2 = P::C2, R = P::R, L = P::L, TC = G::TC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Y = reinterpret_cast<float2*>(smem_raw);            // [ka][b][col]
    const int row = blockIdx.y;
    const int col = threadIdx.x % TC, q = threadIdx.x / TC;
    const int pp = blockIdx.x * TC + col;
    const bool in = pp < R;
    const int base = in ? P::row_index(pp) : 0;
    if (q < C2 && in) {                                         // phase 1: q = b
        float2 x[C1];
        if (MODE == 0) {
            const int k = row / p.nonCoh, m = row % p.nonCoh;
            const uint64_t dphi = p.dphi[k];
            const long long x0 = p.winStart + (long long)m * p.N;
#pragma unroll
            for (int a = 0; a < C1; ++a) {
                const int n = P::index(C2 * a + q, base);
                const short2 sv = p.rec.load(x0 + n);
                float sn, cs;
                fix_sincos(dphi * (uint64_t)n, &sn, &cs);
                const float I = p.swapIQ ? (float)sv.y : (float)sv.x, Q = p.swapIQ ? (float)sv.x : (float)sv.y;
                x[a] = make_float2(fmaf(cs, I, sn * Q), fmaf(cs, Q, -sn * I));
            }
        } else {
            const int8_t* code = p.codeTab + (size_t)row * p.N;
#pragma unroll
            for (int a = 0; a < C1; ++a) {
                const int n = P::index(C2 * a + q, base);
                x[a] = make_float2(n < p.N ? (float)code[n] : 0.f, 0.f);
            }
        }
        codelet::dft<C1, false>(x, [&](int ka, float re, float im) {
            Y[(ka * C2 + q) * TC + col] = cmul(make_float2(re, im), unit_root(ka * q, C, false));
        });
    }
    __syncthreads();
    if (q < C1 && in) {                                         // phase 2: q = ka
        float2 y[C2];
#pragma unroll
        for (int b = 0; b < C2; ++b) y[b] = Y[(q * C2 + b) * TC + col];
        float2* dst = p.out + (size_t)row * L + pp;
        const float2* tw = p.tw + pp;
        codelet::dft<C2, false>(y, [&](int kb, float re, float im) {
            const int k1 = q + C1 * kb;
            dst[(size_t)k1 * R] = cmul(make_float2(re, im), __ldg(tw + (size_t)k1 * R));      // w_L^(j1*m)
        });
    }
}

// inverse column pass + |.| + sum over blocks/replicas + max: grid (ceil(R/TC), nBins, nPrnChunk)
// SINGLE: one transform per cell (variant B: GPS L2C, BDS B1I; nothing to accumulate, so no accumulator array stays live across the
// loads of the next pass and the maximum is taken as the outputs appear); MAGOUT: the magnitudes are also written in natural lag
// order (corrVec of variant B's winner rows).  The w_C^(-ta*beta) table comes from global memory (InvColsParams::colTw, float64
// on the host) instead of one sincospif per entry and CTA.
template <class P, bool SINGLE, bool MAGOUT>
__global__ void __launch_bounds__(BigGeo<P>::NT, BigGeo<P>::kMinCtas)
inv_cols_big_kernel(InvColsParams p)
{
    using G = BigGeo<P>;
    constexpr int C = P::C, C1 = P::C1, C2 = P::C2, R = P::R, L = P::L, TC = G::TC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Y = reinterpret_cast<float2*>(smem_raw);            // [ta][beta][col]
    float2* TW = Y + C * TC;                                     // [ta][beta] = w_C^(-ta*beta)
    for (int i = threadIdx.x; i < C; i += G::NT) TW[i] = __ldg(p.colTw + i);
    const int col = threadIdx.x % TC, q = threadIdx.x / TC;
    const int pp = blockIdx.x * TC + col;
    const bool in = pp < R;
    const int k = blockIdx.y, pi = blockIdx.z;
    const float2* base = p.W + ((size_t)(pi * p.nBins + k) * p.nonCoh) * L + pp;
    const bool weighted = !SINGLE && p.weighted;
    float best = -1.f;
    int btb = 0x7fffffff;                                        // output tb of this thread's maximum (lag index grows with tb)
    float acc[SINGLE ? 1 : C2];
    if (!SINGLE) {
#pragma unroll
        for (int i = 0; i < C2; ++i) acc[i] = 0.f;
    }
    float* mag = MAGOUT ? p.magOut + (size_t)(pi * p.nBins + k) * L + P::index(q, in ? P::row_index(pp) : 0) : nullptr;
    const int nPass = SINGLE ? 1 : p.nonCoh;
    const bool ph1 = q < C2 && in, ph2 = q < C1 && in;          // (every __syncthreads below is reached by the whole CTA)
    if (!SINGLE) __syncthreads();                                // the twiddle table
    for (int m = 0; m < nPass; ++m) {
        if constexpr (SINGLE) {                                  // one short pass per CTA: the loads go out before the table barrier
            float2 x[C1];
            if (ph1) {                                          // phase 1: q = beta, input k1 = C2*alpha + beta
                const float2* src = base + (size_t)q * R;
#pragma unroll
                for (int a = 0; a < C1; ++a) x[a] = __ldcs(src + (size_t)(C2 * a) * R);
            }
            __syncthreads();
            if (ph1)
                codelet::dft<C1, true>(x, [&](int ta, float re, float im) {
                    Y[(ta * C2 + q) * TC + col] = cmul(make_float2(re, im), TW[ta * C2 + q]);
                });
        } else if (ph1) {
            float2 x[C1];
            const float2* src = base + (size_t)m * L + (size_t)q * R;
#pragma unroll
            for (int a = 0; a < C1; ++a) x[a] = __ldcs(src + (size_t)(C2 * a) * R);
            codelet::dft<C1, true>(x, [&](int ta, float re, float im) {
                Y[(ta * C2 + q) * TC + col] = cmul(make_float2(re, im), TW[ta * C2 + q]);
            });
        }
        __syncthreads();
        if (ph2) {                                              // phase 2: q = ta, outputs t1 = ta + C1*tb
            float2 y[C2];
#pragma unroll
            for (int b = 0; b < C2; ++b) y[b] = Y[(q * C2 + b) * TC + col];
            if (SINGLE) {
                codelet::dft<C2, true>(y, [&](int tb, float re, float im) {
                    const float v = cabs_fast(re, im);
                    if (MAGOUT) mag[(size_t)(C1 * tb) * R] = v;
                    const bool take = v > best || (v == best && tb < btb);
                    best = take ? v : best; btb = take ? tb : btb;
                });
            } else {
                const float wm = weighted ? ((m & 1) ? p.w1 : p.w0) : 1.f;
                codelet::dft<C2, true>(y, [&](int tb, float re, float im) { acc[tb] = fmaf(wm, cabs_fast(re, im), acc[tb]); });
            }
        }
        if (m + 1 < nPass) __syncthreads();
    }
    int bidx = 0x7fffffff;
    if (ph2) {
        if (!SINGLE) {
#pragma unroll
            for (int tb = 0; tb < C2; ++tb) {
                if (weighted) acc[tb] *= p.wScale;
                if (MAGOUT) mag[(size_t)(C1 * tb) * R] = acc[tb];
                best = fmaxf(best, acc[tb]);
            }
#pragma unroll
            for (int tb = C2 - 1; tb >= 0; --tb) btb = (acc[tb] == best) ? tb : btb;      // first (smallest) lag holding the maximum
        }
        bidx = P::index(q + C1 * btb, P::row_index(pp));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    __shared__ float s_b[32];
    __shared__ int s_i[32];
    if ((threadIdx.x & 31) == 0) { s_b[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_b[w] > best || (s_b[w] == best && s_i[w] < bidx)) { best = s_b[w]; bidx = s_i[w]; }
        const size_t o = ((size_t)(p.prnSlot0 + pi) * (p.nBinsTotal ? p.nBinsTotal : p.nBins) + p.bin0 + k) * gridDim.x + blockIdx.x;
        p.partMax[o] = best;
        p.partIdx[o] = bidx;
    }
}

// ------------------------------------------------------------------ correlation stage as ONE persistent kernel (experimental)
// The split stage writes the inverse-row output W to HBM and reads it back (2 x 8 B per transform point: the bound of the
// two-kernel design).  Here CTAs pull work items from one ordered queue: "rows" items (cell, row j1, group of 5 transforms) write
// into a ring of nSlots cells, "cols" items (cell, tile of 160 row positions) consume a cell `lag` cells later, so that W lives
// in the 126 MB L2 only.  Every item's prerequisites sit EARLIER in the queue and items are taken in order, so the oldest
// unfinished item never waits on a younger one (no deadlock); waits are bounded and raise an abort flag instead of hanging.
__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <class P>
__global__ void __launch_bounds__(160, 4)
corr_queue_kernel(QueueParams p)
{
    constexpr int C = P::C, RA = P::RA, RB = P::RB, R = P::R, L = P::L;
    constexpr int kPitchI = RB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_item, s_ok;
    __shared__ float s_b[5];
    __shared__ int s_i[5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* s_x = reinterpret_cast<float2*>(smem_raw) + warp * (RA * kPitchI);
    const int M = p.nonCoh * p.nRep;
    const int mGroups = (M + 4) / 5;
    // a column item covers one tile of 160 row positions and one GROUP of 5 transforms (short items: the ring slot of a cell is
    // free again only when its last column item has finished); the group partial sums meet in an L2-resident buffer and the
    // item that finishes a tile last adds them in group order - the sum does not depend on who that is
    const int rowItems = C * mGroups, colItems = p.parts * mGroups, stepItems = rowItems + colItems;
    const int nCells = p.nPrn * p.nBins;
    const int total = (nCells + p.lag) * stepItems;
    int* rowsDone = p.ctrl + 2;
    int* colsDone = rowsDone + nCells;
    int* tileDone = colsDone + nCells;
    __shared__ int s_last;
    auto wait_for = [&](const int* ctr, int want) {             // thread 0 spins (bounded), then everybody knows the outcome
        if (tid == 0) {
            int ok = 1;
            long long spins = 0;
            while (ld_acquire(ctr) < want) {
                if (++spins > 4000000 || ld_acquire(p.ctrl + 1) != 0) { atomicExch(p.ctrl + 1, 1); ok = 0; break; }
                __nanosleep(64);
            }
            s_ok = ok;
        }
        __syncthreads();
        return s_ok != 0;
    };
    for (;;) {
        __syncthreads();                                         // s_item / s_ok / shared buffers of the previous item are free
        if (tid == 0) s_item = atomicAdd(p.ctrl, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        const int step = item / stepItems, r = item - step * stepItems;
        if (r < rowItems) {
            // ---- rows item: X .* Cc and the inverse RB x RA row DFTs of 5 transforms of cell `step`, row j1
            const int c = step;
            if (c >= nCells) continue;
            if (c >= p.nSlots && !wait_for(colsDone + (c - p.nSlots), p.parts)) break;   // the ring slot is free again
            const int k1 = r / mGroups, mg = r - k1 * mGroups;
            const int mv = mg * 5 + warp;
            const int k = c / p.nPrn, pi = c - k * p.nPrn;
            if (mv < M) {
                const int m = (p.nRep == 1) ? mv : mv / p.nRep, rep = mv - m * p.nRep;
                const float2* src = p.X + ((size_t)(k * p.nonCoh + m) * C + k1) * R;
                const float2* mul = p.Cc + ((size_t)(p.prnList[p.prnSlot0 + pi] + rep * p.repStride) * C + k1) * R;
                float2* dst = p.W + (((size_t)(c % p.nSlots) * M + mv) * C + k1) * R;
                {
                    float2 u[RB];
#pragma unroll
                    for (int kb = 0; kb < RB; ++kb) u[kb] = cmul(__ldcs(src + kb * RA + lane), __ldg(mul + kb * RA + lane));
                    codelet::dft<RB, true>(u, [&](int tb, float re, float im) { s_x[lane * kPitchI + tb] = make_float2(re, im); });
                }
                __syncwarp();
                if (lane < RB) {
                    float2 v[RA];
#pragma unroll
                    for (int ka = 0; ka < RA; ++ka) v[ka] = s_x[ka * kPitchI + lane];
                    const float2* tw = p.tw + (size_t)k1 * R + lane;
                    codelet::dft<RA, true>(v, [&](int ta, float re, float im) {
                        float2 t = make_float2(re, im);
                        if (!P::kPfa) t = cmul_conj(t, __ldg(tw + ta * RB));
                        __stcg(dst + ta * RB + lane, t);          // L2 only: the column item of this cell reads it back from there
                    });
                }
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(rowsDone + c, 1);
        } else {
            // ---- cols item: inverse column DFTs, |.|, sum over the cell's transforms, tile maximum
            const int c = step - p.lag;
            if (c < 0 || c >= nCells) continue;
            if (!wait_for(rowsDone + c, rowItems)) break;
            const int tq = r - rowItems;
            const int t = tq / mGroups, grp = tq - t * mGroups;
            const int k = c / p.nPrn, pi = c - k * p.nPrn;
            const int pp = t * 160 + tid;
            const int slot = c % p.nSlots;
            float acc[C];
#pragma unroll
            for (int i = 0; i < C; ++i) acc[i] = 0.f;
            if (pp < R) {
                const float2* base = p.W + ((size_t)slot * M) * L + pp;
                const int m1 = min(M, grp * 5 + 5);
                for (int m = grp * 5; m < m1; ++m) {
                    float2 x[C];
#pragma unroll
                    for (int k1 = 0; k1 < C; ++k1) x[k1] = __ldcg(base + (size_t)m * L + (size_t)k1 * R);
                    codelet::dft<C, true>(x, [&](int t1, float re, float im) { acc[t1] += cabs_fast(re, im); });
                }
            }
            bool last = true;
            if (mGroups > 1) {
                float* pb = p.partial + ((size_t)slot * mGroups) * L;
                if (pp < R)
#pragma unroll
                    for (int t1 = 0; t1 < C; ++t1) __stcg(pb + ((size_t)grp * C + t1) * R + pp, acc[t1]);
                __threadfence();
                __syncthreads();
                if (tid == 0) { s_last = (atomicAdd(tileDone + (size_t)c * p.parts + t, 1) == mGroups - 1); __threadfence(); }
                __syncthreads();
                last = s_last != 0;
                if (last && pp < R) {
#pragma unroll
                    for (int t1 = 0; t1 < C; ++t1) {
                        float a = 0.f;
                        for (int g = 0; g < mGroups; ++g) a += __ldcg(pb + ((size_t)g * C + t1) * R + pp);
                        acc[t1] = a;
                    }
                }
            }
            if (last) {
                float best = -1.f;
                int bidx = 0x7fffffff;
                if (pp < R) {
                    const int rest = P::row_index(pp);
#pragma unroll
                    for (int t1 = 0; t1 < C; ++t1) {
                        const int idx = P::index(t1, rest);
                        if (acc[t1] > best || (acc[t1] == best && idx < bidx)) { best = acc[t1]; bidx = idx; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ob = __shfl_down_sync(0xffffffffu, best, o);
                    const int oi = __shfl_down_sync(0xffffffffu, bidx, o);
                    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
                }
                if (lane == 0) { s_b[warp] = best; s_i[warp] = bidx; }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 1; w < 5; ++w)
                        if (s_b[w] > best || (s_b[w] == best && s_i[w] < bidx)) { best = s_b[w]; bidx = s_i[w]; }
                    const size_t o = ((size_t)(p.prnSlot0 + pi) * p.nBins + k) * p.parts + t;
                    p.partMax[o] = best;
                    p.partIdx[o] = bidx;
                    __threadfence();
                    atomicAdd(colsDone + c, 1);
                }
            }
        }
    }
}

// ------------------------------------------------------------------ host side, per plan
template <class P>
struct Launch {
    static cudaError_t fwd_cols(const FwdColsParams& p, int nRows, bool codeMode, cudaStream_t s)
    {
        if constexpr (P::kBig) {
            using G = BigGeo<P>;
            dim3 grid((P::R + G::TC - 1) / G::TC, nRows);
            cudaError_t e = cudaFuncSetAttribute(codeMode ? fwd_cols_big_kernel<P, 1> : fwd_cols_big_kernel<P, 0>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem);
            if (e != cudaSuccess) return e;
            if (codeMode) fwd_cols_big_kernel<P, 1><<<grid, G::NT, G::kSmem, s>>>(p);
            else fwd_cols_big_kernel<P, 0><<<grid, G::NT, G::kSmem, s>>>(p);
            return cudaGetLastError();
        } else {
            dim3 grid((P::R + 127) / 128, nRows);
            if (codeMode) fwd_cols_kernel<P, 1><<<grid, 128, 0, s>>>(p);
            else fwd_cols_kernel<P, 0><<<grid, 128, 0, s>>>(p);
            return cudaGetLastError();
        }
    }
    static cudaError_t fwd_rows(const RowsParams& p, cudaStream_t s)
    {
        const int smem = (int)(sizeof(float2) * kRowWarps * P::RB * (P::RA + 1));
        cudaError_t e = cudaFuncSetAttribute(fwd_rows_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)((p.nRows + kRowWarps - 1) / kRowWarps);
        fwd_rows_kernel<P><<<grid, kRowWarps * 32, smem, s>>>(p);
        return cudaGetLastError();
    }
    template <int WARPS, int MINB, bool LOOP>
    static cudaError_t inv_rows_t(const RowsParams& p, cudaStream_t s)
    {
        const int smem = (int)(sizeof(float2) * WARPS * P::RA * P::RB);
        cudaError_t e = cudaFuncSetAttribute(inv_rows_kernel<P, WARPS, MINB, LOOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        const int pGroups = (p.nPrnChunk + p.prnPerCta - 1) / p.prnPerCta;
        const int bpc = p.binPerCta > 1 ? p.binPerCta : 1;
        RowsParams q = p;
        q.mGroups = (p.nonCoh * p.nRep + p.mPerCta - 1) / p.mPerCta;
        q.zTotal = ((p.nBins + bpc - 1) / bpc) * q.mGroups;
        q.zLoop = LOOP ? p.zLoop : 1;
        if (LOOP) {
            // never fewer than ~12 waves of CTAs: a small share of the grid (a few SVs per GPU of a multi-GPU run, GLONASS's
            // small grids) keeps short CTAs so that its last wave stays a small part of the launch
            const long long ctas1 = (long long)P::C * pGroups * q.zTotal;
            const long long cap = ctas1 / (148LL * MINB * 12);
            if (q.zLoop > cap) q.zLoop = (int)(cap > 1 ? cap : 1);
        }
        dim3 grid(P::C, pGroups, (q.zTotal + q.zLoop - 1) / q.zLoop);
        inv_rows_kernel<P, WARPS, MINB, LOOP><<<grid, WARPS * 32, smem, s>>>(q);
        return cudaGetLastError();
    }
    // p.prnPerCta * p.mPerCta warps per CTA: 5 (96 registers, 20 warps/SM) or 8 (128 registers, 16 warps/SM)
    static cudaError_t inv_rows(const RowsParams& p, cudaStream_t s)
    {
        const bool eight = (p.binPerCta > 1 ? p.binPerCta : 1) * p.prnPerCta * p.mPerCta == 8;
        if (p.zLoop > 1) return eight ? inv_rows_t<8, 2, true>(p, s) : inv_rows_t<5, 4, true>(p, s);
        return eight ? inv_rows_t<8, 2, false>(p, s) : inv_rows_t<5, 4, false>(p, s);
    }
    static cudaError_t corr_queue(const QueueParams& p, cudaStream_t s)
    {
        if constexpr (P::kBig) return cudaErrorNotSupported;
        else {
            const int smem = (int)(sizeof(float2) * 5 * P::RA * P::RB);
            cudaError_t e = cudaFuncSetAttribute(corr_queue_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            int perSm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, corr_queue_kernel<P>, 160, smem);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            corr_queue_kernel<P><<<sms * (perSm > 0 ? perSm : 1), 160, smem, s>>>(p);   // all CTAs resident: the queue order is the schedule
            return cudaGetLastError();
        }
    }
    static cudaError_t inv_cols(const InvColsParams& p, cudaStream_t s)
    {
        if constexpr (P::kBig) {
            using G = BigGeo<P>;
            dim3 grid((P::R + G::TC - 1) / G::TC, p.nBins, p.nPrnChunk);
            if (!p.colTw) return cudaErrorInvalidValue;
            auto go = [&](auto kern) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmemInv);
                if (e != cudaSuccess) return e;
                kern<<<grid, G::NT, G::kSmemInv, s>>>(p);
                return cudaGetLastError();
            };
            const bool single = p.nonCoh == 1;
            if (p.magOut) return single ? go(inv_cols_big_kernel<P, true, true>) : go(inv_cols_big_kernel<P, false, true>);
            return single ? go(inv_cols_big_kernel<P, true, false>) : go(inv_cols_big_kernel<P, false, false>);
        } else {
            dim3 grid((P::R + 127) / 128, p.nBins, p.nPrnChunk);
            if (p.persist > 0) grid = dim3((unsigned)p.persist, 1, 1);
            inv_cols_kernel<P><<<grid, 128, 0, s>>>(p);
            return cudaGetLastError();
        }
    }
};

template <class P>
void fill_info(FusedPlanInfo* o)
{
    o->L = P::L; o->C = P::C; o->RA = P::RA; o->RB = P::RB; o->R = P::R; o->pfa = P::kPfa ? 1 : 0;
    o->C1 = P::C1; o->C2 = P::C2;
    if constexpr (P::kBig) o->parts = (P::R + BigGeo<P>::TC - 1) / BigGeo<P>::TC;
    else o->parts = (P::R + 127) / 128;
}

}  // namespace

bool fused_plan_info(int L, FusedPlanInfo* o)
{
    switch (L) {
        case P32736::L: fill_info<P32736>(o); return true;
        case P36000::L: fill_info<P36000>(o); return true;
        case P24000::L: fill_info<P24000>(o); return true;
        case P32000::L: fill_info<P32000>(o); return true;
        case P40000::L: fill_info<P40000>(o); return true;
        case P160000::L: fill_info<P160000>(o); return true;
        case P144000::L: fill_info<P144000>(o); return true;
        case P320000::L: fill_info<P320000>(o); return true;
        case P360000::L: fill_info<P360000>(o); return true;
        case P72000::L: fill_info<P72000>(o); return true;
        default: return false;
    }
}

cudaError_t launch_fwd_cols(int L, const FwdColsParams& p, int nRows, bool codeMode, cudaStream_t s) { GC_PLAN_DISPATCH(L, fwd_cols(p, nRows, codeMode, s)) }
cudaError_t launch_fwd_rows(int L, const RowsParams& p, cudaStream_t s) { GC_PLAN_DISPATCH(L, fwd_rows(p, s)) }
cudaError_t launch_inv_rows(int L, const RowsParams& p, cudaStream_t s) { GC_PLAN_DISPATCH(L, inv_rows(p, s)) }
cudaError_t launch_inv_cols(int L, const InvColsParams& p, cudaStream_t s) { GC_PLAN_DISPATCH(L, inv_cols(p, s)) }
cudaError_t launch_corr_queue(int L, const QueueParams& p, cudaStream_t s) { GC_PLAN_DISPATCH(L, corr_queue(p, s)) }

cudaError_t launch_finish_replica(float2* Cc, size_t n, int L, cudaStream_t s)
{
    finish_replica_kernel<<<148 * 2, 256, 0, s>>>(Cc, n, 1.0f / (float)L);
    return cudaGetLastError();
}

}  // namespace gc
