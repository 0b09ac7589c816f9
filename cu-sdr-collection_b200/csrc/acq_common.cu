// Acquisition kernels shared by the fused and the generic path: input power, 2-D peak pick,
// fine-frequency search.  GPS/GPS_L1CA/include/acquisition.m:151, :196-200, :206-260.
#include "acq.h"
#include "common.cuh"

namespace gc {

namespace {

// sigPower = sqrt(var(longSignal(1:N)) * N), var of a complex vector with N-1 (acquisition.m:151).
// The samples are small integers, so the three sums are exact in 64-bit integers.
__global__ void __launch_bounds__(1024)
sig_power_kernel(Rec rec, long long winStart, int N, double* out)
{
    long long sI = 0, sQ = 0, s2 = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const short2 v = rec.load(winStart + i);
        sI += v.x; sQ += v.y; s2 += (long long)((int)v.x * v.x) + (long long)((int)v.y * v.y);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sI += __shfl_down_sync(0xffffffffu, sI, o);
        sQ += __shfl_down_sync(0xffffffffu, sQ, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    __shared__ long long sh[3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = sI; sh[1][warp] = sQ; sh[2][warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        sI = sQ = s2 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sI += sh[0][w]; sQ += sh[1][w]; s2 += sh[2][w]; }
        const double n = (double)N;
        const double var = ((double)s2 - ((double)sI * (double)sI + (double)sQ * (double)sQ) / n) / (n - 1.0);
        *out = sqrt(var * n);
    }
}

// [~,bin] = max(max(results,[],2)); [peak,codePhase] = max(max(results))  (acquisition.m:196-198)
// from the per-(bin, column-tile) partial maxima.  MATLAB's max returns the first maximal index:
// bin = first row holding the global maximum, codePhase = first column holding it.
__global__ void __launch_bounds__(32)
peak_select_kernel(const float* partMax, const int* partIdx, int nBins, int parts, PeakOut* out)
{
    const int slot = blockIdx.x, lane = threadIdx.x;
    float best = -1.f;      // global max value
    int bbin = 0x7fffffff;  // first row with it
    int bcol = 0x7fffffff;  // first column with it
    for (int k = lane; k < nBins; k += 32) {
        float rmax = -1.f; int ridx = 0x7fffffff;
        for (int q = 0; q < parts; ++q) {
            const size_t o = ((size_t)slot * nBins + k) * parts + q;
            const float v = partMax[o]; const int i = partIdx[o];
            if (v > rmax || (v == rmax && i < ridx)) { rmax = v; ridx = i; }
        }
        if (rmax > best) { best = rmax; bbin = k; bcol = ridx; }
        else if (rmax == best) { bbin = min(bbin, k); bcol = min(bcol, ridx); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int obin = __shfl_down_sync(0xffffffffu, bbin, o);
        const int ocol = __shfl_down_sync(0xffffffffu, bcol, o);
        if (ob > best) { best = ob; bbin = obin; bcol = ocol; }
        else if (ob == best) { bbin = min(bbin, obin); bcol = min(bcol, ocol); }
    }
    if (lane == 0) {
        out[slot].peak = (double)best;
        out[slot].bin = bbin + 1;
        out[slot].codePhase = bcol + 1;
    }
}

// ---- fine frequency search (acquisition.m:211-250) -----------------------------------------
// prep: sig40cm .* caCode40ms for one acquired PRN, stored as int16 pairs.
__global__ void fine_prep_kernel(FineParams p)
{
    short2* prod = p.prod;
    const int a = blockIdx.y;
    if (a >= *p.nAcqDev * p.nCodes) return;
    const long long total = (long long)p.nPeriods * p.N;
    const long long x0 = p.winStart + (p.codePhase[a] - 1);                                      // :221
    const int8_t* chips = p.chips + (size_t)p.chipRow[a] * p.codeLen;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
        const int c = chips[p.chipIdx[gi]];                      // caCode40ms (:215-218; GLO generateCAcode.m:110-116)
        short2 v = p.rec.load(x0 + gi);
        if (p.swapIQ) v = make_short2(v.y, v.x);
        prod[(size_t)a * total + gi] = make_short2((short)(v.x * c), (short)(v.y * c));
    }
}

constexpr int kFineBins = 8;   // fine bins handled per thread
// grid (nPeriods, nAcq, ceil(nFine/8)), block 256: sumPerCode(index) for 8 fine bins (:232-238)
__global__ void __launch_bounds__(256)
fine_sum_kernel(FineParams p)
{
    const short2* prod = p.prod;
    const int c = blockIdx.x, a = blockIdx.y, j0 = blockIdx.z * kFineBins;
    if (a >= *p.nAcqDev * p.nCodes) return;
    const long long total = (long long)p.nPeriods * p.N;
    const short2* x = prod + (size_t)a * total + (size_t)c * p.N;
    uint64_t dphi[kFineBins];
#pragma unroll
    for (int j = 0; j < kFineBins; ++j) dphi[j] = (j0 + j < p.nFine) ? p.dphi[a * p.nFine + j0 + j] : 0;
    float ar[kFineBins], ai[kFineBins];
#pragma unroll
    for (int j = 0; j < kFineBins; ++j) ar[j] = ai[j] = 0.f;
    // A thread visits samples n = tid, tid+256, ...: the carrier of bin j advances by the constant
    // rotation e^{-i*256*dphi_j} between visits.  The phasor is re-seeded from the exact fixed-point
    // phase every 8 visits, so the recurrence never accumulates more than 7 roundings.
    float rc[kFineBins], rs[kFineBins], wc[kFineBins], ws[kFineBins];
#pragma unroll
    for (int j = 0; j < kFineBins; ++j) { fix_sincos(dphi[j] * 256ull, &rs[j], &rc[j]); wc[j] = 1.f; ws[j] = 0.f; }
    int visit = 0;
    for (int n = threadIdx.x; n < p.N; n += 256, ++visit) {
        const short2 v = x[n];
        const float I = (float)v.x, Q = (float)v.y;
        const uint64_t gi = (uint64_t)c * p.N + n;     // finePhasePoints index (:148)
        const bool reseed = (visit & 7) == 0;
#pragma unroll
        for (int j = 0; j < kFineBins; ++j) {
            if (reseed) fix_sincos(dphi[j] * gi, &ws[j], &wc[j]);   // exp(-1i*f*finePhasePoints), :230
            ar[j] += fmaf(wc[j], I, ws[j] * Q);
            ai[j] += fmaf(wc[j], Q, -ws[j] * I);
            const float nc = fmaf(wc[j], rc[j], -ws[j] * rs[j]);    // advance the phase by 256 samples
            ws[j] = fmaf(wc[j], rs[j], ws[j] * rc[j]);
            wc[j] = nc;
        }
    }
    __shared__ double sh[8][kFineBins][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kFineBins; ++j) {
        double r = ar[j], i = ai[j];
        for (int o = 16; o > 0; o >>= 1) {
            r += __shfl_down_sync(0xffffffffu, r, o);
            i += __shfl_down_sync(0xffffffffu, i, o);
        }
        if (lane == 0) { sh[warp][j][0] = r; sh[warp][j][1] = i; }
    }
    __syncthreads();
    if (threadIdx.x < kFineBins && j0 + threadIdx.x < p.nFine) {
        double r = 0, i = 0;
        for (int w = 0; w < 8; ++w) { r += sh[w][threadIdx.x][0]; i += sh[w][threadIdx.x][1]; }
        double* o = p.sums + (((size_t)a * p.nFine + j0 + threadIdx.x) * p.nPeriods + c) * 2;
        o[0] = r; o[1] = i;
    }
}

// The same sums through moments (the default where the fine band is narrow, see launch_fine): the fine bins of an SV differ
// from its centre bin jc by d_j = dphi_j - dphi_jc, at most a few 1e-5 turns per sample, so over a run of 256 samples around
// g_c the bin's carrier is the centre bin's times e^{-i 2 pi d_j g_c} * (1 - i t r - t^2 r^2 / 2 + i t^3 r^3 / 6), t = 2 pi d_j,
// r = g - g_c in [-128, 128) (|t r| <= 0.025: the truncation is below 2e-8 of a term).  A warp wipes the centre carrier off its
// 256 samples ONCE, forms the four moments sum z r^m, reduces them over its lanes, and lane j turns them into bin j's sum:
// ~1.5 instructions per sample and warp instead of ~11.  The products d_j * g_c and dphi_jc * g are exact 64-bit fixed point.
constexpr int kFineRun = 256;      // samples per run
constexpr int kFineRounds = 4;     // fine bins per lane: nFine <= 128
__global__ void __launch_bounds__(256)
fine_sum_moments_kernel(FineParams p)
{
    const int c = blockIdx.x, a = blockIdx.y;
    if (a >= *p.nAcqDev * p.nCodes) return;
    const long long total = (long long)p.nPeriods * p.N;
    const short2* x = p.prod + (size_t)a * total + (size_t)c * p.N;
    const uint64_t* dphiA = p.dphi + (size_t)a * p.nFine;
    const uint64_t dc = dphiA[p.nFine / 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t dj[kFineRounds];
    float th[kFineRounds], ar[kFineRounds], ai[kFineRounds];
#pragma unroll
    for (int q = 0; q < kFineRounds; ++q) {
        const int j = lane + 32 * q;
        dj[q] = (j < p.nFine) ? dphiA[j] - dc : 0;
        th[q] = (float)((double)(int64_t)dj[q] * (6.283185307179586 / 18446744073709551616.0));
        ar[q] = ai[q] = 0.f;
    }
    float rs, rc;
    fix_sincos(dc * 32ull, &rs, &rc);                            // the centre carrier over 32 samples
    const int nRuns = (p.N + kFineRun - 1) / kFineRun;
    short2 nxt[kFineRun / 32];                                   // the next run's samples are in flight while this one is summed
#pragma unroll
    for (int i = 0; i < kFineRun / 32; ++i) {
        const int n = warp * kFineRun + lane + 32 * i;
        nxt[i] = (warp < nRuns && n < p.N) ? x[n] : make_short2(0, 0);
    }
    for (int b = warp; b < nRuns; b += 8) {
        const int n0 = b * kFineRun;
        const uint64_t g0 = (uint64_t)c * p.N + n0;             // finePhasePoints index of the run's first sample (:148)
        short2 cur[kFineRun / 32];
#pragma unroll
        for (int i = 0; i < kFineRun / 32; ++i) {
            cur[i] = nxt[i];
            const int n = n0 + 8 * kFineRun + lane + 32 * i;
            nxt[i] = (b + 8 < nRuns && n < p.N) ? x[n] : make_short2(0, 0);
        }
        float ws, wc;
        fix_sincos(dc * (g0 + lane), &ws, &wc);                  // exp(-1i*f*finePhasePoints) of the centre bin, :230
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = 0.f;
#pragma unroll
        for (int i = 0; i < kFineRun / 32; ++i) {                // (samples past the end of the code period were loaded as zero)
            const float I = (float)cur[i].x, Q = (float)cur[i].y;
            const float zr = fmaf(wc, I, ws * Q), zi = fmaf(wc, Q, -ws * I);
            const float r = (float)(lane + 32 * i - kFineRun / 2);
            m[0] += zr; m[1] += zi;
            float tr = zr * r, ti = zi * r;
            m[2] += tr; m[3] += ti;
            tr *= r; ti *= r;
            m[4] += tr; m[5] += ti;
            m[6] = fmaf(tr, r, m[6]); m[7] = fmaf(ti, r, m[7]);
            const float nc = fmaf(wc, rc, -ws * rs);             // advance the phase by 32 samples (7 roundings per run)
            ws = fmaf(wc, rs, ws * rc);
            wc = nc;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m[i] += __shfl_xor_sync(0xffffffffu, m[i], o);
        const uint64_t gc = g0 + kFineRun / 2;
#pragma unroll
        for (int q = 0; q < kFineRounds; ++q) {
            if (32 * q >= p.nFine) break;
            float es, ec;
            fix_sincos(dj[q] * gc, &es, &ec);
            const float t = th[q], t2 = 0.5f * t * t, t3 = t2 * t * (1.f / 3.f);
            const float pr = m[0] + t * m[3] - t2 * m[4] - t3 * m[7];
            const float pi = m[1] - t * m[2] - t2 * m[5] + t3 * m[6];
            ar[q] += fmaf(pr, ec, pi * es);
            ai[q] += fmaf(pi, ec, -pr * es);
        }
    }
    extern __shared__ double sh_m[];                             // [8][nFine][2]
#pragma unroll
    for (int q = 0; q < kFineRounds; ++q) {
        const int j = lane + 32 * q;
        if (j < p.nFine) { sh_m[(warp * p.nFine + j) * 2] = ar[q]; sh_m[(warp * p.nFine + j) * 2 + 1] = ai[q]; }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.nFine; j += 256) {
        double r = 0, i = 0;
        for (int w = 0; w < 8; ++w) { r += sh_m[(w * p.nFine + j) * 2]; i += sh_m[(w * p.nFine + j) * 2 + 1]; }
        double* o = p.sums + (((size_t)a * p.nFine + j) * p.nPeriods + c) * 2;
        o[0] = r; o[1] = i;
    }
}

// nav-bit-edge search and arg-max over the fine bins (acquisition.m:240-253): for each bin the
// maximum over the 20 start offsets of |sum of 20 consecutive per-code sums| (GLONASS: the two
// 10 ms meander halves enter with opposite sign).  B3I (BDS/B3I/include/acquisition.m:193-211):
// GEO PRNs (1-5, 59-63) carry 2 ms bits, the others the 20-bit Neumann-Hoffman code.
// One (bin, alignment) pair per thread, each evaluated with the reference's own summation order; the maximum over the
// alignments of a bin is order independent (non-negative doubles compare like their bit patterns: atomicMax in shared memory).
__device__ __forceinline__ int fine_alignments(const FineParams& p, int prn)
{
    switch (p.combine) {
        case 2: return ((prn >= 1 && prn <= 5) || (prn >= 59 && prn <= 63)) ? 2 : 20;
        case 3: return 25;
        case 4: return p.nPeriods;
        case 5: return 1;
        default: return p.nPeriods / 2;
    }
}

__device__ double fine_power(const FineParams& p, int a, int nAcq, int j, int c, int prn)
{
    const double* s = p.sums + ((size_t)a * p.nFine + j) * p.nPeriods * 2;
    const int half = p.nPeriods / 2;
    if (p.combine == 4) {
        // secondary code of the pilot, every circular alignment (GPS_L5C acquisition.m:214-219: the code is rotated
        // by one element per step, after step c it holds sec(q - c))
        const int8_t* sec = p.secondary + (size_t)a * p.nPeriods;
        double r = 0, i = 0;
        for (int q = 0; q < p.nPeriods; ++q) {
            const double sc = (double)sec[(q - c + p.nPeriods) % p.nPeriods];
            r += s[2 * q] * sc; i += s[2 * q + 1] * sc;
        }
        return hypot(r, i);
    }
    if (p.combine == 5) {
        const double* s2 = p.sums + ((size_t)(a + nAcq) * p.nFine + j) * p.nPeriods * 2;
        double t1 = 0, t2 = 0;
        for (int q = 0; q < p.nPeriods; ++q) { t1 += hypot(s[2 * q], s[2 * q + 1]); t2 += hypot(s2[2 * q], s2[2 * q + 1]); }
        return t1 + t2;
    }
    if (p.combine == 3 || (p.combine == 2 && fine_alignments(p, prn) == 20)) {
        // Galileo E1: 25 code periods against the 25-chip pilot secondary code '380AD90', aligned and at the 24 other edges
        // (GAL_E1C/include/acquisition.m:135, 236-252); B3I MEO / IGSO: the 20-bit Neumann-Hoffman code the same way
        // (BDS/B3I/include/acquisition.m:127, 199-210).  code2ndShift = circshift(code', c)': element q takes code[(q - c) mod n]
        const double SEC[25] = {1, 1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, -1, 1, -1, 1, -1, -1, 1, -1, -1, 1, 1, -1, 1};
        const double NH[20] = {1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1};
        const int n = p.combine == 3 ? 25 : 20;
        const double* code = p.combine == 3 ? SEC : NH;
        if (c == 0) {
            double r = 0, i = 0;
            for (int q = 0; q < n; ++q) { r += s[2 * q] * code[q]; i += s[2 * q + 1] * code[q]; }
            return hypot(r, i);
        }
        double r1 = 0, i1 = 0, r2 = 0, i2 = 0;
        for (int q = 0; q < n; ++q) {
            const double sc = code[(q - c + n) % n];
            if (q < c) { r1 += s[2 * q] * sc; i1 += s[2 * q + 1] * sc; }
            else { r2 += s[2 * q] * sc; i2 += s[2 * q + 1] * sc; }
        }
        return hypot(r1, i1) + hypot(r2, i2);
    }
    if (p.combine == 2) {                                                        // B3I GEO: 2 ms bits (:193-198)
        if (c == 0) {
            double c1 = 0;
            for (int q = 0; q < 20; q += 2) c1 += hypot(s[2 * q] + s[2 * q + 2], s[2 * q + 1] + s[2 * q + 3]);
            return c1;
        }
        double c2 = hypot(s[0], s[1]) + hypot(s[38], s[39]);
        for (int q = 1; q < 19; q += 2) c2 += hypot(s[2 * q] + s[2 * q + 2], s[2 * q + 1] + s[2 * q + 3]);
        return c2;
    }
    double r = 0, i = 0;
    if (p.combine == 0) {
        for (int q = c; q < c + half; ++q) { r += s[2 * q]; i += s[2 * q + 1]; }
    } else {                                             // 10 ms meander halves of opposite sign (GLO :246-252)
        for (int q = c; q < c + half / 2; ++q) { r += s[2 * q]; i += s[2 * q + 1]; }
        for (int q = c + half / 2; q < c + half; ++q) { r -= s[2 * q]; i -= s[2 * q + 1]; }
    }
    return sqrt(r * r + i * i);
}

__global__ void __launch_bounds__(256)
fine_select_kernel(FineParams p)
{
    const int a = blockIdx.x;
    const int nAcq = *p.nAcqDev;
    if (a >= nAcq) return;
    extern __shared__ unsigned long long s_best[];               // [nFine] bit pattern of the bin's maximum (>= 0)
    for (int j = threadIdx.x; j < p.nFine; j += blockDim.x) s_best[j] = 0ull;
    __syncthreads();
    const int prn = p.combine == 2 ? p.svId[a] : 0;
    const int nC = fine_alignments(p, prn);
    for (int w = threadIdx.x; w < p.nFine * nC; w += blockDim.x) {
        const int j = w / nC, c = w - j * nC;
        const double pw = fine_power(p, a, nAcq, j, c, prn);
        if (pw > 0) atomicMax(&s_best[j], (unsigned long long)__double_as_longlong(pw));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.nFine; j += blockDim.x) p.fineResult[a * p.nFine + j] = __longlong_as_double((long long)s_best[j]);
    if (threadIdx.x == 0) {
        int best = 0; double bv = __longlong_as_double((long long)s_best[0]);
        for (int j = 1; j < p.nFine; ++j) {
            const double v = __longlong_as_double((long long)s_best[j]);
            if (v > bv) { bv = v; best = j; }
        }
        p.best[a] = best;
    }
}

}  // namespace

cudaError_t launch_sig_power(Rec rec, long long winStart, int N, double* out, cudaStream_t s)
{
    sig_power_kernel<<<1, 1024, 0, s>>>(rec, winStart, N, out);
    return cudaGetLastError();
}

cudaError_t launch_peak_select(const float* partMax, const int* partIdx, int nPrnSlots, int nBins, int parts,
                               PeakOut* out, cudaStream_t s)
{
    peak_select_kernel<<<nPrnSlots, 32, 0, s>>>(partMax, partIdx, nBins, parts, out);
    return cudaGetLastError();
}

// one block: metric and threshold per slot (acquisition.m:200, 206), the list of acquired slots in list order, and per
// fine-search entry its chip row, code phase and the phase increments of the fine bins (:221-227)
__global__ void __launch_bounds__(128)
fine_setup_kernel(FineSetup p)
{
    // one thread per list slot (at most 63: two warps); the list of acquired slots in list order from the warps' ballots
    __shared__ unsigned s_mask[4];
    const int slot = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool above = false;
    if (slot < p.nSv) {
        const double m = __ddiv_rn(__ddiv_rn(p.peaks[slot].peak, *p.sigPower), (double)p.nonCoh);     // :200
        p.metric[slot] = m;
        above = m > p.threshold;                                                                        // :206
    }
    const unsigned mask = __ballot_sync(0xffffffffu, above);
    if (lane == 0) s_mask[warp] = mask;
    __syncthreads();
    {
        int before = __popc(mask & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) before += __popc(s_mask[w]);
        if (above) p.acqSlot[before] = slot;
    }
    int nAcq = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) nAcq += __popc(s_mask[w]);
    if (threadIdx.x == 0) *p.nAcq = nAcq;
    __syncthreads();                                             // acqSlot is read back below
    for (int e = threadIdx.x; e < nAcq * p.nCodes; e += blockDim.x) {
        const int a = e % nAcq, comp = p.nCodes == 2 ? e / nAcq : p.pilotComp;
        const int s = p.acqSlot[a];
        p.chipRow[e] = p.slotChipRow[s] + comp;
        p.codePhase[e] = p.peaks[s].codePhase;
        if (e < nAcq) p.svId[a] = p.slotSv[s];
    }
    for (int i = threadIdx.x; i < nAcq * p.nCodes * p.nFine; i += blockDim.x) {
        const int e = i / p.nFine, j = i - e * p.nFine;
        const int s = p.acqSlot[e % nAcq];
        // coarseFreqBin(bin) + acqSearchStep/2 - fineSearchStep*(j-1), every operation rounded as the host does (:169, :227)
        const double coarse = __dsub_rn(p.slotFreq0[s], __dmul_rn(p.step, (double)(p.peaks[s].bin - 1)));
        const double f = __dsub_rn(__dadd_rn(coarse, __ddiv_rn(p.step, 2.0)), __dmul_rn(p.fineStep, (double)j));
        p.dphi[i] = turns_to_fix(__dmul_rn(f, p.ts));
    }
    if (p.slotSecondary)
        for (int i = threadIdx.x; i < nAcq * p.nPeriods; i += blockDim.x)
            p.secondary[i] = p.slotSecondary[(size_t)p.acqSlot[i / p.nPeriods] * p.nPeriods + i % p.nPeriods];
}

cudaError_t launch_fine_setup(const FineSetup& p, cudaStream_t s)
{
    fine_setup_kernel<<<1, 128, 0, s>>>(p);
    return cudaGetLastError();
}

// acqResults on the device (gc_acquire_device): the host expressions of acquire_impl, operation for operation
// (acquisition.m:200, 206, 227, 254-260), so that what a collective gathers from device memory is what gc_acquire returns
__global__ void pack_results_kernel(PackParams p)
{
    for (int i = threadIdx.x; i < 4 * p.resultLen; i += blockDim.x) p.out[i] = 0.0;
    // one thread per list slot; a = the slot's position among the acquired ones (the order fine_setup_kernel numbered them in)
    __shared__ unsigned s_mask[4];
    const int s = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool above = false;
    double m = 0.0;
    if (s < p.nSv) {
        m = __ddiv_rn(__ddiv_rn(p.peaks[s].peak, *p.sigPower), (double)p.nonCoh);                     // :200
        above = m > p.threshold;                                                                      // :206
    }
    const unsigned mask = __ballot_sync(0xffffffffu, above);
    if (lane == 0) s_mask[warp] = mask;
    __syncthreads();                                             // (also orders the zero fill above before the entries below)
    if (s >= p.nSv) return;
    int a = __popc(mask & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) a += __popc(s_mask[w]);
    const int ri = p.slotResult[s];
    p.out[ri] = m;
    p.out[3 * p.resultLen + ri] = (double)p.peaks[s].bin;
    if (above) {
        const double coarse = __dsub_rn(p.slotFreq0[s], __dmul_rn(p.step, (double)(p.peaks[s].bin - 1)));   // :169
        double f = coarse;
        if (!p.noFine) {
            f = __dsub_rn(__dadd_rn(coarse, __ddiv_rn(p.step, 2.0)), __dmul_rn(p.fineStep, (double)p.best[a]));   // :227, :254
            if (f == 0.0) f = 1.0;                                                                    // :258
        }
        p.out[2 * p.resultLen + ri] = f;
        p.out[p.resultLen + ri] = (double)p.peaks[s].codePhase;                                       // :256
    }
}

cudaError_t launch_pack_results(const PackParams& p, cudaStream_t s)
{
    pack_results_kernel<<<1, 128, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fine(const FineParams& p, int nEntries, int nAcq, cudaStream_t s)
{
    dim3 g1(148 * 2, nEntries);
    fine_prep_kernel<<<g1, 256, 0, s>>>(p);
    if (p.moments) {
        dim3 g2(p.nPeriods, nEntries, 1);
        fine_sum_moments_kernel<<<g2, 256, sizeof(double) * 16 * p.nFine, s>>>(p);
    } else {
        dim3 g2(p.nPeriods, nEntries, (p.nFine + kFineBins - 1) / kFineBins);
        fine_sum_kernel<<<g2, 256, 0, s>>>(p);
    }
    fine_select_kernel<<<nAcq, 256, sizeof(unsigned long long) * p.nFine, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gc
