// Tracking: correlate-and-dump with the DLL/PLL closed on the device.
//
// Replaces the epoch loop of GPS/GPS_L1CA/include/tracking.m:184-360.  One persistent CTA per
// channel walks the resident IF record; the 1 ms sample block of epoch e+1 is staged into shared
// memory by the TMA unit (cp.async.bulk + mbarrier, double buffered) while epoch e is correlated.
// Its start is known as soon as epoch e begins (start + blksize); only its length changes by a
// sample or so, so a fixed-size window is fetched.
//
// Numerics (targets: I/Q sums within 1e-6 relative of the float64 reference over a minute-long closed loop - which in practice needs
// them at ~1e-13, DESIGN.md section 2 - and every recorded state variable computed in float64 as the reference does):
//   * code phase: t(k) in float64 with the reference's own operation order (MATLAB colon vector a:d:b built from both ends, then
//     ceil) because one chip flip changes a sum by ~1e-3 relative.  ceil() is a round-up add of 1.5*2^52 (FP64 pipe) instead of a
//     float->int conversion; where a table entry spans >= 8 samples the per-sample evaluation is replaced by a per-chunk edge
//     prediction that falls back to it whenever a sample is within 2^-24 samples of an edge;
//   * carrier: phase kept as a 64-bit fixed-point fraction of a turn; per 8-sample chunk one float64 phasor from a 1024-entry
//     table + Taylor remainder (~2e-16), the 8 in-chunk rotations e^{-i*j*dphi} are per-epoch float64 constants;
//   * samples, wipe-off, replica entries and sums: float64 throughout (byte -> double by PRMT, entries as double high words);
//   * discriminators atan(Q/I), sqrt, divide: float64 as the reference evaluates them (GC_PARAM_TRACK_FAST_DISC selects fp32 forms,
//     1e-7 relative, for callers that want the last 1.5 %); recorded pllDiscr / dllDiscr rows follow the oracle to ~1e-13;
//   * remCarrPhase is the fixed-point phase in radians with the sign of carrFreq: the reference's rem(trigarg, 2*pi) recurrence to
//     ~4e-8 rad after 60000 epochs, and modulo 2*pi where carrFreq changes sign on a baseband record.
//
// With few channels (BASELINE: 12) one CTA per channel leaves most of the chip idle and every
// epoch is a dependent step, so a channel can instead be spread over a thread-block cluster of
// G = 2/4/8 CTAs: each CTA stages and correlates 1/G of the block, pushes its partial sums
// into every peer's shared memory (DSMEM, st.async + mbarrier complete_tx), and then every CTA closes the
// loops redundantly from the same numbers (bit-identical state everywhere, no broadcast step).
//
// The per-epoch scalar work is split so that little of it sits on the critical path:
//   carrier-loop thread : PLL (float64 atan), the next block's carrier constants, the recorded values    after the sums are reduced
//   code-loop thread    : DLL and the geometry of the next block                                          concurrently
// Everything that feeds a ceil() (codePhaseStep, blksize) is exact IEEE.
#include <type_traits>

#include "common.cuh"
#include "track.h"

namespace gc {

namespace {

constexpr int kMaxWarps = 16;
constexpr int kMaxCluster = 8;
constexpr int kPad = 16;     // zero entries on both sides of the code table (masked out-of-block samples index them)
constexpr int kPad61 = 96;   // same for the BOC(6,1) table, whose index runs six times as fast
constexpr int kStage = 16;   // epochs of results staged in smem before a coalesced flush
constexpr double kCeilMagic = 6755399441055744.0;   // 1.5 * 2^52: t + magic stays in [2^52, 2^53) for |t| < 2^51

struct alignas(16) EpochParams {   // parameters of the block being correlated (read by every thread)
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
    int fast;                 // every chunk of 8 samples holds at most one table-entry edge per correlator (8*d*subChip <= 1) and the
                              // three colon vectors share their length: indices by the per-chunk edge prediction
    uint32_t rinv;            // floor(2^24 / (d*subChip)): samples per table entry in 8.24 fixed point
    int pad;
    double codeFreq, carrFreq, remCodePhase, remCarrPhase;   // values used in this block (recorded)
    // in-chunk carrier rotations e^{-i*2*pi*j*dphi}, j = 0..7, and the bias terms of the byte -> double trick (written by eight
    // lanes of the carrier-loop warp): rot[j] = {cos, sin, B*(cos + sin), B*(cos - sin)}
    alignas(16) double rot[8][4];
};

struct NextPhases {           // NCO phases at the end of the block being correlated (warp 2)
    double remCodePhase, remCarrPhase;
    uint64_t phase0;
};

struct LoopMem {              // loop-filter memories
    double oldCodeNco, oldCodeError, oldCarrNco, oldCarrError, carrFreqBasis;
    double d2CarrError, dCarrError;   // loopType 1 (GLO tracking.m:171-172)
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
                           double remCarrPhase, uint64_t phase0, long long pos, EpochParams& ep, double& invStep)
{
    // :219 codeFreq / samplingFreq, correctly rounded: q0 = a*RN(1/b), r = a - q0*b (exact, FMA),
    // q = RN(q0 + r*RN(1/b))  (Markstein); three dependent operations instead of a full division
    const double q0 = __dmul_rn(codeFreq, p.invFs);
    const double step = __fma_rn(__fma_rn(-q0, p.fs, codeFreq), p.invFs, q0);
    // :222 blksize = ceil((codeLength - remCodePhase) / codePhaseStep).  The quotient is formed with a
    // reciprocal carried from the previous block and refreshed by a Newton step (the step moves by
    // < 1e-6 relative per epoch, so the refreshed value is good to ~1e-12); if the quotient is within 1e-6
    // of an integer the exact IEEE division decides, so the result always equals the reference's.
    const double x = __dsub_rn(p.codeLength, remCodePhase);
    double y = (invStep != 0.0) ? invStep : __ddiv_rn(1.0, step);
    y = fma(y, fma(-step, y, 1.0), y);
    invStep = y;
    double q = x * y;
    if (!(fabs(q - rint(q)) > 1e-6)) q = __ddiv_rn(x, step);
    const int blk = (int)ceil(q);
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
    // fast path: the three vectors share n, and the samples masked just outside the block still index the
    // zero padding of the code table
    ep.generic = !(ep.nE == blk - 1 && ep.nP == blk - 1 && ep.nL == blk - 1) || !((8.0 * step + p.spc) * (double)p.subChip + 2.0 < (double)kPad) ||
                 (p.pilot == 5 && !((8.0 * step + p.spc) * (double)p.subChip * 6.0 + 2.0 < (double)kPad61));
    {   // per-chunk edge prediction (track_kernel, fast chunks): needs at most one table-entry edge per correlator in 8 samples and a
        // samples-per-entry count that fits 8.24 fixed point
        const double dt = step * (double)p.subChip;
        ep.fast = (!ep.generic && 8.0 * dt <= 1.0 && dt >= 0.0078125 && p.pilot != 5 && p.subChip <= 2) ? 1 : 0;
        ep.rinv = ep.fast ? (uint32_t)(y * (p.subChip == 2 ? 8388608.0 : 16777216.0)) : 0u;   // y = 1/step to ~1e-12: far inside the margin
    }
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
    ep.dphi = turns_to_fix(carrFreq * p.invFs);
}

// NCO phases after the block described by ep (tracking.m:273, :280-283).  The carrier phase is
// advanced in the 64-bit fixed-point domain (exact to 2^-64 turn per sample); remCarrPhase, the
// recorded value, is that phase in radians with the sign rem(trigarg, 2*pi) would have (it agrees
// with the reference's float64 recurrence to ~1e-12 rad, far inside every tolerance).
__device__ __forceinline__ void end_phases(const TrackParams& p, const EpochParams& ep, NextPhases& nx)
{
    const double lastP = ep.generic ? colon_elem(ep.aP, ep.d, ep.cP, ep.nP, ep.blk - 1) : ep.cP;
    nx.remCodePhase = __dsub_rn(__dadd_rn(lastP, ep.d), p.codeLength);                        // :273
    nx.phase0 = ep.phase0 + ep.dphi * (uint64_t)ep.blk;                                       // :280-283
    const double frac = (double)(nx.phase0 >> 11) * 1.1102230246251565e-16;                   // [0,1) turns, 53 bits
    nx.remCarrPhase = (ep.carrFreq < 0.0 && frac != 0.0) ? (frac - 1.0) * kTwoPi : frac * kTwoPi;
    if (p.exact) {                                               // rem(trigarg(blksize+1), 2*pi) in float64 as written (:281-283)
        const double w = __dmul_rn(__dmul_rn(ep.carrFreq, 2.0), 3.141592653589793);
        const double trigEnd = __dadd_rn(__dmul_rn(w, __ddiv_rn((double)ep.blk, p.fs)), ep.remCarrPhase);
        nx.remCarrPhase = fmod(trigEnd, __dmul_rn(2.0, 3.141592653589793));
    }
}

// ---- float64 carrier phasors from the 64-bit fixed-point phase -------------------------------------------------------------
// cos / sin(2*pi*phase) = table entry of the nearest 1/1024 turn rotated by the remainder (|alpha| <= 2*pi/2048: Taylor terms
// up to alpha^5, next one < 1e-18).  About 15 float64 instructions and one 16-byte load; error ~2e-16.
__device__ const double2 g_sincos1024[1024] = {
#include "sincos1024.inc"
};
__device__ __forceinline__ void fix_sincos_f64(uint64_t phase, double* sn, double* cs)
{
    const uint64_t rounded = phase + (1ull << 53);
    const uint32_t idx = (uint32_t)(rounded >> 54);                                // nearest 1/1024 turn (1024 wraps to 0 below)
    const long long r = (long long)(phase - (rounded & ~((1ull << 54) - 1)));      // signed remainder, |r| <= 2^53
    const double alpha = __dmul_rn(__ll2double_rn(r), 3.4061215800865545e-19);     // 2*pi * 2^-64
    const double a2 = __dmul_rn(alpha, alpha);
    const double c = __fma_rn(a2, __fma_rn(a2, 4.1666666666666664e-2, -0.5), 1.0);
    const double t = __dmul_rn(alpha, __fma_rn(a2, __fma_rn(a2, 8.3333333333333332e-3, -1.6666666666666666e-1), 1.0));
    const double2 T = g_sincos1024[idx & 1023u];
    *cs = __fma_rn(-T.y, t, __dmul_rn(T.x, c));
    *sn = __fma_rn(T.x, t, __dmul_rn(T.y, c));
}

// ---- byte -> double without a conversion instruction (I2F.F64 and F2F.F64 issue at 1/8 rate on this part) ------------------
// The sample byte, already xor-ed with 0x80 (b' = x + 128), is permuted into bits 8-15 of the high word 0x40B0xx00 of a double
// whose low word is zero: that double is exactly 4096 + b' = kByteBias + x.  The bias is removed inside the wipe-off FMA chain:
// wc*(B + xr) + ws*(B + xq) - B*(wc + ws), with B*(wc + ws) a per-epoch constant (EpochParams::rot).  One PRMT per component.
constexpr double kByteBias = 4224.0;
__device__ __forceinline__ double byte_to_biased_double(uint32_t wx, uint32_t sel)
{
    return __hiloint2double((int)__byte_perm(wx, 0x40B00000u, sel), 0);
}

// ---- code tables in shared memory: the HIGH WORD of the entry as a double (low word zero), so that a looked-up entry enters a
// DFMA without any conversion.  uint32_t: 0x3FF00000 / 0xBFF00000 / 0 = +1 / -1 / 0.  uint8_t (where three tables have to fit,
// B1C full band): the top byte 0x40 / 0xC0 / 0 = +2 / -2 / 0, the sums are halved after the reduction (exact).
template <typename TT> struct CodeEnc;
template <> struct CodeEnc<uint32_t> {
    static __device__ __forceinline__ uint32_t store(int8_t v) { return v > 0 ? 0x3FF00000u : v < 0 ? 0xBFF00000u : 0u; }
    static __device__ __forceinline__ uint32_t hi(uint32_t e) { return e; }
    static constexpr double kScale = 1.0;
};
template <> struct CodeEnc<uint8_t> {
    static __device__ __forceinline__ uint8_t store(int8_t v) { return v > 0 ? 0x40 : v < 0 ? 0xC0 : 0; }
    static __device__ __forceinline__ uint32_t hi(uint8_t e) { return (uint32_t)e << 24; }
    static constexpr double kScale = 0.5;
};
__device__ __forceinline__ double hi2d(uint32_t hi) { return __hiloint2double((int)hi, 0); }

// in-chunk rotations of the block whose carrier plan_carrier just wrote: lane j < 8 of the calling warp fills rot[j]
__device__ __forceinline__ void plan_rotations(EpochParams& ep, int lane, double bias)
{
    if (lane < 8) {
        double sn, cs;
        fix_sincos_f64(ep.dphi * (uint64_t)lane, &sn, &cs);
        ep.rot[lane][0] = cs; ep.rot[lane][1] = sn;
        ep.rot[lane][2] = __dmul_rn(bias, __dadd_rn(cs, sn)); ep.rot[lane][3] = __dmul_rn(bias, __dsub_rn(cs, sn));
    }
}


__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a double into the same shared-memory location of CTA `rank` of this cluster and credit 8
// bytes to that CTA's mbarrier `bar` (same offset): data and signal travel together, so the
// consumer needs no cluster-wide barrier, only a wait on its own mbarrier.
__device__ __forceinline__ void dsmem_push(double* local, uint64_t* bar, uint32_t rank, double v)
{
    uint32_t raddr, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}

// two doubles in one DSMEM transaction (16 bytes credited)
__device__ __forceinline__ void dsmem_push2(double* local, uint64_t* bar, uint32_t rank, double v0, double v1)
{
    uint32_t raddr, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "l"(__double_as_longlong(v0)), "l"(__double_as_longlong(v1)), "r"(rbar) : "memory");
}

}  // namespace

// G = CTAs per channel (cluster size), T = threads per CTA, NSET = replicas correlated (1 data; 2 data + pilot, 12 sums;
// 3 data + pilot BOC(1,1) + pilot BOC(6,1), 18 sums - B1C WB_tracking.m), TT = storage type of the code tables in
// shared memory (CodeEnc: uint32_t, or uint8_t where three tables have to fit)
// FMT0 = the record is int8 I,Q (bulk-copied windows, byte-permute conversion); false = the per-sample accessor for the
// int16 / real formats (a separate instantiation, so that the fast path's code is not touched by it)
// EXACT = the float64 checking mode (TrackParams::exact, GC_PARAM_TRACK_EXACT_SUMS): carrier exp(-1i*trigarg) per sample in float64 from
// the reference's own expression, float64 products and sums, rem(trigarg, 2*pi) recurrence for remCarrPhase, float64 discriminators - the
// loop state then follows the float64 reference to ~1e-13 instead of ~1e-10, which is what the parity tests use to show that the
// windows they skip at 18 Msps are conditioning (a sample within 1e-9 chips of a chip edge) and not an error of this kernel.
template <int G, int T, int NSET, typename TT, bool FMT0, bool EXACT>
__global__ void __launch_bounds__(T, T <= 256 ? 2 : 1)
track_kernel(TrackParams p)
{
    constexpr bool PILOT = NSET >= 2;
    using Enc = CodeEnc<TT>;
    constexpr int NS = 6 * NSET;                                 // correlator sums per epoch
    constexpr int kThreads = T;
    constexpr int kWarps = T / 32;
    // Role threads.  The 8-CTA variant carries three spare warps so that the TMA issue, the one
    // non-uniform chunk and the loop closure never serialise with a warp full of regular chunks.
    constexpr int kSpecialTid = (G == 8) ? 9 * 32 : T - 1;      // chunk holding the middle of the colon vector
    constexpr int kLoader = (G == 8) ? 10 * 32 : T - 1;         // issues the bulk copies
    constexpr int kPllTid = (G == 8) ? 10 * 32 : 0;             // carrier loop
    constexpr int kDllTid = (G == 8) ? 9 * 32 : 32;             // code loop + geometry of the next block
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [buf0 | buf1 | code table (float) | warp partials | staging | params | mbarriers]
    int8_t* const buf0 = reinterpret_cast<int8_t*>(smem_raw);
    TT* s_code_raw = reinterpret_cast<TT*>(smem_raw + (p.singleBuf ? 1 : 2) * (size_t)p.bufBytes);
    TT* s_code = s_code_raw + kPad;                              // index 0 = c(L) of the wrapped table
    const int tabFloats = (p.codeLen + 2 + 2 * kPad + 15) & ~15; // entries per table
    TT* s_pilot = s_code + (PILOT ? tabFloats : 0);              // pilot table right behind the data table
    const int tab61 = (NSET == 3) ? ((p.codeLen * 6 + 2 + 2 * kPad61 + 15) & ~15) : 0;
    TT* s_p61 = s_code_raw + 2 * tabFloats + kPad61;             // BOC(6,1) pilot table, 6 entries per BOC(1,1) entry (NSET == 3)
    double* s_part = reinterpret_cast<double*>(s_code_raw + (PILOT ? 2 : 1) * tabFloats + tab61);
    double* s_cl = s_part + kMaxWarps * NS;                      // [2][kMaxCluster][NS] per-CTA partial sums (pushed by peers)
    double* s_stage = s_cl + 2 * kMaxCluster * NS;               // [15][kStage]
    EpochParams* s_ep = reinterpret_cast<EpochParams*>(s_stage + GC_TRACK_ROWS * kStage);   // [2]
    NextPhases* s_nx = reinterpret_cast<NextPhases*>(s_ep + 2);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_nx + 1);     // 2 mbarriers (TMA stages) + 2 (partial-sum exchange)
    uint64_t* s_xbar = s_bar + 2;
    int* s_issued = reinterpret_cast<int*>(s_bar + 4);           // per stage: 1 + epoch whose window was requested (0 = none)
    // [7][T] running sums of the fast chunks (TrackParams::preSlots), one 16-byte column per thread
    double2* s_pre = reinterpret_cast<double2*>((reinterpret_cast<uintptr_t>(s_issued + 2) + 15) & ~(uintptr_t)15);

    const int ch = blockIdx.x / G;
    const uint32_t crank = (G > 1) ? cluster_ctarank() : 0u;
    const bool leader = (crank == 0);
    const TrackChan cinfo = p.chans[ch];
    if (cinfo.pad == 0) {                                        // channel off (tracking.m:136; GLO tracking.m:137)
        if (threadIdx.x == 0 && leader) p.epochsDone[ch] = 0;
        return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* out = p.out + (size_t)ch * p.nRows * p.nEpochs;

    // wrapped code table [c(L) c(1..L) c(1)]  (tracking.m:156-158)
    for (int i = tid; i < p.codeLen + 2 + 2 * kPad; i += kThreads) {
        const int j = i - kPad;
        s_code_raw[i] = (j >= 0 && j < p.codeLen + 2) ? Enc::store(p.codeTables[(size_t)ch * p.codeStride + j]) : (TT)0;
        if (PILOT) s_code_raw[tabFloats + i] = (p.pilot != 4 && j >= 0 && j < p.codeLen + 2) ? Enc::store(p.pilotTables[(size_t)ch * p.pilotStride + j]) : (TT)0;
    }
    if (NSET == 3)                                               // [p61(12L) p61 p61(1)] (B1C WB_tracking.m:181-183)
        for (int i = tid; i < p.codeLen * 6 + 2 + 2 * kPad61; i += kThreads) {
            const int j = i - kPad61;
            s_code_raw[2 * tabFloats + i] = (j >= 0 && j < p.codeLen * 6 + 2) ? Enc::store(p.p61Tables[(size_t)ch * p.p61Stride + j]) : (TT)0;
        }

    LoopMem lm;   // PLL thread: carrier memories; DLL thread: code memories
    lm.oldCodeNco = lm.oldCodeError = lm.oldCarrNco = lm.oldCarrError = 0.0;   // :173-178
    lm.carrFreqBasis = cinfo.acqFreq;                                          // :168
    lm.d2CarrError = lm.dCarrError = 0.0;
    double invStep = 0.0;                                        // DLL thread: 1/codePhaseStep of the previous block
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_xbar[0], 1);
        mbar_init(&s_xbar[1], 1);
        mbar_fence_init();
        s_issued[0] = s_issued[1] = 0;
        // :163-170 codeFreq = codeFreqBasis, remCodePhase = 0, carrFreq = acquiredFreq, remCarrPhase = 0; :150 fseek
        double inv0 = 0.0;
        plan_epoch(p, cinfo.codeFreq0, 0.0, 0.0, 0ull, cinfo.startSample, s_ep[0], inv0);
        plan_carrier(p, cinfo.acqFreq, s_ep[0]);
    }
    __syncthreads();
    if (warp == 0) plan_rotations(s_ep[0], lane, FMT0 ? kByteBias : 0.0);
    __syncthreads();

    const long long recBytesUp = (p.recSamples * 2 + 15) & ~15LL;
    // this CTA correlates the 16-byte chunks [c_lo, c_lo + cpc) of the block window
    const int cpc = p.bufBytes / 16;                             // chunks per CTA (bufBytes is per CTA)
    const int c_lo = (int)crank * cpc;
    auto prefetch = [&](long long startSample, int stage, int epoch) {      // one thread only
        if (!FMT0) return false;                                 // only int8 I,Q windows are staged
        const long long b0 = ((startSample * 2) & ~15LL) + (long long)c_lo * 16;
        long long n = p.bufBytes;
        if (b0 + n > recBytesUp) n = recBytesUp - b0;
        if (b0 < 0 || n <= 0) return false;                      // this CTA's slice lies beyond the record
        mbar_expect_tx(&s_bar[stage], (uint32_t)n);
        bulk_g2s(buf0 + (p.singleBuf ? 0 : (size_t)stage * p.bufBytes), p.rec + b0, (uint32_t)n, &s_bar[stage]);
        s_issued[stage] = epoch + 1;
        return true;
    };
    if (tid == kLoader && !s_ep[0].stop) prefetch(s_ep[0].pos, 0, 0);
    __syncthreads();
    if (G > 1) cluster_sync_all();                               // every peer's exchange barriers are initialised
    uint32_t xphase[2] = {0u, 0u};

    uint32_t phase[2] = {0u, 0u};                                // mbarrier phase parity per stage
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};                // optional phase timing (p.dbg != nullptr)
    const bool timing = (p.dbg != nullptr) && blockIdx.x < G;
#define GC_TICK(i) if (timing) { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; }
    long long tprev = clock64();
    int e = 0;
    for (; e < p.nEpochs; ++e) {
        const int stage = e & 1;
        const EpochParams& ep = s_ep[stage];
        const bool staged = (s_issued[stage] == e + 1);          // was this block's window requested?
        if (ep.stop) {
            // never leave a bulk copy in flight into this CTA's shared memory
            if (staged) mbar_wait(&s_bar[stage], phase[stage]);
            break;
        }
        const int blk = ep.blk, n = ep.n;
        const long long pos = ep.pos;
        const uint64_t dphi = ep.dphi, phase0 = ep.phase0;
        if (PILOT && p.pilot == 4) {
            // GPS L2C: the pilot table of this epoch is one 20 ms segment of the padded CL sequence,
            // CLCode(tcode2 + codeLength*(CLCodePhase-1)), CLCodePhase stepping 1..75 with the epochs (GPS_L2C tracking.m:261, 363-366).
            // Every thread finished the previous epoch's lookups at the barrier that ended it.
            const int seg = (cinfo.clPhase - 1 + e) % 75;
            const int8_t* cl = p.pilotTables + (size_t)ch * p.pilotStride + (size_t)seg * p.codeLen;
            for (int i = tid; i < p.codeLen + 2; i += kThreads) s_pilot[i] = Enc::store(cl[i]);
            __syncthreads();
        }
        // window of epoch e+1 starts where this one ends; fetch it while we correlate
        if (!p.singleBuf && tid == kLoader && e + 1 < p.nEpochs) prefetch(pos + blk, stage ^ 1, e + 1);
        if (G > 1 && tid == kPllTid) mbar_expect_tx(&s_xbar[e & 1], 8u * NS * G);   // G CTAs x NS doubles will arrive
        GC_TICK(0)
        if (staged) { mbar_wait(&s_bar[stage], phase[stage]); phase[stage] ^= 1u; }
        GC_TICK(1)

        const int off = (int)((pos * 2) & 15) >> 1;              // samples skipped in the first 16-byte chunk
        const int nChunks = (off + blk + 7) >> 3;
        const bool fits = (nChunks <= cpc * G);
        // oversize block (never with sane loop settings): CTA rank r takes chunks r, r+G, ... from L2
        const int c_begin = fits ? c_lo : (int)crank, c_end = fits ? min(nChunks, c_lo + cpc) : nChunks;
        const int c_step = fits ? kThreads : kThreads * G;
        const bool inBuf = fits && staged;
        const int8_t* src = buf0 + (p.singleBuf ? 0 : (size_t)stage * p.bufBytes) - (size_t)c_lo * 16;
        const int8_t* gsrc = p.rec + ((pos * 2) & ~15LL);
        // table index = ceil(tcode * subChip): the scaling by 1 or 2 is exact, so the scaled colon vector is
        // element for element the reference's (rem -/+ spc)*2 : step*2 : (...)*2  (GAL_E1C tracking.m:236-262)
        const double sc = (double)p.subChip;
        const double d = ep.d * sc;
        const double aE = ep.aE * sc, aP = ep.aP * sc, aL = ep.aL * sc, cE = ep.cE * sc, cP = ep.cP * sc, cL = ep.cL * sc;
        const bool generic = ep.generic != 0;
        const double mE = ep.mE * sc, mP = ep.mP * sc, mL = ep.mL * sc;
        const int nE_ = ep.nE, nP_ = ep.nP, nL_ = ep.nL;
        const uint32_t rinv = ep.rinv;
        const double2* rot = reinterpret_cast<const double2*>(&ep.rot[0][0]);   // [j][0] = {cos, sin}, [j][1] = {B(cos+sin), B(cos-sin)}

        // float64 sums of this thread: {I_E, Q_E, I_P, Q_P, I_L, Q_L} of the data replica, then of the pilot replica(s)
        double acc[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q) acc[q] = 0.0;
        // One 16-byte chunk = 8 consecutive samples.  Everything that enters a sum is float64: the wipe-off by the per-epoch
        // in-chunk rotations (two DFMA per component, the byte -> double bias folded in), the +-1 replica entries as doubles built
        // from their high word, the chunk's carrier phasor from the exact fixed-point phase.  The sums follow the float64 reference
        // to ~1e-14 of |P|, which a minute-long closed loop needs (profiles/r02_parity_60000.md).  Three ways to the table indices:
        //   MODE 0  per-chunk edge prediction: with at least 8 samples per table entry a chunk holds at most one edge per
        //           correlator; the index of the first sample comes from the reference's own expression (exact), the position of
        //           the edge from a 32-bit fixed-point division.  A sample closer than 2^-24 samples to an edge (or a first sample
        //           closer than 2^-26 entries) sends the chunk to MODE 1, so the prediction never decides a close call.
        //   MODE 1  per-sample index from the reference's expression, whole chunk in the left or in the right half of the colon vector
        //   MODE 2  per-sample left / right / middle selection (the chunk holding the middle, or every chunk of a `generic` block)
        auto do_chunk = [&](int c, auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            const int k0 = c * 8 - off;
            const bool masked = (k0 < 0 || k0 + 7 >= blk);        // first / last chunk of the block: samples outside it count as zero
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            if (FMT0) {
                int4 raw;
                if (inBuf) raw = *reinterpret_cast<const int4*>(src + (size_t)c * 16);
                else raw = __ldg(reinterpret_cast<const int4*>(gsrc + (size_t)c * 16));   // not staged: straight from L2
                w[0] = (uint32_t)raw.x ^ 0x80808080u; w[1] = (uint32_t)raw.y ^ 0x80808080u;
                w[2] = (uint32_t)raw.z ^ 0x80808080u; w[3] = (uint32_t)raw.w ^ 0x80808080u;
                if (p.swapIQ) {                                   // GLONASS: rawSignal = Q + 1i*I (GLO tracking.m:227)
#pragma unroll
                    for (int q = 0; q < 4; ++q) w[q] = __byte_perm(w[q], 0u, 0x2301u);
                }
                if (masked) {                                     // a sample outside the block becomes 0x80 0x80, i.e. x = 0
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((unsigned)(k0 + j) >= (unsigned)blk) w[j >> 1] = (j & 1) ? ((w[j >> 1] & 0x0000FFFFu) | 0x80800000u) : ((w[j >> 1] & 0xFFFF0000u) | 0x00008080u);
                }
            }
            // sample j of the chunk as (biased) doubles: rawSignal = I + 1i*Q (tracking.m:233-235), Q + 1i*I for GLONASS
            auto sample = [&](int j, double& mr, double& mq) {
                if (FMT0) {
                    mr = byte_to_biased_double(w[j >> 1], (j & 1) ? 0x7624u : 0x7604u);
                    mq = byte_to_biased_double(w[j >> 1], (j & 1) ? 0x7634u : 0x7614u);
                } else {                                          // int16 and / or real samples (tracking.m:145-149, 229-240)
                    const Rec rec{p.rec, p.fmt};
                    const long long si = min(max(pos + k0 + j, 0LL), p.recSamples - 1);
                    const short2 v = rec.load(si);
                    const bool in = (unsigned)(k0 + j) < (unsigned)blk;
                    mr = in ? (double)(p.swapIQ ? v.y : v.x) : 0.0;
                    mq = in ? (double)(p.swapIQ ? v.x : v.y) : 0.0;
                }
            };
            if constexpr (EXACT) {
                // tracking.m:249-300 sample by sample in float64 as written (the checking mode)
                const double w2pi = __dmul_rn(__dmul_rn(ep.carrFreq, 2.0), 3.141592653589793);         // carrFreq * 2.0 * pi (:281)
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const int kc = k0 + j;
                    if ((unsigned)kc >= (unsigned)blk) continue;
                    double tE, tP, tL;
                    if (!generic) {
                        const bool left = 2 * kc < n, mid = 2 * kc == n;
                        const double st = __dmul_rn((double)(left ? kc : n - kc), left ? d : -d);
                        tE = mid ? mE : __dadd_rn(left ? aE : cE, st);
                        tP = mid ? mP : __dadd_rn(left ? aP : cP, st);
                        tL = mid ? mL : __dadd_rn(left ? aL : cL, st);
                    } else {
                        tE = colon_elem(aE, d, cE, nE_, kc);
                        tP = colon_elem(aP, d, cP, nP_, kc);
                        tL = colon_elem(aL, d, cL, nL_, kc);
                    }
                    const int iE = ceil_idx(tE), iP = ceil_idx(tP), iL = ceil_idx(tL);
                    // time = (0:blksize)./fs; trigarg = ((carrFreq*2.0*pi).*time) + remCarrPhase; carrsig = exp(-1i.*trigarg)  (:280-287)
                    const double trig = __dadd_rn(__dmul_rn(w2pi, __ddiv_rn((double)kc, p.fs)), ep.remCarrPhase);
                    double sn, cs;
                    sincos(trig, &sn, &cs);
                    double xr, xm;
                    {
                        double mr = 0, mq = 0;
                        switch (j) {                              // (compile-time sample numbers for the byte selectors)
                            case 0: sample(0, mr, mq); break; case 1: sample(1, mr, mq); break; case 2: sample(2, mr, mq); break;
                            case 3: sample(3, mr, mq); break; case 4: sample(4, mr, mq); break; case 5: sample(5, mr, mq); break;
                            case 6: sample(6, mr, mq); break; default: sample(7, mr, mq); break;
                        }
                        xr = FMT0 ? mr - kByteBias : mr; xm = FMT0 ? mq - kByteBias : mq;
                    }
                    const double ci = -sn;                                                                // carrsig = cs + 1i*ci
                    const double ur = __dsub_rn(__dmul_rn(cs, xr), __dmul_rn(ci, xm));                  // real(carrsig .* rawSignal) (:291)
                    const double ui = __dadd_rn(__dmul_rn(cs, xm), __dmul_rn(ci, xr));                  // imag(...)                  (:292)
                    const double vE = hi2d(Enc::hi(s_code[iE])), vP = hi2d(Enc::hi(s_code[iP])), vL = hi2d(Enc::hi(s_code[iL]));
                    acc[0] += vE * ur; acc[1] += vE * ui; acc[2] += vP * ur; acc[3] += vP * ui; acc[4] += vL * ur; acc[5] += vL * ui;
                    if constexpr (PILOT) {
                        const double uE = hi2d(Enc::hi(s_pilot[iE])), uP = hi2d(Enc::hi(s_pilot[iP])), uL = hi2d(Enc::hi(s_pilot[iL]));
                        acc[6] += uE * ur; acc[7] += uE * ui; acc[8] += uP * ur; acc[9] += uP * ui; acc[10] += uL * ur; acc[11] += uL * ui;
                    }
                    if constexpr (NSET == 3) {
                        const double wE = hi2d(Enc::hi(s_p61[ceil_idx(__dmul_rn(tE, 6.0))])), wP = hi2d(Enc::hi(s_p61[ceil_idx(__dmul_rn(tP, 6.0))])),
                                     wL = hi2d(Enc::hi(s_p61[ceil_idx(__dmul_rn(tL, 6.0))]));
                        acc[12] += wE * ur; acc[13] += wE * ui; acc[14] += wP * ur; acc[15] += wP * ui; acc[16] += wL * ur; acc[17] += wL * ui;
                    }
                }
                return;
            }
            double ps[NS];                                        // chunk sums before the chunk's carrier phasor
#pragma unroll
            for (int q = 0; q < NS; ++q) ps[q] = 0.0;
            // wipe-off of sample j by the in-chunk rotation and accumulation against the three replica entries (high words)
            auto mac = [&](int j, uint32_t hE, uint32_t hP, uint32_t hL, uint32_t gE, uint32_t gP, uint32_t gL,
                           uint32_t fE, uint32_t fP, uint32_t fL) {
                double mr, mq;
                sample(j, mr, mq);
                const double2 r0 = rot[2 * j], r1 = rot[2 * j + 1];
                // x * e^{-i*j*dphi}   (tracking.m:287-292 with the chunk phase factored out)
                const double ur = __fma_rn(r0.x, mr, __fma_rn(r0.y, mq, -r1.x));
                const double ui = __fma_rn(r0.x, mq, __fma_rn(-r0.y, mr, -r1.y));
                const double vE = hi2d(hE), vP = hi2d(hP), vL = hi2d(hL);
                ps[0] = __fma_rn(vE, ur, ps[0]); ps[1] = __fma_rn(vE, ui, ps[1]);                     // :295-300
                ps[2] = __fma_rn(vP, ur, ps[2]); ps[3] = __fma_rn(vP, ui, ps[3]);
                ps[4] = __fma_rn(vL, ur, ps[4]); ps[5] = __fma_rn(vL, ui, ps[5]);
                if constexpr (PILOT) {                            // same code phase, pilot table (GAL_E1C tracking.m:241-262)
                    const double uE = hi2d(gE), uP = hi2d(gP), uL = hi2d(gL);
                    ps[6] = __fma_rn(uE, ur, ps[6]); ps[7] = __fma_rn(uE, ui, ps[7]);
                    ps[8] = __fma_rn(uP, ur, ps[8]); ps[9] = __fma_rn(uP, ui, ps[9]);
                    ps[10] = __fma_rn(uL, ur, ps[10]); ps[11] = __fma_rn(uL, ui, ps[11]);
                }
                if constexpr (NSET == 3) {                        // pilotBOC61(ceil(tcode * 6) + 1), B1C WB_tracking.m:283,294,305
                    const double wE = hi2d(fE), wP = hi2d(fP), wL = hi2d(fL);
                    ps[12] = __fma_rn(wE, ur, ps[12]); ps[13] = __fma_rn(wE, ui, ps[13]);
                    ps[14] = __fma_rn(wP, ur, ps[14]); ps[15] = __fma_rn(wP, ui, ps[15]);
                    ps[16] = __fma_rn(wL, ur, ps[16]); ps[17] = __fma_rn(wL, ui, ps[17]);
                }
            };
            // per-sample indices from the reference's expression (tracking.m:252-270): ceil(tcode) indexes [c(L) c c(1)] 0-based
            auto by_sample = [&](auto special) {
                constexpr bool SPECIAL = decltype(special)::value;
                const bool allLeft = (2 * (k0 + 7) < n);
                const double sg = allLeft ? 1.0 : -1.0;
                const double f0 = (double)(allLeft ? k0 : (n - k0));
                const double ds = allLeft ? d : -d;
                const double bE = allLeft ? aE : cE, bP = allLeft ? aP : cP, bL = allLeft ? aL : cL;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    double tE, tP, tL;
                    if (!SPECIAL) {
                        // whole chunk in the left half (t = a + k*d) or in the right half (t = c - (n-k)*d).  Samples masked above may
                        // have k < 0 or k >= blk; their code index stays inside the padded table (t > -1 and t < codeLength + 1).
                        const double st = __dmul_rn(__fma_rn(sg, (double)j, f0), ds);   // (+-)(k or n-k)*d, exact integer factor
                        tE = __dadd_rn(bE, st); tP = __dadd_rn(bP, st); tL = __dadd_rn(bL, st);
                    } else {
                        const int kc = min(max(k0 + j, 0), blk - 1);
                        if (!generic) {                           // the three vectors share n: one index conversion
                            const bool left = 2 * kc < n, mid = 2 * kc == n;
                            const double st = __dmul_rn((double)(left ? kc : n - kc), left ? d : -d);
                            tE = mid ? mE : __dadd_rn(left ? aE : cE, st);
                            tP = mid ? mP : __dadd_rn(left ? aP : cP, st);
                            tL = mid ? mL : __dadd_rn(left ? aL : cL, st);
                        } else {
                            tE = colon_elem(aE, d, cE, nE_, kc);
                            tP = colon_elem(aP, d, cP, nP_, kc);
                            tL = colon_elem(aL, d, cL, nL_, kc);
                        }
                    }
                    const int iE = ceil_idx(tE), iP = ceil_idx(tP), iL = ceil_idx(tL);
                    uint32_t gE = 0, gP = 0, gL = 0, fE = 0, fP = 0, fL = 0;
                    if constexpr (PILOT) { gE = Enc::hi(s_pilot[iE]); gP = Enc::hi(s_pilot[iP]); gL = Enc::hi(s_pilot[iL]); }
                    if constexpr (NSET == 3) {
                        fE = Enc::hi(s_p61[ceil_idx(__dmul_rn(tE, 6.0))]); fP = Enc::hi(s_p61[ceil_idx(__dmul_rn(tP, 6.0))]);
                        fL = Enc::hi(s_p61[ceil_idx(__dmul_rn(tL, 6.0))]);
                    }
                    mac(j, Enc::hi(s_code[iE]), Enc::hi(s_code[iP]), Enc::hi(s_code[iL]), gE, gP, gL, fE, fP, fL);
                }
            };
            if constexpr (MODE == 0) {
                const bool allLeft = (2 * (k0 + 7) < n);
                const double st0 = __dmul_rn((double)(allLeft ? k0 : (n - k0)), allLeft ? d : -d);
                const double t0[3] = {__dadd_rn(allLeft ? aE : cE, st0), __dadd_rn(allLeft ? aP : cP, st0), __dadd_rn(allLeft ? aL : cL, st0)};
                int idxA[3], eN[3];
                bool susp = false;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    // t0 + 1.5*2^20 has an ulp of 2^-32: low word = fraction of t0 in 0.32 fixed point, high word = floor(t0) + 2^19
                    const double r = __dadd_rn(t0[q], 1572864.0);
                    const uint32_t lo = (uint32_t)__double2loint(r);
                    const int fl = (int)((uint32_t)__double2hiint(r) & 0xFFFFFu) - 0x80000;
                    idxA[q] = fl + (lo != 0u);                    // ceil(t0): index of the chunk's first sample
                    const uint32_t Q = __umulhi(0u - lo, rinv);   // (distance to the next entry) / d in 8.24 fixed point
                    eN[q] = (int)(Q >> 24) + 1;                   // samples of this chunk that still have the first index
                    susp |= (eN[q] <= 8 && ((Q + 16u) & 0xFFFFFFu) < 32u) || (lo + 64u < 128u);
                }
                if (susp) {
                    by_sample(std::false_type{});
                } else {
                    // running sums S_(j+1) = u_0 + ... + u_j of the wiped-off samples, parked in this thread's column of s_pre; the sum
                    // against a replica whose entry changes from A to B after eN samples is A*S_eN + B*(S_8 - S_eN)
                    double sr = 0.0, si = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        double mr, mq;
                        sample(j, mr, mq);
                        const double2 r0 = rot[2 * j], r1 = rot[2 * j + 1];
                        sr = __dadd_rn(sr, __fma_rn(r0.x, mr, __fma_rn(r0.y, mq, -r1.x)));
                        si = __dadd_rn(si, __fma_rn(r0.x, mq, __fma_rn(-r0.y, mr, -r1.y)));
                        if (j < 7) s_pre[j * kThreads + tid] = make_double2(sr, si);
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        double2 H = make_double2(sr, si);
                        if (eN[q] < 8) H = s_pre[(eN[q] - 1) * kThreads + tid];
                        const double tr_ = __dsub_rn(sr, H.x), ti_ = __dsub_rn(si, H.y);
                        const double vA = hi2d(Enc::hi(s_code[idxA[q]])), vB = hi2d(Enc::hi(s_code[idxA[q] + 1]));
                        ps[2 * q] = __fma_rn(vA, H.x, __dmul_rn(vB, tr_));
                        ps[2 * q + 1] = __fma_rn(vA, H.y, __dmul_rn(vB, ti_));
                        if constexpr (PILOT) {
                            const double uA = hi2d(Enc::hi(s_pilot[idxA[q]])), uB = hi2d(Enc::hi(s_pilot[idxA[q] + 1]));
                            ps[6 + 2 * q] = __fma_rn(uA, H.x, __dmul_rn(uB, tr_));
                            ps[6 + 2 * q + 1] = __fma_rn(uA, H.y, __dmul_rn(uB, ti_));
                        }
                    }
                }
            } else if constexpr (MODE == 1) {
                by_sample(std::false_type{});
            } else {
                by_sample(std::true_type{});
            }
            // rotate the chunk sums by e^{-i*phase(k0)} (float64 phasor from the exact fixed-point phase)
            double s0, c0;
            fix_sincos_f64(phase0 + dphi * (uint64_t)(long long)k0, &s0, &c0);
#pragma unroll
            for (int q = 0; q < NS; q += 2) {
                acc[q] = __fma_rn(c0, ps[q], __fma_rn(s0, ps[q + 1], acc[q]));
                acc[q + 1] = __fma_rn(c0, ps[q + 1], __fma_rn(-s0, ps[q], acc[q + 1]));
            }
        };
        using Mode0 = std::integral_constant<int, 0>;
        using Mode1 = std::integral_constant<int, 1>;
        using Mode2 = std::integral_constant<int, 2>;
        if (!generic) {
            // the one chunk that straddles the middle of the colon vector is left to the thread with
            // the least regular work (a divergent special chunk would otherwise double its warp's time)
            const int cMid = (off + (n >> 1)) >> 3;
            const bool midUniform = (2 * (cMid * 8 - off + 7) < n) || (2 * (cMid * 8 - off) > n);
            if (ep.fast && p.preSlots) {
                for (int c = c_begin + (fits ? tid : tid * G); c < c_end; c += c_step)
                    if (c != cMid || midUniform) do_chunk(c, Mode0{});
            } else {
                for (int c = c_begin + (fits ? tid : tid * G); c < c_end; c += c_step)
                    if (c != cMid || midUniform) do_chunk(c, Mode1{});
            }
            if (tid == kSpecialTid && !midUniform && cMid >= c_begin && cMid < c_end && (fits || cMid % G == (int)crank))
                do_chunk(cMid, Mode2{});
        } else {
            for (int c = c_begin + (fits ? tid : tid * G); c < c_end; c += c_step) do_chunk(c, Mode2{});
        }
        GC_TICK(2)
        // cross-thread reduction in float64: warp shuffle, then warps 0 and 1 over the warp partials
        double v[NS];
#pragma unroll
        for (int q = 0; q < NS; ++q) v[q] = acc[q] * Enc::kScale;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < NS; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
        if (lane == 0)
#pragma unroll
            for (int q = 0; q < NS; ++q) s_part[warp * NS + q] = v[q];
        __syncthreads();
        GC_TICK(3)
        // one-window mode: every thread of this CTA has consumed its samples, the window of the next epoch may land now
        if (p.singleBuf && tid == kLoader && e + 1 < p.nEpochs) prefetch(pos + blk, stage ^ 1, e + 1);
        if (G > 1) {
            // CTA partial -> every CTA of the cluster (slot [epoch parity][my rank]), then one barrier
            double* slot = s_cl + ((e & 1) * kMaxCluster + (int)crank) * NS;
            if (warp == 0) {
#pragma unroll
                for (int q = 0; q < NS; ++q) v[q] = (lane < kWarps) ? s_part[lane * NS + q] : 0.0;
#pragma unroll
                for (int o = kMaxWarps / 2; o > 0; o >>= 1)
#pragma unroll
                    for (int q = 0; q < NS; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
                double b[NS];
#pragma unroll
                for (int q = 0; q < NS; ++q) b[q] = __shfl_sync(0xffffffffu, v[q], 0);
                // pairs of sums per transaction: NS / 2 x G items of 16 bytes (24 for six sums and eight CTAs: one per lane)
                for (int i = lane; i < (NS / 2) * G; i += 32) {
                    const int qp = i % (NS / 2);
                    double v0 = b[0], v1 = b[1];
#pragma unroll
                    for (int t = 1; t < NS / 2; ++t) { v0 = (qp == t) ? b[2 * t] : v0; v1 = (qp == t) ? b[2 * t + 1] : v1; }
                    dsmem_push2(slot + 2 * qp, &s_xbar[e & 1], (uint32_t)(i / (NS / 2)), v0, v1);
                }
            }
        }
        GC_TICK(4)
        if (warp == kPllTid / 32 || warp == kDllTid / 32) {
            NextPhases nx;                                       // NCO phases at the end of this block: known before its sums are
            if (tid == kDllTid) end_phases(p, ep, nx);
            if (G > 1) {
                mbar_wait(&s_xbar[e & 1], xphase[e & 1]);       // all G partial sets have landed here
                xphase[e & 1] ^= 1u;
                GC_TICK(7)
                // every CTA adds the G partials in the same (pairwise) order -> identical sums everywhere
                const double* sl = s_cl + (e & 1) * kMaxCluster * NS;
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                    double t[G];
#pragma unroll
                    for (int r = 0; r < G; ++r) t[r] = sl[r * NS + q];
#pragma unroll
                    for (int w = G / 2; w > 0; w >>= 1)
#pragma unroll
                        for (int r = 0; r < w; ++r) t[r] += t[r + w];
                    v[q] = t[0];
                }
            } else {
#pragma unroll
                for (int q = 0; q < NS; ++q) v[q] = (lane < kWarps) ? s_part[lane * NS + q] : 0.0;
#pragma unroll
                for (int o = kMaxWarps / 2; o > 0; o >>= 1)
#pragma unroll
                    for (int q = 0; q < NS; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
            }
            const double I_E = v[0], Q_E = v[1], I_P = v[2], Q_P = v[3], I_L = v[4], Q_L = v[5];
            double pv[6] = {0, 0, 0, 0, 0, 0};                   // pilot I_E, Q_E, I_P, Q_P, I_L, Q_L as the discriminators see them
            if (PILOT) {
                if (NSET == 3) {
                    // composite pilot: -sqrt(4/33) * p61 +- sqrt(29/33) * p11 cross terms (B1C WB_tracking.m:339-344)
                    const double a61 = -0.3481553119113957, b11 = 0.937436866561092;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        pv[2 * q] = __dadd_rn(__dmul_rn(a61, v[12 + 2 * q]), __dmul_rn(b11, v[7 + 2 * q]));
                        pv[2 * q + 1] = __dsub_rn(__dmul_rn(a61, v[13 + 2 * q]), __dmul_rn(b11, v[6 + 2 * q]));
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 6; ++q) pv[q] = v[6 + q];
                }
            }
            double* sg = s_stage + (e % kStage);
            if (tid == kPllTid) {
                // PLL (tracking.m:305-317)
                // float64 atan / divide as the reference evaluates it (p.exactDisc, the default); GC_PARAM_TRACK_FAST_DISC selects the
                // fp32 form (~1e-7 relative in the recorded pllDiscr row and, through the filter, in carrFreq)
                double carrError = p.exactDisc ? atan(__ddiv_rn(Q_P, I_P)) / kTwoPi
                                               : (double)atanf((float)Q_P / (float)I_P) * 0.15915494309189535;
                if (PILOT) {                                     // GAL_E1C tracking.m:297-300
                    double num = pv[3], den = pv[2];             // Q_P, I_P of the pilot
                    if (p.pilot == 2) {
                        // QI = (I + 1i*Q) * exp(-1i*pi/2), exp(-1i*pi/2) = 6.123e-17 - 1i in float64 (GPS_L5C tracking.m:278-279)
                        const double eps = 6.123233995736766e-17;
                        num = __dsub_rn(__dmul_rn(pv[3], eps), pv[2]);    // imag(QI) = Q*eps - I
                        den = __dadd_rn(__dmul_rn(pv[2], eps), pv[3]);    // real(QI) = I*eps + Q
                    }
                    if (p.pilot == 3) { num = -pv[2]; den = pv[3]; }      // atan(-p11_I_P/p11_Q_P), B1C NB_tracking.m:301
                    const double cP2 = p.exactDisc ? atan(__ddiv_rn(num, den)) / kTwoPi
                                                   : (double)atanf((float)num / (float)den) * 0.15915494309189535;
                    carrError = (p.pilot == 3) ? __ddiv_rn(__dadd_rn(__dmul_rn(carrError, 11.0), __dmul_rn(cP2, 29.0)), 40.0)   // :302
                              : (p.pilot == 5) ? __ddiv_rn(__dadd_rn(carrError, __dmul_rn(cP2, 3.0)), 4.0)                     // B1C WB_tracking.m:356
                                               : __dmul_rn(__dadd_rn(carrError, cP2), 0.5);
                    if (p.pilot >= 2) { sg[GC_F_PILOT_I_P * kStage] = pv[2]; sg[GC_F_PILOT_Q_P * kStage] = pv[3]; }   // Pilot_I_P, Pilot_Q_P (GPS_L5C :323-324)
                    if (p.pilot >= 4) {                          // GPS_L2C tracking.m:396-402; B1C WB_tracking.m:409-414
                        sg[GC_F_PILOT_I_E * kStage] = pv[0]; sg[GC_F_PILOT_I_L * kStage] = pv[4];
                        sg[GC_F_PILOT_Q_E * kStage] = pv[1]; sg[GC_F_PILOT_Q_L * kStage] = pv[5];
                    }
                }
                double carrNco;
                if (p.loopType == 0) {
                    carrNco = __dadd_rn(__dadd_rn(lm.oldCarrNco, __dmul_rn(p.pA, __dsub_rn(carrError, lm.oldCarrError))),
                                        __dmul_rn(carrError, p.pB));
                    lm.oldCarrNco = carrNco; lm.oldCarrError = carrError;
                } else {                                         // GLO_GL1/include/tracking.m:281-285
                    lm.d2CarrError = __dadd_rn(lm.d2CarrError, __dmul_rn(carrError, p.pf3));
                    lm.dCarrError = __dadd_rn(__dadd_rn(lm.d2CarrError, __dmul_rn(carrError, p.pf2)), lm.dCarrError);
                    carrNco = __dadd_rn(lm.dCarrError, __dmul_rn(carrError, p.pf1));
                }
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
            } else if (tid == kDllTid) {
                // DLL (tracking.m:322-335)
                const double pE = __dadd_rn(__dmul_rn(I_E, I_E), __dmul_rn(Q_E, Q_E));
                const double pL = __dadd_rn(__dmul_rn(I_L, I_L), __dmul_rn(Q_L, Q_L));
                double codeError;
                if (p.exactDisc) {
                    const double sE = sqrt(pE), sL = sqrt(pL);
                    codeError = __ddiv_rn(__dsub_rn(sE, sL), __dadd_rn(sE, sL));
                } else {
                    const float sE = sqrtf((float)pE), sL = sqrtf((float)pL);
                    codeError = (double)((sE - sL) / (sE + sL));
                }
                if (PILOT) {                                     // GAL_E1C tracking.m:327-333
                    const double qE = __dadd_rn(__dmul_rn(pv[0], pv[0]), __dmul_rn(pv[1], pv[1]));
                    const double qL = __dadd_rn(__dmul_rn(pv[4], pv[4]), __dmul_rn(pv[5], pv[5]));
                    double ce2;
                    if (p.exactDisc) {
                        const double sE = sqrt(qE), sL = sqrt(qL);
                        ce2 = __ddiv_rn(__dsub_rn(sE, sL), __dadd_rn(sE, sL));
                    } else {
                        const float sE = sqrtf((float)qE), sL = sqrtf((float)qL);
                        ce2 = (double)((sE - sL) / (sE + sL));
                    }
                    if (p.pilot == 3) {                          // B1C NB_tracking.m:313-318: both scaled by (1 - spacing), weights 11/40, 29/40
                        const double sc1 = __dsub_rn(1.0, p.spc);
                        codeError = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(codeError, sc1), 11.0), __dmul_rn(__dmul_rn(ce2, sc1), 29.0)), 40.0);
                    } else if (p.pilot == 5) {                   // B1C WB_tracking.m:366-374: weighted by CalcWeighingFactor.m's factor
                        const double sc1 = __dsub_rn(1.0, p.spc);
                        codeError = __dadd_rn(__dmul_rn(__dmul_rn(codeError, sc1), p.wbFactor), __dmul_rn(__dmul_rn(ce2, sc1), __dsub_rn(1.0, p.wbFactor)));
                    } else {
                        codeError = __dmul_rn(__dadd_rn(codeError, ce2), 0.5);
                    }
                }
                const double codeNco = __dadd_rn(__dadd_rn(lm.oldCodeNco, __dmul_rn(p.cA, __dsub_rn(codeError, lm.oldCodeError))),
                                                 __dmul_rn(codeError, p.cB));
                lm.oldCodeNco = codeNco; lm.oldCodeError = codeError;
                sg[GC_F_DLL_DISCR * kStage] = codeError;                            // :338-339
                sg[GC_F_DLL_DISCR_FILT * kStage] = codeNco;
                // :335 codeFreq of the next block, then its geometry
                plan_epoch(p, __dsub_rn(cinfo.codeFreq0, codeNco), nx.remCodePhase, nx.remCarrPhase, nx.phase0,
                           pos + blk, s_ep[stage ^ 1], invStep);
            }
            if (warp == kPllTid / 32) {                          // the next block's in-chunk rotations, on eight lanes of the carrier warp
                __syncwarp();
                plan_rotations(s_ep[stage ^ 1], lane, FMT0 ? kByteBias : 0.0);
            }
        }
        GC_TICK(5)
        __syncthreads();
        GC_TICK(6)
        // coalesced flush of the staged rows every kStage epochs
        if (leader && (e % kStage) == kStage - 1) {
            const int e0 = e - (kStage - 1);
            for (int i = tid; i < p.nRows * kStage; i += kThreads) {
                const int f = i / kStage, q = i % kStage;
                out[(size_t)f * p.nEpochs + e0 + q] = s_stage[f * kStage + q];
            }
        }
    }
    __syncthreads();
    // tail flush (e = number of completed epochs)
    const int rem = e % kStage;
    if (leader && rem) {
        const int e0 = e - rem;
        for (int i = tid; i < p.nRows * kStage; i += kThreads) {
            const int f = i / kStage, q = i % kStage;
            if (q < rem) out[(size_t)f * p.nEpochs + e0 + q] = s_stage[f * kStage + q];
        }
    }
    if (tid == 0 && leader) p.epochsDone[ch] = e;
    if (timing && blockIdx.x == 0 && (tid == kPllTid || tid == kDllTid || tid == 96))
        for (int i = 0; i < 8; ++i) p.dbg[(tid == kPllTid ? 0 : tid == kDllTid ? 1 : 3) * 8 + i] = tacc[i];
    if (timing && tid == 0)                                      // per-rank view of regular warp 0
        for (int i = 0; i < 8; ++i) p.dbg[32 + blockIdx.x * 8 + i] = tacc[i];
    if (G > 1) cluster_sync_all();                               // nobody leaves while a peer may still push to it
}

size_t track_smem_bytes(int bufBytes, int codeLen, int pilot, int singleBuf, int preThreads)
{
    const int nset = pilot == 5 ? 3 : pilot ? 2 : 1;
    const int ns = 6 * nset;
    size_t s = (singleBuf ? 1 : 2) * (size_t)bufBytes;
    s += (pilot == 5 ? sizeof(int8_t) : sizeof(float)) * ((codeLen + 2 + 2 * kPad + 15) & ~15) * (pilot ? 2 : 1);
    if (pilot == 5) s += (codeLen * 6 + 2 + 2 * kPad61 + 15) & ~15;
    s += sizeof(double) * (kMaxWarps * ns + 2 * kMaxCluster * ns + GC_TRACK_ROWS * kStage);
    s += 2 * sizeof(EpochParams) + sizeof(NextPhases) + 4 * sizeof(uint64_t) + 2 * sizeof(int) + 64;
    s += (size_t)preThreads * 7 * sizeof(double2);               // running sums of the fast chunks
    return s;
}

template <int G, int T, int NSET, typename TT, bool FMT0, bool EXACT = false>
static cudaError_t launch_track_f(const TrackParams& p, int nCh, cudaStream_t stream)
{
    const size_t smem = track_smem_bytes(p.bufBytes, p.codeLen, p.pilot, p.singleBuf, p.preSlots ? T : 0);
    cudaError_t err = cudaFuncSetAttribute(track_kernel<G, T, NSET, TT, FMT0, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nCh * G);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, track_kernel<G, T, NSET, TT, FMT0, EXACT>, p);
}

template <int G, int T, int NSET, typename TT = uint32_t>
static cudaError_t launch_track_g(const TrackParams& p, int nCh, cudaStream_t stream)
{
    // the float64 checking mode reads every format through the per-sample accessor (one instantiation per geometry)
    if (p.exact) return launch_track_f<G, T, NSET, TT, false, true>(p, nCh, stream);
    return p.fmt == 0 ? launch_track_f<G, T, NSET, TT, true>(p, nCh, stream) : launch_track_f<G, T, NSET, TT, false>(p, nCh, stream);
}

// p.bufBytes must be the per-CTA staging size for `cluster` CTAs per channel (track_buf_bytes)
cudaError_t launch_track(const TrackParams& p, int nCh, int cluster, int batch, cudaStream_t stream)
{
    if (batch && cluster == 1 && p.pilot != 5) return p.pilot ? launch_track_g<1, 256, 2>(p, nCh, stream) : launch_track_g<1, 256, 1>(p, nCh, stream);
    if (p.pilot == 5) {                                          // three int8 tables; 18 accumulators want the 352-thread register budget
        if (cluster != 8) return cudaErrorInvalidValue;
        return launch_track_g<8, 352, 3, uint8_t>(p, nCh, stream);
    }
    if (p.pilot) {
        switch (cluster) {
            case 8: return launch_track_g<8, 352, 2>(p, nCh, stream);
            case 4: return launch_track_g<4, 512, 2>(p, nCh, stream);
            case 2: return launch_track_g<2, 512, 2>(p, nCh, stream);
            default: return launch_track_g<1, 512, 2>(p, nCh, stream);
        }
    }
    switch (cluster) {
        case 8: return launch_track_g<8, 352, 1>(p, nCh, stream);
        case 4: return launch_track_g<4, 512, 1>(p, nCh, stream);
        case 2: return launch_track_g<2, 512, 1>(p, nCh, stream);
        default: return launch_track_g<1, 512, 1>(p, nCh, stream);
    }
}

// threads per CTA of the kernel launch_track picks for this cluster size; `batch`: more channels than SMs and one CTA per channel -
// 256-thread CTAs, two to an SM, so that one channel's loop closure (a single-warp section) overlaps another channel's samples
int track_threads(int cluster, int batch) { return cluster == 8 ? 352 : (cluster == 1 && batch) ? 256 : 512; }

// bytes each CTA stages per epoch: ceil(maxChunks / cluster) 16-byte chunks
int track_buf_bytes(int maxBlockSamples, int cluster)
{
    const int maxChunks = (2 * maxBlockSamples + 16 + 15) / 16;
    return ((maxChunks + cluster - 1) / cluster) * 16;
}

// out rows pre-fill (tracking.m:51-77): zeros for absoluteSample and the six I/Q rows, +inf elsewhere
__global__ void track_fill_kernel(double* out, int nCh, int nRows, int nEpochs)
{
    const size_t total = (size_t)nCh * nRows * nEpochs;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)((i / nEpochs) % nRows);
        out[i] = (f == GC_F_ABSOLUTE_SAMPLE || (f >= GC_F_I_P && f <= GC_F_Q_L) || f >= GC_TRACK_NFIELDS) ? 0.0 : inf;
    }
}

// C/N0 by the variance summing method on the recorded prompt rows (Common/CNoVSM.m:38-47; tracking.m:351-358): one thread per
// (channel, VSM interval), the sums in the reference's order
__global__ void cno_vsm_kernel(const double* out, int nCh, int nRows, int nEpochs, int vint, double T, const int32_t* epochsDone,
                               double* vsmValue, double* vsmIndex)
{
    const int nV = nEpochs / vint;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nCh * nV) return;
    const int ch = i / nV, v = i % nV + 1;
    double val = 0.0, idx = 0.0;
    if (v * vint <= epochsDone[ch]) {
        const double* I = out + ((size_t)ch * nRows + GC_F_I_P) * nEpochs + (size_t)(v - 1) * vint;
        const double* Q = out + ((size_t)ch * nRows + GC_F_Q_P) * nEpochs + (size_t)(v - 1) * vint;
        double Zm = 0;
        for (int k = 0; k < vint; ++k) Zm = __dadd_rn(Zm, __dadd_rn(__dmul_rn(I[k], I[k]), __dmul_rn(Q[k], Q[k])));   // Z = I.^2 + Q.^2
        Zm = __ddiv_rn(Zm, (double)vint);                                                                       // mean(Z)
        double Zv = 0;
        for (int k = 0; k < vint; ++k) {
            const double d = __dsub_rn(__dadd_rn(__dmul_rn(I[k], I[k]), __dmul_rn(Q[k], Q[k])), Zm);
            Zv = __dadd_rn(Zv, __dmul_rn(d, d));
        }
        Zv = __ddiv_rn(Zv, (double)(vint - 1));                                                                 // var(Z)
        // MATLAB's sqrt of a negative number is complex: Pav = sqrt(Zm^2 - Zv), Nv = 0.5*(Zm - Pav), CNo = 10*log10(abs((1/T)*Pav/(2*Nv)))
        const double d = __dsub_rn(__dmul_rn(Zm, Zm), Zv);
        const double pr = d >= 0 ? sqrt(d) : 0.0, pi = d >= 0 ? 0.0 : sqrt(-d);
        const double nr = __dmul_rn(0.5, __dsub_rn(Zm, pr)), ni = __dmul_rn(0.5, -pi);
        const double num = __dmul_rn(hypot(pr, pi), __ddiv_rn(1.0, T)), den = __dmul_rn(2.0, hypot(nr, ni));
        val = __dmul_rn(10.0, log10(__ddiv_rn(num, den)));
        idx = (double)(v * vint);
    }
    vsmValue[i] = val;
    vsmIndex[i] = idx;
}

// DataCNo / DataPLD / PilotCNo / PilotPLD / total C/N0 of BDS B2a and B1C (BDS/B2a/include/Calc_CNo_PLD.m:38-100, B1C's twin) with
// the 0.5/0.5 smoothing of tracking.m:409-431: one thread per (channel, interval); it evaluates its own interval and the one before
// (the smoothing partner) from the recorded prompt rows.  pilotMode: 0 none, 1 pilot rows swap roles (Calc_CNo_PLD.m:74-75), 2 as recorded.
__device__ void cno_pld_one(const double* I, const double* Q, int n, double T, double* cno, double* pld)
{
    double Zm = 0, sp = 0, sn = 0, sq = 0;
    for (int k = 0; k < n; ++k) {
        Zm = __dadd_rn(Zm, __dadd_rn(__dmul_rn(I[k], I[k]), __dmul_rn(Q[k], Q[k])));              // Z = I.^2 + Q.^2
        if (I[k] > 0) sp = __dadd_rn(sp, I[k]);
        if (I[k] < 0) sn = __dadd_rn(sn, I[k]);
        sq = __dadd_rn(sq, Q[k]);
    }
    Zm = __ddiv_rn(Zm, (double)n);
    double Zv = 0;
    for (int k = 0; k < n; ++k) {
        const double d = __dsub_rn(__dadd_rn(__dmul_rn(I[k], I[k]), __dmul_rn(Q[k], Q[k])), Zm);
        Zv = __dadd_rn(Zv, __dmul_rn(d, d));
    }
    Zv = __ddiv_rn(Zv, (double)(n - 1));
    const double d = __dsub_rn(__dmul_rn(Zm, Zm), Zv);                                           // Pav = sqrt(Zm^2 - Zv), complex when negative
    const double pr = d >= 0 ? sqrt(d) : 0.0, pi = d >= 0 ? 0.0 : sqrt(-d);
    const double nr = __dmul_rn(0.5, __dsub_rn(Zm, pr)), ni = __dmul_rn(0.5, -pi);               // Nv = 0.5*(Zm - Pav)
    *cno = __ddiv_rn(__dmul_rn(hypot(pr, pi), __ddiv_rn(1.0, T)), __dmul_rn(2.0, hypot(nr, ni)));  // abs((1/T)*Pav/(2*Nv))
    const double a = __dsub_rn(sp, sn), a2 = __dmul_rn(a, a), q2 = __dmul_rn(sq, sq);            // (sum(I>0) - sum(I<0))^2, sum(Q)^2
    *pld = __ddiv_rn(__dsub_rn(a2, q2), __dadd_rn(a2, q2));                                      // NBD / NBP
}

__global__ void cno_pld_kernel(const double* out, int nCh, int nRows, int nEpochs, int vint, double T, int pilotMode,
                               const int32_t* epochsDone, double* res)
{
    const int nV = nEpochs / vint;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nCh * nV) return;
    const int ch = i / nV, v = i % nV + 1;
    double* r = res + (size_t)ch * 5 * nV + (v - 1);
    for (int q = 0; q < 5; ++q) r[(size_t)q * nV] = 0.0;
    if (v * vint > epochsDone[ch]) return;
    double cur[3] = {0, 0, 0}, prev[3] = {0, 0, 0}, pldD = 0, pldP = 0;
    for (int w = (v > 1 ? v - 1 : v); w <= v; ++w) {
        const size_t e0 = (size_t)(w - 1) * vint;
        const double* I = out + ((size_t)ch * nRows + GC_F_I_P) * nEpochs + e0;
        const double* Q = out + ((size_t)ch * nRows + GC_F_Q_P) * nEpochs + e0;
        double val[3] = {0, 0, 0}, dC, dP, pC = 0.0, pP = 0.0;
        cno_pld_one(I, Q, vint, T, &dC, &dP);
        val[0] = __dmul_rn(10.0, log10(dC));
        if (pilotMode) {
            const double* PI = out + ((size_t)ch * nRows + GC_F_PILOT_I_P) * nEpochs + e0;
            const double* PQ = out + ((size_t)ch * nRows + GC_F_PILOT_Q_P) * nEpochs + e0;
            if (pilotMode == 1) cno_pld_one(PQ, PI, vint, T, &pC, &pP);                          // I_P = Pilot_Q_P, Q_P = Pilot_I_P
            else cno_pld_one(PI, PQ, vint, T, &pC, &pP);
            val[1] = __dmul_rn(10.0, log10(pC));
        }
        val[2] = __dmul_rn(10.0, log10(__dadd_rn(dC, pC)));
        if (w == v) { cur[0] = val[0]; cur[1] = val[1]; cur[2] = val[2]; pldD = dP; pldP = pP; }
        else { prev[0] = val[0]; prev[1] = val[1]; prev[2] = val[2]; }
    }
    r[0] = __dadd_rn(__dmul_rn(cur[0], 0.5), __dmul_rn(prev[0], 0.5));
    r[(size_t)nV] = pldD;
    if (pilotMode) {
        r[(size_t)2 * nV] = __dadd_rn(__dmul_rn(cur[1], 0.5), __dmul_rn(prev[1], 0.5));
        r[(size_t)3 * nV] = pldP;
        r[(size_t)4 * nV] = __dadd_rn(__dmul_rn(cur[2], 0.5), __dmul_rn(prev[2], 0.5));
    }
}

cudaError_t launch_cno_pld(const double* out, int nCh, int nRows, int nEpochs, int vint, double T, int pilotMode,
                           const int32_t* epochsDone, double* res, cudaStream_t stream)
{
    const int n = nCh * (nEpochs / vint);
    if (n > 0) cno_pld_kernel<<<(n + 127) / 128, 128, 0, stream>>>(out, nCh, nRows, nEpochs, vint, T, pilotMode, epochsDone, res);
    return cudaGetLastError();
}

cudaError_t launch_cno_vsm(const double* out, int nCh, int nRows, int nEpochs, int vint, double T, const int32_t* epochsDone,
                           double* vsmValue, double* vsmIndex, cudaStream_t stream)
{
    const int n = nCh * (nEpochs / vint);
    if (n > 0) cno_vsm_kernel<<<(n + 127) / 128, 128, 0, stream>>>(out, nCh, nRows, nEpochs, vint, T, epochsDone, vsmValue, vsmIndex);
    return cudaGetLastError();
}

cudaError_t launch_track_fill(double* out, int nCh, int nRows, int nEpochs, cudaStream_t stream)
{
    track_fill_kernel<<<148 * 4, 256, 0, stream>>>(out, nCh, nRows, nEpochs);
    return cudaGetLastError();
}

}  // namespace gc
