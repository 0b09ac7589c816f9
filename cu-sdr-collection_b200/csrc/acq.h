// Host/device interface of the acquisition kernels (internal to libgnsscorr).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rec.h"

namespace gc {

// ---- fused plans L = C x 32 x RB (acq_fused.cu) ------------------------------------------------
struct FusedPlanInfo {
    int L, C, RA, RB, R;
    int pfa;                  // 1: gcd(C, R) = 1, no twiddle between the passes; 0: [C][R] twiddle table needed
    int parts;                // column tiles per (SV, bin) in the partial-maximum arrays
    int C1, C2;               // two-level column pass C = C1 x C2 (0: the column is one register codelet)
};
bool fused_plan_info(int L, FusedPlanInfo* out);   // false: no fused plan for this length

struct FwdColsParams {
    Rec rec;                  // resident record
    long long winStart;       // sample index of longSignal(1) in the record
    int N;                    // samplesPerCode
    int nonCoh;
    int swapIQ;               // GLONASS: rawSignal = Q + 1i*I (GLO_GL1/include/postProcessing.m:94)
    const uint64_t* dphi;     // [nBins] carrier phase increment per sample (turns, 0.64 fixed point)
    const int8_t* codeTab;    // [nReplicas][N] +-1 resampled replicas (code mode)
    float2* out;              // [nRows][C][R]
    const float2* tw;         // [C][R] w_L^(j1*m(p)) (plans with pfa == 0), else unused
};

struct RowsParams {
    float2* X;                // forward: rows transformed in place; inverse: spectra [nBins*nonCoh][C][R]
    const float2* Cc;         // [nReplicas][C][R] conj(FFT(code))/L
    float2* W;                // inverse output [nPrnChunk][nBins][nonCoh][C][R]
    const float2* tw;         // as above
    long long nRows;          // forward only
    int nonCoh, nBins;
    int nRep, repStride;      // inverse: replicas summed per SV (1, or 2 = data + pilot) and their distance in Cc; the
                              // work buffer then holds nonCoh*nRep transforms per (SV, bin), replica index fastest
    int prnPerCta, mPerCta;   // warps of an inverse CTA = binPerCta x prnPerCta x mPerCta (same row j1)
    int zLoop;                // consecutive grid-z positions (block groups of mPerCta, then bins) a warp walks; 0 / 1 = one row per warp
    int mGroups, zTotal;      // filled by the launcher: ceil(nonCoh*nRep / mPerCta), bin groups x mGroups
    const int* slotGroup;     // optional [nSv]: carrier grid of each list slot (GLONASS frequency numbers): the spectra of grid g start
    int groupRows;            // groupRows rows into X (nBins x nonCoh); nullptr = one grid
    int bin0;                 // first bin of this launch: nBins counts the bins of the launch (and of the W layout), the spectra
                              // and binMap are addressed with bin0 + local bin
    int binPerCta;            // bins per CTA (0 = 1); > 1 where a (SV, bin) cell has too few transforms to fill a CTA
    const int2* binMap;       // optional [nBins]: .x = spectrum row in X (instead of bin*nonCoh + block), .y = circular shift of
                              // the spectrum, circshift(IQfreqDom, y) (acquisition variants B and C); nullptr = variant A
    int binMapSlotStride;     // 0: one map for every list slot; else slot s uses binMap[s * binMapSlotStride + bin]
    int nPrnChunk, prnSlot0;  // list slots [prnSlot0, prnSlot0 + nPrnChunk) are processed by this launch
    const int* prnList;       // [nSv] replica index per list slot (index into Cc)
};

struct InvColsParams {
    const float2* W;
    int nBins, nonCoh, nPrnChunk, prnSlot0;
    int bin0, nBinsTotal;     // the launch covers bins [bin0, bin0 + nBins) of nBinsTotal (0 = nBins) in partMax / partIdx
    float* partMax;           // [nSv][nBinsTotal][parts]
    int* partIdx;
    int weighted;             // 1: magnitudes of even / odd transforms are weighted w0 / w1 and the sum scaled by wScale
    float w0, w1, wScale;     // (BDS B1C: (|data|*sqrt(11) + |pilot|*sqrt(29)) / sqrt(40), acquisition.m:213-214)
    float* magOut;            // optional [nPrnChunk][nBins][L]: the summed magnitudes in natural lag order (corrVec of variant B)
    const float2* colTw;      // two-level column pass: [C1][C2] w_C^(-ta*beta) (gc_handle::twCols)
    int persist;              // > 0: that many CTAs walk all (tile, bin, SV) items (register-codelet columns only)
};

// correlation stage as one persistent kernel with an ordered work queue (acq_fused.cu, experimental: GC_ACQ_PATH=queue)
struct QueueParams {
    const float2* X; const float2* Cc; float2* W; const float2* tw;
    int nonCoh, nBins, nRep, repStride;
    int nPrn, prnSlot0;       // cells = nPrn x nBins, bin-major (the spectra of a bin are shared by consecutive cells)
    const int* prnList;
    int nSlots, lag;          // W is a ring of nSlots cells; the column items of a cell are queued `lag` cells after its row items
    float* partMax; int* partIdx; int parts;   // [nSv][nBins][parts], parts = ceil(R / 160)
    float* partial;           // [nSlots][ceil(M/5)][C][R] partial magnitude sums of the column items (L2 resident)
    int* ctrl;                // zeroed by the caller: [0] queue head, [1] abort flag, [2 ..) rowsDone[nCells], colsDone[nCells],
                              // tileDone[nCells * parts]
};
cudaError_t launch_corr_queue(int L, const QueueParams& p, cudaStream_t s);
cudaError_t launch_fwd_cols(int L, const FwdColsParams& p, int nRows, bool codeMode, cudaStream_t s);
cudaError_t launch_fwd_rows(int L, const RowsParams& p, cudaStream_t s);
cudaError_t launch_inv_rows(int L, const RowsParams& p, cudaStream_t s);
cudaError_t launch_finish_replica(float2* Cc, size_t n, int L, cudaStream_t s);
cudaError_t launch_inv_cols(int L, const InvColsParams& p, cudaStream_t s);

// ---- correlation stage as one cluster kernel (acq_cluster.cu) ----------------------------------
struct CorrParams {
    const float2* X;          // [nGroups][nBins*nonCoh][C][R] wiped-off spectra per carrier grid
    const float2* Cc;         // [nReplicas][C][R] conj(FFT(code))/L
    const float2* tw;         // [C][R] (plans with pfa == 0)
    int nonCoh, nBins, nSlots;
    int nRep, repStride;      // replicas summed per SV (1, or 2 = data + pilot) and their distance in Cc
    const int* slotReplica;   // [nSlots] first replica of the SV in list slot s
    const int* slotGroup;     // [nSlots] carrier grid (X group) of slot s
    float* partMax;           // [nSlots][nBins][parts]
    int* partIdx;
};
constexpr int kCorrClusterParts = 4;   // partial maxima per (SV, bin): one per CTA of the largest cluster
cudaError_t launch_corr_cluster(int L, const CorrParams& p, cudaStream_t s);

// ---- generic mixed-radix path (any length whose prime factors are <= 64) --------------------
struct GenericPlan {
    int L;                    // FFT length
    int nf;
    int fac[32];
    const float2* tw;         // [L] w_L^t forward sign
};
// one Stockham pass over `batch` transforms: src -> dst
cudaError_t launch_generic_stage(const GenericPlan& pl, int stage, int n, int s, bool inverse,
                                 const float2* src, float2* dst, long long batch, cudaStream_t st);
cudaError_t launch_generic_wipe(Rec rec, long long winStart, int N, int nonCoh, int nBins, int swapIQ,
                                const uint64_t* dphi, float2* out, int L, cudaStream_t st);
cudaError_t launch_generic_code(const int8_t* codeTab, int N, int nPrn, float2* out, int L, cudaStream_t st);
cudaError_t launch_generic_mul(const float2* X, const float2* Cc, float2* out, int L, long long nKm, cudaStream_t st);
// W2 != nullptr: second replica of a data + pilot pair, abs(ifft(.)) of both summed (GAL_E1C acquisition.m:192)
cudaError_t launch_generic_absacc(const float2* W, const float2* W2, int L, int nBins, int nonCoh, int parts,
                                  float* partMax, int* partIdx, size_t outBase, cudaStream_t st);
cudaError_t launch_generic_conj_scale(float2* Cc, size_t n, float scale, cudaStream_t st);

// ---- acquisition variant B (acq_varb.cu): circularly shifted spectra, best row kept ----------------
struct VarbRow {
    int src;                  // spectrum row in X: shift * nBlocks + block
    int rep;                  // replica row in Cc
    int shift;                // circshift amount = frqBinIndex - 1
    int pad;
};
cudaError_t launch_varb_mulshift(const float2* X, const float2* Cc, const VarbRow* rows, int nRows, float2* out, int L, cudaStream_t st);
cudaError_t launch_varb_rowpeak(const float2* W, int nRows, int L, float* peak, int* idx, cudaStream_t st);
cudaError_t launch_varb_segmax(const float2* W, int nRows, int L, const int4* seg, float* out, cudaStream_t st);
cudaError_t launch_varb_segmax_mag(const float* mag, int nRows, int L, const int4* seg, float* out, cudaStream_t st);
cudaError_t launch_varb_pad(const int8_t* tab, int n, int nRows, float2* out, int L, cudaStream_t st);
// GPS L2C: |sum((x - mean) .* CL segment .* carrier)| for the 75 CL segments (acquisition.m:100-137); codeIdx 1-based [N]
cudaError_t launch_l2c_clphase(Rec rec, long long start, int N, const int8_t* cl, int segLen, const int* codeIdx,
                               uint64_t dphi, double* power, cudaStream_t st);
// variant C (BDS B1C): weighted data + pilot magnitudes per Doppler bin -> (max, first index); one-period fine search
cudaError_t launch_varc_combine(const float2* W, int nBins, int L, int nRep, float* partMax, int* partIdx, size_t outBase, cudaStream_t st);
cudaError_t launch_varc_fine(Rec rec, long long winStart, int N, int nRep, const int8_t* tabs, const int* tabSlot,
                             const int* codePhase, const uint64_t* dphi, int nFine, int nAcq, double* fineResult, cudaStream_t st);

// ---- shared by both paths -------------------------------------------------------------------
struct PeakOut {              // one per PRN slot
    double peak;              // max(max(results))                      acquisition.m:198
    int bin;                  // acqCoarseBin, 1-based                  :196
    int codePhase;            // 1-based                                :198
};
cudaError_t launch_sig_power(Rec rec, long long winStart, int N, double* out, cudaStream_t s);
cudaError_t launch_peak_select(const float* partMax, const int* partIdx, int nPrnSlots, int nBins, int parts,
                               PeakOut* out, cudaStream_t s);

struct FineParams {
    Rec rec;
    long long winStart;
    int N;                    // samplesPerCode
    int nPeriods;             // 40   (acquisition.m:146-148)
    int nFine;                // numOfFineBins (:140)
    int codeLen;              // 1023 (511 GLONASS)
    int swapIQ;               // GLONASS I/Q swap
    int combine;              // 0: max_c |sum of 20 codes| (acquisition.m:243-248); 1: |sum of 10 - sum of next 10| (GLO :246-252);
                              // 2: B3I NH-code / GEO 2-ms-bit search over 20 codes (BDS/B3I/include/acquisition.m:193-211)
                              // 3: Galileo E1 25-chip secondary code, 25 alignments (GAL_E1C/include/acquisition.m:236-252)
                              // 4: pilot secondary code, max over all circular shifts of |sum_q s(q)*sec(q - c)| (GPS_L5C
                              //    acquisition.m:214-219 with NH20, GAL_E5a :211-216 with the per-PRN 100-chip code)
                              // 5: sum_q |s_data(q)| + sum_q |s_pilot(q)| (BDS/B2a/include/acquisition.m:226-228)
    const int* nAcqDev;       // device: number of acquired SVs (written by fine_setup_kernel); entries [nAcq, 2*nAcq) of
                              // chipRow/prod/sums are the pilot-code ones of combine 5
    int nCodes;               // codes wiped off per acquired SV (2 for combine 5, else 1)
    const int8_t* secondary;  // [nAcq][nPeriods] +-1 secondary code (combine 4)
    const int* svId;          // [nAcq] PRN of each acquired SV (combine 2 depends on it)
    const int16_t* chipIdx;   // [nPeriods*N] sample -> chip index of the 40 ms replica (host table, :215-218)
    const int8_t* chips;      // [rows][codeLen] +-1 chips of every SV / component of the handle
    const int* chipRow;       // [entries] row of `chips` each entry wipes off
    const int* codePhase;     // [entries] 1-based coarse code phase (:221)
    const uint64_t* dphi;     // [entries][nFine] fine-bin phase increments
    short2* prod;             // [nAcq][nPeriods*N] scratch: sig40cm .* caCode40ms (:232)
    double* sums;             // [nAcq][nFine][nPeriods][2]  sumPerCode (:235-238)
    int* best;                // [nAcq] arg-max fine bin, 0-based (:253)
    double* fineResult;       // [nAcq][nFine]
    int moments;              // 1: per-code sums through moments around the centre bin (fine_sum_moments_kernel)
};
cudaError_t launch_fine(const FineParams& p, int maxEntries, int maxAcq, cudaStream_t s);

// Threshold test and fine-search set-up on the device (acquisition.m:200-206, 221-227), so that gc_acquire needs no
// host round trip between the coarse and the fine stage.
struct FineSetup {
    const PeakOut* peaks;     // [nSv] from peak_select_kernel
    const double* sigPower;
    int nSv, nonCoh, nFine, nCodes, nPeriods;
    int pilotComp;            // component wiped off when nCodes == 1 (nRep - 1)
    double threshold, step, fineStep, ts;
    const double* slotFreq0;  // [nSv] (IF + offset) + acqSearchBand: frequency of coarse bin 1 for the slot
    const int* slotChipRow;   // [nSv] first chip row of the slot's SV (component 0; component c at +c)
    const int* slotSv;        // [nSv] SV id of the slot
    const int8_t* slotSecondary;   // [nSv][nPeriods] or nullptr
    // outputs
    double* metric;           // [nSv] peakMetric
    int* nAcq;                // [1]
    int* acqSlot;             // [nSv] list slots above threshold, ascending
    int* chipRow; int* codePhase; uint64_t* dphi; int* svId; int8_t* secondary;   // the FineParams arrays
};
cudaError_t launch_fine_setup(const FineSetup& p, cudaStream_t s);

// acqResults of variant A assembled on the device (gc_acquire_device): [peakMetric | codePhase | carrFreq | coarseBin]
struct PackParams {
    const PeakOut* peaks; const double* sigPower;
    const double* slotFreq0;  // [nSv] frequency of coarse bin 1 of the slot
    const int* slotResult;    // [nSv] index of the slot's SV in the result vectors
    const int* best;          // [nAcq] arg-max fine bin per acquired slot, in slot order (FineParams::best)
    int nSv, nonCoh, resultLen, noFine;
    double threshold, step, fineStep;
    double* out;              // [4][resultLen]
};
cudaError_t launch_pack_results(const PackParams& p, cudaStream_t s);

}  // namespace gc
