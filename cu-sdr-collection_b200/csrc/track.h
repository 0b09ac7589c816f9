// Host/device interface of the tracking kernels (internal to libgnsscorr).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gnsscorr.h"
#include "rec.h"

#define GC_TRACK_ROWS GC_TRACK_NFIELDS_PILOT6   /* most rows a channel records per epoch (staging size) */

namespace gc {

struct TrackChan {
    int32_t prn;             // PRN (GPS) or frequency number K (GLONASS)
    int32_t pad;             // 1 = channel active, 0 = channel off
    double acqFreq;          // channel.acquiredFreq
    double codeFreq0;        // centre of the code NCO: settings.codeFreqBasis, or channel.codeFreq (B3I tracking.m:57,146)
    long long startSample;   // skipNumberOfBytes + codePhase - 1   (tracking.m:150)
    int32_t clPhase;         // GPS L2C with the CL pilot: channel.CLCodePhase, 1..75 (GPS_L2C tracking.m:162)
    int32_t pad2;
};

struct TrackParams {
    const int8_t* rec;       // resident IF record, 16-byte aligned
    int fmt;                 // Rec::fmt; 0 (int8 I,Q) is staged through shared memory by bulk copies, the other formats
                             // (int16, real: tracking.m:141-153, 229-240) are read through the generic accessor from L2
    long long recSamples;    // samples in the record
    double fs, invFs, codeFreqBasis, codeLength, spc;
    double cA, cB;           // tau2code/tau1code, PDIcode/tau1code  (tracking.m:326)
    double pA, pB;           // tau2carr/tau1carr, PDIcarr/tau1carr  (tracking.m:308)
    double pf1, pf2, pf3;    // carrier filter of Common/calcLoopCoefCarr.m (loopType 1)
    int loopType;            // 0: GPS L1CA second-order form (tracking.m:308); 1: pf3/pf2/pf1 form (GLO tracking.m:281-285)
    int swapIQ;              // GLONASS: rawSignal = Q + 1i*I (GLO tracking.m:227)
    int nEpochs;
    int exactDisc;           // 1: float64 atan/sqrt/divide in the discriminators (GC_TRACK_EXACT_DISC), 0: fp32-seeded
    int exact;               // 1: the float64 checking mode (GC_PARAM_TRACK_EXACT_SUMS): per-sample float64 carrier, products and sums,
                             // rem(trigarg, 2*pi) recurrence; implies exactDisc
    int bufBytes;            // bytes staged per epoch by ONE CTA (multiple of 16)
    int codeLen;             // entries of one code period in the tables: chips x subChip
    int subChip;             // table entries per chip: 1, or 2 for the BOC(1,1) tables of Galileo E1
                             // (tcode*2 in GAL/GAL_E1C/include/tracking.m:236-262)
    int pilot;               // 1: a second (pilot) replica is correlated with the same code phase and both
                             // discriminators are averaged (GAL_E1C tracking.m:241-333, settings.pilotTRKflag);
                             // 2: same, the pilot is in quadrature - its prompt is rotated by -pi/2 before the atan
                             // (GPS_L5C tracking.m:277-281) - and Pilot_I_P / Pilot_Q_P are recorded (rows 15, 16)
                             // 3: quadrature pilot weighted 11/40 : 29/40, pilot discriminator atan(-I/Q), code discriminators
                             // scaled by (1 - spacing) (BDS/B1C/include/NB_tracking.m:300-320); Pilot rows recorded
                             // 4: GPS L2C CL pilot - the pilot table is reloaded every epoch with the CL segment CLCodePhase points at
                             // (pilotTables = [nCh][pilotStride] padded CL sequences), plain average of both discriminators,
                             // six Pilot rows recorded (GPS_L2C tracking.m:259-366)
                             // 5: BDS B1C full band - data BOC(1,1), pilot BOC(1,1) and pilot BOC(6,1) tables (18 sums), composite
                             // pilot correlations, carrier 1:3, code weighted by wbFactor (B1C WB_tracking.m:270-374)
    int singleBuf;           // 1: one sample window in shared memory instead of two (long epochs whose double-buffered window
                             // would not fit next to the code tables: B1C, 10 ms = 180000 samples with two 20460-entry tables)
    int preSlots;            // 1: shared memory holds the [7][threads] running-sum columns of the fast chunks (low-rate codes: at least
                             // 8 samples per table entry), 0: no room - every chunk takes its indices sample by sample
    int nRows;               // rows recorded per epoch: GC_TRACK_NFIELDS, GC_TRACK_NFIELDS_PILOT (pilot 2, 3) or GC_TRACK_NFIELDS_PILOT6 (pilot 4, 5)
    int codeStride;          // bytes between channels in codeTables
    int pilotStride;         // bytes between channels in pilotTables (codeStride, or the padded CL length for pilot == 4)
    int p61Stride;           // bytes between channels in p61Tables
    double wbFactor;         // pilot == 5: weight of the data channel in the code discriminator (CalcWeighingFactor.m)
    const int8_t* p61Tables; // pilot == 5: [nCh][p61Stride] wrapped pilot BOC(6,1) tables, 12 entries per chip
    const int8_t* codeTables;   // [nCh][codeStride]: wrapped +-1 table [c(L) c(1..L) c(1)]
    const int8_t* pilotTables;  // same layout, pilot component (pilot == 1)
    const TrackChan* chans;
    double* out;             // [nCh][nRows][nEpochs]
    int32_t* epochsDone;
    long long* dbg;          // optional [4][8] phase-timing accumulators (GC_TRACK_DEBUG), else nullptr
};

size_t track_smem_bytes(int bufBytes, int codeLen, int pilot, int singleBuf = 0, int preThreads = 0);
int track_threads(int cluster, int batch);
cudaError_t launch_track(const TrackParams& p, int nCh, int cluster, int batch, cudaStream_t stream);
int track_buf_bytes(int maxBlockSamples, int cluster);
cudaError_t launch_track_fill(double* out, int nCh, int nRows, int nEpochs, cudaStream_t stream);
// trackResults.CNo.VSMValue / VSMIndex from the recorded prompt rows, on the device
cudaError_t launch_cno_vsm(const double* out, int nCh, int nRows, int nEpochs, int vint, double T, const int32_t* epochsDone,
                           double* vsmValue, double* vsmIndex, cudaStream_t stream);

// DataCNo / DataPLD / PilotCNo / PilotPLD / total C/N0 of BDS B2a and B1C on the device (Calc_CNo_PLD.m + the smoothing of tracking.m)
cudaError_t launch_cno_pld(const double* out, int nCh, int nRows, int nEpochs, int vint, double T, int pilotMode,
                           const int32_t* epochsDone, double* res, cudaStream_t stream);

}  // namespace gc
