/*
 * gnsscorr.h — C ABI of the B200 GNSS correlator engine (libgnsscorr.so).
 *
 * This is the drop-in boundary for the two hot functions of every per-signal folder of
 * gnsscusdr/CU-SDR-Collection.  The reference has no FFI of its own; the interfaces replaced are
 * two MATLAB function signatures, each called once from postProcessing.m:
 *
 *   acqResults              = acquisition(longSignal, settings)   GPS/GPS_L1CA/include/acquisition.m:1
 *                                                                 (called at include/postProcessing.m:100)
 *   [trackResults, channel] = tracking(fid, channel, settings)    GPS/GPS_L1CA/include/tracking.m:1
 *                                                                 (called at include/postProcessing.m:124)
 *
 * A thin MEX gateway (matlab/gnsscorr_mex.c) marshals mxArray <-> these POD arguments; the same
 * entry points are bound from Python with ctypes for all testing.
 *
 * Conventions: plain pointers and sizes only; the caller allocates every output; the library owns
 * all device memory behind the handle; integer return codes (0 = ok, <0 = error, text via
 * gc_last_error), never exceptions across the ABI; one gc_handle = one GPU = one host thread (gc_multi: one handle per
 * GPU with the fan-out inside the library).
 * All indices in results are 1-based exactly as the MATLAB code returns them.
 */
#ifndef GNSSCORR_H
#define GNSSCORR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GC_ABI_VERSION 5

/* signal ids (reference folders).  Implemented: GPS/GPS_L1CA, GLO/GLO_GL1 + GLO/GLO_GL2 (the two
 * GLONASS folders differ only in settings.freqSpacing and the file name), BDS/B3I and GAL/GAL_E1C
 * (E1B + E1C BOC(1,1) replicas summed in acquisition, 25-chip secondary-code fine search, 4 ms
 * data + pilot tracking). */
enum { GC_SIG_GPS_L1CA = 0, GC_SIG_GLO_G1G2 = 1, GC_SIG_BDS_B3I = 2, GC_SIG_GAL_E1C = 3,
       /* the four 10230-chip data + pilot signals (variant A with two replicas, quadrature-pilot tracking): */
       GC_SIG_GPS_L5C = 4, GC_SIG_GAL_E5A = 5, GC_SIG_GAL_E5B = 6, GC_SIG_BDS_B2A = 7,
       /* acquisition variant B (Doppler bins by circshift of one spectrum, best row kept, peak / second-peak
        * metric): BDS B1I (two 4 ms blocks, 1 ms tracking) and GPS L2C (20 ms tracking, with and without the CL pilot).
        * For these acq_search_band is in kHz as in their initSettings.m and acq_search_step is the sub-bin
        * step (settings.stepSize resolved as BDS/B1I/include/acquisition.m:24-39 / settings.acqStep). */
       GC_SIG_BDS_B1I = 8, GC_SIG_GPS_L2C = 9,
       /* acquisition variant C (BDS/B1C/include/acquisition.m:128-276): ONE wipe-off + FFT of (10 + acqCohT) ms, Doppler
        * bins by circshift, data and pilot BOC(1,1) replicas combined (d*sqrt(11) + p*sqrt(29))/sqrt(40), 2-D maximum,
        * 25 Hz fine search over one 10 ms period; NB_tracking.m (pilot_trk_flag 1) and WB_tracking.m (pilot_trk_flag 2). */
       GC_SIG_BDS_B1C = 10 };

/* "no satellite on this channel" for gc_track: GPS uses PRN 0 (tracking.m:136); a GLONASS channel is
 * identified by its frequency number K, for which 0 is valid, so unused channels carry GC_SV_NONE
 * (GLO_GL1/include/tracking.m:137 tests channel.status ~= '-'). */
#define GC_SV_NONE (-2147483647 - 1)

/* error codes */
enum {
    GC_OK = 0,
    GC_ERR_ARG = -1,          /* bad argument / unsupported configuration                          */
    GC_ERR_CUDA = -2,         /* CUDA runtime failure (text in gc_last_error)                      */
    GC_ERR_NO_RECORD = -3,    /* no IF record resident                                             */
    GC_ERR_SHORT_RECORD = -4, /* record shorter than max(42, nonCoh+2) code periods (acquisition)  */
    GC_ERR_UNSUPPORTED = -5,  /* e.g. resamplingflag==1, FFT length with a prime factor > 64       */
    GC_ERR_IO = -6            /* gc_track_file: open/read failure (postProcessing.m:155-158)       */
};

#define GC_FILE_PACKED2 3   /* file_type: two complex samples per byte, values +-1 / +-3 (GPS_L2C/include/unpack_cplx.m:17-20) */

/* POD image of the hot-path fields of the reference's `settings` struct
 * (GPS/GPS_L1CA/initSettings.m:44-136).  Filled by the MATLAB wrapper / Python mirror. */
typedef struct gc_config {
    int32_t abi_version;         /* GC_ABI_VERSION                                                  */
    int32_t device;              /* CUDA device ordinal                                             */
    int32_t signal;              /* GC_SIG_*                                                        */
    int32_t file_type;           /* settings.fileType: 1 = real, 2 = I/Q interleaved (:68); GC_FILE_PACKED2 = the 2-bit packed
                                    I/Q records that include/unpack_cplx.m converts to 'schar' files, decoded on the fly */
    int32_t sample_bytes;        /* settings.dataType: 1 = 'schar', 2 = 'int16' (:63)               */
    int32_t code_length;         /* settings.codeLength (:76)                                       */
    int32_t acq_noncoh_time;     /* settings.acqNonCohTime (:88)                                    */
    int32_t cno_vsm_interval;    /* settings.CNo.VSMinterval (:135)                                 */
    int64_t skip_number_of_bytes;/* settings.skipNumberOfBytes (:56) (the reference multiplies it
                                    by dataAdaptCoeff, i.e. it counts samples)                      */
    double sampling_freq;        /* settings.samplingFreq (:72)                                     */
    double IF;                   /* settings.IF (:71)                                               */
    double code_freq_basis;      /* settings.codeFreqBasis (:73)                                    */
    double acq_search_band;      /* settings.acqSearchBand (:86)                                    */
    double acq_search_step;      /* settings.acqSearchStep (:92)                                    */
    double acq_threshold;        /* settings.acqThreshold (:90)                                     */
    double dll_damping_ratio;    /* settings.dllDampingRatio (:100)                                 */
    double dll_noise_bandwidth;  /* settings.dllNoiseBandwidth (:101)                               */
    double dll_correlator_spacing;/* settings.dllCorrelatorSpacing (:102)                           */
    double pll_damping_ratio;    /* settings.pllDampingRatio (:105)                                 */
    double pll_noise_bandwidth;  /* settings.pllNoiseBandwidth (:106)                               */
    double int_time;             /* settings.intTime (:108)                                         */
    double cno_acc_time;         /* settings.CNo.accTime (:133)                                     */
    double freq_spacing;         /* GLONASS only: settings.freqSpacing, FDMA channel spacing in Hz
                                    (GLO/GLO_GL1/initSettings.m: 562.5e3, GLO_GL2: 437.5e3)         */
    int32_t pilot_trk_flag;      /* settings.pilotTRKflag (GAL/GAL_E1C/initSettings.m:113): 1 = track the
                                    pilot component too and average the discriminators                */
    int32_t acq_coh_t;           /* settings.acqCohT in ms (BDS/B1C/initSettings.m:97; B1C only)                     */
    int32_t pilot_acq_flag;      /* settings.pilotACQflag (BDS/B1C/initSettings.m:74): 1 = pilot replica joins the search */
    int32_t reserved1;
    double carr_freq_basis;      /* settings.carrFreqBasis: the RF carrier of the carrier-aided signals (BDS/B3I/initSettings.m:132, GPS_L5C,
                                    GAL_E5a/E5b, BDS_B2a/B1C); channel.codeFreq = codeFreqBasis + (acquiredFreq - IF) / carrFreqBasis *
                                    codeFreqBasis (GPS_L5C/include/preRun.m:69-71), used by gc_acquire_track; 0 for the other signals   */
} gc_config;

typedef struct gc_handle gc_handle;

/* Number of per-epoch result rows gc_track writes per channel and their order
 * (trackResults fields, GPS/GPS_L1CA/include/tracking.m:48-77). */
#define GC_TRACK_NFIELDS 15
/* With a quadrature pilot tracked (GPS L5C, GAL E5a/E5b, BDS B2a and pilot_trk_flag == 1) two more rows follow:
 * Pilot_I_P, Pilot_Q_P (GPS/GPS_L5C/include/tracking.m:57-60, 323-324); gc_track_nfields() tells which. */
#define GC_TRACK_NFIELDS_PILOT 17
/* GPS L2C with the CL pilot (pilot_trk_flag == 1, GPS/GPS_L2C/include/tracking.m:72-83, 396-402) and BDS B1C full-band
 * tracking (pilot_trk_flag == 2, BDS/B1C/include/WB_tracking.m:60-67, 409-414) record all six pilot correlators:
 * four more rows Pilot_I_E, Pilot_I_L, Pilot_Q_E, Pilot_Q_L. */
#define GC_TRACK_NFIELDS_PILOT6 21
enum {
    GC_F_ABSOLUTE_SAMPLE = 0, GC_F_CODE_FREQ, GC_F_CARR_FREQ, GC_F_I_P, GC_F_I_E, GC_F_I_L,
    GC_F_Q_E, GC_F_Q_P, GC_F_Q_L, GC_F_DLL_DISCR, GC_F_DLL_DISCR_FILT, GC_F_PLL_DISCR,
    GC_F_PLL_DISCR_FILT, GC_F_REM_CODE_PHASE, GC_F_REM_CARR_PHASE,
    GC_F_PILOT_I_P, GC_F_PILOT_Q_P, GC_F_PILOT_I_E, GC_F_PILOT_I_L, GC_F_PILOT_Q_E, GC_F_PILOT_Q_L
};

/* Length of the acqResults vectors for a signal: 32 for GPS L1CA indexed PRN-1 (acquisition.m:130-134),
 * 21 for GLONASS indexed K+7, i.e. MATLAB's K+8 (GLO_GL1/include/acquisition.m:138-142,212),
 * 63 for BeiDou B3I indexed PRN-1 (BDS/B3I/include/acquisition.m:118-122),
 * 50 for Galileo E1 indexed PRN-1 (GAL/GAL_E1C/include/acquisition.m:127-131). */
int gc_acq_result_len(int32_t signal);

/* Create / destroy an engine bound to one GPU.  Builds the FFT plan and twiddle tables for
 * 2*samplesPerCode (acquisition.m:116-122) and the loop coefficients (tracking.m:100-110). */
int  gc_create(gc_handle** out, const gc_config* cfg);
void gc_destroy(gc_handle* h);
/* Text of the last error on this handle (or of the last failed gc_create when h == NULL). */
const char* gc_last_error(const gc_handle* h);

/* Memory-code signals (Galileo E1): the primary codes are ICD tables that the reference reads at run time
 * from include/E1b.dat / include/E1c.dat (GAL/GAL_E1C/include/generateE1Bcode.m:44-55,
 * generateE1Ccode.m), so the caller hands them over instead of the library embedding them:
 *   sv         PRN 1..50
 *   component  0 = data (E1B), 1 = pilot (E1C)
 *   chips      the +-1 primary chips (1 - 2*bit, generateE1Bcode.m:55), nChips == code_length (4092);
 *              the BOC(1,1) sub-carrier (:58-64) is applied by the library.
 * The 10230-chip data + pilot signals (GPS L5C I5/Q5, GAL E5a/E5b I/Q, BDS B2a data/pilot) take their codes
 * the same way - the wrapper passes what the reference's own generateL5Icode / generateL5Qcode /
 * generateE5aIcode ... return (component 0 = data, 1 = pilot, nChips == 10230); GAL E5a also takes
 * component 2 = the PRN's 100-chip pilot secondary code (generateE5aQ_secondary.m) for the fine search.
 * GPS L2C takes component 0 = the 20460-entry return-to-zero CM sequence (generateCMcode.m) and, when pilot_trk_flag == 1,
 * component 1 = the 1534500-entry return-to-zero CL sequence (generateCLcode.m; entries +-1 and 0).  BDS B1C takes
 * components 0 / 1 = the 20460 BOC(1,1) sub-chips of generateDataBOC11.m / generatePilotBOC11.m and, for full-band
 * tracking (pilot_trk_flag == 2), component 2 = the 122760-entry pilot BOC(6,1) sequence of generatePilotBOC61.m.
 * Components that are not set are generated by the library (gc_generate_code below): gc_set_code is an override.  Signals whose
 * codes are always generated (GPS L1CA, GLONASS, B3I) return GC_ERR_ARG. */
int gc_set_code(gc_handle* h, int32_t sv, int32_t component, const int8_t* chips, int32_t nChips);

/* The primary codes themselves.  The reference builds them at run time inside acquisition.m / tracking.m (generateL5Icode.m,
 * generateE5aIcode.m ..., generateB2aDataCode.m, generateCAcode53.m, generateCMcode.m / generateCLcode.m, generateDataBOC11.m /
 * generatePilotBOC11.m / generatePilotBOC61.m with JacobiSymbol.m, generateE1Bcode.m reading E1b.dat); the library has the same
 * generators (csrc/codegen.h, bit-packed registers; ICD tables extracted from those files) and runs them ON THE DEVICE for every
 * (SV, component) the caller has not supplied through gc_set_code - gc_set_code stays as the override.
 *   gc_code_entries          entries of a component in gc_set_code's layout (0 = no such component)
 *   gc_generate_code         one code on the host (no GPU needed): returns the number of entries written, < 0 on error
 *   gc_generate_code_device  the same generators as one kernel on GPU `device`, one thread per SV; out = nSv x entries, host memory
 * Components: as gc_set_code (GAL E1C 0/1 = E1-B / E1-C primary chips; GPS L5C 0/1 = I5 / Q5; GAL E5a, E5b 0/1/2 = I, Q, the Q
 * secondary code; BDS B2a 0/1 = data / pilot; BDS B1I 0; GPS L2C 0/1 = return-to-zero CM / CL; BDS B1C 0/1/2 = data BOC(1,1),
 * pilot BOC(1,1), pilot BOC(6,1)); GPS L1CA, GLONASS and BDS B3I: component 0 (host only). */
int gc_code_entries(int32_t signal, int32_t component);
int gc_generate_code(int32_t signal, int32_t sv, int32_t component, int8_t* out, int32_t nOut);
int gc_generate_code_device(int32_t device, int32_t signal, int32_t nSv, const int32_t* svList, int32_t component, int8_t* out);

/* Scalar settings outside gc_config.  GC_PARAM_B1C_WB_FACTOR: the data-channel weight of the composite code
 * discriminator in B1C full-band tracking, `factor = CalcWeighingFactor(settings)` (BDS/B1C/include/WB_tracking.m:124,
 * CalcWeighingFactor.m:45-82 - adaptive quadrature of the BOC / QMBOC spectra over settings.FEBW, done by the caller);
 * required before gc_track when pilot_trk_flag == 2. */
enum { GC_PARAM_B1C_WB_FACTOR = 1,
       /* 1 = tracking in the float64 CHECKING mode: carrier exp(-1i*trigarg) per sample in float64 from the reference's own
        * expression (tracking.m:280-287), float64 products and sums, rem(trigarg, 2*pi) recurrence for remCarrPhase.  Several times
        * slower; the recorded loop state then follows the float64 reference to ~1e-13 instead of ~1e-10, which the parity tests
        * use to show that what separates the default mode from the reference at 18 Msps is the conditioning of ceil(tcode) alone.
        * Also switched on by GC_TRACK_EXACT_SUMS=1 in the environment when the handle is created. */
       GC_PARAM_TRACK_EXACT_SUMS = 2,
       /* 1 = discriminators atan(Q/I), sqrt, divide evaluated in fp32 (their inputs are fp32-accumulated sums anyway); the default
        * is float64 as the reference evaluates them (tracking.m:305, 322), which costs 1.5 % */
       GC_PARAM_TRACK_FAST_DISC = 3 };
int gc_set_param(gc_handle* h, int32_t key, double value);

/* GPS L2C with pilot_trk_flag == 1: acqResults.CLCodePhase (1..75, 0 = not acquired; GPS_L2C/include/acquisition.m:100-137)
 * of the last gc_acquire, indexed PRN-1 (32 entries), and channel(ch).CLCodePhase for the next gc_track (nCh entries,
 * GPS_L2C/include/tracking.m:162; preRun.m copies it from acqResults). */
int gc_get_cl_code_phase(const gc_handle* h, int32_t* clCodePhase);
int gc_set_cl_code_phase(gc_handle* h, int32_t nCh, const int32_t* clCodePhase);

/* Make an IF record resident in HBM.  `bytes` is the raw file image from byte 0
 * (what fopen/fread see, postProcessing.m:59-96): I,Q interleaved for fileType 2, one value per sample for fileType 1,
 * int8 for dataType 'schar' (sample_bytes 1), little-endian int16 for 'int16' (sample_bytes 2; the dataAdaptCoeff and
 * int16 branches of postProcessing.m:66-96 and tracking.m:141-153, 229-240).  int8 I,Q is the fast path (bulk-copied
 * windows in tracking); the other formats go through a per-sample accessor.
 * _host copies host->device (pinned or pageable memory); _device adopts a device pointer without
 * copying (must be 16-byte aligned with its capacity rounded up to a multiple of 16 bytes, which
 * every cudaMalloc/torch allocation satisfies) — the caller keeps it alive. */
int gc_set_record_host(gc_handle* h, const void* bytes, size_t nbytes);
int gc_set_record_device(gc_handle* h, const void* dptr, size_t nbytes);

/* acquisition(longSignal, settings) on the resident record — replaces
 * GPS/GPS_L1CA/include/acquisition.m:113-292 (resamplingflag must be 0).
 * longSignal = the max(42, nonCoh+2) code periods starting at skip_number_of_bytes
 * (postProcessing.m:74,83-96).  svList = settings.acqSatelliteList (GPS: PRNs 1..32; GLONASS: frequency
 * numbers K = -7..13, GLO_GL1/include/acquisition.m:172-183 — one shared 511-chip replica, carrier grid
 * shifted by -freqSpacing*K per channel, I/Q swapped on input).
 * Outputs (length gc_acq_result_len, indexed PRN-1 / K+7, zero for SVs not searched):
 *   carrFreq, codePhase, peakMetric  = acqResults fields (acquisition.m:130-134,200,254-260)
 *   coarseBin, coarseCodePhase       = acqCoarseBin / codePhase of :196-198 for every searched
 *                                      PRN (1-based; diagnostic, may be NULL). */
int gc_acquire(gc_handle* h, int32_t nSv, const int32_t* svList,
               double* carrFreq, double* codePhase, double* peakMetric,
               int32_t* coarseBin, int32_t* coarseCodePhase);

/* gc_acquire with the results left ON THE DEVICE for a collective that follows (one process per GPU: every rank searches its
 * share of settings.acqSatelliteList and one ncclAllGather of these buffers merges them - entries of SVs a rank did not search
 * are zero, so the merge is a sum).  dResults: device pointer, 4 * gc_acq_result_len doubles laid out
 * [peakMetric | codePhase | carrFreq | coarseBin], each indexed like the host arrays of gc_acquire; written on the handle's
 * stream and complete when the call returns.  Bit-identical to what gc_acquire returns on the host (same expressions). */
int gc_acquire_device(gc_handle* h, int32_t nSv, const int32_t* svList, double* dResults);
/* The same without the final synchronisation: returns as soon as the search is enqueued (one cudaGraphLaunch once the call has
 * been seen twice); dResults is complete in STREAM ORDER on gc_get_stream(h), so a collective enqueued behind that stream (or a
 * stream that waits on it) needs no host round trip between the search and the gather.  svList is copied before the call
 * returns.  gc_get_stats synchronises the stream when it is asked for the timings of such a call.  Signals whose acquisition
 * finishes on the host (variants B and C: BDS B1I, GPS L2C, BDS B1C) behave like gc_acquire_device. */
int gc_acquire_device_async(gc_handle* h, int32_t nSv, const int32_t* svList, double* dResults);

/* Same, but with longSignal supplied from HOST memory in the record's own sample format - int8 I,Q pairs for
 * fileType 2 / 'schar' (what the MEX gateway passes after checking the complex-double longSignal is integer valued),
 * int16 pairs for 'int16', single values for fileType 1; nSamples counts samples, not bytes: copies it to the GPU,
 * runs gc_acquire on it (skip ignored — longSignal already starts at the skip point), copies the
 * results back.  This is the reference-facing call and the one bench.py times as `e2e`. */
int gc_acquire_host(gc_handle* h, const int8_t* iq, size_t nSamples,
                    int32_t nSv, const int32_t* svList,
                    double* carrFreq, double* codePhase, double* peakMetric,
                    int32_t* coarseBin, int32_t* coarseCodePhase);

/* tracking(fid, channel, settings) on the resident record — replaces
 * GPS/GPS_L1CA/include/tracking.m:88-368.
 *   sv[ch]        channel(ch).PRN (0 = channel off, tracking.m:136); GLONASS: channel(ch).K
 *                 (GC_SV_NONE = channel off).  GLONASS uses the 3-coefficient carrier filter of
 *                 Common/calcLoopCoefCarr.m (GLO_GL1/include/tracking.m:281-285) and swaps I/Q (:227).
 *   acqFreq[ch]   channel(ch).acquiredFreq        codePhase[ch]  channel(ch).codePhase (1-based)
 *   codeFreq0[ch] channel(ch).codeFreq, the carrier-aided centre of the code NCO that preRun.m computes
 *                 for B3I (BDS/B3I/include/preRun.m:71-73, tracking.m:57,146); NULL = settings.codeFreqBasis
 *                 (GPS L1CA, GLONASS)
 *   nEpochs       integration periods to process: settings.msToProcess for the 1 ms signals,
 *                 round(msToProcess/1000/intTime) for Galileo E1 (GAL_E1C/include/tracking.m:48)
 *   out           [nCh][gc_track_nfields(h)][nEpochs] doubles; rows pre-filled like tracking.m:51-77
 *                 (zeros for absoluteSample and I/Q, +inf for the rest)
 *   vsmValue/vsmIndex  [nCh][nEpochs / cno_vsm_interval]  (trackResults.CNo, tracking.m:80-83,351-358)
 *   epochsDone    [nCh] completed epochs; < nEpochs means the record ran out (tracking.m:241-245):
 *                 as in the reference the whole call stops there and later channels stay untouched,
 *                 and `status` must be left '-' for every channel with epochsDone < nEpochs. */
int gc_track_nfields(const gc_handle* h);   /* rows per channel in `out`: GC_TRACK_NFIELDS, GC_TRACK_NFIELDS_PILOT or GC_TRACK_NFIELDS_PILOT6 */
int gc_track(gc_handle* h, int32_t nCh, const int32_t* sv, const double* acqFreq,
             const double* codePhase, const double* codeFreq0, int32_t nEpochs,
             double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone);

/* BDS B2a and B1C record DataCNo / DataPLD / PilotCNo / PilotPLD / B2a_CNo (B1C_CNo) every settings.CNoInterval epochs instead of
 * CNo.VSMValue (BDS/B2a/include/tracking.m:409-431 with Calc_CNo_PLD.m:38-100: C/N0 by the variance-summing method and the
 * narrow-band PLL lock detector on the data prompt and - pilot_trk_flag 1: with its I/Q roles swapped, 2: as recorded - the pilot
 * prompt, C/N0 values smoothed 0.5/0.5 with the previous interval's).  gc_track computes them on the device from the rows it has
 * just written (cfg.cno_vsm_interval = settings.CNoInterval); this returns them:
 *   out   [nCh][5][nEpochs / cno_vsm_interval] doubles of the last gc_track: rows DataCNo, DataPLD, PilotCNo, PilotPLD, total C/N0
 *         (zeros for intervals a channel did not complete, and for the pilot rows when pilot_trk_flag == 0) */
#define GC_CNO_PLD_ROWS 5
int gc_get_cno_pld(const gc_handle* h, int32_t nCh, int32_t nIntervals, double* out);

/* Convenience for the MATLAB wrapper: tracking() receives an open fid, which a MEX cannot use;
 * the wrapper recovers the path with fopen(fid) and calls this.  Reads the file (from byte 0)
 * into pinned memory, makes it resident, then gc_track. */
int gc_track_file(gc_handle* h, const char* path,
                  int32_t nCh, const int32_t* sv, const double* acqFreq,
                  const double* codePhase, const double* codeFreq0, int32_t nEpochs,
                  double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone);

/* acquisition -> preRun -> tracking in ONE call, the hand-off done inside the library (postProcessing.m:100-124 with
 * preRun.m:44-72: channels in descending peakMetric order, first index wins ties, at most nChannels of the acquired SVs,
 * the other channels off).  For callers that do not need to stop between the two hot functions (the MATLAB drop-in keeps
 * the two separate signatures).  Every signal: GLONASS channels carry the frequency number K (GLO_GL1/include/preRun.m), the
 * carrier-aided signals get channel.codeFreq from cfg.carr_freq_basis (GPS_L5C/include/preRun.m:69-71, BDS/B3I :71-73), GPS L2C
 * with the CL pilot hands acqResults.CLCodePhase on to the channels.
 *   carrFreq, codePhase, peakMetric   acqResults (length gc_acq_result_len)
 *   chanSv, chanAcqFreq, chanCodePhase   [nChannels] channel(ch).PRN (K) / acquiredFreq / codePhase as preRun.m leaves them
 *                                        (channel off: PRN 0, K GC_SV_NONE)
 *   out, vsmValue, vsmIndex, epochsDone  as gc_track */
int gc_acquire_track(gc_handle* h, int32_t nSv, const int32_t* svList, int32_t nChannels, int32_t nEpochs,
                     double* carrFreq, double* codePhase, double* peakMetric,
                     int32_t* chanSv, double* chanAcqFreq, double* chanCodePhase,
                     double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone);

/* Bit and frame synchronisation front end of postNavigation for GPS L1 C/A - replaces
 * GPS/GPS_L1CA/include/NAVdecoding.m:69-170 (with Common/navPartyChk.m): sign of the prompt outputs, cross-correlation with the
 * 8-bit TLM preamble at 20 values per bit, candidates |xcorr| > 153 with 40 < index < msToProcess - 1199, for every
 * candidate that has another one 6000 ms later the parity of the TLM and HOW words on 20 ms bit sums, and the navigation
 * bits summed from subFrameStart - 20 (GC_NAV_BITS = 1501: the last bit of the previous subframe and five subframes).
 *   I_P            [nCh][nEpochs] trackResults(ch).I_P (host)
 *   subFrameStart  [nCh] first index (1-based) that passes, 0 = 'Could not find valid preambles in channel!' (:143-146)
 *   navBits        [nCh][GC_NAV_BITS] bits as 0/1 (:152-166), zero when bitsValid[ch] == 0
 *   bitsValid      [nCh] 1 when subFrameStart - 20 .. subFrameStart + 29999 lies inside the record (the reference
 *                  indexes out of range otherwise)
 * GPS L1 C/A only (GC_ERR_UNSUPPORTED for the other signals, whose messages have their own framing). */
#define GC_NAV_BITS 1501
int gc_nav_sync(gc_handle* h, int32_t nCh, int32_t nEpochs, const double* I_P,
                int32_t* subFrameStart, uint8_t* navBits, int32_t* bitsValid);

/* ---- several GPUs behind one handle -----------------------------------------------------------------------------------------
 * the PRN loop (acquisition.m:155; frequency numbers for GLONASS, GLO_GL1/include/acquisition.m:172-183) and
 * the channel loop (tracking.m:133) are dealt over nGpus B200s inside the library - one gc_handle, one stream and one host thread
 * per GPU; the SV list round-robin, the channels in contiguous blocks - and the per-SV results are merged on the host, so that
 * the MATLAB caller gets every GPU of the box through the same two calls.  Results are bit-identical to one GPU's.
 *   nGpus <= 0 = every visible device; the devices used are cfg->device .. cfg->device + nGpus - 1.
 * Every function mirrors the single-GPU one of the same name (arguments, result layout, error codes, short-record semantics). */
typedef struct gc_multi gc_multi;
int  gc_multi_create(gc_multi** out, const gc_config* cfg, int32_t nGpus);
void gc_multi_destroy(gc_multi* m);
const char* gc_multi_last_error(const gc_multi* m);
int gc_multi_n_gpus(const gc_multi* m);
gc_handle* gc_multi_handle(gc_multi* m, int32_t gpu);                  /* the per-GPU handle (statistics, stream) */
int gc_multi_set_code(gc_multi* m, int32_t sv, int32_t component, const int8_t* chips, int32_t nChips);
int gc_multi_set_param(gc_multi* m, int32_t key, double value);
int gc_multi_set_cl_code_phase(gc_multi* m, int32_t nCh, const int32_t* clCodePhase);
int gc_multi_get_cl_code_phase(const gc_multi* m, int32_t* clCodePhase);
int gc_multi_set_record_host(gc_multi* m, const void* bytes, size_t nbytes);      /* one H2D copy per GPU, in parallel */
int gc_multi_acquire(gc_multi* m, int32_t nSv, const int32_t* svList, double* carrFreq, double* codePhase, double* peakMetric,
                     int32_t* coarseBin, int32_t* coarseCodePhase);
int gc_multi_acquire_host(gc_multi* m, const int8_t* iq, size_t nSamples, int32_t nSv, const int32_t* svList,
                          double* carrFreq, double* codePhase, double* peakMetric, int32_t* coarseBin, int32_t* coarseCodePhase);
int gc_multi_track(gc_multi* m, int32_t nCh, const int32_t* sv, const double* acqFreq, const double* codePhase,
                   const double* codeFreq0, int32_t nEpochs, double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone);
int gc_multi_track_file(gc_multi* m, const char* path, int32_t nCh, const int32_t* sv, const double* acqFreq,
                        const double* codePhase, const double* codeFreq0, int32_t nEpochs, double* out, double* vsmValue,
                        double* vsmIndex, int32_t* epochsDone);
/* device time of the slowest GPU in the last gc_multi_acquire* / gc_multi_track* call (ms, CUDA events on each GPU's stream) */
int gc_multi_get_times(const gc_multi* m, double* acqMs, double* trackMs);

/* Device-side timing of the most recent gc_acquire / gc_track, measured with CUDA events on the
 * engine's own stream (the stream the kernels are launched on). */
typedef struct gc_stats {
    float acq_total_ms;        /* all acquisition kernels of the last gc_acquire                    */
    float acq_fwd_ms;          /* wipe-off + forward FFTs (PRN independent)                         */
    float acq_corr_ms;         /* spectrum multiply + inverse FFT + |.| + non-coherent sum + argmax */
    float acq_fine_ms;         /* fine-frequency search                                             */
    float track_kernel_ms;     /* correlate-and-dump + loop closure kernel of the last gc_track     */
    int32_t acq_launches;      /* kernels launched by the last gc_acquire                           */
    int32_t track_launches;    /* kernels launched by the last gc_track                             */
    int32_t fft_len;           /* 2*samplesPerCode                                                  */
    int32_t acq_path;          /* 0 = generic mixed-radix passes, 1 = fused C x 32 x RB plan (lengths
                                  32736, 36000, 24000, 32000, 40000, 72000, 144000, 160000, 320000, 360000),
                                  2 = fused plan with the one-kernel cluster correlation stage (GC_ACQ_PATH=cluster),
                                  3 = fused plan with the persistent work-queue correlation kernel (GC_ACQ_PATH=queue) */
    int32_t n_acquired;        /* PRNs above threshold in the last gc_acquire                       */
    float corr_rows_ms;        /* dominant kernel: spectrum multiply + inverse row FFT              */
    float corr_cols_ms;        /* inverse column DFT + |.| + non-coherent sum + row max             */
    int32_t corr_row_launches; /* launches of the dominant (inverse row FFT) kernel in the last gc_acquire */
} gc_stats;
int gc_get_stats(const gc_handle* h, gc_stats* out);

/* The CUDA stream (cudaStream_t) every kernel of this handle is launched on, for callers that want
 * to bracket work with their own events. */
void* gc_get_stream(const gc_handle* h);

/* Library-level: ABI version and the GPU architecture the kernels were compiled for ("sm_100a"). */
int gc_abi_version(void);
const char* gc_build_arch(void);

#ifdef __cplusplus
}
#endif
#endif /* GNSSCORR_H */
