/*
 * Second, independent CPU restatement of the acquisition variants that gnss_oracle.c does not cover - TEST INFRASTRUCTURE, like
 * the rest of oracle/: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may call it.
 *
 *   orc_acquisition_fam5   GPS/GPS_L5C, GAL/GAL_E5a, GAL/GAL_E5b, BDS/B2a  include/acquisition.m  (variant A, two replicas)
 *   orc_acquisition_varb   BDS/B1I, GPS/GPS_L2C                            include/acquisition.m  (variant B)
 *   orc_acquisition_b1c    BDS/B1C                                         include/acquisition.m  (variant C)
 *
 * Written from the reference's .m files (cited file:line), not from oracle/np_oracle.py: together with the NumPy restatement they
 * are the two witnesses tests/test_oracle.py compares (indices exactly, floats to 1e-8).  The primary codes are inputs (the
 * generators are pinned separately, tests/test_codegen.py); the sampled tables, the search and the fine search are restated here.
 * MATLAB semantics: 1-based indices, first-index max, var with N-1 on complex data, half-away round, unscaled fft / scaled ifft.
 * resamplingflag == 0 only.  "Parity unpinned": the reference ships no vectors and cannot run here (no MATLAB / Octave).
 */
#include "gnss_oracle.c"   /* orc_settings, m_round, the float64 mixed-radix FFT (one translation unit; nothing else is shared) */

enum { ORC_SIG_L5C = 4, ORC_SIG_E5A = 5, ORC_SIG_E5B = 6, ORC_SIG_B2A = 7 };

static double sig_power_of(const cplx* x, int n)            /* sqrt(var(x(1:n)) * n), e.g. GPS_L5C acquisition.m:167 */
{
    cplx mean = 0;
    for (int i = 0; i < n; i++) mean += x[i];
    mean /= n;
    double v = 0;
    for (int i = 0; i < n; i++) { cplx d = x[i] - mean; v += creal(d) * creal(d) + cimag(d) * cimag(d); }
    return sqrt(v / (n - 1) * n);
}
static int small_factors(const fftplan* p)
{
    for (int i = 0; i < p->nf; i++) if (p->fac[i] > 64) return 0;
    return 1;
}
/* [~, row] = max(max(results, [], 2)) and [peak, col] = max(max(results)), both 1-based, first maximal index */
static void peak_2d(const double* results, int nRows, int nCols, int* row, int* col, double* peak)
{
    int r1 = 1; double best = -1;
    for (int k = 0; k < nRows; k++) {
        double rm = results[(size_t)k * nCols];
        for (int n = 1; n < nCols; n++) if (results[(size_t)k * nCols + n] > rm) rm = results[(size_t)k * nCols + n];
        if (rm > best) { best = rm; r1 = k + 1; }
    }
    int c1 = 1; double pk = -1;
    for (int n = 0; n < nCols; n++) {
        double cm = results[n];
        for (int k = 1; k < nRows; k++) if (results[(size_t)k * nCols + n] > cm) cm = results[(size_t)k * nCols + n];
        if (cm > pk) { pk = cm; c1 = n + 1; }
    }
    *row = r1; *col = c1; *peak = pk;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * GPS_L5C/include/acquisition.m:127-290 and its twins GAL_E5a (:120-300), GAL_E5b (:118-250), BDS/B2a (:122-300).
 * iq: longSignal as int8 I,Q pairs from the skip point; it must hold codePhase + nFinePeriods*N samples for every acquired PRN
 * (the reference would index out of range otherwise).  dataCodes / pilotCodes: [nPrn][10230] +-1 chips of generateL5Icode /
 * generateL5Qcode (E5aI/E5aQ, E5bI/E5bQ, B2a data / pilot); secondary: [nPrn][100] generateE5aQ_secondary (E5a only).
 * nRes: length of the result vectors (32 L5C, 50 E5a / E5b, max(acqSatelliteList) B2a). */
int orc_acquisition_fam5(const int8_t* iq, size_t nSamplesAvail, const orc_settings* s, int signal,
                         const int* prnList, int nPrn, const int8_t* dataCodes, const int8_t* pilotCodes, const int8_t* secondary,
                         int nRes, double* carrFreq, double* codePhaseOut, double* peakMetric, int* coarseBin, int* coarseCodePhase)
{
    const int codeLength = (int)s->codeLength;
    const int N = (int)m_round(s->samplingFreq / (s->codeFreqBasis / s->codeLength));        /* GPS_L5C :131-132 samplesPerCode */
    const int L2 = 2 * N;
    const double ts = 1 / s->samplingFreq;                                                   /* :134 */
    const int nBins = (int)m_round(s->acqSearchBand * 2 / s->acqSearchStep) + 1;             /* :139 numberOfFreqBins */
    const int nonCoh = s->acqNonCohTime;
    /* fine acquisition: L5C 20 codes on a 25 Hz grid with the NH20 code (:153-165); E5a 100 codes on a 5 Hz grid with the PRN's
     * secondary code (GAL_E5a :146-158); B2a max(10, acqNonCohTime) codes on a 25 Hz grid, data and pilot (B2a :147-160); E5b none */
    const double fineStep = signal == ORC_SIG_E5A ? 5 : 25;
    const int nFine = (int)m_round(s->acqSearchStep / fineStep) + 1;
    const int nFinePer = signal == ORC_SIG_L5C ? 20 : signal == ORC_SIG_E5A ? 100 : signal == ORC_SIG_B2A ? (10 > nonCoh ? 10 : nonCoh) : 0;
    if (nSamplesAvail < (size_t)(nonCoh + 1) * N) return -1;
    for (int i = 0; i < nRes; i++) { carrFreq[i] = codePhaseOut[i] = peakMetric[i] = 0; coarseBin[i] = coarseCodePhase[i] = 0; }
    cplx* sig = (cplx*)malloc(sizeof(cplx) * nSamplesAvail);
    for (size_t i = 0; i < nSamplesAvail; i++) sig[i] = (double)iq[2 * i] + I * (double)iq[2 * i + 1];
    const double sigPower = sig_power_of(sig, N);                                            /* :167 */
    fftplan plan; plan_make(&plan, L2);
    if (!small_factors(&plan)) { plan_free(&plan); free(sig); return -2; }
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ip = 0; ip < nPrn; ip++) {
        const int PRN = prnList[ip];
        const int8_t* dcode = dataCodes + (size_t)ip * codeLength;
        const int8_t* pcode = pilotCodes + (size_t)ip * codeLength;
        cplx* fI = (cplx*)malloc(sizeof(cplx) * L2), *fQ = (cplx*)malloc(sizeof(cplx) * L2);
        cplx* buf = (cplx*)malloc(sizeof(cplx) * L2), *buf2 = (cplx*)malloc(sizeof(cplx) * L2), *tmp = (cplx*)malloc(sizeof(cplx) * L2);
        cplx* carr = (cplx*)malloc(sizeof(cplx) * L2);
        double* results = (double*)calloc((size_t)nBins * L2, sizeof(double));               /* :180 */
        double* coarseFreqBin = (double*)malloc(sizeof(double) * nBins);
        /* makeL5ITable.m:43-66 / makeL5QTable.m: codeValueIndex = ceil((ts * (1:N)) / tc), last index = codeLength; then the
         * zero-padded local duplicate (:176-177) and its conjugated spectrum (:183-184) */
        const double tc = 1 / s->codeFreqBasis;
        for (int n = 1; n <= N; n++) {
            int idx = (int)ceil((ts * (double)n) / tc);
            if (n == N) idx = codeLength;
            fI[n - 1] = (double)dcode[idx - 1];
            fQ[n - 1] = (double)pcode[idx - 1];
        }
        for (int n = N; n < L2; n++) fI[n] = fQ[n] = 0.0;
        fft_exec(&plan, fI, tmp, -1);
        fft_exec(&plan, fQ, tmp, -1);
        for (int n = 0; n < L2; n++) { fI[n] = conj(fI[n]); fQ[n] = conj(fQ[n]); }
        for (int k = 1; k <= nBins; k++) {                                                   /* :187 */
            coarseFreqBin[k - 1] = s->IF + s->acqSearchBand - s->acqSearchStep * (k - 1);    /* :189-190 */
            for (int n = 0; n < L2; n++) {
                const double a = coarseFreqBin[k - 1] * ((double)n * 2 * M_PI * ts);         /* :136 phasePoints, :192 sigCarr */
                carr[n] = cos(a) - I * sin(a);
            }
            for (int m = 1; m <= nonCoh; m++) {                                              /* :195 */
                const cplx* w = sig + (size_t)(m - 1) * N;                                   /* :197-198 */
                for (int n = 0; n < L2; n++) buf[n] = carr[n] * w[n];                        /* :200-204 */
                fft_exec(&plan, buf, tmp, -1);
                for (int n = 0; n < L2; n++) { buf2[n] = buf[n] * fQ[n]; buf[n] *= fI[n]; }  /* :208-209 */
                fft_exec(&plan, buf, tmp, +1);
                fft_exec(&plan, buf2, tmp, +1);
                double* row = results + (size_t)(k - 1) * L2;
                for (int n = 0; n < L2; n++) row[n] += cabs(buf[n]) / L2 + cabs(buf2[n]) / L2;   /* :212-214 */
            }
        }
        int bin, cp; double peak;
        peak_2d(results, nBins, L2, &bin, &cp, &peak);                                       /* :220-222 */
        const int ri = PRN - 1;
        peakMetric[ri] = peak / sigPower / nonCoh;                                           /* :224 */
        coarseBin[ri] = bin; coarseCodePhase[ri] = cp;
        if (peakMetric[ri] > s->acqThreshold) {                                              /* :228 */
            if (signal == ORC_SIG_E5B) {                                                     /* GAL_E5b :227-229: no fine search */
                carrFreq[ri] = coarseFreqBin[bin - 1];
                codePhaseOut[ri] = cp;
            } else if ((size_t)(cp - 1) + (size_t)nFinePer * N > nSamplesAvail) {
                rc = -4;                                                                     /* (MATLAB: index exceeds array bounds) */
            } else {
                cplx* sum1 = (cplx*)malloc(sizeof(cplx) * nFinePer), *sum2 = (cplx*)malloc(sizeof(cplx) * nFinePer);
                double bestFine = -1, bestFreq = 0;
                for (int j = 1; j <= nFine; j++) {                                           /* :246 */
                    const double f = coarseFreqBin[bin - 1] + s->acqSearchStep / 2 - fineStep * (j - 1);   /* :249-250 */
                    for (int c = 0; c < nFinePer; c++) {                                     /* :259-262 */
                        cplx a1 = 0, a2 = 0;
                        for (int n = 0; n < N; n++) {
                            const long gi = (long)c * N + n;                                 /* 0-based sample of the 20 (100, ...) codes */
                            /* codeValueIndex = floor((ts * (1:n)) / (1/codeFreqBasis)); code(rem(idx, codeLength) + 1)  (:236-238) */
                            const long idx = (long)floor((ts * (double)(gi + 1)) / (1 / s->codeFreqBasis));
                            const double a = f * ((double)gi * 2 * M_PI * ts);               /* :165 finePhasePoints, :252 */
                            const cplx cw = cos(a) - I * sin(a);
                            const cplx x = sig[(size_t)(cp - 1) + gi];                       /* :241 sig20cm */
                            /* basebandSig = longCode .* sigCarr .* sig (:256); L5C / E5a wipe the PILOT code (Q) off, B2a both */
                            a2 += ((double)pcode[idx % codeLength] * cw) * x;
                            if (signal == ORC_SIG_B2A) a1 += ((double)dcode[idx % codeLength] * cw) * x;   /* B2a :258-262 */
                        }
                        sum1[c] = a1; sum2[c] = a2;
                    }
                    double power = 0;
                    if (signal == ORC_SIG_B2A) {                                             /* B2a :273: sum(abs(.)) + sum(abs(.)) */
                        double p1 = 0, p2 = 0;
                        for (int c = 0; c < nFinePer; c++) { p1 += cabs(sum1[c]); p2 += cabs(sum2[c]); }
                        power = p1 + p2;
                    } else {
                        /* the secondary code circularly shifted right by one element per combination (:266-276, GAL_E5a :259-267) */
                        static const double NH[20] = {1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1};   /* :151 */
                        for (int com = 0; com < nFinePer; com++) {
                            cplx t = 0;
                            for (int q = 0; q < nFinePer; q++) {
                                const int src = ((q - com) % nFinePer + nFinePer) % nFinePer;
                                const double sc = signal == ORC_SIG_L5C ? NH[src] : (double)secondary[(size_t)ip * 100 + src];
                                t += sum2[q] * sc;
                            }
                            const double pw = cabs(t);
                            if (pw > power) power = pw;
                        }
                    }
                    if (power > bestFine) { bestFine = power; bestFreq = f; }                /* :283 [~, maxFinBin] = max(FineResult) */
                }
                carrFreq[ri] = bestFreq;                                                     /* :284 */
                codePhaseOut[ri] = cp;                                                       /* :287 */
                if (carrFreq[ri] == 0) carrFreq[ri] = 1;                                     /* :290-292 */
                free(sum1); free(sum2);
            }
        }
        free(fI); free(fQ); free(buf); free(buf2); free(tmp); free(carr); free(results); free(coarseFreqBin);
    }
    plan_free(&plan); free(sig);
    return rc;
}

/* second peak of variant B outside the +-1 chip range around the peak, in the first N1 = samplesPerBlock/Nblocks lags
 * (BDS/B1I acquisition.m:126-140, GPS_L2C :84-99); 1-based ranges exactly as written */
static double varb_second_peak(const double* corrVec, int codePhase, int chip, int N1)
{
    const int e1 = codePhase - chip, e2 = codePhase + chip;
    double m = -1;
    if (e1 < 2) {
        for (int i = e2; i <= N1 + e1; i++) if (corrVec[i - 1] > m) m = corrVec[i - 1];
    } else if (e2 >= N1) {
        for (int i = e2 - N1 + 1; i <= e1; i++) if (corrVec[i - 1] > m) m = corrVec[i - 1];
    } else {
        for (int i = 1; i <= e1; i++) if (corrVec[i - 1] > m) m = corrVec[i - 1];
        for (int i = e2; i <= N1; i++) if (corrVec[i - 1] > m) m = corrVec[i - 1];
    }
    return m;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * BDS/B1I/include/acquisition.m:4-150 (l2c == 0) and GPS/GPS_L2C/include/acquisition.m:4-99 (l2c == 1; the CL code phase search
 * of :100-137 is not restated here).  acqSearchBand is in kHz as in those initSettings.m.  stepSize: settings.stepSize of B1I
 * (0 = empty) resolved as :24-39; settings.acqStep of L2C.  codes: [nPrn][2046] chips (generateCAcode53) or [nPrn][20460] entries of
 * the return-to-zero CM sequence (generateCMcode).  Results: 58 entries (B1I :46-50) or 32 (L2C :28-32), indexed PRN-1. */
int orc_acquisition_varb(const int8_t* iq, size_t nSamplesAvail, const orc_settings* s, int l2c, double stepSizeIn,
                         const int* prnList, int nPrn, const int8_t* codes,
                         double* carrFreq, double* codePhaseOut, double* peakMetric, int* coarseBin, int* coarseCodePhase)
{
    const int Ncodes = 2, Nblocks = l2c ? 2 : 4;                                             /* B1I :5-7 ; L2C :4 */
    const int N = (int)m_round(s->samplingFreq / (s->codeFreqBasis / s->codeLength));        /* samplesPerCode */
    const int spb = l2c ? N * Nblocks                                                        /* L2C :10 */
                        : (int)m_round(s->samplingFreq / (s->codeFreqBasis / (Nblocks * s->codeLength)));   /* B1I :8-9 */
    const int nSig = l2c ? 1 : 2;                                                            /* B1I :12-13 signal1, signal2 */
    const int nRes = l2c ? 32 : 58;
    if (nSamplesAvail < (size_t)spb * nSig) return -1;
    const double ts = 1 / s->samplingFreq;
    const double freqResolution = s->samplingFreq / spb;                                     /* B1I :20 ; L2C :19 */
    const int nBins = (int)m_round(s->acqSearchBand * 1e3 / freqResolution) + 1;             /* B1I :24 ; L2C :21 */
    double stepSize = stepSizeIn;
    if (!l2c) {                                                                              /* B1I :29-49 */
        if (stepSizeIn == 0) stepSize = 0.5 / (Nblocks * s->codeLength / s->codeFreqBasis);
        else if (stepSizeIn != freqResolution) {
            /* steps = 1:0.25:freqResolution/2 with rem(freqResolution, steps) == 0; the one closest to settings.stepSize, the
             * next smaller one when that is larger than asked for */
            double bestDiff = 0, chosen = 0, prevValid = 0;
            int have = 0;
            const int nSteps = (int)floor((freqResolution / 2 - 1) / 0.25 + 1e-9) + 1;
            double* valid = (double*)malloc(sizeof(double) * (nSteps > 0 ? nSteps : 1));
            int nv = 0;
            for (int i = 0; i < nSteps; i++) {
                const double st = 1 + 0.25 * i;
                if (fmod(freqResolution, st) == 0) valid[nv++] = st;
            }
            int minDiv = 0;
            for (int i = 0; i < nv; i++) {
                const double d = fabs(valid[i] - stepSizeIn);
                if (!have || d < bestDiff) { bestDiff = d; minDiv = i; have = 1; }
            }
            (void)prevValid;
            chosen = (valid[minDiv] - stepSizeIn > 0) ? valid[minDiv - 1] : valid[minDiv];
            free(valid);
            stepSize = chosen;
        }
    }
    const int Nshifts = (int)m_round(freqResolution / stepSize);                             /* B1I :53 ; L2C :23 */
    const double initFreq = s->IF + (s->acqSearchBand / 2) * 1000;                           /* B1I :62 ; L2C :34 */
    const int chip = (int)m_round(s->samplingFreq / s->codeFreqBasis);                       /* B1I :126 ; L2C :7 */
    for (int i = 0; i < nRes; i++) { carrFreq[i] = codePhaseOut[i] = peakMetric[i] = 0; coarseBin[i] = coarseCodePhase[i] = 0; }
    cplx* sig = (cplx*)malloc(sizeof(cplx) * (size_t)spb * nSig);
    for (size_t i = 0; i < (size_t)spb * nSig; i++) sig[i] = (double)iq[2 * i] + I * (double)iq[2 * i + 1];
    fftplan plan; plan_make(&plan, spb);
    if (!small_factors(&plan)) { plan_free(&plan); free(sig); return -2; }
    /* the wiped-off spectra do not depend on the PRN: IQfreqDom of every (sub-bin shift, signal block) once (B1I :83-95 ; L2C :52-60) */
    cplx* F = (cplx*)malloc(sizeof(cplx) * (size_t)Nshifts * nSig * spb);
    {
        cplx* tmp = (cplx*)malloc(sizeof(cplx) * spb);
        for (int b = 1; b <= Nshifts; b++) {
            const double f0 = l2c ? initFreq - (b - 1) * (freqResolution / Nshifts)          /* L2C :52 */
                                  : initFreq + (b - 1) * (freqResolution / Nshifts);         /* B1I :83 */
            for (int g = 0; g < nSig; g++) {
                cplx* x = F + ((size_t)(b - 1) * nSig + g) * spb;
                for (int n = 0; n < spb; n++) {
                    const double a = f0 * ((double)n * 2 * M_PI * ts);                       /* phasePoints (B1I :17 ; L2C :17) */
                    x[n] = (cos(a) - I * sin(a)) * sig[(size_t)g * spb + n];
                }
                fft_exec(&plan, x, tmp, -1);
            }
        }
        free(tmp);
    }
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ip = 0; ip < nPrn; ip++) {
        const int PRN = prnList[ip];
        cplx* codeF = (cplx*)malloc(sizeof(cplx) * spb), *buf = (cplx*)malloc(sizeof(cplx) * spb), *tmp = (cplx*)malloc(sizeof(cplx) * spb);
        double* corrVec = (double*)calloc(spb, sizeof(double)), *acq = (double*)malloc(sizeof(double) * spb);
        if (l2c) {
            /* makeCMTable.m: codeValueIndex = ceil((ts * (0:N-1)) / tc), tc = 1/(2*codeFreqBasis); first = 1, last = 2*codeLength;
             * localCode = [cmCodesTable(1:N) zeros(1, N)] (L2C :46-48) */
            const int8_t* cm = codes + (size_t)ip * 20460;
            const double tc = 1 / (s->codeFreqBasis * 2);
            for (int n = 0; n < N; n++) {
                int idx = (int)ceil((ts * (double)n) / tc);
                if (n == N - 1) idx = (int)s->codeLength * 2;
                if (n == 0) idx = 1;
                codeF[n] = (double)cm[idx - 1];
            }
            for (int n = N; n < spb; n++) codeF[n] = 0.0;
        } else {
            /* makeCaTableDMA.m: samplesPerCode = round(fs / (codeFreqBasis / (Ncodes*codeLength))), caCode = [caCode caCode],
             * codeValueIndex = ceil((ts * (1:n)) / tc), last = Ncodes*2046; then [table zeros(1, samplesPerBlock/Ncodes)] (B1I :78) */
            const int8_t* ca = codes + (size_t)ip * 2046;
            const int n2 = (int)m_round(s->samplingFreq / (s->codeFreqBasis / (Ncodes * s->codeLength)));
            const double tc = 1 / s->codeFreqBasis;
            for (int n = 1; n <= n2; n++) {
                int idx = (int)ceil((ts * (double)n) / tc);
                if (n == n2) idx = Ncodes * 2046;
                codeF[n - 1] = (double)ca[(idx - 1) % 2046];
            }
            for (int n = n2; n < spb; n++) codeF[n] = 0.0;
        }
        fft_exec(&plan, codeF, tmp, -1);
        for (int n = 0; n < spb; n++) codeF[n] = conj(codeF[n]);
        double prevmax = 0;
        int freqShift = 0, frequencyBinIndex = 0;
        for (int b = 1; b <= Nshifts; b++) {
            for (int k = 1; k <= nBins; k++) {
                if (k == nBins && b > 1) continue;                                           /* B1I :100-102 ; L2C :66-68 */
                double peakOf[2] = {0, 0};
                for (int g = 0; g < nSig; g++) {
                    const cplx* x = F + ((size_t)(b - 1) * nSig + g) * spb;
                    /* circshift(IQfreqDom, k - 1): element n takes element n - (k-1) (B1I :103-104 ; L2C :70) */
                    for (int n = 0; n < spb; n++) buf[n] = x[((n - (k - 1)) % spb + spb) % spb] * codeF[n];
                    fft_exec(&plan, buf, tmp, +1);
                    double mx = -1;
                    for (int n = 0; n < spb; n++) { const double v = cabs(buf[n]) / spb; if (g == 0) acq[n] = v; if (v > mx) mx = v; }
                    peakOf[g] = mx;
                    if (l2c) {                                                               /* L2C :77-83 */
                        if (mx > prevmax) { prevmax = mx; memcpy(corrVec, acq, sizeof(double) * spb); frequencyBinIndex = k; freqShift = b; }
                    } else if (g == 1) {                                                     /* B1I :116-128 */
                        if (peakOf[0] > prevmax || peakOf[1] > prevmax) {
                            if (peakOf[0] > peakOf[1]) { prevmax = peakOf[0]; memcpy(corrVec, acq, sizeof(double) * spb); }
                            else { prevmax = peakOf[1]; for (int n = 0; n < spb; n++) corrVec[n] = cabs(buf[n]) / spb; }
                            freqShift = b; frequencyBinIndex = k;
                        }
                    }
                }
            }
        }
        int codePhase = 1; double maxPeak = corrVec[0];                                      /* B1I :133 ; L2C :87 */
        for (int n = 1; n < spb; n++) if (corrVec[n] > maxPeak) { maxPeak = corrVec[n]; codePhase = n + 1; }
        const double second = varb_second_peak(corrVec, codePhase, chip, spb / Nblocks);
        const int ri = PRN - 1;
        peakMetric[ri] = maxPeak / second;                                                   /* B1I :157 ; L2C :109 */
        coarseBin[ri] = frequencyBinIndex; coarseCodePhase[ri] = codePhase;
        if (maxPeak / second > s->acqThreshold) {                                            /* B1I :160-167 ; L2C :111-115 */
            codePhaseOut[ri] = codePhase;
            carrFreq[ri] = l2c ? initFreq - freqResolution * (frequencyBinIndex - 1) - (freqResolution / Nshifts) * (freqShift - 1)
                               : initFreq - freqResolution * (frequencyBinIndex - 1) + (freqResolution / Nshifts) * (freqShift - 1);
        }
        free(codeF); free(buf); free(tmp); free(corrVec); free(acq);
    }
    plan_free(&plan); free(sig); free(F);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * BDS/B1C/include/acquisition.m:128-262.  nSamples = length(longSignal) (the fine search moves codePhase back by one code period
 * when it would run off the end, :221-223).  dataBoc / pilotBoc: [nPrn][20460] BOC(1,1) sub-chips of generateDataBOC11 /
 * generatePilotBOC11.  Results: nRes = max(acqSatelliteList) entries indexed PRN-1 (:149-153). */
int orc_acquisition_b1c(const int8_t* iq, size_t nSamples, const orc_settings* s, double acqStep, int acqCohT, int pilotACQflag,
                        const int* prnList, int nPrn, const int8_t* dataBoc, const int8_t* pilotBoc, int nRes,
                        double* carrFreq, double* codePhaseOut, double* peakMetric, int* coarseBin, int* coarseCodePhase)
{
    const int N = (int)m_round(s->samplingFreq / (s->codeFreqBasis / s->codeLength));        /* :128-129 samplesPerCode (10 ms) */
    const int xLen = (int)m_round((double)N / 10 * acqCohT);                                 /* :131 samplesXmsLen */
    const int Lc = (int)m_round((double)N / 10 * (10 + acqCohT));                            /* :134 len10PlusXms */
    if (nSamples < (size_t)Lc) return -1;
    const double ts = 1 / s->samplingFreq;
    const int nBins = (int)m_round(s->acqSearchBand * 2 / acqStep) + 1;                      /* :142 */
    const double fineStep = 25;                                                              /* :159 */
    const int nFine = (int)m_round(acqStep / 25) * 2 + 1;                                    /* :160 */
    for (int i = 0; i < nRes; i++) { carrFreq[i] = codePhaseOut[i] = peakMetric[i] = 0; coarseBin[i] = coarseCodePhase[i] = 0; }
    cplx* sig = (cplx*)malloc(sizeof(cplx) * nSamples);
    for (size_t i = 0; i < nSamples; i++) sig[i] = (double)iq[2 * i] + I * (double)iq[2 * i + 1];
    const double sigPower = sig_power_of(sig, xLen);                                         /* :169 */
    const double initFreq = s->IF + s->acqSearchBand;                                        /* :171 */
    fftplan plan; plan_make(&plan, Lc);
    if (!small_factors(&plan)) { plan_free(&plan); free(sig); return -2; }
    cplx* F = (cplx*)malloc(sizeof(cplx) * Lc);                                              /* :173-178 IQfreqDom, once */
    {
        cplx* tmp = (cplx*)malloc(sizeof(cplx) * Lc);
        for (int n = 0; n < Lc; n++) {
            const double a = initFreq * ((double)n * 2 * M_PI * ts);
            F[n] = (cos(a) - I * sin(a)) * sig[n];
        }
        fft_exec(&plan, F, tmp, -1);
        free(tmp);
    }
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ip = 0; ip < nPrn; ip++) {
        const int PRN = prnList[ip];
        double* dataTab = (double*)malloc(sizeof(double) * N), *pilotTab = (double*)malloc(sizeof(double) * N);
        cplx* fD = (cplx*)malloc(sizeof(cplx) * Lc), *fP = (cplx*)malloc(sizeof(cplx) * Lc);
        cplx* buf = (cplx*)malloc(sizeof(cplx) * Lc), *tmp = (cplx*)malloc(sizeof(cplx) * Lc);
        double* results = (double*)malloc(sizeof(double) * (size_t)nBins * Lc);
        /* makeDataTable.m / makePilotTable.m: codeValueIndex = ceil((ts * (1:N)) / tc), tc = 1/codeFreqBasis/2; first = 1,
         * last = 2*codeLength */
        const double tc = 1 / s->codeFreqBasis / 2;
        for (int n = 1; n <= N; n++) {
            int idx = (int)ceil((ts * (double)n) / tc);
            if (n == N) idx = (int)s->codeLength * 2;
            if (n == 1) idx = 1;
            dataTab[n - 1] = (double)dataBoc[(size_t)ip * 20460 + idx - 1];
            pilotTab[n - 1] = (double)pilotBoc[(size_t)ip * 20460 + idx - 1];
        }
        for (int n = 0; n < Lc; n++) { fD[n] = n < xLen ? dataTab[n] : 0.0; fP[n] = n < xLen ? pilotTab[n] : 0.0; }   /* :186-187, :194-195 */
        fft_exec(&plan, fD, tmp, -1);
        for (int n = 0; n < Lc; n++) fD[n] = conj(fD[n]);                                    /* :190 */
        if (pilotACQflag == 1) { fft_exec(&plan, fP, tmp, -1); for (int n = 0; n < Lc; n++) fP[n] = conj(fP[n]); }   /* :196 */
        for (int k = 1; k <= nBins; k++) {                                                   /* :199 */
            double* row = results + (size_t)(k - 1) * Lc;
            for (int n = 0; n < Lc; n++) buf[n] = F[((n - (k - 1)) % Lc + Lc) % Lc] * fD[n]; /* :200-202 */
            fft_exec(&plan, buf, tmp, +1);
            for (int n = 0; n < Lc; n++) row[n] = cabs(buf[n]) / Lc;                         /* :204 */
            if (pilotACQflag == 1) {                                                         /* :207-212 */
                for (int n = 0; n < Lc; n++) buf[n] = F[((n - (k - 1)) % Lc + Lc) % Lc] * fP[n];
                fft_exec(&plan, buf, tmp, +1);
                for (int n = 0; n < Lc; n++) row[n] = (row[n] * sqrt(11.0) + (cabs(buf[n]) / Lc) * sqrt(29.0)) / sqrt(40.0);
            }
        }
        int bin, cp; double peak;
        peak_2d(results, nBins, Lc, &bin, &cp, &peak);                                       /* :221-225 */
        const double selFreq = initFreq - (bin - 1) * acqStep;                               /* :222 */
        const int ri = PRN - 1;
        peakMetric[ri] = peak / sigPower;                                                    /* :227 */
        coarseBin[ri] = bin; coarseCodePhase[ri] = cp;
        if ((size_t)(cp + N - 1) > nSamples) cp -= N;                                        /* :231-233 */
        if (peakMetric[ri] > s->acqThreshold) {                                              /* :236 */
            double bestFine = -1, bestFreq = 0;
            for (int j = 1; j <= nFine; j++) {                                               /* :252 */
                const double f = selFreq + acqStep - fineStep * (j - 1);                     /* :254 */
                cplx a1 = 0, a2 = 0;
                for (int n = 0; n < N; n++) {
                    const double a = f * ((double)n * 2 * M_PI * ts);                        /* :164 finePhasePoints, :256 */
                    const cplx cw = cos(a) - I * sin(a);
                    const cplx x = sig[(size_t)(cp - 1) + n];                                /* :241 signal0DC */
                    a1 += (x * dataTab[n]) * cw;                                             /* :243, :257 */
                    if (pilotACQflag == 1) a2 += (x * pilotTab[n]) * cw;                     /* :247, :261-262 */
                }
                double r = cabs(a1);
                if (pilotACQflag == 1) r = (r * 11 + cabs(a2) * 29) / 40;
                if (r > bestFine) { bestFine = r; bestFreq = f; }                            /* :268 */
            }
            carrFreq[ri] = bestFreq;                                                         /* :269 */
            if (carrFreq[ri] == 0) carrFreq[ri] = 1;                                         /* :272-274 */
            codePhaseOut[ri] = cp;                                                           /* :275 */
        }
        free(dataTab); free(pilotTab); free(fD); free(fP); free(buf); free(tmp); free(results);
    }
    plan_free(&plan); free(sig); free(F);
    return rc;
}
