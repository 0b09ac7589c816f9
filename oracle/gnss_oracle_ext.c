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

/* ------------------------------------------------------------------------------------------------------------------------------
 * Tracking of the folders whose loop is BDS/B3I/include/tracking.m's with other codes: GPS/GPS_L5C/include/tracking.m:136-330 (and
 * its twins GAL_E5a, GAL_E5b, BDS/B2a: the same loop with the pilot always on) and BDS/B1I/include/tracking.m:44-160.
 *   dataCodes / pilotCodes  [nCh][codeLength] +-1 primary chips of the channel's PRN (pilotCodes NULL = no pilot)
 *   codeFreq0               channel.codeFreq, the carrier-aided centre of the code NCO (GPS_L5C :150, :299); NULL = settings.codeFreqBasis
 *                           (BDS/B1I tracking.m:52, :139)
 *   quadPilot               1: settings.pilotTRKflag == 1 - the pilot replica on the same code phase, its prompt rotated by exp(-1i*pi/2)
 *                           before the atan, both discriminator pairs averaged, Pilot_I_P / Pilot_Q_P recorded (GPS_L5C :262-268, :277-281,
 *                           :291-295, :323-324)
 *   out                     [nCh][17][nEpochs]: the 15 rows of orc_tracking, then Pilot_I_P, Pilot_Q_P (zero without a pilot)
 * A channel with PRN 0 is skipped (:139); a short read ends the function (:203-207): later channels stay as initialised. */
#define ORC_NFIELDS_PILOT 17
int orc_tracking_codes(const int8_t* iq, size_t nBytes, const orc_settings* s, int nCh, const int* PRN,
                       const double* acquiredFreq, const double* codePhase, const double* codeFreq0,
                       const int8_t* dataCodes, const int8_t* pilotCodes, int quadPilot,
                       int nEpochs, double* out, int* epochsDone)
{
    const double earlyLateSpc = s->dllCorrelatorSpacing;                                     /* :99 */
    const double PDIcode = s->intTime;                                                       /* :102 */
    double tau1code, tau2code;
    calcLoopCoef(s->dllNoiseBandwidth, s->dllDampingRatio, 1.0, &tau1code, &tau2code);       /* :105-106 */
    /* Common/calcLoopCoefCarr.m:41-56 (:110) */
    const double Wn = 1.2 * s->pllNoiseBandwidth;
    const double pf3 = pow(Wn, 3) * pow(s->intTime, 2), pf2 = 2 * pow(Wn, 2) * s->intTime, pf1 = 2 * Wn;
    const int L = (int)s->codeLength;
    const int pilot = quadPilot && pilotCodes != NULL;
    for (int ch = 0; ch < nCh; ch++) {                                                       /* :48-86 */
        double* o = out + (size_t)ch * ORC_NFIELDS_PILOT * nEpochs;
        for (int f = 0; f < ORC_NFIELDS_PILOT; f++) {
            const double fill = (f == 0 || (f >= 3 && f <= 8) || f >= 15) ? 0.0 : INFINITY;
            for (int e = 0; e < nEpochs; e++) o[(size_t)f * nEpochs + e] = fill;
        }
        epochsDone[ch] = 0;
    }
    for (int ch = 0; ch < nCh; ch++) {                                                       /* :136 */
        if (PRN[ch] == 0) continue;                                                          /* :139 */
        double* o = out + (size_t)ch * ORC_NFIELDS_PILOT * nEpochs;
#define F(i) (o + (size_t)(i) * nEpochs)
        size_t pos = (size_t)(2 * ((long)s->skipNumberOfBytes + (long)codePhase[ch] - 1));   /* :141-143 */
        double* code = (double*)malloc(sizeof(double) * (L + 2)), *codeQ = (double*)malloc(sizeof(double) * (L + 2));
        const int8_t* dc = dataCodes + (size_t)ch * L;
        code[0] = dc[L - 1]; for (int i = 0; i < L; i++) code[i + 1] = dc[i]; code[L + 1] = dc[0];          /* :144-145 */
        if (pilot) {
            const int8_t* pc = pilotCodes + (size_t)ch * L;
            codeQ[0] = pc[L - 1]; for (int i = 0; i < L; i++) codeQ[i + 1] = pc[i]; codeQ[L + 1] = pc[0];   /* :147-148 */
        }
        const double centre = codeFreq0 ? codeFreq0[ch] : s->codeFreqBasis;
        double codeFreq = centre, remCodePhase = 0.0;                                        /* :150-151 */
        double carrFreq = acquiredFreq[ch], carrFreqBasis = acquiredFreq[ch], remCarrPhase = 0.0;   /* :152-154 */
        double oldCodeNco = 0, oldCodeError = 0, d2CarrError = 0, dCarrError = 0;            /* :155-158 */
        int stop = 0;
        for (int loopCnt = 1; loopCnt <= nEpochs; loopCnt++) {                               /* :160 */
            F(0)[loopCnt - 1] = (double)pos / 2;                                             /* :181 */
            const double codePhaseStep = codeFreq / s->samplingFreq;                         /* :182 */
            const int blksize = (int)ceil((s->codeLength - remCodePhase) / codePhaseStep);   /* :183 */
            if (pos + 2 * (size_t)blksize > nBytes) { stop = 1; break; }                     /* :192-196 */
            const int8_t* raw = iq + pos; pos += 2 * (size_t)blksize;                        /* :184-191 */
            F(13)[loopCnt - 1] = remCodePhase;                                               /* :197 */
            const double span = (blksize - 1) * codePhaseStep;
            const double aE = remCodePhase - earlyLateSpc, bE = span + remCodePhase - earlyLateSpc;   /* :198-200 */
            const double aL = remCodePhase + earlyLateSpc, bL = span + remCodePhase + earlyLateSpc;   /* :206-208 */
            const double aP = remCodePhase, bP = span + remCodePhase;                                 /* :214-216 */
            int nE_, nL_, nP_; double cE, cL, cP;
            colon_setup(aE, codePhaseStep, bE, &nE_, &cE);
            colon_setup(aL, codePhaseStep, bL, &nL_, &cL);
            colon_setup(aP, codePhaseStep, bP, &nP_, &cP);
            F(14)[loopCnt - 1] = remCarrPhase;                                               /* :223 */
            const double w = carrFreq * 2.0 * M_PI;                                          /* :225 */
            double I_E = 0, Q_E = 0, I_P = 0, Q_P = 0, I_L = 0, Q_L = 0, I_EQ = 0, Q_EQ = 0, I_PQ = 0, Q_PQ = 0, I_LQ = 0, Q_LQ = 0;
            for (int n = 0; n < blksize; n++) {
                const int iE = (int)ceil(colon_elem(aE, codePhaseStep, cE, nE_, n)) + 1;     /* tcode2 = ceil(tcode) + 1 (1-based) */
                const int iL = (int)ceil(colon_elem(aL, codePhaseStep, cL, nL_, n)) + 1;
                const int iP = (int)ceil(colon_elem(aP, codePhaseStep, cP, nP_, n)) + 1;
                const double trig = (w * ((double)n / s->samplingFreq)) + remCarrPhase;      /* :224-225 */
                const double c = cos(trig), sn = sin(trig);                                  /* :227 exp(-1i*trig) */
                const double xr = raw[2 * n], xi = raw[2 * n + 1];                           /* :186-190 */
                const double iBB = c * xr + sn * xi, qBB = c * xi - sn * xr;                 /* :228-229 */
                I_E += code[iE - 1] * iBB; Q_E += code[iE - 1] * qBB;                        /* :230-235 */
                I_P += code[iP - 1] * iBB; Q_P += code[iP - 1] * qBB;
                I_L += code[iL - 1] * iBB; Q_L += code[iL - 1] * qBB;
                if (pilot) {                                                                 /* :236-243 */
                    I_EQ += codeQ[iE - 1] * iBB; Q_EQ += codeQ[iE - 1] * qBB;
                    I_PQ += codeQ[iP - 1] * iBB; Q_PQ += codeQ[iP - 1] * qBB;
                    I_LQ += codeQ[iL - 1] * iBB; Q_LQ += codeQ[iL - 1] * qBB;
                }
            }
            remCodePhase = (colon_elem(aP, codePhaseStep, cP, nP_, blksize - 1) + codePhaseStep) - s->codeLength;   /* :221 */
            remCarrPhase = fmod((w * ((double)blksize / s->samplingFreq)) + remCarrPhase, 2 * M_PI);                /* :226 */
            double carrError = atan(Q_P / I_P) / (2.0 * M_PI);                               /* :245 */
            if (pilot) {                                                                     /* :246-250 */
                /* QI = (I_PQ + 1i*Q_PQ) * exp(-1i*pi/2); exp(-1i*pi/2) = cos(pi/2) - 1i = 6.123233995736766e-17 - 1i in float64 */
                const double er = cos(M_PI / 2), ei = -sin(M_PI / 2);
                const double qiRe = I_PQ * er - Q_PQ * ei, qiIm = I_PQ * ei + Q_PQ * er;
                const double carrErrorQ = atan(qiIm / qiRe) / (2.0 * M_PI);
                carrError = (carrError + carrErrorQ) / 2;
            }
            d2CarrError = d2CarrError + carrError * pf3;                                     /* :251-253 */
            dCarrError = d2CarrError + carrError * pf2 + dCarrError;
            const double carrNco = dCarrError + carrError * pf1;
            F(2)[loopCnt - 1] = carrFreq;                                                    /* :254 */
            carrFreq = carrFreqBasis + carrNco;                                              /* :255 */
            double codeError = (sqrt(I_E * I_E + Q_E * Q_E) - sqrt(I_L * I_L + Q_L * Q_L)) /
                               (sqrt(I_E * I_E + Q_E * Q_E) + sqrt(I_L * I_L + Q_L * Q_L));  /* :256-257 */
            if (pilot) {                                                                     /* :258-262 */
                const double codeErrorQ = (sqrt(I_EQ * I_EQ + Q_EQ * Q_EQ) - sqrt(I_LQ * I_LQ + Q_LQ * Q_LQ)) /
                                          (sqrt(I_EQ * I_EQ + Q_EQ * Q_EQ) + sqrt(I_LQ * I_LQ + Q_LQ * Q_LQ));
                codeError = (codeError + codeErrorQ) / 2;
            }
            const double codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code);   /* :263-264 */
            oldCodeNco = codeNco; oldCodeError = codeError;
            F(1)[loopCnt - 1] = codeFreq;                                                    /* :267 */
            codeFreq = centre - codeNco;                                                     /* :268 ; B1I :139 */
            F(9)[loopCnt - 1] = codeError; F(10)[loopCnt - 1] = codeNco;                     /* :269-272 */
            F(11)[loopCnt - 1] = carrError; F(12)[loopCnt - 1] = carrNco;
            F(4)[loopCnt - 1] = I_E; F(3)[loopCnt - 1] = I_P; F(5)[loopCnt - 1] = I_L;       /* :273-278 */
            F(6)[loopCnt - 1] = Q_E; F(7)[loopCnt - 1] = Q_P; F(8)[loopCnt - 1] = Q_L;
            if (pilot) { F(15)[loopCnt - 1] = I_PQ; F(16)[loopCnt - 1] = Q_PQ; }             /* :280-281 */
            epochsDone[ch] = loopCnt;
        }
        free(code); free(codeQ);
#undef F
        if (stop) break;                                                                     /* `return` (:195) */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * GPS/GPS_L2C/include/tracking.m:45-402: 20 ms epochs in HALF-chip units on the return-to-zero CM sequence (earlyLateSpc*2 :107,
 * codeLength*2 :109, codeFreq = 2*codeFreqBasis :171), fseek to codePhase without the -1 (:153), fractional absoluteSample (:223),
 * halved recorded code quantities (:250, :376, :382-383); with settings.pilotTRKflag the CL pilot on the same code phase, its table
 * the CLCodePhase-th 20 ms segment of the CL sequence, CLCodePhase stepping 1..75 (:253-288, :363-366), discriminators averaged.
 *   cmCodes [nCh][2*codeLength] return-to-zero CM entries; clCodes [nCh][2*CLCodeLength] or NULL; clCodePhase [nCh] (1..75)
 *   out     [nCh][21][nEpochs]: the 15 common rows, then Pilot_I_P, Pilot_Q_P, Pilot_I_E, Pilot_I_L, Pilot_Q_E, Pilot_Q_L */
#define ORC_NFIELDS_PILOT6 21
int orc_tracking_l2c(const int8_t* iq, size_t nBytes, const orc_settings* s, int nCh, const int* PRN,
                     const double* acquiredFreq, const double* codePhase, const int8_t* cmCodes, const int8_t* clCodes,
                     const int* clCodePhase, long CLCodeLength, int nEpochs, double* out, int* epochsDone)
{
    const double earlyLateSpc = s->dllCorrelatorSpacing * 2;                                 /* :107 */
    const int codeLength = (int)s->codeLength * 2;                                           /* :109 */
    const double PDIcode = s->intTime;                                                       /* :112 */
    double tau1code, tau2code;
    calcLoopCoef(s->dllNoiseBandwidth, s->dllDampingRatio, 1.0, &tau1code, &tau2code);       /* :115 */
    const double Wn = 1.2 * s->pllNoiseBandwidth;                                            /* Common/calcLoopCoefCarr.m (:120) */
    const double pf3 = pow(Wn, 3) * pow(s->intTime, 2), pf2 = 2 * pow(Wn, 2) * s->intTime, pf1 = 2 * Wn;
    const int pilot = clCodes != NULL;
    for (int ch = 0; ch < nCh; ch++) {
        double* o = out + (size_t)ch * ORC_NFIELDS_PILOT6 * nEpochs;
        for (int f = 0; f < ORC_NFIELDS_PILOT6; f++) {
            const double fill = (f == 0 || (f >= 3 && f <= 8) || f >= 15) ? 0.0 : INFINITY;
            for (int e = 0; e < nEpochs; e++) o[(size_t)f * nEpochs + e] = fill;
        }
        epochsDone[ch] = 0;
    }
    for (int ch = 0; ch < nCh; ch++) {                                                       /* :144 */
        if (PRN[ch] == 0) continue;                                                          /* :147 */
        double* o = out + (size_t)ch * ORC_NFIELDS_PILOT6 * nEpochs;
#define F(i) (o + (size_t)(i) * nEpochs)
        size_t pos = (size_t)(2 * ((long)s->skipNumberOfBytes + (long)codePhase[ch]));       /* :153 */
        const int8_t* cm = cmCodes + (size_t)ch * codeLength;
        double* cmCode = (double*)malloc(sizeof(double) * (codeLength + 2));
        cmCode[0] = cm[codeLength - 1]; for (int i = 0; i < codeLength; i++) cmCode[i + 1] = cm[i]; cmCode[codeLength + 1] = cm[0];   /* :157 */
        const long clLen = 2 * CLCodeLength;
        const int8_t* cl = pilot ? clCodes + (size_t)ch * clLen : NULL;                      /* CLCode = [CLCode(end) CLCode CLCode(1)] (:165) */
        int CLCodePhase = pilot ? clCodePhase[ch] : 0;                                       /* :162 */
        double codeFreq = s->codeFreqBasis * 2, remCodePhase = 0.0;                          /* :171-173 */
        double carrFreq = acquiredFreq[ch], carrFreqBasis = acquiredFreq[ch], remCarrPhase = 0.0;
        double oldCodeNco = 0, oldCodeError = 0, d2CarrError = 0, dCarrError = 0;
        int stop = 0;
        for (int loopCnt = 1; loopCnt <= nEpochs; loopCnt++) {                               /* :190 */
            const double codePhaseStep = codeFreq / s->samplingFreq;                         /* :219 */
            F(0)[loopCnt - 1] = ((double)pos / 2) / 1 + 1 - remCodePhase / codePhaseStep;    /* :223 */
            const int blksize = (int)ceil((codeLength - remCodePhase) / codePhaseStep);      /* :226 */
            if (pos + 2 * (size_t)blksize > nBytes) { stop = 1; break; }                     /* :243-247 */
            const int8_t* raw = iq + pos; pos += 2 * (size_t)blksize;
            F(13)[loopCnt - 1] = remCodePhase / 2;                                           /* :250 */
            const double span = (blksize - 1) * codePhaseStep;
            const double aE = remCodePhase - earlyLateSpc, bE = span + remCodePhase - earlyLateSpc;   /* :253-255 */
            const double aL = remCodePhase + earlyLateSpc, bL = span + remCodePhase + earlyLateSpc;   /* :264-266 */
            const double aP = remCodePhase, bP = span + remCodePhase;                                 /* :275-277 */
            int nE_, nL_, nP_; double cE, cL, cP;
            colon_setup(aE, codePhaseStep, bE, &nE_, &cE);
            colon_setup(aL, codePhaseStep, bL, &nL_, &cL);
            colon_setup(aP, codePhaseStep, bP, &nP_, &cP);
            F(14)[loopCnt - 1] = remCarrPhase;                                               /* :291 */
            const double w = carrFreq * 2.0 * M_PI;
            double I_E = 0, Q_E = 0, I_P = 0, Q_P = 0, I_L = 0, Q_L = 0, I_ECL = 0, Q_ECL = 0, I_PCL = 0, Q_PCL = 0, I_LCL = 0, Q_LCL = 0;
            for (int n = 0; n < blksize; n++) {
                const int iE = (int)ceil(colon_elem(aE, codePhaseStep, cE, nE_, n)) + 1;     /* tcode2, 1-based */
                const int iL = (int)ceil(colon_elem(aL, codePhaseStep, cL, nL_, n)) + 1;
                const int iP = (int)ceil(colon_elem(aP, codePhaseStep, cP, nP_, n)) + 1;
                const double trig = (w * ((double)n / s->samplingFreq)) + remCarrPhase;      /* :293-295 */
                const double c = cos(trig), sn = sin(trig);
                const double xr = raw[2 * n], xi = raw[2 * n + 1];
                const double iBB = c * xr + sn * xi, qBB = c * xi - sn * xr;                 /* :301-302 */
                I_E += cmCode[iE - 1] * iBB; Q_E += cmCode[iE - 1] * qBB;                    /* :305-310 */
                I_P += cmCode[iP - 1] * iBB; Q_P += cmCode[iP - 1] * qBB;
                I_L += cmCode[iL - 1] * iBB; Q_L += cmCode[iL - 1] * qBB;
                if (pilot) {
                    /* CLCode(tcode2 + codeLength*(CLCodePhase-1)) in the padded sequence [CL(end) CL CL(1)] (:258, :269, :280) */
                    const long base = (long)codeLength * (CLCodePhase - 1);
#define CLV(i1) ({ long k_ = base + (i1); (double)(k_ == 1 ? cl[clLen - 1] : k_ == clLen + 2 ? cl[0] : cl[k_ - 2]); })
                    const double e = CLV(iE), l = CLV(iL), p = CLV(iP);
#undef CLV
                    I_ECL += e * iBB; Q_ECL += e * qBB; I_PCL += p * iBB; Q_PCL += p * qBB; I_LCL += l * iBB; Q_LCL += l * qBB;   /* :313-318 */
                }
            }
            remCodePhase = (colon_elem(aP, codePhaseStep, cP, nP_, blksize - 1) + codePhaseStep) - codeLength;   /* :286 */
            remCarrPhase = fmod((w * ((double)blksize / s->samplingFreq)) + remCarrPhase, 2 * M_PI);             /* :297 */
            double carrError = atan(Q_P / I_P) / (2.0 * M_PI);                               /* :324 */
            if (pilot) carrError = (carrError + atan(Q_PCL / I_PCL) / (2.0 * M_PI)) / 2;     /* :327-329 */
            d2CarrError = d2CarrError + carrError * pf3;                                     /* :333-335 */
            dCarrError = d2CarrError + carrError * pf2 + dCarrError;
            const double carrNco = dCarrError + carrError * pf1;
            F(2)[loopCnt - 1] = carrFreq;                                                    /* :338 */
            carrFreq = carrFreqBasis + carrNco;
            double codeError = (sqrt(I_E * I_E + Q_E * Q_E) - sqrt(I_L * I_L + Q_L * Q_L)) /
                               (sqrt(I_E * I_E + Q_E * Q_E) + sqrt(I_L * I_L + Q_L * Q_L));  /* :346-347 */
            if (pilot) {                                                                     /* :350-366 */
                const double ceCL = (sqrt(I_ECL * I_ECL + Q_ECL * Q_ECL) - sqrt(I_LCL * I_LCL + Q_LCL * Q_LCL)) /
                                    (sqrt(I_ECL * I_ECL + Q_ECL * Q_ECL) + sqrt(I_LCL * I_LCL + Q_LCL * Q_LCL));
                codeError = (codeError + ceCL) / 2;
                CLCodePhase = CLCodePhase + 1;
                if (CLCodePhase >= 76) CLCodePhase = 1;
            }
            const double codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code);   /* :369-370 */
            oldCodeNco = codeNco; oldCodeError = codeError;
            F(1)[loopCnt - 1] = codeFreq / 2;                                                /* :376 */
            codeFreq = s->codeFreqBasis * 2 - codeNco;                                       /* :377 */
            F(9)[loopCnt - 1] = codeError / 2; F(10)[loopCnt - 1] = codeNco / 2;             /* :382-383 */
            F(11)[loopCnt - 1] = carrError; F(12)[loopCnt - 1] = carrNco;
            F(4)[loopCnt - 1] = I_E; F(3)[loopCnt - 1] = I_P; F(5)[loopCnt - 1] = I_L;       /* :388-393 */
            F(6)[loopCnt - 1] = Q_E; F(7)[loopCnt - 1] = Q_P; F(8)[loopCnt - 1] = Q_L;
            if (pilot) {                                                                     /* :396-402 */
                F(15)[loopCnt - 1] = I_PCL; F(16)[loopCnt - 1] = Q_PCL; F(17)[loopCnt - 1] = I_ECL;
                F(18)[loopCnt - 1] = I_LCL; F(19)[loopCnt - 1] = Q_ECL; F(20)[loopCnt - 1] = Q_LCL;
            }
            epochsDone[ch] = loopCnt;
        }
        free(cmCode);
#undef F
        if (stop) break;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * BDS/B1C/include/NB_tracking.m:44-431 (the loop; DataCNo / PLD block not restated): 10 ms epochs, BOC(1,1) sub-chip tables indexed
 * by ceil(tcode*2)+1 with the colon vectors scaled by 2 (:265-297), carrier-aided code NCO centre channel.codeFreq (:222, :354),
 * pilot discriminator atan(-I/Q) (:301), weights 11/40 and 29/40 (:302, :318), code discriminators scaled by (1 - earlyLateSpc)
 * (:313-317).  dataBoc / pilotBoc [nCh][2*codeLength].  out [nCh][17][nEpochs]: 15 common rows, Pilot_I_P, Pilot_Q_P. */
/* WB_tracking.m (pilotBoc61 != NULL): the same loop with the pilot's BOC(6,1) component on ceil(tcode*6)+1 (:283, :294, :305), the
 * composite pilot -sqrt(4/33)*p61 +- sqrt(29/33)*p11 cross terms (:339-344), carrier error (data + 3*pilot)/4 with atan(Q/I) of the
 * composite (:355-356), code error weighted by CalcWeighingFactor's factor (:124, :374) and all six composite Pilot_* rows
 * (:409-414): out is then [nCh][21][nEpochs] with Pilot_I_P, Pilot_Q_P, Pilot_I_E, Pilot_I_L, Pilot_Q_E, Pilot_Q_L after the 15. */
static int tracking_b1c(const int8_t* iq, size_t nBytes, const orc_settings* s, int nCh, const int* PRN,
                        const double* acquiredFreq, const double* codePhase, const double* codeFreq0,
                        const int8_t* dataBoc, const int8_t* pilotBoc, const int8_t* pilotBoc61, double factor,
                        int nEpochs, double* out, int* epochsDone);
int orc_tracking_b1c_nb(const int8_t* iq, size_t nBytes, const orc_settings* s, int nCh, const int* PRN,
                        const double* acquiredFreq, const double* codePhase, const double* codeFreq0,
                        const int8_t* dataBoc, const int8_t* pilotBoc, int nEpochs, double* out, int* epochsDone)
{
    return tracking_b1c(iq, nBytes, s, nCh, PRN, acquiredFreq, codePhase, codeFreq0, dataBoc, pilotBoc, NULL, 0.0, nEpochs, out, epochsDone);
}
int orc_tracking_b1c_wb(const int8_t* iq, size_t nBytes, const orc_settings* s, int nCh, const int* PRN,
                        const double* acquiredFreq, const double* codePhase, const double* codeFreq0,
                        const int8_t* dataBoc, const int8_t* pilotBoc, const int8_t* pilotBoc61, double factor,
                        int nEpochs, double* out, int* epochsDone)
{
    return tracking_b1c(iq, nBytes, s, nCh, PRN, acquiredFreq, codePhase, codeFreq0, dataBoc, pilotBoc, pilotBoc61, factor, nEpochs, out, epochsDone);
}
static int tracking_b1c(const int8_t* iq, size_t nBytes, const orc_settings* s, int nCh, const int* PRN,
                        const double* acquiredFreq, const double* codePhase, const double* codeFreq0,
                        const int8_t* dataBoc, const int8_t* pilotBoc, const int8_t* pilotBoc61, double factor,
                        int nEpochs, double* out, int* epochsDone)
{
    const int wb = pilotBoc61 != NULL;
    const int NF = wb ? ORC_NFIELDS_PILOT6 : ORC_NFIELDS_PILOT;
    const int L12 = (int)s->codeLength * 12;
    const double earlyLateSpc = s->dllCorrelatorSpacing;                                     /* :186 */
    const double codeLength = s->codeLength;                                                 /* :188 */
    const int L2c = (int)s->codeLength * 2;
    const double PDIcode = s->intTime;
    double tau1code, tau2code;
    calcLoopCoef(s->dllNoiseBandwidth, s->dllDampingRatio, 1.0, &tau1code, &tau2code);       /* :194-196 */
    const double Wn = 1.2 * s->pllNoiseBandwidth;                                            /* Common/calcLoopCoefCarr.m (:200) */
    const double pf3 = pow(Wn, 3) * pow(s->intTime, 2), pf2 = 2 * pow(Wn, 2) * s->intTime, pf1 = 2 * Wn;
    for (int ch = 0; ch < nCh; ch++) {
        double* o = out + (size_t)ch * NF * nEpochs;
        for (int f = 0; f < NF; f++) {
            const double fill = (f == 0 || (f >= 3 && f <= 8) || f >= 15) ? 0.0 : INFINITY;
            for (int e = 0; e < nEpochs; e++) o[(size_t)f * nEpochs + e] = fill;
        }
        epochsDone[ch] = 0;
    }
    for (int ch = 0; ch < nCh; ch++) {                                                       /* :204 */
        if (PRN[ch] == 0) continue;                                                          /* :207 */
        double* o = out + (size_t)ch * NF * nEpochs;
#define F(i) (o + (size_t)(i) * nEpochs)
        size_t pos = (size_t)(2 * ((long)s->skipNumberOfBytes + (long)codePhase[ch] - 1));   /* :211-213 */
        double* dat = (double*)malloc(sizeof(double) * (L2c + 2)), *p11 = (double*)malloc(sizeof(double) * (L2c + 2));
        const int8_t* d = dataBoc + (size_t)ch * L2c, *p = pilotBoc + (size_t)ch * L2c;
        dat[0] = d[L2c - 1]; p11[0] = p[L2c - 1];                                            /* :216-220 */
        for (int i = 0; i < L2c; i++) { dat[i + 1] = d[i]; p11[i + 1] = p[i]; }
        dat[L2c + 1] = d[0]; p11[L2c + 1] = p[0];
        double* p61 = NULL;
        if (wb) {                                                                            /* WB :181-183 */
            const int8_t* q = pilotBoc61 + (size_t)ch * L12;
            p61 = (double*)malloc(sizeof(double) * (L12 + 2));
            p61[0] = q[L12 - 1]; for (int i = 0; i < L12; i++) p61[i + 1] = q[i]; p61[L12 + 1] = q[0];
        }
        double codeFreq = codeFreq0[ch], remCodePhase = 0.0;                                 /* :222-224 */
        double carrFreq = acquiredFreq[ch], carrFreqBasis = acquiredFreq[ch], remCarrPhase = 0.0;
        double oldCodeNco = 0, oldCodeError = 0, d2CarrError = 0, dCarrError = 0;
        int stop = 0;
        for (int loopCnt = 1; loopCnt <= nEpochs; loopCnt++) {                               /* :238 */
            F(0)[loopCnt - 1] = (double)pos / 2;                                             /* :256 */
            const double codePhaseStep = codeFreq / s->samplingFreq;                         /* :258 */
            const int blksize = (int)ceil((codeLength - remCodePhase) / codePhaseStep);      /* :259 */
            if (pos + 2 * (size_t)blksize > nBytes) { stop = 1; break; }                     /* :255-259 */
            const int8_t* raw = iq + pos; pos += 2 * (size_t)blksize;
            F(13)[loopCnt - 1] = remCodePhase;                                               /* :262 */
            const double span = (blksize - 1) * codePhaseStep;
            /* (remCodePhase -/+ spc)*2 : codePhaseStep*2 : ((blksize-1)*codePhaseStep + remCodePhase -/+ spc)*2  (:265-267, :275-277, :285-287) */
            const double aE = (remCodePhase - earlyLateSpc) * 2, bE = (span + remCodePhase - earlyLateSpc) * 2;
            const double aL = (remCodePhase + earlyLateSpc) * 2, bL = (span + remCodePhase + earlyLateSpc) * 2;
            const double aP = remCodePhase * 2, bP = (span + remCodePhase) * 2;
            const double tstep = codePhaseStep * 2;
            int nE_, nL_, nP_; double cE, cL, cP;
            colon_setup(aE, tstep, bE, &nE_, &cE);
            colon_setup(aL, tstep, bL, &nL_, &cL);
            colon_setup(aP, tstep, bP, &nP_, &cP);
            F(14)[loopCnt - 1] = remCarrPhase;                                               /* :300 */
            const double w = carrFreq * 2.0 * M_PI;
            double I_E = 0, Q_E = 0, I_P = 0, Q_P = 0, I_L = 0, Q_L = 0, pI_E = 0, pQ_E = 0, pI_P = 0, pQ_P = 0, pI_L = 0, pQ_L = 0;
            double sI_E = 0, sQ_E = 0, sI_P = 0, sQ_P = 0, sI_L = 0, sQ_L = 0;                 /* p61_* (WB :321-326) */
            for (int n = 0; n < blksize; n++) {
                const double tE = colon_elem(aE, tstep, cE, nE_, n), tL = colon_elem(aL, tstep, cL, nL_, n), tP = colon_elem(aP, tstep, cP, nP_, n);
                const int iE = (int)ceil(tE) + 1;                                            /* tcode2 = ceil(tcode) + 1 */
                const int iL = (int)ceil(tL) + 1;
                const int iP = (int)ceil(tP) + 1;
                const double trig = (w * ((double)n / s->samplingFreq)) + remCarrPhase;
                const double c = cos(trig), sn = sin(trig);
                const double xr = raw[2 * n], xi = raw[2 * n + 1];
                const double iBB = c * xr + sn * xi, qBB = c * xi - sn * xr;                 /* :310-311 */
                I_E += dat[iE - 1] * iBB; Q_E += dat[iE - 1] * qBB; I_P += dat[iP - 1] * iBB; Q_P += dat[iP - 1] * qBB;
                I_L += dat[iL - 1] * iBB; Q_L += dat[iL - 1] * qBB;                          /* :314-319 */
                pI_E += p11[iE - 1] * iBB; pQ_E += p11[iE - 1] * qBB; pI_P += p11[iP - 1] * iBB; pQ_P += p11[iP - 1] * qBB;
                pI_L += p11[iL - 1] * iBB; pQ_L += p11[iL - 1] * qBB;                        /* :321-326 */
                if (wb) {                                                                    /* pilotBOC61(ceil(tcode * 6) + 1) */
                    const double e6 = p61[(int)ceil(tE * 6)], l6 = p61[(int)ceil(tL * 6)], p6 = p61[(int)ceil(tP * 6)];
                    sI_E += e6 * iBB; sQ_E += e6 * qBB; sI_P += p6 * iBB; sQ_P += p6 * qBB; sI_L += l6 * iBB; sQ_L += l6 * qBB;
                }
            }
            remCodePhase = colon_elem(aP, tstep, cP, nP_, blksize - 1) / 2 + codePhaseStep - codeLength;   /* :297 */
            remCarrPhase = fmod((w * ((double)blksize / s->samplingFreq)) + remCarrPhase, 2 * M_PI);       /* :305 */
            double carrError = atan(Q_P / I_P) / (2.0 * M_PI);                               /* :330 */
            double cI_E = 0, cQ_E = 0, cI_P = 0, cQ_P = 0, cI_L = 0, cQ_L = 0;               /* composite pilot (WB :339-344) */
            if (wb) {
                cI_E = -sqrt(4.0 / 33) * sI_E + sqrt(29.0 / 33) * pQ_E; cQ_E = -sqrt(4.0 / 33) * sQ_E - sqrt(29.0 / 33) * pI_E;
                cI_P = -sqrt(4.0 / 33) * sI_P + sqrt(29.0 / 33) * pQ_P; cQ_P = -sqrt(4.0 / 33) * sQ_P - sqrt(29.0 / 33) * pI_P;
                cI_L = -sqrt(4.0 / 33) * sI_L + sqrt(29.0 / 33) * pQ_L; cQ_L = -sqrt(4.0 / 33) * sQ_L - sqrt(29.0 / 33) * pI_L;
                const double p_carrError = atan(cQ_P / cI_P) / (2.0 * M_PI);                 /* WB :355 */
                carrError = (carrError * 1 + p_carrError * 3) / 4;                           /* WB :356 */
            } else {
                const double p11_carrError = atan(-pI_P / pQ_P) / (2.0 * M_PI);              /* :331 */
                carrError = (carrError * 11 + p11_carrError * 29) / 40;                      /* :332 */
            }
            d2CarrError = d2CarrError + carrError * pf3;                                     /* :335-337 */
            dCarrError = d2CarrError + carrError * pf2 + dCarrError;
            const double carrNco = dCarrError + carrError * pf1;
            F(2)[loopCnt - 1] = carrFreq;
            carrFreq = carrFreqBasis + carrNco;
            double codeError = (sqrt(I_E * I_E + Q_E * Q_E) - sqrt(I_L * I_L + Q_L * Q_L)) /
                               (sqrt(I_E * I_E + Q_E * Q_E) + sqrt(I_L * I_L + Q_L * Q_L)) * (1 - earlyLateSpc);   /* :343-344 */
            const double p11_codeError = (sqrt(pI_E * pI_E + pQ_E * pQ_E) - sqrt(pI_L * pI_L + pQ_L * pQ_L)) /
                                         (sqrt(pI_E * pI_E + pQ_E * pQ_E) + sqrt(pI_L * pI_L + pQ_L * pQ_L)) * (1 - earlyLateSpc);   /* :345-346 */
            if (wb) {                                                                        /* WB :371-374 */
                const double p_codeError = (sqrt(cI_E * cI_E + cQ_E * cQ_E) - sqrt(cI_L * cI_L + cQ_L * cQ_L)) /
                                           (sqrt(cI_E * cI_E + cQ_E * cQ_E) + sqrt(cI_L * cI_L + cQ_L * cQ_L)) * (1 - earlyLateSpc);
                codeError = codeError * factor + p_codeError * (1 - factor);
            } else
            codeError = (codeError * 11 + p11_codeError * 29) / 40;                          /* :347 */
            const double codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code);
            oldCodeNco = codeNco; oldCodeError = codeError;
            F(1)[loopCnt - 1] = codeFreq;                                                    /* :353 */
            codeFreq = codeFreq0[ch] - codeNco;                                              /* :354 */
            F(9)[loopCnt - 1] = codeError; F(10)[loopCnt - 1] = codeNco;
            F(11)[loopCnt - 1] = carrError; F(12)[loopCnt - 1] = carrNco;
            F(4)[loopCnt - 1] = I_E; F(3)[loopCnt - 1] = I_P; F(5)[loopCnt - 1] = I_L;
            F(6)[loopCnt - 1] = Q_E; F(7)[loopCnt - 1] = Q_P; F(8)[loopCnt - 1] = Q_L;
            if (wb) {                                                                        /* WB :409-414 */
                F(15)[loopCnt - 1] = cI_P; F(16)[loopCnt - 1] = cQ_P; F(17)[loopCnt - 1] = cI_E;
                F(18)[loopCnt - 1] = cI_L; F(19)[loopCnt - 1] = cQ_E; F(20)[loopCnt - 1] = cQ_L;
            } else { F(15)[loopCnt - 1] = pI_P; F(16)[loopCnt - 1] = pQ_P; }                 /* :364-365 */
            epochsDone[ch] = loopCnt;
        }
        free(dat); free(p11); free(p61);
#undef F
        if (stop) break;
    }
    return 0;
}
