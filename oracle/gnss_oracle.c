/*
 * gnss_oracle.c — plain-C float64 restatement of the reference's GPS L1 C/A hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libgnsscorr.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference is 100 % MATLAB with no tests/golden vectors and
 * cannot be executed in this image (no MATLAB/Octave).  This file is an independent
 * second restatement (the first is oracle/np_oracle.py); the two must agree, and both
 * are pinned against the IS-GPS-200 C/A first-chip octals and closed-loop KATs.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/GPS/GPS_L1CA/).
 *
 * Build: gcc -O2 -fopenmp -fPIC -shared -o oracle/_build/libgnss_oracle.so oracle/gnss_oracle.c -lm
 *        (-ffp-contract=off so a*b+c is never fused: MATLAB rounds each op)
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

typedef struct {
    double samplingFreq, IF, codeFreqBasis, codeLength;            /* initSettings.m:71-76 */
    double acqSearchBand, acqSearchStep, acqThreshold;             /* :86,:92,:90 */
    int    acqNonCohTime;                                          /* :88 */
    int    skipNumberOfBytes;                                      /* :56 */
    double dllDampingRatio, dllNoiseBandwidth, dllCorrelatorSpacing; /* :100-102 */
    double pllDampingRatio, pllNoiseBandwidth, intTime;            /* :105-108 */
    double CNo_accTime; int CNo_VSMinterval;                       /* :133-135 */
    double freqSpacing;                                            /* GLO/GLO_GL1/initSettings.m:72 */
    int    glo;            /* 0: GPS/GPS_L1CA files; 1: GLO/GLO_GL1 (= GLO_GL2), cited "GLO :line"; 2: BDS/B3I, cited "B3I :line";
                              3: GAL/GAL_E1C, cited "E1C :line" (pilotTRKflag below) */
    int    pilotTRKflag;   /* GAL/GAL_E1C/initSettings.m:113 */
} orc_settings;

/* ---------------------------------------------------------------- helpers */
static double m_round(double x) { return x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5); }

int orc_samples_per_code(const orc_settings* s)                    /* acquisition.m:116-117 */
{
    return (int)m_round(s->samplingFreq / (s->codeFreqBasis / s->codeLength));
}

/* generateCAcode.m:42-90 — ±1 chips */
static const int G2S[51] = {5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258, 469, 470, 471, 472,
                            473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862,
                            145, 175, 52, 21, 237, 235, 886, 657, 634, 762, 355, 1012, 176, 603, 130, 359, 595, 68, 386};

void orc_generateCAcode(int PRN, double* CAcode /*1023*/)
{
    double g1[1023], g2[1023], reg[10];
    int g2shift = G2S[PRN - 1];
    for (int i = 0; i < 10; i++) reg[i] = -1;
    for (int i = 0; i < 1023; i++) {
        g1[i] = reg[9];
        double saveBit = reg[2] * reg[9];
        for (int j = 9; j >= 1; j--) reg[j] = reg[j - 1];
        reg[0] = saveBit;
    }
    for (int i = 0; i < 10; i++) reg[i] = -1;
    for (int i = 0; i < 1023; i++) {
        g2[i] = reg[9];
        double saveBit = reg[1] * reg[2] * reg[5] * reg[7] * reg[8] * reg[9];
        for (int j = 9; j >= 1; j--) reg[j] = reg[j - 1];
        reg[0] = saveBit;
    }
    /* g2 = [g2(1023-g2shift+1 : 1023), g2(1 : 1023-g2shift)] */
    for (int i = 0; i < 1023; i++) {
        int src = (i < g2shift) ? (1023 - g2shift + i) : (i - g2shift);
        CAcode[i] = -(g1[i] * g2[src]);
    }
}

/* BDS/B3I/include/generateB3Icode.m:33-86 — +-1 chips (10230) */
static const int B3I_INIT[63] = {4, 11, 13, 22, 30, 36, 44, 48, 88, 104, 116, 129, 376, 418, 458, 682, 696, 707, 1078, 2069,
                                 2248, 2574, 2596, 2731, 4294, 4436, 4647, 4978, 4986, 1, 5209, 5539, 6061, 6488, 7130, 7165,
                                 7403, 5879, 1681, 5080, 5938, 3983, 6208, 7223, 2996, 1814, 6906, 6144, 4713, 7406, 7264, 1766,
                                 5347, 3515, 7951, 7054, 3884, 6067, 4230, 3803, 869, 3683, 1205};
void orc_generateB3Icode(int PRN, double* code /*10230*/)
{
    static const double reset_state[13] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 1, 1};
    double reg[13], CA[10230];
    for (int i = 0; i < 13; i++) reg[i] = -1;
    for (int ind = 0; ind < 10230; ind++) {                        /* :40-49 */
        CA[ind] = reg[12];
        int eq = 1; for (int i = 0; i < 13; i++) if (reg[i] != reset_state[i]) eq = 0;
        if (eq) { for (int i = 0; i < 13; i++) reg[i] = -1; }
        else {
            double fb = reg[0] * reg[2] * reg[3] * reg[12];
            for (int j = 12; j >= 1; j--) reg[j] = reg[j - 1];
            reg[0] = fb;
        }
    }
    for (int i = 0; i < 13; i++) reg[i] = -1;
    for (int step = 0; step < B3I_INIT[PRN - 1] + 10230; step++) { /* :68-80 */
        if (step >= B3I_INIT[PRN - 1]) code[step - B3I_INIT[PRN - 1]] = reg[12] * CA[step - B3I_INIT[PRN - 1]];
        double fb = reg[0] * reg[4] * reg[5] * reg[6] * reg[8] * reg[9] * reg[11] * reg[12];
        for (int j = 12; j >= 1; j--) reg[j] = reg[j - 1];
        reg[0] = fb;
    }
}

/* GLO/GLO_GL1/include/generateCAcode.m:95-108 — 511-chip ST code (PRN == 0 branch) */
void orc_glo_code(double* code /*511*/)
{
    double reg[9];
    for (int i = 0; i < 9; i++) reg[i] = -1;
    for (int i = 0; i < 511; i++) {
        code[i] = reg[6];
        double save1 = reg[4] * reg[8];
        for (int j = 8; j >= 1; j--) reg[j] = reg[j - 1];
        reg[0] = save1;
    }
}

/* element k of MATLAB 0:d:b (colonop), n+1 elements, last element c */
static void colon0_setup(double d, double b, long* n_out, double* c_out)
{
    double tol = 2.0 * 2.220446049250313e-16 * fabs(b);
    long n;
    if (d == 1) n = (long)floor(b);
    else if (d == floor(d)) n = (long)trunc(b / d);
    else { n = (long)m_round(b / d); if ((0.0 + n * d - b) > tol) n -= 1; }
    double c = 0.0 + n * d;
    if ((c - b) > -tol) c = b;
    *n_out = n; *c_out = c;
}
/* GLO generateCAcode.m:110-116 — samples = floor(0:stepSize:(num*stepSize)-stepSize) wrapped to 511 */
void orc_glo_sampled_code(double sampFreq, long numSamples, double* out)
{
    double code[511]; orc_glo_code(code);
    double stepSize = 511e3 / sampFreq;
    long n; double c;
    colon0_setup(stepSize, (numSamples * stepSize) - stepSize, &n, &c);
    for (long k = 0; k <= n && k < numSamples; k++) {
        double v = (2 * k < n) ? 0.0 + (double)k * stepSize : (2 * k > n) ? c - (double)(n - k) * stepSize : (0.0 + c) / 2;
        out[k] = code[(long)fmod(floor(v), 511)];
    }
}

/* makeCaTable.m:43-67 */
void orc_makeCaTable(int PRN, const orc_settings* s, double* table /*N*/)
{
    int N = orc_samples_per_code(s);
    double ts = 1 / s->samplingFreq, tc = 1 / s->codeFreqBasis;
    double ca[1023];
    orc_generateCAcode(PRN, ca);
    for (int n = 1; n <= N; n++) {
        int idx = (int)ceil((ts * (double)n) / tc);
        if (n == N) idx = 1023;
        table[n - 1] = ca[idx - 1];
    }
}

/* ------------------------------------------------- mixed-radix FFT (f64)
 * Stands in for MATLAB's fft/ifft built-ins (closed FFTW/MKL inside MATLAB; call sites
 * acquisition.m:164,183,188).  Stockham autosort, generic radix butterfly; any length
 * whose prime factors are small. */
typedef struct { int n, nf, fac[40]; cplx* tw; } fftplan;

static void plan_make(fftplan* p, int n)
{
    p->n = n; p->nf = 0;
    int m = n;
    while (m % 4 == 0) { p->fac[p->nf++] = 4; m /= 4; }
    for (int f = 2; m > 1; f++) while (m % f == 0) { p->fac[p->nf++] = f; m /= f; }
    p->tw = (cplx*)malloc(sizeof(cplx) * (size_t)n);
    for (int k = 0; k < n; k++) {
        /* octant-exact table */
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        p->tw[k] = (double)cosl(a) + I * (double)sinl(a);
    }
}
static void plan_free(fftplan* p) { free(p->tw); }

/* out-of-place ping-pong; result returned in x; sign=-1 forward, +1 inverse (unscaled) */
static void fft_exec(const fftplan* p, cplx* x, cplx* y, int sign)
{
    int N = p->n, n = N, s = 1;
    cplx *src = x, *dst = y;
    for (int f = 0; f < p->nf; f++) {
        int r = p->fac[f], m = n / r, tstep = N / n, rstep = N / r;
        for (int pp = 0; pp < m; pp++) {
            for (int q = 0; q < s; q++) {
                cplx a[64];
                for (int i = 0; i < r; i++) a[i] = src[q + s * (pp + i * m)];
                for (int j = 0; j < r; j++) {
                    cplx acc = a[0];
                    for (int i = 1; i < r; i++) {
                        cplx w = p->tw[((long)i * j % r) * rstep];
                        if (sign > 0) w = conj(w);
                        acc += a[i] * w;
                    }
                    cplx w = p->tw[((long)pp * j * tstep) % N];
                    if (sign > 0) w = conj(w);
                    dst[q + s * (r * pp + j)] = acc * w;
                }
            }
        }
        n = m; s *= r;
        cplx* t = src; src = dst; dst = t;
    }
    if (src != x) memcpy(x, src, sizeof(cplx) * (size_t)N);
}

/* exported for the unit tests: dir=-1 fft, +1 ifft (scaled 1/n like MATLAB) */
int orc_fft(double* reim /*interleaved, n*/, int n, int dir)
{
    fftplan p; plan_make(&p, n);
    for (int i = 0; i < p.nf; i++) if (p.fac[i] > 64) { plan_free(&p); return -1; }
    cplx* x = (cplx*)malloc(sizeof(cplx) * n), *y = (cplx*)malloc(sizeof(cplx) * n);
    for (int i = 0; i < n; i++) x[i] = reim[2 * i] + I * reim[2 * i + 1];
    fft_exec(&p, x, y, dir);
    double sc = dir > 0 ? 1.0 / n : 1.0;
    for (int i = 0; i < n; i++) { reim[2 * i] = creal(x[i]) * sc; reim[2 * i + 1] = cimag(x[i]) * sc; }
    free(x); free(y); plan_free(&p);
    return 0;
}

/* Galileo E1 memory codes: DATA the reference reads at run time from include/E1b.dat / E1c.dat
 * (E1C generateE1Bcode.m:44-55); the caller hands the 0/1 tables [50][4092] over before using mode 3. */
static const int8_t* g_e1bits[2] = {NULL, NULL};
void orc_set_e1_codes(const int8_t* e1b, const int8_t* e1c) { g_e1bits[0] = e1b; g_e1bits[1] = e1c; }
/* E1C generateE1Bcode.m:55-64 / generateE1Ccode.m: 1-2*bit with the BOC(1,1) sub-carrier [c -c], 8184 values */
static void e1_code(int PRN, int comp, double* out /*8184*/)
{
    const int8_t* b = g_e1bits[comp] + (size_t)(PRN - 1) * 4092;
    for (int i = 0; i < 4092; i++) { double c = 1.0 - 2.0 * (double)b[i]; out[2 * i] = c; out[2 * i + 1] = -c; }
}
/* E1C makeE1BTable.m:38-58 / makeE1CTable.m */
static void e1_table(const double* code, const orc_settings* s, int N, double* table)
{
    double ts = 1 / s->samplingFreq, tc = 1 / s->codeFreqBasis / 2;
    for (int n = 1; n <= N; n++) {
        int idx = (int)ceil((ts * (double)n) / tc);
        if (n == N) idx = (int)s->codeLength * 2;
        if (n == 1) idx = 1;
        table[n - 1] = code[idx - 1];
    }
}

/* --------------------------------------------------------------- acquisition
 * acquisition.m:113-292 (resampling branch :50-111 not restated, resamplingflag==0).
 * iq: int8 interleaved I,Q record bytes AFTER the fseek of postProcessing.m:74; the first
 * max(42,nonCoh+2) code periods are longSignal (postProcessing.m:83-96).
 * Outputs indexed by PRN-1 (length 32): carrFreq, codePhase, peakMetric (acquisition.m:130-134)
 * plus coarseBin/coarseCodePhase (1-based) for every searched PRN. */
int orc_acquisition(const int8_t* iq, size_t nSamplesAvail, const orc_settings* s,
                    const int* prnList, int nPrn,
                    double* carrFreq, double* codePhaseOut, double* peakMetric,
                    int* coarseBin, int* coarseCodePhase, double* sigPowerOut)
{
    const int N = orc_samples_per_code(s);
    const int L2 = 2 * N;
    const int b3i = (s->glo == 2);
    const int e1c = (s->glo == 3);
    /* postProcessing.m:86 max(42, nonCoh+2); B3I postProcessing.m:86 max(22, nonCoh+1) */
    const int codeLen = b3i ? ((22 > s->acqNonCohTime + 1) ? 22 : s->acqNonCohTime + 1)
                            : ((42 > s->acqNonCohTime + 2) ? 42 : s->acqNonCohTime + 2);
    const int nFinePer = b3i ? 20 : e1c ? 25 : 40;                               /* B3I acquisition.m:131-133; E1C :148 */
    if (nSamplesAvail < (size_t)codeLen * N) return -1;
    const double ts = 1 / s->samplingFreq;                                       /* :119 */
    const int nBins = (int)m_round(s->acqSearchBand * 2 / s->acqSearchStep) + 1; /* :124 */
    const double fineSearchStep = e1c ? 10 : 25;                                 /* :138; E1C :138 */
    const int nFine = (int)m_round(s->acqSearchStep / fineSearchStep) + 1;       /* :140 */
    const int nonCoh = s->acqNonCohTime;
    const int nRes = s->glo == 1 ? 21 : b3i ? 63 : e1c ? 50 : 32;   /* GLO acquisition.m:138-142 ; B3I :118-122 ; E1C :127-131 */
    if (e1c && (!g_e1bits[0] || !g_e1bits[1])) return -3;
    for (int i = 0; i < nRes; i++) { carrFreq[i] = codePhaseOut[i] = peakMetric[i] = 0; coarseBin[i] = coarseCodePhase[i] = 0; }

    size_t Ltot = (size_t)codeLen * N;
    cplx* sig = (cplx*)malloc(sizeof(cplx) * Ltot);
    for (size_t i = 0; i < Ltot; i++)                        /* GLO postProcessing.m:94: data2 + 1i*data1 */
        sig[i] = s->glo == 1 ? (double)iq[2 * i + 1] + I * (double)iq[2 * i] : (double)iq[2 * i] + I * (double)iq[2 * i + 1];
    double* glo40 = NULL;                                    /* GLO acquisition.m:164 caCode40ms */
    if (s->glo == 1) { glo40 = (double*)malloc(sizeof(double) * 40 * (size_t)N); orc_glo_sampled_code(s->samplingFreq, 40L * N, glo40); }
    double* phasePoints = (double*)malloc(sizeof(double) * L2);
    for (int n = 0; n < L2; n++) phasePoints[n] = (double)n * 2 * M_PI * ts;     /* :122 */

    /* :151 sigPower = sqrt(var(x(1:N))*N), var of complex with N-1 */
    cplx mean = 0; for (int i = 0; i < N; i++) mean += sig[i]; mean /= N;
    double v = 0; for (int i = 0; i < N; i++) { cplx d = sig[i] - mean; v += creal(d) * creal(d) + cimag(d) * cimag(d); }
    double sigPower = sqrt(v / (N - 1) * N);
    if (sigPowerOut) *sigPowerOut = sigPower;

    fftplan plan; plan_make(&plan, L2);
    for (int i = 0; i < plan.nf; i++) if (plan.fac[i] > 64) { plan_free(&plan); free(sig); free(phasePoints); return -2; }

    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ip = 0; ip < nPrn; ip++) {
        int PRN = prnList[ip];
        double* table = (double*)malloc(sizeof(double) * N);
        cplx* codeF = (cplx*)malloc(sizeof(cplx) * L2);
        cplx* buf = (cplx*)malloc(sizeof(cplx) * L2);
        cplx* tmp = (cplx*)malloc(sizeof(cplx) * L2);
        cplx* carr = (cplx*)malloc(sizeof(cplx) * L2);
        double* results = (double*)calloc((size_t)nBins * L2, sizeof(double));   /* :162 */
        double* coarseFreqBin = (double*)malloc(sizeof(double) * nBins);
        double* b3code = NULL;
        double* e1code = NULL; cplx* codeF2 = NULL; cplx* buf2 = NULL;          /* E1C: pilot (E1C) code and its spectrum */
        if (e1c) {                                                               /* E1C :158-171 */
            e1code = (double*)malloc(sizeof(double) * 8184);
            codeF2 = (cplx*)malloc(sizeof(cplx) * L2);
            buf2 = (cplx*)malloc(sizeof(cplx) * L2);
            e1_code(PRN, 1, e1code);
            e1_table(e1code, s, N, table);
            for (int n = 0; n < L2; n++) codeF2[n] = n < N ? table[n] : 0.0;
            fft_exec(&plan, codeF2, tmp, -1);
            for (int n = 0; n < L2; n++) codeF2[n] = conj(codeF2[n]);
            double* e1b = (double*)malloc(sizeof(double) * 8184);
            e1_code(PRN, 0, e1b);
            e1_table(e1b, s, N, table);
            free(e1b);
        }
        else if (s->glo == 1) orc_glo_sampled_code(s->samplingFreq, N, table);   /* GLO :145 */
        else if (b3i) {                                                          /* B3I makeB3ITable.m:38-52 */
            b3code = (double*)malloc(sizeof(double) * 10230);
            orc_generateB3Icode(PRN, b3code);
            double ts_ = 1 / s->samplingFreq, tc_ = 1 / s->codeFreqBasis;
            for (int n = 1; n <= N; n++) { int idx = (int)ceil((ts_ * (double)n) / tc_); if (n == N) idx = 10230; table[n - 1] = b3code[idx - 1]; }
        }
        else orc_makeCaTable(PRN, s, table);                                     /* :158 */
        const int ri = s->glo == 1 ? PRN + 7 : PRN - 1;                          /* result slot: K+8 / PRN (1-based) */
        for (int n = 0; n < L2; n++) codeF[n] = n < N ? table[n] : 0.0;          /* :160 */
        fft_exec(&plan, codeF, tmp, -1);
        for (int n = 0; n < L2; n++) codeF[n] = conj(codeF[n]);                  /* :164 */
        for (int k = 1; k <= nBins; k++) {                                       /* :167 */
            coarseFreqBin[k - 1] = s->glo == 1 ? s->IF - s->freqSpacing * PRN + s->acqSearchBand - s->acqSearchStep * (k - 1)   /* GLO :181 */
                                          : s->IF + s->acqSearchBand - s->acqSearchStep * (k - 1);                   /* :169 */
            for (int n = 0; n < L2; n++) {
                double a = coarseFreqBin[k - 1] * phasePoints[n];
                carr[n] = cos(a) - I * sin(a);                                   /* :172 */
            }
            for (int m = 1; m <= nonCoh; m++) {                                  /* :175 */
                const cplx* w = sig + (size_t)(m - 1) * N;                       /* :177 */
                for (int n = 0; n < L2; n++) buf[n] = carr[n] * w[n];            /* :180-181 */
                fft_exec(&plan, buf, tmp, -1);                                   /* :183 */
                if (e1c) { for (int n = 0; n < L2; n++) buf2[n] = buf[n] * codeF2[n]; fft_exec(&plan, buf2, tmp, +1); }   /* E1C :194 */
                for (int n = 0; n < L2; n++) buf[n] *= codeF[n];                 /* :186 */
                fft_exec(&plan, buf, tmp, +1);
                double* row = results + (size_t)(k - 1) * L2;
                if (e1c) for (int n = 0; n < L2; n++) row[n] += cabs(buf[n]) / L2 + cabs(buf2[n]) / L2;   /* E1C :196-198 */
                else for (int n = 0; n < L2; n++) row[n] += cabs(buf[n]) / L2;   /* :188-190 */
            }
        }
        /* :196  [~,bin] = max(max(results,[],2))   — first maximal row */
        int bin = 1; double best = -1;
        for (int k = 0; k < nBins; k++) {
            double rm = results[(size_t)k * L2];
            for (int n = 1; n < L2; n++) if (results[(size_t)k * L2 + n] > rm) rm = results[(size_t)k * L2 + n];
            if (rm > best) { best = rm; bin = k + 1; }
        }
        /* :198  [peak,codePhase] = max(max(results)) — column max then first maximal column */
        int cp = 1; double peak = -1;
        for (int n = 0; n < L2; n++) {
            double cm = results[n];
            for (int k = 1; k < nBins; k++) if (results[(size_t)k * L2 + n] > cm) cm = results[(size_t)k * L2 + n];
            if (cm > peak) { peak = cm; cp = n + 1; }
        }
        peakMetric[ri] = peak / sigPower / nonCoh;                               /* :200 */
        coarseBin[ri] = bin; coarseCodePhase[ri] = cp;
        if (peakMetric[ri] > s->acqThreshold) {                                  /* :206 */
            double ca[1023]; if (s->glo == 0) orc_generateCAcode(PRN, ca);       /* :213 */
            if (e1c) e1_code(PRN, 1, e1code);                                    /* E1C :209 */
            double bestFine = -1; int bestJ = 1; double bestFreq = 0;
            for (int j = 1; j <= nFine; j++) {                                   /* :224 */
                double f = coarseFreqBin[bin - 1] + s->acqSearchStep / 2 - fineSearchStep * (j - 1);  /* :227 */
                cplx sumPerCode[40];
                for (int c = 0; c < nFinePer; c++) {
                    cplx acc = 0;
                    for (int n = 0; n < N; n++) {
                        long gi = (long)c * N + n;
                        double chip;
                        if (s->glo == 1) chip = glo40[gi];                                     /* GLO :164,236 */
                        else if (e1c) {                                                        /* E1C :211-214 */
                            long idx = (long)floor((ts * (double)gi) / (1 / s->codeFreqBasis / 2));
                            chip = e1code[idx % 8184];
                        }
                        else if (b3i) {                                                        /* B3I :174-177 */
                            long idx = (long)floor((ts * (double)gi) / (1 / s->codeFreqBasis));
                            chip = b3code[idx % 10230];
                        }
                        else {
                            long idx = (long)floor((ts * (double)gi) / (1 / s->codeFreqBasis));  /* :215 */
                            chip = ca[idx % (long)s->codeLength];                              /* :218 */
                        }
                        double a = f * ((double)gi * 2 * M_PI * ts);                           /* :148,:230 */
                        cplx cw = cos(a) - I * sin(a);
                        acc += (sig[(size_t)(cp - 1) + gi] * chip) * cw;                       /* :221,:232,:236 */
                    }
                    sumPerCode[c] = acc;
                }
                double maxPower = 0;
                if (e1c) {                                                       /* E1C :236-252 */
                    static const double SEC[25] = {1, 1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, -1, 1, -1, 1, -1, -1, 1, -1, -1, 1, 1, -1, 1};
                    cplx t = 0; for (int q = 0; q < 25; q++) t += sumPerCode[q] * SEC[q];
                    maxPower = cabs(t);
                    for (int ci = 1; ci <= 24; ci++) {
                        cplx t1 = 0, t2 = 0;
                        for (int q = 0; q < 25; q++) { cplx v_ = sumPerCode[q] * SEC[(q - ci + 25) % 25]; if (q < ci) t1 += v_; else t2 += v_; }
                        double pw = cabs(t1) + cabs(t2);
                        if (pw > maxPower) maxPower = pw;
                    }
                }
                if (b3i) {                                                       /* B3I :193-211 */
                    static const double NH[20] = {1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1};
                    if ((PRN >= 1 && PRN <= 5) || (PRN >= 59 && PRN <= 63)) {
                        double c1 = 0, c2 = cabs(sumPerCode[0]) + cabs(sumPerCode[19]);
                        for (int q = 0; q < 20; q += 2) c1 += cabs(sumPerCode[q] + sumPerCode[q + 1]);
                        for (int q = 1; q < 19; q += 2) c2 += cabs(sumPerCode[q] + sumPerCode[q + 1]);
                        maxPower = c1 > c2 ? c1 : c2;
                    } else {
                        cplx t = 0; for (int q = 0; q < 20; q++) t += sumPerCode[q] * NH[q];
                        maxPower = cabs(t);
                        for (int ci = 1; ci <= 19; ci++) {
                            cplx t1 = 0, t2 = 0;
                            for (int q = 0; q < 20; q++) { cplx v_ = sumPerCode[q] * NH[(q - ci + 20) % 20]; if (q < ci) t1 += v_; else t2 += v_; }
                            double pw = cabs(t1) + cabs(t2);
                            if (pw > maxPower) maxPower = pw;
                        }
                    }
                }
                for (int c = 0; c < 20 && !b3i && !e1c; c++) {                   /* :243 */
                    cplx t = 0;
                    if (s->glo == 0) { for (int q = c; q < c + 20; q++) t += sumPerCode[q]; }
                    else {                                                       /* GLO :250-251: sum(c:c+9) - sum(c+10:c+19) */
                        cplx t1 = 0, t2 = 0;
                        for (int q = c; q < c + 10; q++) t1 += sumPerCode[q];
                        for (int q = c + 10; q < c + 20; q++) t2 += sumPerCode[q];
                        t = t1 - t2;
                    }
                    double pw = cabs(t);                                         /* :245 */
                    if (pw > maxPower) maxPower = pw;                            /* :247 */
                }
                if (maxPower > bestFine) { bestFine = maxPower; bestJ = j; bestFreq = f; }       /* :253 */
            }
            (void)bestJ;
            carrFreq[ri] = bestFreq;                                             /* :254 */
            codePhaseOut[ri] = cp;                                               /* :256 */
            if (carrFreq[ri] == 0) carrFreq[ri] = 1;                             /* :258 */
        }
        free(table); free(codeF); free(buf); free(tmp); free(carr); free(results); free(coarseFreqBin); free(b3code);
        free(e1code); free(codeF2); free(buf2);
    }
    plan_free(&plan); free(sig); free(phasePoints); free(glo40);
    return rc;
}

/* ------------------------------------------------------------------ tracking */
/* Common/calcLoopCoef.m:41-45 */
static void calcLoopCoef(double LBW, double zeta, double k, double* tau1, double* tau2)
{
    double Wn = LBW * 8 * zeta / (4 * zeta * zeta + 1);
    *tau1 = k / (Wn * Wn);
    *tau2 = 2.0 * zeta / Wn;
}

/* Common/CNoVSM.m:38-47 */
double orc_CNoVSM(const double* Ip, const double* Qp, int n, double T)
{
    double Zm = 0; double Z[4096];
    for (int i = 0; i < n; i++) { Z[i] = Ip[i] * Ip[i] + Qp[i] * Qp[i]; Zm += Z[i]; }
    Zm /= n;
    double Zv = 0; for (int i = 0; i < n; i++) Zv += (Z[i] - Zm) * (Z[i] - Zm); Zv /= (n - 1);
    cplx Pav = csqrt((cplx)(Zm * Zm - Zv));
    cplx Nv = 0.5 * (Zm - Pav);
    return 10 * log10(cabs((1 / T) * Pav / (2 * Nv)));
}

/* element idx (0-based) of MATLAB's a:d:b with n+1 elements, last element c (colonop) */
static inline double colon_elem(double a, double d, double c, int n, int idx)
{
    if (2 * idx < n) return a + (double)idx * d;
    if (2 * idx > n) return c - (double)(n - idx) * d;
    /* 2*idx == n: middle element of an odd-length vector */
    return (a + c) / 2;
}
/* n (count-1) and c (last element) of MATLAB's a:d:b for non-integer a/d (colonop) */
static void colon_setup(double a, double d, double b, int* n_out, double* c_out)
{
    double tol = 2.0 * 2.220446049250313e-16 * fmax(fabs(a), fabs(b));
    int n;
    if (a == floor(a) && d == 1) n = (int)(floor(b) - a);
    else if (a == floor(a) && d == floor(d)) n = (int)trunc((b - a) / d);
    else {
        n = (int)m_round((b - a) / d);
        if ((a + n * d - b) > tol) n -= 1;
    }
    double c = a + n * d;
    if ((c - b) > -tol) c = b;
    *n_out = n; *c_out = c;
}

#define ORC_NFIELDS 15
/* field order of out[ch][field][epoch]: absoluteSample, codeFreq, carrFreq, I_P, I_E, I_L, Q_E, Q_P, Q_L,
 * dllDiscr, dllDiscrFilt, pllDiscr, pllDiscrFilt, remCodePhase, remCarrPhase (tracking.m:48-77) */

/* tracking.m:88-368 for fileType 2 / schar.  iq = whole file bytes (int8 I,Q interleaved), nBytes its size.
 * Returns 0; epochsDone[ch] = completed epochs; a short read stops the WHOLE call (tracking.m:241-245). */
int orc_tracking(const int8_t* iq, size_t nBytes, const orc_settings* s, int nCh,
                 const int* PRN, const double* acquiredFreq, const double* codePhase, const double* codeFreq0 /* B3I channel.codeFreq, else NULL */,
                 int nEpochs, double* out /* nCh*15*nEpochs */, double* vsmValue, double* vsmIndex /* nCh*floor(nE/VSMint) */,
                 int* epochsDone, int parallel)
{
    const double earlyLateSpc = s->dllCorrelatorSpacing;                         /* :94 */
    const double PDIcode = s->intTime, PDIcarr = s->intTime;                     /* :97,:106 */
    double tau1code, tau2code, tau1carr, tau2carr;
    calcLoopCoef(s->dllNoiseBandwidth, s->dllDampingRatio, 1.0, &tau1code, &tau2code);   /* :100 */
    calcLoopCoef(s->pllNoiseBandwidth, s->pllDampingRatio, 0.25, &tau1carr, &tau2carr);  /* :109 */
    /* GLO Common/calcLoopCoefCarr.m:41-56 (GLO tracking.m:110) */
    const double WnC = 1.2 * s->pllNoiseBandwidth;
    const double pf3 = pow(WnC, 3) * pow(s->intTime, 2), pf2 = 2 * pow(WnC, 2) * s->intTime, pf1 = 2 * WnC;
    const int e1c = (s->glo == 3);
    const double sub = e1c ? 2.0 : 1.0;                  /* E1C tracking.m:236-262: tcode*2 into the BOC(1,1) sub-chip table */
    const int pilot = e1c && s->pilotTRKflag == 1;       /* E1C :127 */
    const int L = (int)s->codeLength * (e1c ? 2 : 1);    /* table entries per code period */
    const int nV = nEpochs / s->CNo_VSMinterval;
    if (e1c && (!g_e1bits[0] || !g_e1bits[1])) return -3;
    /* result init (tracking.m:48-83): zeros for absoluteSample and I/Q, inf elsewhere */
    for (int ch = 0; ch < nCh; ch++) {
        double* o = out + (size_t)ch * ORC_NFIELDS * nEpochs;
        for (int f = 0; f < ORC_NFIELDS; f++) {
            double fill = (f == 0 || (f >= 3 && f <= 8)) ? 0.0 : INFINITY;
            for (int e = 0; e < nEpochs; e++) o[(size_t)f * nEpochs + e] = fill;
        }
        for (int i = 0; i < nV; i++) { vsmValue[(size_t)ch * nV + i] = 0; vsmIndex[(size_t)ch * nV + i] = 0; }
        epochsDone[ch] = 0;
    }
    volatile int abortAll = 0;
#pragma omp parallel for schedule(dynamic, 1) if (parallel)
    for (int ch = 0; ch < nCh; ch++) {                                           /* :133 */
        if (s->glo == 1 ? (PRN[ch] == INT32_MIN) : (PRN[ch] == 0)) continue;          /* :136 ; GLO :137 status ~= '-' (INT32_MIN = off) */
        if (!parallel && abortAll) continue;   /* sequential semantics: return ends all later channels */
        double* o = out + (size_t)ch * ORC_NFIELDS * nEpochs;
#define F(i) (o + (size_t)(i) * nEpochs)
        size_t pos = (size_t)(2 * ((long)s->skipNumberOfBytes + (long)codePhase[ch] - 1));   /* :150 */
        double* ca = (double*)malloc(sizeof(double) * 10230);
        double* caCode = (double*)malloc(sizeof(double) * 10232);
        double* pCode = pilot ? (double*)malloc(sizeof(double) * 10232) : NULL;
        if (e1c) {                                                               /* E1C :125-130 */
            e1_code(PRN[ch], 0, ca);
            if (pilot) { double* t_ = (double*)malloc(sizeof(double) * 8184); e1_code(PRN[ch], 1, t_);
                         pCode[0] = t_[L - 1]; memcpy(pCode + 1, t_, sizeof(double) * L); pCode[L + 1] = t_[0]; free(t_); }
        }
        else if (s->glo == 1) orc_glo_code(ca);                                  /* GLO :88 */
        else if (s->glo == 2) orc_generateB3Icode(PRN[ch], ca);                  /* B3I :55 */
        else orc_generateCAcode(PRN[ch], ca);                                    /* :156 */
        const double codeFreqCentre = (s->glo == 2 && codeFreq0) ? codeFreq0[ch] : s->codeFreqBasis;   /* B3I :57 */
        double d2CarrError = 0, dCarrError = 0;                                  /* GLO :171-172 */
        caCode[0] = ca[L - 1]; memcpy(caCode + 1, ca, sizeof(double) * L); caCode[L + 1] = ca[0];   /* :158 */
        double codeFreq = codeFreqCentre, remCodePhase = 0.0;                    /* :163-165 */
        double carrFreq = acquiredFreq[ch], carrFreqBasis = acquiredFreq[ch], remCarrPhase = 0.0;   /* :167-170 */
        double oldCodeNco = 0, oldCodeError = 0, oldCarrNco = 0, oldCarrError = 0;           /* :173-178 */
        int vsmCnt = 0;
        for (int loopCnt = 1; loopCnt <= nEpochs; loopCnt++) {                   /* :184 */
            F(0)[loopCnt - 1] = (double)pos / 2;                                 /* :215 */
            double codePhaseStep = codeFreq / s->samplingFreq;                   /* :219 */
            int blksize = (int)ceil((s->codeLength - remCodePhase) / codePhaseStep);         /* :222 */
            if (pos + 2 * (size_t)blksize > nBytes) { abortAll = 1; break; }     /* :241-245 */
            const int8_t* raw = iq + pos; pos += 2 * (size_t)blksize;            /* :226 */
            F(13)[loopCnt - 1] = remCodePhase;                                   /* :249 */
            double aE = remCodePhase - earlyLateSpc, bE = (blksize - 1) * codePhaseStep + remCodePhase - earlyLateSpc;   /* :252 */
            double aL = remCodePhase + earlyLateSpc, bL = (blksize - 1) * codePhaseStep + remCodePhase + earlyLateSpc;   /* :259 */
            double aP = remCodePhase, bP = (blksize - 1) * codePhaseStep + remCodePhase;                                 /* :266 */
            int nE_, nL_, nP_; double cE, cL, cP;
            /* E1C :236-256: the three vectors are (rem -/+ spc)*2 : step*2 : (...)*2 */
            aE *= sub; bE *= sub; aL *= sub; bL *= sub; aP *= sub; bP *= sub;
            const double tstep = codePhaseStep * sub;
            colon_setup(aE, tstep, bE, &nE_, &cE);
            colon_setup(aL, tstep, bL, &nL_, &cL);
            colon_setup(aP, tstep, bP, &nP_, &cP);
            F(14)[loopCnt - 1] = remCarrPhase;                                   /* :277 */
            double w = carrFreq * 2.0 * M_PI;                                    /* :281 */
            double I_E = 0, Q_E = 0, I_P = 0, Q_P = 0, I_L = 0, Q_L = 0;
            double I_Ec = 0, Q_Ec = 0, I_Pc = 0, Q_Pc = 0, I_Lc = 0, Q_Lc = 0;                /* E1C pilot sums :285-290 */
            for (int n = 0; n < blksize; n++) {
                const int iE = (int)ceil(colon_elem(aE, tstep, cE, nE_, n));
                const int iL = (int)ceil(colon_elem(aL, tstep, cL, nL_, n));
                const int iP = (int)ceil(colon_elem(aP, tstep, cP, nP_, n));
                double e = caCode[iE];                                           /* :255-256 */
                double l = caCode[iL];                                           /* :262-263 */
                double p = caCode[iP];                                           /* :269-270 */
                double trig = (w * ((double)n / s->samplingFreq)) + remCarrPhase;            /* :280-281 */
                double c = cos(trig), sn = sin(trig);                            /* :287 exp(-1i*trig) = c - i*sn */
                double xr = raw[2 * n], xi = raw[2 * n + 1];                     /* :233-235 */
                if (s->glo == 1) { double t_ = xr; xr = xi; xi = t_; }                /* GLO :227 rawSignal2 + 1i*rawSignal1 */
                double iBB = c * xr + sn * xi, qBB = c * xi - sn * xr;           /* :291-292 */
                I_E += e * iBB; Q_E += e * qBB; I_P += p * iBB; Q_P += p * qBB; I_L += l * iBB; Q_L += l * qBB;   /* :295-300 */
                if (pilot) {
                    I_Ec += pCode[iE] * iBB; Q_Ec += pCode[iE] * qBB; I_Pc += pCode[iP] * iBB; Q_Pc += pCode[iP] * qBB;
                    I_Lc += pCode[iL] * iBB; Q_Lc += pCode[iL] * qBB;
                }
            }
            remCodePhase = (colon_elem(aP, tstep, cP, nP_, blksize - 1) / sub + codePhaseStep) - s->codeLength;  /* :273; E1C :263 */
            double trigEnd = (w * ((double)blksize / s->samplingFreq)) + remCarrPhase;
            remCarrPhase = fmod(trigEnd, 2 * M_PI);                              /* :283 */
            double carrError = atan(Q_P / I_P) / (2.0 * M_PI);                   /* :305 */
            if (pilot) carrError = (carrError + atan(Q_Pc / I_Pc) / (2.0 * M_PI)) / 2;       /* E1C :297-300 */
            double carrNco;
            if (s->glo == 0) {
                carrNco = oldCarrNco + (tau2carr / tau1carr) * (carrError - oldCarrError) + carrError * (PDIcarr / tau1carr);   /* :308 */
                oldCarrNco = carrNco; oldCarrError = carrError;
            } else {                                                             /* GLO :282-285 */
                d2CarrError = d2CarrError + carrError * pf3;
                dCarrError = d2CarrError + carrError * pf2 + dCarrError;
                carrNco = dCarrError + carrError * pf1;
            }
            F(2)[loopCnt - 1] = carrFreq;                                        /* :314 */
            carrFreq = carrFreqBasis + carrNco;                                  /* :317 */
            double sE = sqrt(I_E * I_E + Q_E * Q_E), sL = sqrt(I_L * I_L + Q_L * Q_L);
            double codeError = (sE - sL) / (sE + sL);                            /* :322 */
            if (pilot) {                                                         /* E1C :327-333 */
                double sEc = sqrt(I_Ec * I_Ec + Q_Ec * Q_Ec), sLc = sqrt(I_Lc * I_Lc + Q_Lc * Q_Lc);
                codeError = (codeError + (sEc - sLc) / (sEc + sLc)) / 2;
            }
            double codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code);   /* :326 */
            oldCodeNco = codeNco; oldCodeError = codeError;
            F(1)[loopCnt - 1] = codeFreq;                                        /* :332 */
            codeFreq = codeFreqCentre - codeNco;                                 /* :335 ; B3I :146 */
            F(9)[loopCnt - 1] = codeError; F(10)[loopCnt - 1] = codeNco;         /* :338-341 */
            F(11)[loopCnt - 1] = carrError; F(12)[loopCnt - 1] = carrNco;
            F(4)[loopCnt - 1] = I_E; F(3)[loopCnt - 1] = I_P; F(5)[loopCnt - 1] = I_L;       /* :343-348 */
            F(6)[loopCnt - 1] = Q_E; F(7)[loopCnt - 1] = Q_P; F(8)[loopCnt - 1] = Q_L;
            if (loopCnt % s->CNo_VSMinterval == 0) {                             /* :351 */
                vsmCnt++;
                int lo = loopCnt - s->CNo_VSMinterval;
                vsmValue[(size_t)ch * nV + vsmCnt - 1] = orc_CNoVSM(F(3) + lo, F(7) + lo, s->CNo_VSMinterval, s->CNo_accTime);   /* :353 */
                vsmIndex[(size_t)ch * nV + vsmCnt - 1] = loopCnt;                /* :356 */
            }
            epochsDone[ch] = loopCnt;
        }
        free(ca); free(caCode); free(pCode);
#undef F
    }
    /* MATLAB `return` on a short read ends the whole function: channels after the first one that
     * stopped early stay untouched (init fills).  Restore that when channels ran concurrently. */
    if (parallel) {
        int failed = -1;
        for (int ch = 0; ch < nCh && failed < 0; ch++)
            if (!(s->glo == 1 ? (PRN[ch] == INT32_MIN) : (PRN[ch] == 0)) && epochsDone[ch] < nEpochs) failed = ch;
        for (int ch = failed + 1; failed >= 0 && ch < nCh; ch++) {
            double* o = out + (size_t)ch * ORC_NFIELDS * nEpochs;
            for (int f = 0; f < ORC_NFIELDS; f++) {
                double fill = (f == 0 || (f >= 3 && f <= 8)) ? 0.0 : INFINITY;
                for (int e = 0; e < nEpochs; e++) o[(size_t)f * nEpochs + e] = fill;
            }
            for (int i = 0; i < nV; i++) { vsmValue[(size_t)ch * nV + i] = 0; vsmIndex[(size_t)ch * nV + i] = 0; }
            epochsDone[ch] = 0;
        }
    }
    return 0;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
