/*
 * gnsscorr_mex.c — thin MEX gateway: marshals mxArray <-> the POD arguments of the C ABI in
 * include/gnsscorr.h and nothing else.  Build (MATLAB R2018a+, interleaved complex):
 *
 *     mex -R2018a -I../include gnsscorr_mex.c -L../cu-sdr-collection_b200 -lgnsscorr
 *
 * Called by the drop-in wrappers matlab/acquisition.m and matlab/tracking.m, which keep the
 * reference signatures (GPS/GPS_L1CA/include/acquisition.m:1, tracking.m:1).
 *
 *   r = gnsscorr_mex('acquire', cfg, iq_int8, svList)
 *         cfg     : struct with the gc_config field names (doubles)
 *         iq_int8 : int8 vector, I,Q interleaved (longSignal as stored in the file)
 *         svList  : double vector of PRNs (GLONASS: frequency numbers K)
 *         r       : struct carrFreq, codePhase, peakMetric (1x32 double; GLONASS 1x21, index K+8)
 *   r = gnsscorr_mex('track', cfg, path, prn, acqFreq, codePhase, nEpochs [, codeFreq0 [, codes]])
 *         r       : struct out (nEpochs x 15 x nCh double, MATLAB column-major view of the
 *                   C [nCh][15][nEpochs] block), vsmValue, vsmIndex, epochsDone
 *   Signals with caller-supplied codes (gc_set_code: GAL E1C, GPS L5C, GAL E5a/E5b, BDS B2a): 'acquire' takes a
 *   5th and 'track' a 9th argument
 *         codes   : struct sv (double vector of PRNs), data, pilot (int8, codeLength x numel(sv), one
 *                   column of +-1 primary chips per PRN - generateE1Bcode(PRN)(1:2:end), generateL5Icode(PRN, settings),
 *                   ...) and for GAL E5a secondary (int8, 100 x numel(sv), generateE5aQ_secondary(PRN));
 *         with a quadrature pilot tracked r.out is nEpochs x 17 x nCh (rows 16, 17 = Pilot_I_P, Pilot_Q_P)
 *
 * The engine behind the gateway is a gc_multi handle (cfg.n_gpus GPUs, default 1; 0 = every visible GPU) that is CREATED ONCE and
 * kept between calls: acquisition() and tracking() are called once each per run (postProcessing.m:100, 124) and a handle costs
 * ~0.8 s to build (CUDA context, FFT plan, twiddles, replica spectra, work buffer), so the gateway caches it keyed on the whole
 * gc_config + n_gpus + a hash of the codes struct, rebuilds it when any of those change, and frees it in mexAtExit
 * ('gnsscorr_mex(''reset'')' frees it on demand).
 *
 * This file cannot be exercised in the build image (no MATLAB); it is compile-checked against
 * matlab/stub/mex.h and the same C entry points are exercised from Python (ctypes).
 */
#include <string.h>
#include "mex.h"
#include "gnsscorr.h"

static double field(const mxArray* s, const char* name)
{
    const mxArray* f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("gnsscorr:cfg", "settings field %s is missing", name);
    return mxGetScalar(f);
}

static void fill_config(const mxArray* s, gc_config* c)
{
    memset(c, 0, sizeof(*c));
    c->abi_version = GC_ABI_VERSION;
    c->device = (int32_t)field(s, "device");
    c->signal = (int32_t)field(s, "signal");
    c->freq_spacing = field(s, "freq_spacing");
    c->file_type = (int32_t)field(s, "file_type");
    c->sample_bytes = (int32_t)field(s, "sample_bytes");
    c->code_length = (int32_t)field(s, "code_length");
    c->acq_noncoh_time = (int32_t)field(s, "acq_noncoh_time");
    c->cno_vsm_interval = (int32_t)field(s, "cno_vsm_interval");
    c->skip_number_of_bytes = (int64_t)field(s, "skip_number_of_bytes");
    c->sampling_freq = field(s, "sampling_freq");
    c->IF = field(s, "IF");
    c->code_freq_basis = field(s, "code_freq_basis");
    c->acq_search_band = field(s, "acq_search_band");
    c->acq_search_step = field(s, "acq_search_step");
    c->acq_threshold = field(s, "acq_threshold");
    c->dll_damping_ratio = field(s, "dll_damping_ratio");
    c->dll_noise_bandwidth = field(s, "dll_noise_bandwidth");
    c->dll_correlator_spacing = field(s, "dll_correlator_spacing");
    c->pll_damping_ratio = field(s, "pll_damping_ratio");
    c->pll_noise_bandwidth = field(s, "pll_noise_bandwidth");
    c->int_time = field(s, "int_time");
    c->cno_acc_time = field(s, "cno_acc_time");
    c->pilot_trk_flag = mxGetField(s, 0, "pilot_trk_flag") ? (int32_t)field(s, "pilot_trk_flag") : 0;
    c->acq_coh_t = mxGetField(s, 0, "acq_coh_t") ? (int32_t)field(s, "acq_coh_t") : 0;
    c->pilot_acq_flag = mxGetField(s, 0, "pilot_acq_flag") ? (int32_t)field(s, "pilot_acq_flag") : 0;
}

/* ---- the cached engine ------------------------------------------------------------------------------------------------------ */
static gc_multi* g_m = NULL;
static gc_config g_cfg;
static int g_ngpus = -1;
static unsigned long long g_codes_hash = 0;

static void drop_engine(void)
{
    if (g_m) gc_multi_destroy(g_m);
    g_m = NULL;
    g_ngpus = -1;
    g_codes_hash = 0;
}

static void fail_now(int rc, const char* what)
{
    char msg[512];
    strncpy(msg, gc_multi_last_error(g_m), sizeof(msg) - 1);
    msg[sizeof(msg) - 1] = 0;
    drop_engine();                                         /* a failed call leaves no half-configured handle behind */
    mexErrMsgIdAndTxt("gnsscorr:fail", "%s failed (%d): %s", what, rc, msg);
}

static void check(int rc, const char* what)
{
    if (rc != GC_OK) fail_now(rc, what);
}

static unsigned long long fnv(unsigned long long h, const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    size_t i;
    for (i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ULL; }
    return h;
}

static unsigned long long codes_hash(const mxArray* codes)
{
    static const char* f[] = {"sv", "data", "pilot", "secondary", "cl", "boc61"};
    unsigned long long h = 1469598103934665603ULL;
    int i;
    for (i = 0; i < 6; ++i) {
        const mxArray* a = mxGetField(codes, 0, f[i]);
        if (a) h = fnv(h, mxGetData(a), mxGetNumberOfElements(a) * mxGetElementSize(a));
        h = fnv(h, &i, sizeof(i));
    }
    return h;
}

/* the handle for this configuration: the cached one, or a new one */
static void get_engine(const gc_config* cfg, int nGpus)
{
    if (g_m && g_ngpus == nGpus && memcmp(&g_cfg, cfg, sizeof(*cfg)) == 0) return;
    drop_engine();
    {
        const int rc = gc_multi_create(&g_m, cfg, nGpus);
        if (rc != GC_OK) {
            char msg[512];
            strncpy(msg, gc_multi_last_error(NULL), sizeof(msg) - 1);
            msg[sizeof(msg) - 1] = 0;
            g_m = NULL;
            mexErrMsgIdAndTxt("gnsscorr:fail", "gc_multi_create failed (%d): %s", rc, msg);
        }
    }
    g_cfg = *cfg;
    g_ngpus = nGpus;
    mexAtExit(drop_engine);
}

/* codes struct -> gc_multi_set_code for every listed PRN; skipped when the same codes are already set on the cached handle */
static void set_codes(const gc_config* cfg, const mxArray* codes)
{
    const mxArray *sv = mxGetField(codes, 0, "sv"), *d = mxGetField(codes, 0, "data"), *p = mxGetField(codes, 0, "pilot");
    const mxArray* sec = mxGetField(codes, 0, "secondary");   /* GAL E5a only: int8 100 x numel(sv) */
    const mxArray* cl = mxGetField(codes, 0, "cl");           /* GPS L2C with pilotTRKflag: int8 (150*codeLength) x numel(sv) */
    const mxArray* b61 = mxGetField(codes, 0, "boc61");       /* BDS B1C with pilotTRKflag == 2: int8 (12*codeLength) x numel(sv) */
    const mwSize clLen = (mwSize)150 * cfg->code_length, b61Len = (mwSize)12 * cfg->code_length;
    const int single = (cfg->signal == GC_SIG_BDS_B1I || cfg->signal == GC_SIG_GPS_L2C);   /* one code per SV */
    unsigned long long hsh;
    mwSize i, n, len;
    if (!mxIsStruct(codes) || !sv || !d || !p || !mxIsDouble(sv) || !mxIsInt8(d) || !mxIsInt8(p))
        mexErrMsgIdAndTxt("gnsscorr:args", "codes: struct with sv (double), data (int8), pilot (int8)");
    /* (B1C passes the 2*codeLength BOC(1,1) sub-chip sequences of generateDataBOC11 / generatePilotBOC11) */
    n = mxGetNumberOfElements(sv);
    len = n ? mxGetNumberOfElements(d) / n : 0;         /* codeLength, or 2*codeLength for the return-to-zero L2C CM code */
    if (n == 0 || mxGetNumberOfElements(d) != n * len || mxGetNumberOfElements(p) != n * len)
        mexErrMsgIdAndTxt("gnsscorr:args", "codes: data and pilot must be nChips x numel(sv)");
    hsh = codes_hash(codes);
    if (hsh == g_codes_hash) return;
    for (i = 0; i < n; ++i) {
        int rc = gc_multi_set_code(g_m, (int32_t)mxGetDoubles(sv)[i], 0, (const int8_t*)mxGetInt8s(d) + i * len, (int32_t)len);
        if (rc == GC_OK && !single) rc = gc_multi_set_code(g_m, (int32_t)mxGetDoubles(sv)[i], 1, (const int8_t*)mxGetInt8s(p) + i * len, (int32_t)len);
        if (rc == GC_OK && sec && mxIsInt8(sec) && mxGetNumberOfElements(sec) == n * 100)
            rc = gc_multi_set_code(g_m, (int32_t)mxGetDoubles(sv)[i], 2, (const int8_t*)mxGetInt8s(sec) + i * 100, 100);
        if (rc == GC_OK && cl && mxIsInt8(cl) && mxGetNumberOfElements(cl) == n * clLen)      /* GPS L2C CL pilot (generateCLcode.m) */
            rc = gc_multi_set_code(g_m, (int32_t)mxGetDoubles(sv)[i], 1, (const int8_t*)mxGetInt8s(cl) + i * clLen, (int32_t)clLen);
        if (rc == GC_OK && b61 && mxIsInt8(b61) && mxGetNumberOfElements(b61) == n * b61Len)   /* BDS B1C full band (generatePilotBOC61.m) */
            rc = gc_multi_set_code(g_m, (int32_t)mxGetDoubles(sv)[i], 2, (const int8_t*)mxGetInt8s(b61) + i * b61Len, (int32_t)b61Len);
        if (rc != GC_OK) fail_now(rc, "gc_multi_set_code");
    }
    g_codes_hash = hsh;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    char cmd[16];
    gc_config cfg;
    int nGpus = 1;
    (void)nlhs;
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof(cmd))) mexErrMsgIdAndTxt("gnsscorr:args", "usage: gnsscorr_mex(cmd, cfg, ...)");
    if (!strcmp(cmd, "reset")) { drop_engine(); return; }
    if (nrhs < 2 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("gnsscorr:args", "usage: gnsscorr_mex(cmd, cfg, ...) with cfg a struct");
    fill_config(prhs[1], &cfg);
    if (mxGetField(prhs[1], 0, "n_gpus")) nGpus = (int)field(prhs[1], "n_gpus");

    if (!strcmp(cmd, "acquire")) {
        const char* names[] = {"carrFreq", "codePhase", "peakMetric", "CLCodePhase"};
        const int n = gc_acq_result_len(cfg.signal);
        int32_t sv[64];
        mwSize i, nSv, nElem;
        size_t nSamples;
        const int is16 = cfg.sample_bytes == 2, packed = cfg.file_type == GC_FILE_PACKED2;
        /* every argument is checked before it is touched */
        if ((nrhs != 4 && nrhs != 5) || !mxIsDouble(prhs[3]) || mxGetNumberOfElements(prhs[3]) < 1 || mxGetNumberOfElements(prhs[3]) > 64)
            mexErrMsgIdAndTxt("gnsscorr:args", "acquire: gnsscorr_mex('acquire', cfg, samples, svList [, codes]) with 1..64 SVs");
        /* longSignal in the file's own samples: int8 or int16 (cfg.sample_bytes), I,Q pairs (file_type 2), real values (file_type 1),
         * or the 2-bit packed records of unpack_cplx.m as uint8, two complex samples per byte (GC_FILE_PACKED2) */
        if (packed ? !mxIsUint8(prhs[2]) : !(is16 ? mxIsInt16(prhs[2]) : mxIsInt8(prhs[2])))
            mexErrMsgIdAndTxt("gnsscorr:args", "acquire: samples must be int8 / int16 as settings.dataType says (uint8 for the 2-bit packed format)");
        nSv = mxGetNumberOfElements(prhs[3]);
        nElem = mxGetNumberOfElements(prhs[2]);
        nSamples = packed ? (size_t)nElem * 2 : cfg.file_type == 1 ? (size_t)nElem : (size_t)nElem / 2;
        get_engine(&cfg, nGpus);
        if (nrhs == 5) set_codes(&cfg, prhs[4]);
        for (i = 0; i < nSv; ++i) sv[i] = (int32_t)mxGetDoubles(prhs[3])[i];
        plhs[0] = mxCreateStructMatrix(1, 1, 4, names);
        for (i = 0; i < 4; ++i) mxSetField(plhs[0], 0, names[i], mxCreateDoubleMatrix(1, n, mxREAL));
        check(gc_multi_acquire_host(g_m, (const int8_t*)mxGetData(prhs[2]), nSamples, (int32_t)nSv, sv,
                                    mxGetDoubles(mxGetField(plhs[0], 0, "carrFreq")), mxGetDoubles(mxGetField(plhs[0], 0, "codePhase")),
                                    mxGetDoubles(mxGetField(plhs[0], 0, "peakMetric")), NULL, NULL),
              "gc_multi_acquire_host");
        if (cfg.signal == GC_SIG_GPS_L2C && cfg.pilot_trk_flag == 1) {     /* acqResults.CLCodePhase (GPS_L2C acquisition.m:136) */
            int32_t clp[32];
            double* o = mxGetDoubles(mxGetField(plhs[0], 0, "CLCodePhase"));
            check(gc_multi_get_cl_code_phase(g_m, clp), "gc_multi_get_cl_code_phase");
            for (i = 0; i < 32 && i < (mwSize)n; ++i) o[i] = (double)clp[i];
        }
    } else if (!strcmp(cmd, "track")) {
        const char* names[] = {"out", "vsmValue", "vsmIndex", "epochsDone", "cnoPld"};
        char path[4096];
        mwSize nCh, nV, i, dims[3];
        int32_t nEpochs;
        int32_t prn[256];
        mxArray *out, *vv, *vi, *done;
        if (nrhs < 7 || nrhs > 9 || !mxIsChar(prhs[2]) || mxGetString(prhs[2], path, sizeof(path)) || !mxIsDouble(prhs[3]) ||
            !mxIsDouble(prhs[4]) || !mxIsDouble(prhs[5]) || !mxIsDouble(prhs[6]) || mxGetNumberOfElements(prhs[6]) != 1)
            mexErrMsgIdAndTxt("gnsscorr:args", "track: gnsscorr_mex('track', cfg, path, prn, acqFreq, codePhase, nEpochs [, codeFreq0 [, codes]])");
        nCh = mxGetNumberOfElements(prhs[3]);
        nEpochs = (int32_t)mxGetScalar(prhs[6]);
        if (nCh < 1 || nCh > 256 || mxGetNumberOfElements(prhs[4]) != nCh || mxGetNumberOfElements(prhs[5]) != nCh || nEpochs < 1 ||
            (nrhs >= 8 && !mxIsEmpty(prhs[7]) && (!mxIsDouble(prhs[7]) || mxGetNumberOfElements(prhs[7]) != nCh)))
            mexErrMsgIdAndTxt("gnsscorr:args", "track: prn, acqFreq, codePhase (and codeFreq0) need one entry per channel (1..256), nEpochs >= 1");
        nV = nEpochs / cfg.cno_vsm_interval;
        get_engine(&cfg, nGpus);
        if (nrhs == 9) set_codes(&cfg, prhs[8]);
        if (mxGetField(prhs[1], 0, "wb_factor"))                          /* factor = CalcWeighingFactor(settings), B1C WB_tracking.m:124 */
            check(gc_multi_set_param(g_m, GC_PARAM_B1C_WB_FACTOR, field(prhs[1], "wb_factor")), "gc_multi_set_param");
        if (nrhs == 9 && mxGetField(prhs[8], 0, "clCodePhase")) {         /* channel.CLCodePhase (GPS_L2C tracking.m:162) */
            const mxArray* a = mxGetField(prhs[8], 0, "clCodePhase");
            int32_t clp[256];
            if (!mxIsDouble(a) || mxGetNumberOfElements(a) != nCh) mexErrMsgIdAndTxt("gnsscorr:args", "track: clCodePhase must have one entry per channel");
            for (i = 0; i < nCh; ++i) clp[i] = (int32_t)mxGetDoubles(a)[i];
            check(gc_multi_set_cl_code_phase(g_m, (int32_t)nCh, clp), "gc_multi_set_cl_code_phase");
        }
        for (i = 0; i < nCh; ++i) prn[i] = (mxGetDoubles(prhs[3])[i] != mxGetDoubles(prhs[3])[i]) ? GC_SV_NONE : (int32_t)mxGetDoubles(prhs[3])[i];   /* NaN = channel off (GLONASS) */
        dims[0] = nEpochs; dims[1] = gc_track_nfields(gc_multi_handle(g_m, 0)); dims[2] = nCh;   /* 15, 17 with Pilot_I_P / Pilot_Q_P, 21 with all six pilot rows */
        out = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        vv = mxCreateDoubleMatrix(nV, nCh, mxREAL);
        vi = mxCreateDoubleMatrix(nV, nCh, mxREAL);
        done = mxCreateNumericMatrix(1, nCh, mxINT32_CLASS, mxREAL);
        check(gc_multi_track_file(g_m, path, (int32_t)nCh, prn, mxGetDoubles(prhs[4]), mxGetDoubles(prhs[5]),
                                  (nrhs >= 8 && !mxIsEmpty(prhs[7])) ? mxGetDoubles(prhs[7]) : NULL, nEpochs,
                                  mxGetDoubles(out), mxGetDoubles(vv), mxGetDoubles(vi), (int32_t*)mxGetInt32s(done)),
              "gc_multi_track_file");
        plhs[0] = mxCreateStructMatrix(1, 1, 5, names);
        if (cfg.signal == GC_SIG_BDS_B2A || cfg.signal == GC_SIG_BDS_B1C) {
            /* r.cnoPld: nIntervals x 5 x nCh = DataCNo, DataPLD, PilotCNo, PilotPLD, total C/N0 (Calc_CNo_PLD.m on the device); with
             * several GPUs every GPU holds its own block of channels, in channel order */
            mwSize d3[3], c0 = 0;
            mxArray* pld;
            int g;
            const int nG = gc_multi_n_gpus(g_m);
            const mwSize per = (nCh + nG - 1) / nG;
            d3[0] = nV; d3[1] = GC_CNO_PLD_ROWS; d3[2] = nCh;
            pld = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
            for (g = 0; g < nG && c0 < nCh; ++g, c0 += per) {
                const mwSize nc = (nCh - c0 < per) ? nCh - c0 : per;
                if (gc_get_cno_pld(gc_multi_handle(g_m, g), (int32_t)nc, (int32_t)nV, mxGetDoubles(pld) + c0 * GC_CNO_PLD_ROWS * nV) != GC_OK)
                    fail_now(GC_ERR_ARG, "gc_get_cno_pld");
            }
            mxSetField(plhs[0], 0, "cnoPld", pld);
        }
        mxSetField(plhs[0], 0, "out", out);
        mxSetField(plhs[0], 0, "vsmValue", vv);
        mxSetField(plhs[0], 0, "vsmIndex", vi);
        mxSetField(plhs[0], 0, "epochsDone", done);
    } else if (!strcmp(cmd, "navsync")) {
        /* r = gnsscorr_mex('navsync', cfg, I_P)  with I_P = nEpochs x nCh (one column per channel): the front end of
         * NAVdecoding.m:69-170 -> r.subFrameStart (1 x nCh, 0 = none), r.navBits (GC_NAV_BITS x nCh uint8), r.bitsValid */
        const char* names[] = {"subFrameStart", "navBits", "bitsValid"};
        mxArray *sfs, *bits, *valid;
        mwSize nE, nCh;
        if (nrhs != 3 || !mxIsDouble(prhs[2]) || mxIsEmpty(prhs[2])) mexErrMsgIdAndTxt("gnsscorr:args", "navsync: I_P must be a non-empty double matrix");
        nE = mxGetM(prhs[2]); nCh = mxGetN(prhs[2]);
        get_engine(&cfg, nGpus);
        sfs = mxCreateNumericMatrix(1, nCh, mxINT32_CLASS, mxREAL);
        bits = mxCreateNumericMatrix(GC_NAV_BITS, nCh, mxUINT8_CLASS, mxREAL);
        valid = mxCreateNumericMatrix(1, nCh, mxINT32_CLASS, mxREAL);
        {
            gc_handle* h0 = gc_multi_handle(g_m, 0);
            const int rc = gc_nav_sync(h0, (int32_t)nCh, (int32_t)nE, mxGetDoubles(prhs[2]), (int32_t*)mxGetInt32s(sfs), (uint8_t*)mxGetData(bits),
                                       (int32_t*)mxGetInt32s(valid));
            if (rc != GC_OK) {
                char msg[512];
                strncpy(msg, gc_last_error(h0), sizeof(msg) - 1);
                msg[sizeof(msg) - 1] = 0;
                drop_engine();
                mexErrMsgIdAndTxt("gnsscorr:fail", "gc_nav_sync failed (%d): %s", rc, msg);
            }
        }
        plhs[0] = mxCreateStructMatrix(1, 1, 3, names);
        mxSetField(plhs[0], 0, "subFrameStart", sfs);
        mxSetField(plhs[0], 0, "navBits", bits);
        mxSetField(plhs[0], 0, "bitsValid", valid);
    } else {
        mexErrMsgIdAndTxt("gnsscorr:args", "unknown command %s", cmd);
    }
}
