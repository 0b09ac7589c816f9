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

/* codes struct -> gc_set_code for every listed PRN (Galileo E1) */
static void set_codes(gc_handle* h, const gc_config* cfg, const mxArray* codes)
{
    const mxArray *sv = mxGetField(codes, 0, "sv"), *d = mxGetField(codes, 0, "data"), *p = mxGetField(codes, 0, "pilot");
    const mxArray* sec = mxGetField(codes, 0, "secondary");   /* GAL E5a only: int8 100 x numel(sv) */
    const mxArray* cl = mxGetField(codes, 0, "cl");           /* GPS L2C with pilotTRKflag: int8 (150*codeLength) x numel(sv) */
    const mxArray* b61 = mxGetField(codes, 0, "boc61");       /* BDS B1C with pilotTRKflag == 2: int8 (12*codeLength) x numel(sv) */
    const mwSize clLen = (mwSize)150 * cfg->code_length, b61Len = (mwSize)12 * cfg->code_length;
    mwSize i, n;
    if (!sv || !d || !p || !mxIsInt8(d) || !mxIsInt8(p)) { gc_destroy(h); mexErrMsgIdAndTxt("gnsscorr:args", "codes: struct with sv, data (int8), pilot (int8)"); }
    const int single = (cfg->signal == GC_SIG_BDS_B1I || cfg->signal == GC_SIG_GPS_L2C);   /* one code per SV */
    /* (B1C passes the 2*codeLength BOC(1,1) sub-chip sequences of generateDataBOC11 / generatePilotBOC11) */
    mwSize len;
    n = mxGetNumberOfElements(sv);
    len = n ? mxGetNumberOfElements(d) / n : 0;         /* codeLength, or 2*codeLength for the return-to-zero L2C CM code */
    if (n == 0 || mxGetNumberOfElements(d) != n * len || mxGetNumberOfElements(p) != n * len) {
        gc_destroy(h);
        mexErrMsgIdAndTxt("gnsscorr:args", "codes: data and pilot must be nChips x numel(sv)");
    }
    for (i = 0; i < n; ++i) {
        int rc = gc_set_code(h, (int32_t)mxGetDoubles(sv)[i], 0, (const int8_t*)mxGetInt8s(d) + i * len, (int32_t)len);
        if (rc == GC_OK && !single) rc = gc_set_code(h, (int32_t)mxGetDoubles(sv)[i], 1, (const int8_t*)mxGetInt8s(p) + i * len, (int32_t)len);
        if (rc == GC_OK && sec && mxIsInt8(sec) && mxGetNumberOfElements(sec) == n * 100)
            rc = gc_set_code(h, (int32_t)mxGetDoubles(sv)[i], 2, (const int8_t*)mxGetInt8s(sec) + i * 100, 100);
        if (rc == GC_OK && cl && mxIsInt8(cl) && mxGetNumberOfElements(cl) == n * clLen)      /* GPS L2C CL pilot (generateCLcode.m) */
            rc = gc_set_code(h, (int32_t)mxGetDoubles(sv)[i], 1, (const int8_t*)mxGetInt8s(cl) + i * clLen, (int32_t)clLen);
        if (rc == GC_OK && b61 && mxIsInt8(b61) && mxGetNumberOfElements(b61) == n * b61Len)   /* BDS B1C full band (generatePilotBOC61.m) */
            rc = gc_set_code(h, (int32_t)mxGetDoubles(sv)[i], 2, (const int8_t*)mxGetInt8s(b61) + i * b61Len, (int32_t)b61Len);
        if (rc != GC_OK) {
            char msg[512];
            strncpy(msg, gc_last_error(h), sizeof(msg) - 1);
            msg[sizeof(msg) - 1] = 0;
            gc_destroy(h);
            mexErrMsgIdAndTxt("gnsscorr:fail", "gc_set_code failed (%d): %s", rc, msg);
        }
    }
}

static void check(gc_handle* h, int rc, const char* what)
{
    if (rc != GC_OK) {
        char msg[512];
        strncpy(msg, gc_last_error(h), sizeof(msg) - 1);
        msg[sizeof(msg) - 1] = 0;
        if (h) gc_destroy(h);
        mexErrMsgIdAndTxt("gnsscorr:fail", "%s failed (%d): %s", what, rc, msg);
    }
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    char cmd[16];
    gc_config cfg;
    gc_handle* h = NULL;
    (void)nlhs;
    if (nrhs < 2 || mxGetString(prhs[0], cmd, sizeof(cmd))) mexErrMsgIdAndTxt("gnsscorr:args", "usage: gnsscorr_mex(cmd, cfg, ...)");
    fill_config(prhs[1], &cfg);
    check(NULL, gc_create(&h, &cfg), "gc_create");

    if (!strcmp(cmd, "acquire")) {
        const char* names[] = {"carrFreq", "codePhase", "peakMetric", "CLCodePhase"};
        const int n = gc_acq_result_len(cfg.signal);
        const mwSize nSv = mxGetNumberOfElements(prhs[3]);
        const double* svd = mxGetDoubles(prhs[3]);
        int32_t sv[64];
        mwSize i;
        /* longSignal in the file's own samples: int8 or int16 (cfg.sample_bytes), I,Q pairs or real values (cfg.file_type) */
        const int is16 = cfg.sample_bytes == 2;
        const size_t perSample = cfg.file_type == 1 ? 1 : 2;
        if ((nrhs != 4 && nrhs != 5) || !(is16 ? mxIsInt16(prhs[2]) : mxIsInt8(prhs[2])) || nSv > 64) { gc_destroy(h); mexErrMsgIdAndTxt("gnsscorr:args", "acquire: bad arguments"); }
        if (nrhs == 5) set_codes(h, &cfg, prhs[4]);
        for (i = 0; i < nSv; ++i) sv[i] = (int32_t)svd[i];
        plhs[0] = mxCreateStructMatrix(1, 1, 4, names);
        for (i = 0; i < 4; ++i) mxSetField(plhs[0], 0, names[i], mxCreateDoubleMatrix(1, n, mxREAL));
        check(h, gc_acquire_host(h, (const int8_t*)mxGetData(prhs[2]), mxGetNumberOfElements(prhs[2]) / perSample, (int32_t)nSv, sv,
                                 mxGetDoubles(mxGetField(plhs[0], 0, "carrFreq")), mxGetDoubles(mxGetField(plhs[0], 0, "codePhase")),
                                 mxGetDoubles(mxGetField(plhs[0], 0, "peakMetric")), NULL, NULL),
              "gc_acquire_host");
        if (cfg.signal == GC_SIG_GPS_L2C && cfg.pilot_trk_flag == 1) {     /* acqResults.CLCodePhase (GPS_L2C acquisition.m:136) */
            int32_t clp[32];
            double* o = mxGetDoubles(mxGetField(plhs[0], 0, "CLCodePhase"));
            check(h, gc_get_cl_code_phase(h, clp), "gc_get_cl_code_phase");
            for (i = 0; i < 32 && i < (mwSize)n; ++i) o[i] = (double)clp[i];
        }
    } else if (!strcmp(cmd, "track")) {
        const char* names[] = {"out", "vsmValue", "vsmIndex", "epochsDone"};
        char path[4096];
        const mwSize nCh = mxGetNumberOfElements(prhs[3]);
        const int32_t nEpochs = (int32_t)mxGetScalar(prhs[6]);
        const mwSize nV = nEpochs / cfg.cno_vsm_interval;
        const double* prnd = mxGetDoubles(prhs[3]);
        mwSize dims[3];
        int32_t prn[256];
        mxArray *out, *vv, *vi, *done;
        mwSize i;
        if (nrhs < 7 || nrhs > 9 || mxGetString(prhs[2], path, sizeof(path)) || nCh > 256) { gc_destroy(h); mexErrMsgIdAndTxt("gnsscorr:args", "track: bad arguments"); }
        if (nrhs == 9) set_codes(h, &cfg, prhs[8]);
        if (mxGetField(prhs[1], 0, "wb_factor"))                          /* factor = CalcWeighingFactor(settings), B1C WB_tracking.m:124 */
            check(h, gc_set_param(h, GC_PARAM_B1C_WB_FACTOR, field(prhs[1], "wb_factor")), "gc_set_param");
        if (nrhs == 9 && mxGetField(prhs[8], 0, "clCodePhase")) {         /* channel.CLCodePhase (GPS_L2C tracking.m:162) */
            const mxArray* a = mxGetField(prhs[8], 0, "clCodePhase");
            int32_t clp[256];
            if (mxGetNumberOfElements(a) != nCh) { gc_destroy(h); mexErrMsgIdAndTxt("gnsscorr:args", "track: clCodePhase must have one entry per channel"); }
            for (i = 0; i < nCh; ++i) clp[i] = (int32_t)mxGetDoubles(a)[i];
            check(h, gc_set_cl_code_phase(h, (int32_t)nCh, clp), "gc_set_cl_code_phase");
        }
        for (i = 0; i < nCh; ++i) prn[i] = (prnd[i] != prnd[i]) ? GC_SV_NONE : (int32_t)prnd[i];   /* NaN = channel off (GLONASS) */
        dims[0] = nEpochs; dims[1] = gc_track_nfields(h); dims[2] = nCh;   /* 15, 17 with Pilot_I_P / Pilot_Q_P, 21 with all six pilot rows */
        out = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        vv = mxCreateDoubleMatrix(nV, nCh, mxREAL);
        vi = mxCreateDoubleMatrix(nV, nCh, mxREAL);
        done = mxCreateNumericMatrix(1, nCh, mxINT32_CLASS, mxREAL);
        check(h, gc_track_file(h, path, (int32_t)nCh, prn, mxGetDoubles(prhs[4]), mxGetDoubles(prhs[5]),
                               (nrhs >= 8 && !mxIsEmpty(prhs[7])) ? mxGetDoubles(prhs[7]) : NULL, nEpochs,
                               mxGetDoubles(out), mxGetDoubles(vv), mxGetDoubles(vi), (int32_t*)mxGetInt32s(done)),
              "gc_track_file");
        plhs[0] = mxCreateStructMatrix(1, 1, 4, names);
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
        if (nrhs != 3 || !mxIsDouble(prhs[2])) { gc_destroy(h); mexErrMsgIdAndTxt("gnsscorr:args", "navsync: I_P must be a double matrix"); }
        nE = mxGetM(prhs[2]); nCh = mxGetN(prhs[2]);
        sfs = mxCreateNumericMatrix(1, nCh, mxINT32_CLASS, mxREAL);
        bits = mxCreateNumericMatrix(GC_NAV_BITS, nCh, mxUINT8_CLASS, mxREAL);
        valid = mxCreateNumericMatrix(1, nCh, mxINT32_CLASS, mxREAL);
        check(h, gc_nav_sync(h, (int32_t)nCh, (int32_t)nE, mxGetDoubles(prhs[2]), (int32_t*)mxGetInt32s(sfs), (uint8_t*)mxGetData(bits),
                             (int32_t*)mxGetInt32s(valid)), "gc_nav_sync");
        plhs[0] = mxCreateStructMatrix(1, 1, 3, names);
        mxSetField(plhs[0], 0, "subFrameStart", sfs);
        mxSetField(plhs[0], 0, "navBits", bits);
        mxSetField(plhs[0], 0, "bitsValid", valid);
    } else {
        gc_destroy(h);
        mexErrMsgIdAndTxt("gnsscorr:args", "unknown command %s", cmd);
    }
    gc_destroy(h);
}
