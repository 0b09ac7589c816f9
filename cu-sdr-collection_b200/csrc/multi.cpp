// Multi-GPU engine behind the C ABI (include/gnsscorr.h, gc_multi_*): one gc_handle per GPU and one host thread per GPU and call.
//
// The units of both hot paths are independent - the PRN loop of acquisition.m:155 (frequency numbers K for GLONASS,
// GLO_GL1/include/acquisition.m:172-183) and the channel loop of tracking.m:133 - so the SV list is dealt round-robin and the
// channel list in contiguous blocks over the GPUs, every GPU works on its share with its own copy of the record, and the only
// exchange is the merge of the per-SV results (a few doubles per SV, copied back by each GPU's own gc_acquire and merged here).
// Built on the single-GPU entry points only, so that everything they guarantee (result layout, error codes, short-record
// semantics) carries over; the merged results are bit-identical to one GPU's because each SV / channel is computed by exactly
// the same code on exactly the same data.
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <functional>
#include <memory>
#include <mutex>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/gnsscorr.h"

namespace {
// one persistent host thread per GPU (a fresh std::thread per call costs a CUDA per-thread set-up of several milliseconds)
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> task;
    int rc = 0;
    bool busy = false, quit = false;
    void loop()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return busy || quit; });
            if (quit) return;
            lk.unlock();
            const int r = task();
            lk.lock();
            rc = r; busy = false;
            cv.notify_all();
        }
    }
    void start() { th = std::thread([this] { loop(); }); }
    void submit(std::function<int()> f)
    {
        std::lock_guard<std::mutex> lk(mu);
        task = std::move(f); busy = true;
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !busy; });
        return rc;
    }
    void stop()
    {
        { std::lock_guard<std::mutex> lk(mu); quit = true; cv.notify_all(); }
        if (th.joinable()) th.join();
    }
};
}  // namespace

struct gc_multi {
    gc_config cfg{};
    std::vector<gc_handle*> h;
    std::vector<std::unique_ptr<Worker>> workers;   // workers[g] serves GPU g (GPU 0 runs on the calling thread)
    std::string err;
    std::vector<int32_t> clPhase;      // channel.CLCodePhase of the next gc_multi_track (GPS L2C CL pilot)
    double lastAcqMs = 0, lastTrackMs = 0;
};

namespace {

thread_local std::string g_multi_create_error;

int fail(gc_multi* m, int code, const std::string& msg)
{
    if (m) m->err = msg; else g_multi_create_error = msg;
    return code;
}

bool is_glo(const gc_multi* m) { return m->cfg.signal == GC_SIG_GLO_G1G2; }
int result_index(const gc_multi* m, int sv) { return is_glo(m) ? sv + 7 : sv - 1; }
bool live(const gc_multi* m, int sv) { return is_glo(m) ? sv != GC_SV_NONE : sv != 0; }

// run fn(g) on one host thread per GPU in [0, n); returns the first non-zero code and remembers that GPU's message
template <class F>
int for_each_gpu(gc_multi* m, int n, F fn)
{
    std::vector<int> rc(n, GC_OK);
    for (int g = 1; g < n; ++g) m->workers[g]->submit([&fn, g] { return fn(g); });
    rc[0] = fn(0);
    for (int g = 1; g < n; ++g) rc[g] = m->workers[g]->wait();
    for (int g = 0; g < n; ++g)
        if (rc[g] != GC_OK) {
            m->err = "GPU " + std::to_string(g) + ": " + gc_last_error(m->h[g]);
            return rc[g];
        }
    return GC_OK;
}

}  // namespace

extern "C" {

const char* gc_multi_last_error(const gc_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
int gc_multi_n_gpus(const gc_multi* m) { return m ? (int)m->h.size() : 0; }
gc_handle* gc_multi_handle(gc_multi* m, int32_t gpu) { return (m && gpu >= 0 && gpu < (int)m->h.size()) ? m->h[gpu] : nullptr; }

int gc_multi_create(gc_multi** out, const gc_config* cfg, int32_t nGpus)
{
    if (!out || !cfg) return fail(nullptr, GC_ERR_ARG, "gc_multi_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, GC_ERR_CUDA, std::string("gc_multi_create: no CUDA device (") + cudaGetErrorString(ce) + ") - this engine has no CPU fallback");
    if (nGpus <= 0) nGpus = ndev;                                       // all visible GPUs
    if (cfg->device < 0 || cfg->device + nGpus > ndev) return fail(nullptr, GC_ERR_ARG, "gc_multi_create: devices cfg.device .. cfg.device + nGpus - 1 must exist");
    gc_multi* m = new gc_multi();
    m->cfg = *cfg;
    m->h.assign(nGpus, nullptr);
    for (int g = 0; g < nGpus; ++g) {
        m->workers.emplace_back(new Worker());
        if (g > 0) m->workers[g]->start();
    }
    // the handles are created in parallel, each on the thread that will serve its GPU: FFT plan, twiddles, replica spectra
    std::vector<int> rc(nGpus, GC_OK);
    std::vector<std::string> msg(nGpus);
    auto make = [&](int g) {
        gc_config c = *cfg;
        c.device = cfg->device + g;
        rc[g] = gc_create(&m->h[g], &c);
        if (rc[g] != GC_OK) msg[g] = gc_last_error(nullptr);
        return rc[g];
    };
    for (int g = 1; g < nGpus; ++g) m->workers[g]->submit([&make, g] { return make(g); });
    make(0);
    for (int g = 1; g < nGpus; ++g) m->workers[g]->wait();
    for (int g = 0; g < nGpus; ++g)
        if (rc[g] != GC_OK) {
            const int code = rc[g];
            g_multi_create_error = "GPU " + std::to_string(g) + ": " + msg[g];
            gc_multi_destroy(m);
            return code;
        }
    *out = m;
    return GC_OK;
}

void gc_multi_destroy(gc_multi* m)
{
    if (!m) return;
    for (size_t g = 1; g < m->workers.size(); ++g) m->workers[g]->stop();
    for (auto* hh : m->h) gc_destroy(hh);
    delete m;
}

int gc_multi_set_code(gc_multi* m, int32_t sv, int32_t component, const int8_t* chips, int32_t nChips)
{
    if (!m) return GC_ERR_ARG;
    for (size_t g = 0; g < m->h.size(); ++g) {
        const int rc = gc_set_code(m->h[g], sv, component, chips, nChips);
        if (rc != GC_OK) return fail(m, rc, gc_last_error(m->h[g]));
    }
    return GC_OK;
}

int gc_multi_set_param(gc_multi* m, int32_t key, double value)
{
    if (!m) return GC_ERR_ARG;
    for (size_t g = 0; g < m->h.size(); ++g) {
        const int rc = gc_set_param(m->h[g], key, value);
        if (rc != GC_OK) return fail(m, rc, gc_last_error(m->h[g]));
    }
    return GC_OK;
}

int gc_multi_set_cl_code_phase(gc_multi* m, int32_t nCh, const int32_t* clCodePhase)
{
    if (!m || nCh < 1 || !clCodePhase) return fail(m, GC_ERR_ARG, "gc_multi_set_cl_code_phase: bad argument");
    m->clPhase.assign(clCodePhase, clCodePhase + nCh);
    return GC_OK;
}

int gc_multi_get_cl_code_phase(const gc_multi* m, int32_t* clCodePhase)
{
    if (!m || !clCodePhase) return GC_ERR_ARG;
    // every GPU searched its own PRNs: the entries of the others are zero
    std::fill(clCodePhase, clCodePhase + 32, 0);
    for (auto* hh : m->h) {
        int32_t part[32];
        const int rc = gc_get_cl_code_phase(hh, part);
        if (rc != GC_OK) return rc;
        for (int i = 0; i < 32; ++i) if (part[i]) clCodePhase[i] = part[i];
    }
    return GC_OK;
}

int gc_multi_set_record_host(gc_multi* m, const void* bytes, size_t nbytes)
{
    if (!m || !bytes || nbytes == 0) return fail(m, GC_ERR_ARG, "gc_multi_set_record_host: bad argument");
    return for_each_gpu(m, (int)m->h.size(), [&](int g) { return gc_set_record_host(m->h[g], bytes, nbytes); });
}

// acquisition of one SV list dealt round-robin over the GPUs; `host` = longSignal from host memory (gc_acquire_host on every GPU),
// else the resident records (gc_acquire)
static int multi_acquire(gc_multi* m, const int8_t* iq, size_t nSamples, int32_t nSv, const int32_t* svList,
                         double* carrFreq, double* codePhase, double* peakMetric, int32_t* coarseBin, int32_t* coarseCodePhase)
{
    if (!m || nSv < 1 || !svList || !carrFreq || !codePhase || !peakMetric) return fail(m, GC_ERR_ARG, "gc_multi_acquire: bad argument");
    const int nG = (int)m->h.size(), n = gc_acq_result_len(m->cfg.signal);
    for (int i = 0; i < nSv; ++i) {
        const int ri = result_index(m, svList[i]);
        if (ri < 0 || ri >= n) return fail(m, GC_ERR_ARG, "gc_multi_acquire: SV id out of range");
    }
    std::vector<std::vector<int32_t>> part(nG);
    for (int i = 0; i < nSv; ++i) part[i % nG].push_back(svList[i]);
    const int used = std::min(nG, (int)nSv);
    std::vector<std::vector<double>> cf(used, std::vector<double>(n)), cp(used, std::vector<double>(n)), pm(used, std::vector<double>(n));
    std::vector<std::vector<int32_t>> cb(used, std::vector<int32_t>(n)), cc(used, std::vector<int32_t>(n));
    const int rc = for_each_gpu(m, used, [&](int g) {
        return iq ? gc_acquire_host(m->h[g], iq, nSamples, (int32_t)part[g].size(), part[g].data(), cf[g].data(), cp[g].data(), pm[g].data(), cb[g].data(), cc[g].data())
                  : gc_acquire(m->h[g], (int32_t)part[g].size(), part[g].data(), cf[g].data(), cp[g].data(), pm[g].data(), cb[g].data(), cc[g].data());
    });
    if (rc != GC_OK) return rc;
    for (int i = 0; i < n; ++i) {
        carrFreq[i] = codePhase[i] = peakMetric[i] = 0;                  // acquisition.m:130-134
        if (coarseBin) coarseBin[i] = 0;
        if (coarseCodePhase) coarseCodePhase[i] = 0;
    }
    float ms = 0;
    for (int g = 0; g < used; ++g) {
        for (int32_t sv : part[g]) {
            const int ri = result_index(m, sv);
            carrFreq[ri] = cf[g][ri]; codePhase[ri] = cp[g][ri]; peakMetric[ri] = pm[g][ri];
            if (coarseBin) coarseBin[ri] = cb[g][ri];
            if (coarseCodePhase) coarseCodePhase[ri] = cc[g][ri];
        }
        gc_stats st{};
        gc_get_stats(m->h[g], &st);
        ms = std::max(ms, st.acq_total_ms);
    }
    m->lastAcqMs = ms;
    return GC_OK;
}

int gc_multi_acquire(gc_multi* m, int32_t nSv, const int32_t* svList, double* carrFreq, double* codePhase, double* peakMetric,
                     int32_t* coarseBin, int32_t* coarseCodePhase)
{
    return multi_acquire(m, nullptr, 0, nSv, svList, carrFreq, codePhase, peakMetric, coarseBin, coarseCodePhase);
}

int gc_multi_acquire_host(gc_multi* m, const int8_t* iq, size_t nSamples, int32_t nSv, const int32_t* svList,
                          double* carrFreq, double* codePhase, double* peakMetric, int32_t* coarseBin, int32_t* coarseCodePhase)
{
    if (!iq) return fail(m, GC_ERR_ARG, "gc_multi_acquire_host: bad argument");
    return multi_acquire(m, iq, nSamples, nSv, svList, carrFreq, codePhase, peakMetric, coarseBin, coarseCodePhase);
}

int gc_multi_track(gc_multi* m, int32_t nCh, const int32_t* sv, const double* acqFreq, const double* codePhase,
                   const double* codeFreq0, int32_t nEpochs, double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone)
{
    if (!m || nCh < 1 || nEpochs < 1 || !sv || !acqFreq || !codePhase || !out || !epochsDone) return fail(m, GC_ERR_ARG, "gc_multi_track: bad argument");
    const int nG = (int)m->h.size();
    const int per = (nCh + nG - 1) / nG;                                  // contiguous blocks: channel order is kept
    const int used = (nCh + per - 1) / per;
    const int nRows = gc_track_nfields(m->h[0]);
    const int nV = nEpochs / m->cfg.cno_vsm_interval;
    const bool cl = m->cfg.signal == GC_SIG_GPS_L2C && m->cfg.pilot_trk_flag == 1;
    if (cl && (int)m->clPhase.size() != nCh) return fail(m, GC_ERR_ARG, "gc_multi_track: GPS L2C CL pilot needs gc_multi_set_cl_code_phase for these channels first");
    const int rc = for_each_gpu(m, used, [&](int g) {
        const int c0 = g * per, nc = std::min(per, nCh - c0);
        if (cl) {
            const int r = gc_set_cl_code_phase(m->h[g], nc, m->clPhase.data() + c0);
            if (r != GC_OK) return r;
        }
        return gc_track(m->h[g], nc, sv + c0, acqFreq + c0, codePhase + c0, codeFreq0 ? codeFreq0 + c0 : nullptr, nEpochs,
                        out + (size_t)c0 * nRows * nEpochs, vsmValue ? vsmValue + (size_t)c0 * nV : nullptr,
                        vsmIndex ? vsmIndex + (size_t)c0 * nV : nullptr, epochsDone + c0);
    });
    if (rc != GC_OK) return rc;
    float ms = 0;
    for (int g = 0; g < used; ++g) {
        gc_stats st{};
        gc_get_stats(m->h[g], &st);
        ms = std::max(ms, st.track_kernel_ms);
    }
    m->lastTrackMs = ms;
    // A short read makes the reference `return` from tracking() (tracking.m:241-245): every channel after the first one that ran
    // out of data stays as initialised - also the ones another GPU has processed in the meantime
    int failed = -1;
    for (int ch = 0; ch < nCh && failed < 0; ++ch)
        if (live(m, sv[ch]) && epochsDone[ch] < nEpochs) failed = ch;
    const double inf = std::numeric_limits<double>::infinity();
    for (int ch = failed + 1; failed >= 0 && ch < nCh; ++ch) {
        double* o = out + (size_t)ch * nRows * nEpochs;
        for (int f = 0; f < nRows; ++f) {
            const double fill = (f == GC_F_ABSOLUTE_SAMPLE || (f >= GC_F_I_P && f <= GC_F_Q_L) || f >= GC_TRACK_NFIELDS) ? 0.0 : inf;
            std::fill(o + (size_t)f * nEpochs, o + (size_t)(f + 1) * nEpochs, fill);
        }
        epochsDone[ch] = 0;
        if (vsmValue && vsmIndex && nV > 0) {
            std::fill(vsmValue + (size_t)ch * nV, vsmValue + (size_t)(ch + 1) * nV, 0.0);
            std::fill(vsmIndex + (size_t)ch * nV, vsmIndex + (size_t)(ch + 1) * nV, 0.0);
        }
    }
    return GC_OK;
}

int gc_multi_track_file(gc_multi* m, const char* path, int32_t nCh, const int32_t* sv, const double* acqFreq, const double* codePhase,
                        const double* codeFreq0, int32_t nEpochs, double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone)
{
    if (!m || !path) return fail(m, GC_ERR_ARG, "gc_multi_track_file: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(m, GC_ERR_IO, std::string("gc_multi_track_file: unable to read file ") + path);   // postProcessing.m:155-158
    fseek(f, 0, SEEK_END);
    const long long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz <= 0) { fclose(f); return fail(m, GC_ERR_IO, "gc_multi_track_file: empty file"); }
    void* pinned = nullptr;
    cudaSetDevice(m->cfg.device);
    if (cudaHostAlloc(&pinned, (size_t)sz, cudaHostAllocPortable) != cudaSuccess) { fclose(f); return fail(m, GC_ERR_CUDA, "gc_multi_track_file: cudaHostAlloc failed"); }
    const size_t got = fread(pinned, 1, (size_t)sz, f);
    fclose(f);
    int rc = (got == (size_t)sz) ? gc_multi_set_record_host(m, pinned, (size_t)sz) : fail(m, GC_ERR_IO, "gc_multi_track_file: short read");
    cudaFreeHost(pinned);
    if (rc != GC_OK) return rc;
    return gc_multi_track(m, nCh, sv, acqFreq, codePhase, codeFreq0, nEpochs, out, vsmValue, vsmIndex, epochsDone);
}

int gc_multi_get_times(const gc_multi* m, double* acqMs, double* trackMs)
{
    if (!m) return GC_ERR_ARG;
    if (acqMs) *acqMs = m->lastAcqMs;
    if (trackMs) *trackMs = m->lastTrackMs;
    return GC_OK;
}

}  // extern "C"
