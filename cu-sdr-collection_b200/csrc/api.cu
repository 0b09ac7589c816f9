// C ABI of libgnsscorr.so: handle management and host-side orchestration of the acquisition and
// tracking kernels.  Declared in include/gnsscorr.h, which cites the reference interface each
// entry point replaces.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gnsscorr.h"
#include "acq.h"
#include "codes.h"
#include "codegen_jobs.h"
#include "common.cuh"
#include "track.h"
#include "nav.h"

using namespace gc;

namespace {

thread_local std::string g_create_error;

// grid-z positions a row warp walks in one CTA (RowsParams::zLoop); GC_ROWS_ZLOOP overrides (experiments)
int rows_zloop(int L)
{
    static const int env = [] { const char* e = getenv("GC_ROWS_ZLOOP"); return e ? atoi(e) : 0; }();
    if (env > 0) return env;
    // measured (tools/acq_bench.py, tools/big_bench.py; profiles/r02c_rows_experiments.md): 16 rows per warp on the headline plan
    // (rows 1.231 -> 1.194 ms), 8 on the 25-point row plans (E1 20 Msps 5.16 -> 4.67 ms, B1C 47.0 -> 45.3 ms)
    return L == 32736 ? 16 : 8;
}

// fine search: the per-code sums go through moments around the centre bin (fine_sum_moments_kernel) where its third-order
// expansion over a 256-sample run is exact to ~1e-8 of a term: |2 pi (f_j - f_centre) ts| * 128 <= 0.025; GC_FINE_MOMENTS=0
// selects the bin-by-bin kernel
int fine_moments(int nFine, double fineStep, double ts)
{
    static const int env = [] { const char* e = getenv("GC_FINE_MOMENTS"); return e ? atoi(e) : 1; }();
    if (!env || nFine > 128) return 0;
    const double tmax = 6.283185307179586 * fineStep * (nFine - nFine / 2) * ts * 128.0;
    return tmax <= 0.025 ? 1 : 0;
}

double m_round(double x) { return x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5); }

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kMaxSv = 63;            // most SVs a list can hold / replica slots
constexpr int kEvents = 160;
// Cap of the inverse work buffer W (of the 180 GB of HBM).  The spectra X are re-read once per chunk, so fewer, larger chunks
// save traffic: the headline grid (4.9 GB for 32 PRN) runs as one chunk (2.36 -> 2.29 ms), GAL E5b 27.3 -> 24.4 ms.
constexpr double kWorkBytes = 6.0e9;

// GC_HOST_TIMING=1: host wall-clock laps of gc_create / gc_acquire on stderr (where does a one-shot call spend its time)
struct HostLaps {
    bool on;
    const char* tag;
    std::chrono::steady_clock::time_point t0, last;
    explicit HostLaps(const char* tag_) : on(getenv("GC_HOST_TIMING") != nullptr), tag(tag_)
    {
        if (on) t0 = last = std::chrono::steady_clock::now();
    }
    void lap(const char* what)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gc host timing] %s: %-28s %9.3f ms  (+%.3f)\n", tag, what,
                std::chrono::duration<double, std::milli>(now - t0).count(), std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    }
};

// what one enqueue of the variant A kernels recorded: which events bracket which stage, how many kernels
struct AcqEnq {
    int launches = 0, nRowLaunches = 0, e0 = -1, e1 = -1, fa = -1, fb = -1;
    std::vector<std::pair<int, int>> rowEv, colEv, fwdEv;
};
// the kernels and result copies of one variant A acquisition as a CUDA graph: captured the second time the same call (window,
// SV list, record, result buffer, device buffers) comes in, replayed from the third
struct AcqGraph {
    std::vector<char> key;
    int hits = 0;
    cudaGraphExec_t exec = nullptr;
    AcqEnq info;
};

}  // namespace

struct gc_handle {
    gc_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // column passes of the split correlation stage when they overlap the next chunk's row pass
    cudaEvent_t evRows[2]{}, evCols[2]{};    // hand-offs between the two streams (no timing)
    DevBuf<float2> W2;                       // second work buffer of the overlapped pipeline
    cudaEvent_t ev[kEvents]{};
    gc_stats stats{};

    // per-signal description
    bool glo = false;            // GLONASS: FDMA channels K, one shared code, I/Q swapped
    bool b3i = false;            // BeiDou B3I: 63 PRNs, 10230-chip codes, carrier-aided code NCO, NH/GEO fine search
    bool e1c = false;            // Galileo E1: 50 PRNs, caller-supplied 4092-chip memory codes, BOC(1,1) sub-chip tables,
                                 // E1B + E1C replicas summed in acquisition, 25-chip secondary-code fine search, pilot tracking
    int sub = 1;                 // table entries per chip (2 = BOC(1,1) sub-chips)
    int nRep = 1;                // replicas summed per SV in acquisition (2 = data + pilot)
    double fineStep = 25.0;      // fine-search bin width in Hz (acquisition.m:138; GAL_E1C acquisition.m:138: 10)
    bool fam5 = false;           // GPS L5C, GAL E5a, GAL E5b, BDS B2a: 10230-chip data + pilot codes supplied by the caller, two
                                 // replicas summed in acquisition, quadrature pilot tracking, carrier-aided code NCO
    bool varB = false;           // acquisition variant B (BDS B1I, GPS L2C): circularly shifted spectra, best row kept
    struct {                     // variant B geometry (BDS/B1I/include/acquisition.m:6-40, GPS/GPS_L2C/include/acquisition.m:5-24)
        int Lb = 0;              // samplesPerBlock = transform length
        int nSig = 1;            // consecutive signal blocks searched (B1I: 2, the better one is kept)
        int tabLen = 0;          // replica samples before the zero padding
        int nShifts = 1, nBins = 0;
        int sign = 1;            // initFreqShift = initFreq + sign*(binIter-1)*freqResolution/Nshifts
        int chipSamples = 0;     // samplesPerCodeChip: the +-1 chip excluded around the peak
        int N1 = 0;              // samplesPerBlock/Nblocks: the code-phase range of the second-peak search
        double freqRes = 0, initFreq = 0;
    } vb;
    DevBuf<VarbRow> vbRows;
    DevBuf<float> vbPeak;
    DevBuf<int> vbIdx;
    DevBuf<int4> vbSeg;
    DevBuf<int2> vbMap;          // fused variant B / C: spectrum row and circular shift per searched row
    DevBuf<float> vbMag;         // corrVec of the winning rows
    bool varC = false;           // acquisition variant C (BDS B1C): one spectrum, bins by circshift, weighted data + pilot, 2-D max
    struct { int Lc = 0, xLen = 0, nFine = 0; double initFreq = 0; } vc;   // len10PlusXms, samplesXmsLen (B1C acquisition.m:131-134)
    DevBuf<int> vcSlot;
    bool hostCodes = false;      // codes come from gc_set_code (e1c, fam5, varB, varC)
    int acqMinPeriods = 42, acqExtraPeriods = 2;   // longSignal = max(acqMinPeriods, nonCoh + acqExtraPeriods) code periods
    int fineCombine = 0;         // FineParams::combine
    int fineIdx0 = 0;            // first sample index of the fine-search code map (0: ts*(0:n-1), 1: ts*(1:n))
    bool fineTwoCodes = false;   // B2a: data and pilot codes both wiped off, |.| summed per period
    bool noFine = false;         // E5b: carrFreq = coarse bin frequency (GAL_E5b acquisition.m:203)
    int pilotMode = 0;           // tracking: 0 no pilot, 1 same-phase pilot (E1C), 2 quadrature pilot (L5C/E5a/E5b/B2a), 3 B1C narrow band,
                                 // 4 GPS L2C CL pilot, 5 B1C full band (TrackParams::pilot)
    double wbFactor = -1.0;      // GC_PARAM_B1C_WB_FACTOR (CalcWeighingFactor.m), < 0 = not set
    bool trackExact = false;     // GC_PARAM_TRACK_EXACT_SUMS (or GC_TRACK_EXACT_SUMS=1 in the environment at gc_create)
    bool trackFastDisc = false;  // GC_PARAM_TRACK_FAST_DISC
    std::vector<int32_t> clPhaseIn;          // channel.CLCodePhase for the next gc_track (GPS L2C CL pilot)
    int32_t clPhaseOut[32] = {0};            // acqResults.CLCodePhase of the last gc_acquire
    DevBuf<int8_t> clDev; DevBuf<int> clIdx; DevBuf<double> clPower;
    DevBuf<int8_t> trackP61;
    std::vector<int8_t> hostCode[3][63];   // caller-supplied codes [component: 0 data, 1 pilot, 2 pilot secondary][PRN-1]
                                           // (E1: stored as BOC(1,1) sub-chips)
    int nReplicas = 32;          // replica spectra held (32 GPS PRNs; 1 GLONASS; 63 B3I)
    int resultLen = 32;          // length of the acqResults vectors
    int nFinePeriods = 40;       // code periods of the fine-frequency search (40 GPS/GLONASS, 20 B3I)
    DevBuf<int16_t> chipIdx;     // sample -> chip index of the 40-period fine-search replica

    // derived (acquisition.m:116-124,138-140)
    int N = 0, L = 0, nBins = 0, nFine = 0, nonCoh = 0;
    double ts = 0;
    bool fused = false;
    int binShift = 0;            // variant A: binQ * acqSearchStep * L / fs when that is a whole number of FFT bins for a small binQ (else
    int binQ = 1;                // 0): the spectra of Doppler bin k are those of bin k mod binQ shifted by (k div binQ) * binShift, so only
                                 // the first binQ bins are transformed (binQ = 1: only bin 0)
    bool overlap = false;        // split correlation stage pipelined over two streams (GC_ACQ_OVERLAP)
    bool queue = false;          // correlation stage as one persistent kernel with an ordered work queue (GC_ACQ_PATH=queue)
    DevBuf<int> qctrl;
    bool cluster = false;        // correlation stage as one cluster kernel (acq_cluster.cu); else inv_rows + inv_cols
    FusedPlanInfo fp{};
    DevBuf<float2> twFused;      // [C][R] twiddles of fused plans with a Cooley-Tukey column/row link
    DevBuf<float2> twCols;       // [C1][C2] w_C^(-ta*beta) of the two-level inverse column pass (long transforms)

    // resident record
    const int8_t* rec = nullptr;
    size_t recBytes = 0;
    int fmt = 0;                 // Rec::fmt from settings.fileType / dataType
    DevBuf<int8_t> recOwned;

    // acquisition
    DevBuf<float2> twGen, X, T1, T2, T3, Cc, W;
    DevBuf<uint64_t> dphi, fdphi;
    DevBuf<int8_t> codeTab, chips, fineSecondary, slotSecondary;
    DevBuf<double> slotFreq0, metricDev;
    DevBuf<int> slotChipRow, slotSv, nAcqDev, acqSlot, fineChipRow, slotResult;
    DevBuf<int> prnList, slotGroup, partIdx, fineCodePhase, fineBest, fineSv;
    DevBuf<float> partMax;
    DevBuf<PeakOut> peaks;
    DevBuf<double> sigPower, fineSums, fineResult;
    DevBuf<short2> fineProd;
    GenericPlan plan{};
    bool replicasReady = false;
    int parts = 0;

    // tracking
    DevBuf<TrackChan> chans;
    DevBuf<int8_t> trackCodes, trackPilot;
    DevBuf<double> trackOut;
    DevBuf<int32_t> epochsDone;
    DevBuf<uint8_t> navCand, navBits;
    DevBuf<double> vsmDev, cnoPldDev;
    std::vector<double> cnoPld;  // [nCh][5][nV] of the last gc_track (BDS B2a / B1C)
    int cnoPldCh = 0, cnoPldV = 0;
    DevBuf<int> navInt;
    double tau1code = 0, tau2code = 0, tau1carr = 0, tau2carr = 0;
    std::unordered_map<const void*, std::vector<char>> upCache;   // upload_cached: last bytes sent to a buffer
    AcqGraph graph;              // variant A on a fused plan: the whole enqueue as one graph launch
    bool graphOff = false;       // GC_ACQ_GRAPH=0, or a capture failed
    AcqEnq pendingInfo;          // events of the last variant A enqueue whose timings have not been read yet (gc_acquire_device_async)
    bool statsPending = false;
    char* pin = nullptr;         // pinned host staging of the per-call result copies (peaks, sigPower, fine bins, acquired count)
    size_t pinBytes = 0;
};

namespace {

int fail(gc_handle* h, int code, const std::string& msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define GC_CUDA(h, expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(h, GC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

template <class T>
cudaError_t upload(DevBuf<T>& b, const std::vector<T>& v, cudaStream_t s)
{
    cudaError_t e = b.reserve(v.size());
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

// upload that is skipped when the buffer already holds exactly these bytes (the per-call tables of an acquisition - list slots,
// shift map, bin-1 frequencies ... - are the same from call to call for the same SV list; each skipped copy saves a staged
// pageable H2D of a few microseconds in front of the first kernel)
template <class T>
cudaError_t upload_cached(gc_handle* h, DevBuf<T>& b, const std::vector<T>& v, cudaStream_t s);

// w_L^(r*c) = exp(-2*pi*i*r*c/L) tables in float, computed in long double
std::vector<float2> tw_table_2d(int rows, int cols, int L)   // [r][c] -> w_L^(r*c)
{
    std::vector<float2> t((size_t)rows * cols);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const long long k = ((long long)r * c) % L;
            const long double a = -two_pi * (long double)k / (long double)L;
            t[(size_t)r * cols + c] = make_float2((float)cosl(a), (float)sinl(a));
        }
    return t;
}

void calcLoopCoef(double LBW, double zeta, double k, double* tau1, double* tau2)
{   // Common/calcLoopCoef.m:41-45
    const double Wn = LBW * 8 * zeta / (4 * zeta * zeta + 1);
    *tau1 = k / (Wn * Wn);
    *tau2 = 2.0 * zeta / Wn;
}

template <class T>
cudaError_t upload_cached(gc_handle* h, DevBuf<T>& b, const std::vector<T>& v, cudaStream_t s)
{
    std::vector<char>& last = h->upCache[(const void*)&b];
    const size_t nb = v.size() * sizeof(T);
    if (b.p && b.cap >= v.size() && last.size() == nb && (nb == 0 || memcmp(last.data(), v.data(), nb) == 0)) return cudaSuccess;
    cudaError_t e = upload(b, v, s);
    if (e == cudaSuccess) last.assign(reinterpret_cast<const char*>(v.data()), reinterpret_cast<const char*>(v.data()) + nb);
    else last.clear();
    return e;
}

// SV id -> (valid, result index, replica slot, carrier offset)
bool sv_ok(const gc_handle* h, int sv) { return h->glo ? (sv >= -7 && sv <= 13) : (sv >= 1 && sv <= h->resultLen); }
int sv_result_index(const gc_handle* h, int sv) { return h->glo ? sv + 7 : sv - 1; }       // MATLAB K+8 / PRN, 0-based
int sv_replica(const gc_handle* h, int sv) { return h->glo ? 0 : (sv - 1) * h->nRep; }
double sv_freq_offset(const gc_handle* h, int sv) { return h->glo ? -h->cfg.freq_spacing * (double)sv : 0.0; }
Rec rec_of(const gc_handle* h) { return Rec{h->rec, h->fmt}; }
// settings.skipNumberOfBytes as a sample offset: the reference seeks dataAdaptCoeff*skip BYTES (postProcessing.m:74,
// tracking.m:145-151), which is `skip` samples of 'schar' data and skip/2 samples of 'int16' data
long long skip_samples(const gc_handle* h) { return (long long)h->cfg.skip_number_of_bytes / ((h->fmt == 1 || h->fmt == 3) ? 2 : 1); }
long long rec_samples(const gc_handle* h) { return rec_of(h).samples_in((long long)h->recBytes); }

// +-1 table entries (chips, or BOC sub-chips) of one code period for an SV: the tracking replica of
// `component` (0 data, 1 pilot); the fine search uses the pilot component where there is one
void sv_chips(const gc_handle* h, int sv, int8_t* out, int component = 0)
{
    if (h->hostCodes) { const std::vector<int8_t>& c = h->hostCode[component][sv - 1]; std::copy(c.begin(), c.end(), out); }
    else if (h->glo) glo_code(out); else if (h->b3i) b3i_code(sv, out); else ca_code(sv, out);
}
bool sv_has_code(const gc_handle* h, int sv)
{
    if (!h->hostCodes) return true;
    if (h->cfg.signal == GC_SIG_GAL_E5A && h->hostCode[2][sv - 1].empty()) return false;   // per-PRN pilot secondary code
    if (h->varB) return !h->hostCode[0][sv - 1].empty();
    if (h->varC) return !h->hostCode[0][sv - 1].empty() && (h->nRep == 1 || !h->hostCode[1][sv - 1].empty());
    return !h->hostCode[0][sv - 1].empty() && !h->hostCode[1][sv - 1].empty();
}

// Replica spectra conj(fft([code zeros(1,N)]))/L (acquisition.m:158-164; GLO acquisition.m:145-149) and
// the sample -> chip map of the 40-period fine-search replica (acquisition.m:215-218; GLO :164).
int build_replicas(gc_handle* h)
{
    const int N = h->N, L = h->L, nRep = h->nReplicas, codeLen = h->cfg.code_length;
    HostLaps laps("build_replicas");
    std::vector<int8_t> tab((size_t)nRep * N);
    std::vector<int16_t> idx40((size_t)h->nFinePeriods * N);
    if (h->e1c) {                                            // makeE1BTable.m / makeE1CTable.m; GAL_E1C acquisition.m:160-172, 209-214
        for (int prn = 1; prn <= nRep / 2; ++prn) {
            if (!sv_has_code(h, prn)) continue;              // replica stays zero; gc_acquire refuses SVs without codes
            for (int r = 0; r < 2; ++r)
                make_boc_table(h->hostCode[r][prn - 1].data(), h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, N,
                               tab.data() + (size_t)((prn - 1) * 2 + r) * N);
        }
        boc_fine_index(h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, (long long)h->nFinePeriods * N, idx40.data());
    } else if (h->fam5) {                                    // makeL5ITable.m / makeL5QTable.m (and the E5a/E5b/B2a twins)
        for (int prn = 1; prn <= nRep / 2; ++prn) {
            if (!sv_has_code(h, prn)) continue;
            for (int r = 0; r < 2; ++r)
                make_code_table(h->hostCode[r][prn - 1].data(), h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, N,
                                tab.data() + (size_t)((prn - 1) * 2 + r) * N);
        }
        gps_fine_index(h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, (long long)h->nFinePeriods * N, idx40.data(), h->fineIdx0);
    } else if (h->b3i) {                                     // makeB3ITable.m:38-52; acquisition.m:170-173
        std::vector<int8_t> chips(codeLen);
        for (int prn = 1; prn <= nRep; ++prn) {
            b3i_code(prn, chips.data());
            make_code_table(chips.data(), h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, N, tab.data() + (size_t)(prn - 1) * N);
        }
        gps_fine_index(h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, (long long)h->nFinePeriods * N, idx40.data());
    } else if (h->glo) {
        int8_t chips[511];
        glo_code(chips);
        std::vector<int16_t> idx(N);
        glo_sample_index(h->cfg.code_freq_basis, h->cfg.sampling_freq, codeLen, N, idx.data());   // generateCAcode(0, fs, N)
        for (int n = 0; n < N; ++n) tab[n] = chips[idx[n]];
        glo_sample_index(h->cfg.code_freq_basis, h->cfg.sampling_freq, codeLen, 40LL * N, idx40.data());
    } else {
        for (int prn = 1; prn <= nRep; ++prn)
            make_ca_table(prn, h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, N, tab.data() + (size_t)(prn - 1) * N);
        gps_fine_index(h->cfg.sampling_freq, h->cfg.code_freq_basis, codeLen, 40LL * N, idx40.data());
    }
    laps.lap("sampled code tables (host)");
    GC_CUDA(h, upload(h->codeTab, tab, h->stream));
    GC_CUDA(h, upload(h->chipIdx, idx40, h->stream));
    {   // chips of every SV / component for the fine search: row = result index * 2 + component (GLONASS: row 0)
        const int tabLen = codeLen * h->sub;
        const int nSvRows = h->glo ? 1 : h->resultLen;
        std::vector<int8_t> all((size_t)nSvRows * 2 * tabLen, 0);
        for (int i = 0; i < nSvRows; ++i) {
            const int sv = h->glo ? 0 : i + 1;
            if (!h->glo && !sv_has_code(h, sv)) continue;
            for (int comp = 0; comp < (h->nRep == 2 ? 2 : 1); ++comp) sv_chips(h, sv, all.data() + ((size_t)i * 2 + comp) * tabLen, comp);
        }
        GC_CUDA(h, upload(h->chips, all, h->stream));
    }
    GC_CUDA(h, h->Cc.reserve((size_t)nRep * L));
    if (h->fused) {
        FwdColsParams fp{};
        fp.N = N; fp.codeTab = h->codeTab.p; fp.out = h->Cc.p; fp.tw = h->twFused.p;
        GC_CUDA(h, launch_fwd_cols(L, fp, nRep, true, h->stream));
        RowsParams rp{};
        rp.X = h->Cc.p; rp.nRows = (long long)nRep * h->fp.C;
        GC_CUDA(h, launch_fwd_rows(L, rp, h->stream));
        GC_CUDA(h, launch_finish_replica(h->Cc.p, (size_t)nRep * L, L, h->stream));
    } else {
        GC_CUDA(h, h->T1.reserve((size_t)std::max(nRep, h->nBins * h->nonCoh) * L));
        GC_CUDA(h, h->T2.reserve((size_t)std::max(nRep, h->nBins * h->nonCoh) * L));
        GC_CUDA(h, launch_generic_code(h->codeTab.p, N, nRep, h->T1.p, L, h->stream));
        float2 *src = h->T1.p, *dst = h->T2.p;
        int n = L, s = 1;
        for (int f = 0; f < h->plan.nf; ++f) {
            GC_CUDA(h, launch_generic_stage(h->plan, f, n, s, false, src, dst, nRep, h->stream));
            n /= h->plan.fac[f]; s *= h->plan.fac[f];
            std::swap(src, dst);
        }
        GC_CUDA(h, cudaMemcpyAsync(h->Cc.p, src, (size_t)nRep * L * sizeof(float2), cudaMemcpyDeviceToDevice, h->stream));
        GC_CUDA(h, launch_generic_conj_scale(h->Cc.p, (size_t)nRep * L, 1.0f / (float)L, h->stream));
    }
    h->replicasReady = true;
    return GC_OK;
}

}  // namespace

extern "C" {

int gc_abi_version(void) { return GC_ABI_VERSION; }
const char* gc_build_arch(void) { return "sm_100a"; }
int gc_acq_result_len(int32_t signal)
{
    return signal == GC_SIG_GPS_L1CA ? 32 : signal == GC_SIG_GLO_G1G2 ? 21 : signal == GC_SIG_BDS_B3I ? 63 :
           signal == GC_SIG_GAL_E1C ? 50 : signal == GC_SIG_GPS_L5C ? 32 : signal == GC_SIG_GAL_E5A ? 50 :
           signal == GC_SIG_GAL_E5B ? 50 : signal == GC_SIG_BDS_B2A ? 63 : signal == GC_SIG_BDS_B1I ? 58 :
           signal == GC_SIG_GPS_L2C ? 32 : signal == GC_SIG_BDS_B1C ? 63 : 0;
}

const char* gc_last_error(const gc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int gc_create(gc_handle** out, const gc_config* cfg)
{
    if (!out || !cfg) return fail(nullptr, GC_ERR_ARG, "gc_create: null argument");
    *out = nullptr;
    HostLaps laps("gc_create");
    if (cfg->abi_version != GC_ABI_VERSION) return fail(nullptr, GC_ERR_ARG, "gc_create: abi_version mismatch");
    if (cfg->signal < GC_SIG_GPS_L1CA || cfg->signal > GC_SIG_BDS_B1C)
        return fail(nullptr, GC_ERR_UNSUPPORTED, "gc_create: signal not implemented (GPS L1CA/L5C/L2C, GLONASS G1/G2, BDS B1I/B1C/B3I/B2a, GAL E1C/E5a/E5b are)");
    if ((cfg->file_type != 1 && cfg->file_type != 2 && cfg->file_type != GC_FILE_PACKED2) || (cfg->sample_bytes != 1 && cfg->sample_bytes != 2) ||
        (cfg->file_type == GC_FILE_PACKED2 && cfg->sample_bytes != 1))
        return fail(nullptr, GC_ERR_UNSUPPORTED, "gc_create: fileType must be 1 (real), 2 (I/Q) or GC_FILE_PACKED2 and dataType 'schar' (1 byte) or 'int16' (2 bytes)");
    if (!(cfg->sampling_freq > 0) || !(cfg->code_freq_basis > 0) ||
        cfg->code_length != (cfg->signal == GC_SIG_GLO_G1G2 ? 511 : cfg->signal == GC_SIG_GAL_E1C ? 4092 :
                             cfg->signal == GC_SIG_GPS_L1CA ? 1023 : cfg->signal == GC_SIG_BDS_B1I ? 2046 : 10230) ||
        (cfg->acq_noncoh_time < 1 && cfg->signal != GC_SIG_BDS_B1I && cfg->signal != GC_SIG_GPS_L2C && cfg->signal != GC_SIG_BDS_B1C) ||
        !(cfg->acq_search_step > 0) || cfg->cno_vsm_interval < 2)
        return fail(nullptr, GC_ERR_ARG, "gc_create: invalid settings");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, GC_ERR_CUDA, std::string("gc_create: no CUDA device (") + cudaGetErrorString(ce) + ") - this engine has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, GC_ERR_ARG, "gc_create: bad device ordinal");
    int ccMajor = 0, ccMinor = 0;                              // (cudaGetDeviceProperties costs 15-30 ms of a one-shot call)
    cudaDeviceGetAttribute(&ccMajor, cudaDevAttrComputeCapabilityMajor, cfg->device);
    cudaDeviceGetAttribute(&ccMinor, cudaDevAttrComputeCapabilityMinor, cfg->device);
    if (ccMajor != 10)
        return fail(nullptr, GC_ERR_CUDA, "gc_create: kernels are built for sm_100a only; device is sm_" + std::to_string(ccMajor) + std::to_string(ccMinor));

    laps.lap("device query");
    gc_handle* h = new gc_handle();
    h->cfg = *cfg;
    h->fmt = cfg->file_type == GC_FILE_PACKED2 ? 4 : (cfg->sample_bytes == 2 ? 1 : 0) | (cfg->file_type == 1 ? 2 : 0);
    if ((h->fmt == 1 || h->fmt == 3) && (cfg->skip_number_of_bytes & 1)) { delete h; return fail(nullptr, GC_ERR_ARG, "gc_create: with 'int16' data skipNumberOfBytes must be even (the seek would land inside a sample)"); }
    h->glo = (cfg->signal == GC_SIG_GLO_G1G2);
    h->b3i = (cfg->signal == GC_SIG_BDS_B3I);
    h->e1c = (cfg->signal == GC_SIG_GAL_E1C);
    h->fam5 = (cfg->signal == GC_SIG_GPS_L5C || cfg->signal == GC_SIG_GAL_E5A || cfg->signal == GC_SIG_GAL_E5B ||
               cfg->signal == GC_SIG_BDS_B2A);
    h->varB = (cfg->signal == GC_SIG_BDS_B1I || cfg->signal == GC_SIG_GPS_L2C);
    h->varC = (cfg->signal == GC_SIG_BDS_B1C);
    h->hostCodes = h->e1c || h->fam5 || h->varB || h->varC;
    { const char* e = getenv("GC_TRACK_EXACT_SUMS"); h->trackExact = e && atoi(e) != 0; }
    { const char* e = getenv("GC_ACQ_GRAPH"); h->graphOff = e && atoi(e) == 0; }
    h->sub = h->e1c ? 2 : 1;
    h->nRep = (h->e1c || h->fam5) ? 2 : 1;                   // data + pilot replicas (B1C: set below from pilotACQflag) (GAL_E1C acquisition.m:186-192, GPS_L5C :171-175)
    h->fineStep = h->e1c ? 10.0 : 25.0;                      // GAL_E1C acquisition.m:138
    h->nFinePeriods = h->b3i ? 20 : h->e1c ? 25 : 40;        // BDS/B3I/include/acquisition.m:131-133; GAL_E1C :148
    h->fineCombine = h->glo ? 1 : h->b3i ? 2 : h->e1c ? 3 : 0;
    h->pilotMode = (h->e1c && cfg->pilot_trk_flag == 1) ? 1 : 0;
    if (h->b3i) { h->acqMinPeriods = 22; h->acqExtraPeriods = 1; }               // BDS/B3I/include/postProcessing.m:86
    if (h->fam5) {
        h->fineIdx0 = 1;                                     // codeValueIndex = floor(ts*(1:n*samplesPerCode)/tc), GPS_L5C acquisition.m:196
        h->pilotMode = cfg->pilot_trk_flag == 1 ? 2 : 0;     // GPS_L5C tracking.m:277-281
        switch (cfg->signal) {
            case GC_SIG_GPS_L5C: h->nFinePeriods = 20; h->fineCombine = 4; break;                          // GPS_L5C acquisition.m:136-150
            case GC_SIG_GAL_E5A: h->nFinePeriods = 100; h->fineStep = 5.0; h->fineCombine = 4;             // GAL_E5a acquisition.m:136-142
                                 h->acqMinPeriods = 102; break;                                            // GAL_E5a postProcessing.m:88
            case GC_SIG_GAL_E5B: h->noFine = true; h->nFinePeriods = 1; h->acqMinPeriods = 102; break;     // GAL_E5b acquisition.m:203; postProcessing.m:89
            default:             h->nFinePeriods = std::max(10, (int)cfg->acq_noncoh_time); h->fineCombine = 5;   // BDS/B2a acquisition.m:140
                                 h->fineTwoCodes = true; h->acqMinPeriods = 12; break;                     // B2a postProcessing.m:86
        }
    }
    h->nReplicas = h->glo ? 1 : h->b3i ? 63 : h->hostCodes ? 2 * gc_acq_result_len(cfg->signal) : 32;
    h->resultLen = gc_acq_result_len(cfg->signal);
    auto bail = [&](int rc) { g_create_error = h->err; gc_destroy(h); return rc; };
    if (cudaSetDevice(cfg->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return bail(GC_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "cudaStreamCreate failed"; return bail(GC_ERR_CUDA); }
    for (auto& e : h->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { h->err = "cudaEventCreate failed"; return bail(GC_ERR_CUDA); }
    {   // the second stream carries the column passes of the overlapped pipeline (GC_ACQ_OVERLAP): highest priority, so that its few
        // persistent CTAs are placed as soon as a slot frees up next to the row pass of the following chunk
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi) != cudaSuccess) { h->err = "cudaStreamCreate failed"; return bail(GC_ERR_CUDA); }
    }
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&h->evRows[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->evCols[i], cudaEventDisableTiming) != cudaSuccess) { h->err = "cudaEventCreate failed"; return bail(GC_ERR_CUDA); }

    laps.lap("streams, events");
    // acquisition.m:116-124,138-140
    h->N = (int)m_round(cfg->sampling_freq / (cfg->code_freq_basis / (double)cfg->code_length));
    h->L = 2 * h->N;
    if (h->varC) {
        if (cfg->acq_coh_t < 1 || cfg->acq_coh_t > 10) { h->err = "gc_create: acqCohT must be 1..10 ms"; return bail(GC_ERR_ARG); }
        h->vc.xLen = (int)m_round((double)h->N / 10 * cfg->acq_coh_t);                      // samplesXmsLen, B1C acquisition.m:131
        h->vc.Lc = (int)m_round((double)h->N / 10 * (10 + cfg->acq_coh_t));                 // len10PlusXms, :133
        h->vc.nFine = (int)m_round(cfg->acq_search_step / 25) * 2 + 1;                      // :155
        h->vc.initFreq = cfg->IF + cfg->acq_search_band;                                    // :166
        h->L = h->vc.Lc;
        h->nRep = cfg->pilot_acq_flag == 1 ? 2 : 1;
        h->sub = 2;                                                                         // BOC(1,1) sub-chip tables (NB_tracking.m:225-246)
        h->pilotMode = cfg->pilot_trk_flag == 1 ? 3 : cfg->pilot_trk_flag == 2 ? 5 : 0;          // NB_tracking.m / WB_tracking.m (B1C postProcessing.m:34-38)
    }
    if (h->varB) {
        const bool b1i = cfg->signal == GC_SIG_BDS_B1I;
        const int nBlocks = b1i ? 4 : 2;                                                   // B1I :7 ; L2C :5
        h->vb.Lb = b1i ? (int)m_round(cfg->sampling_freq / (cfg->code_freq_basis / (nBlocks * (double)cfg->code_length)))   // B1I :9-10
                       : h->N * nBlocks;                                                    // L2C :10
        h->vb.nSig = b1i ? 2 : 1;                                                           // B1I :13-14
        h->vb.tabLen = b1i ? (int)m_round(cfg->sampling_freq / (cfg->code_freq_basis / (2 * (double)cfg->code_length))) : h->N;   // makeCaTableDMA.m:9-10
        h->vb.freqRes = cfg->sampling_freq / h->vb.Lb;                                      // B1I :20
        h->vb.nBins = (int)m_round(cfg->acq_search_band * 1e3 / h->vb.freqRes) + 1;         // :22 (acqSearchBand in kHz)
        const double ns = h->vb.freqRes / cfg->acq_search_step;                             // :40 Nshifts = freqResolution/stepSize
        h->vb.nShifts = (int)m_round(ns);
        if (h->vb.nShifts < 1 || std::fabs(ns - h->vb.nShifts) > 1e-9) { h->err = "gc_create: freqResolution/stepSize must be a whole number"; return bail(GC_ERR_ARG); }
        h->vb.sign = b1i ? 1 : -1;                                                          // B1I :64 ; L2C :62
        h->vb.chipSamples = (int)m_round(cfg->sampling_freq / cfg->code_freq_basis);        // B1I :126 ; L2C :7
        h->vb.N1 = h->vb.Lb / nBlocks;
        h->vb.initFreq = cfg->IF + (cfg->acq_search_band / 2) * 1000;                       // B1I :50
        h->L = h->vb.Lb;
        if (!b1i && cfg->pilot_trk_flag == 1) h->pilotMode = 4;                            // GPS_L2C tracking.m:160-167
    }
    h->ts = 1 / cfg->sampling_freq;
    h->nBins = h->varB ? h->vb.nBins : (int)m_round(cfg->acq_search_band * 2 / cfg->acq_search_step) + 1;
    h->nFine = (int)m_round(cfg->acq_search_step / h->fineStep) + 1;
    h->nonCoh = (h->varB || h->varC) ? 1 : cfg->acq_noncoh_time;
    h->fused = fused_plan_info(h->L, &h->fp) && !getenv("GC_FORCE_GENERIC") && !((h->varB || h->varC) && h->fp.pfa);   // (spectrum shifts: Cooley-Tukey plans)
    h->stats.fft_len = h->L;
    {   // GC_ACQ_PATH=cluster selects the one-kernel correlation stage (acq_cluster.cu, transform resident
        // in a cluster's shared memory); the default is inverse rows + inverse columns through a work buffer
        const char* e = getenv("GC_ACQ_PATH");
        h->cluster = h->fused && e && strcmp(e, "cluster") == 0;
        const char* o = getenv("GC_ACQ_OVERLAP");
        h->overlap = h->fused && !h->cluster && o && atoi(o) != 0;
        h->queue = h->fused && e && strcmp(e, "queue") == 0 && h->fp.C <= 50 && !h->varB && !h->varC;
    }
    if (h->fused && !h->cluster && !h->queue && !h->varB && !h->varC && !getenv("GC_ACQ_NO_SHIFT")) {
        // acqSearchStep in FFT bins.  A whole number (500 Hz at 16.368, 18 and 12 Msps: 1): every bin is bin 0 shifted.  A fraction
        // with a small denominator q (GAL E5b: 60 Hz = 0.12 bins, q = 25; GAL E1: 150 Hz = 1.2 bins, q = 5): bin k is bin k mod q
        // shifted by (k div q) * (q * step) bins, so q spectra per block serve the whole grid (GC_ACQ_NO_SHIFT_Q=1: whole steps only)
        const double sh = cfg->acq_search_step * (double)h->L / cfg->sampling_freq;
        const int qMax = getenv("GC_ACQ_NO_SHIFT_Q") ? 1 : std::min(h->nBins, 32);
        for (int q = 1; q <= qMax; ++q) {
            const double shq = sh * q, shr = m_round(shq);
            if (shr >= 1 && std::fabs(shq - shr) < 1e-9 * q && shr * ((h->nBins - 1) / q + 1) < h->L) { h->binShift = (int)shr; h->binQ = q; break; }
        }
    }
    h->stats.acq_path = h->cluster ? 2 : h->queue ? 3 : h->fused ? 1 : 0;   // 2 = fused plan + cluster correlation kernel, 1 = fused plan, split
                                                             // correlation stage, 0 = generic mixed-radix passes

    auto setup = [&]() -> int {
        if (h->fused) {
            h->parts = h->cluster ? kCorrClusterParts : h->queue ? (h->fp.R + 159) / 160 : h->fp.parts;
            if (!h->fp.pfa) {   // w_L^(j1 * m(p)), row position p = a*RB + b <-> m = (RB*a + RA*b) mod R
                const FusedPlanInfo& f = h->fp;
                std::vector<float2> tw((size_t)f.C * f.R);
                const long double two_pi = 6.283185307179586476925286766559005768L;
                for (int j1 = 0; j1 < f.C; ++j1)
                    for (int p = 0; p < f.R; ++p) {
                        const int m = (f.RB * (p / f.RB) + f.RA * (p % f.RB)) % f.R;
                        const long long e = ((long long)j1 * m) % f.L;
                        const long double a = -two_pi * (long double)e / (long double)f.L;
                        tw[(size_t)j1 * f.R + p] = make_float2((float)cosl(a), (float)sinl(a));
                    }
                GC_CUDA(h, upload(h->twFused, tw, h->stream));
            }
            if (h->fp.C1 > 0) {   // inverse two-level column pass: w_C^(+ta*beta), ta < C1, beta < C2 (inv_cols_big_kernel)
                const FusedPlanInfo& f = h->fp;
                std::vector<float2> tw((size_t)f.C);
                const long double two_pi = 6.283185307179586476925286766559005768L;
                for (int ta = 0; ta < f.C1; ++ta)
                    for (int be = 0; be < f.C2; ++be) {
                        const long double a = two_pi * (long double)((ta * be) % f.C) / (long double)f.C;
                        tw[(size_t)ta * f.C2 + be] = make_float2((float)cosl(a), (float)sinl(a));
                    }
                GC_CUDA(h, upload(h->twCols, tw, h->stream));
            }
        } else {
            // fewest passes over the radices that have a kernel: register codelets (8 .. 50) and small primes / 4
            h->plan.L = h->L; h->plan.nf = 0;
            {
                static const int kRadix[] = {50, 45, 40, 33, 32, 30, 25, 16, 8, 31, 13, 11, 7, 5, 4, 3, 2};
                std::vector<int> best, cur;
                std::function<void(int)> rec = [&](int m) {
                    if (m == 1) { if (best.empty() || cur.size() < best.size()) best = cur; return; }
                    if (!best.empty() && cur.size() + 1 >= best.size()) return;
                    for (int r : kRadix)
                        if (m % r == 0) { cur.push_back(r); rec(m / r); cur.pop_back(); }
                };
                if (!getenv("GC_GENERIC_SMALL_RADIX")) rec(h->L);
                if (best.empty()) {                              // other primes up to 64 (run-time radix kernel) / the original
                    int m = h->L;                                // pass structure: radix 4, then primes
                    while (m % 4 == 0) { best.push_back(4); m /= 4; }
                    for (int f = 2; m > 1 && f <= 64; ++f) while (m % f == 0) { best.push_back(f); m /= f; }
                    if (m > 1) best.clear();
                }
                if (best.empty() || best.size() > 32) return fail(h, GC_ERR_UNSUPPORTED, "FFT length 2*samplesPerCode has a prime factor > 64");
                for (int r : best) h->plan.fac[h->plan.nf++] = r;
            }
            GC_CUDA(h, upload(h->twGen, tw_table_2d(2, h->L, h->L), h->stream));
            h->plan.tw = h->twGen.p + h->L;   // row 1 of the [2][L] table = w_L^t
            h->parts = 8;
        }
        laps.lap("plan, twiddles");
        calcLoopCoef(cfg->dll_noise_bandwidth, cfg->dll_damping_ratio, 1.0, &h->tau1code, &h->tau2code);    // tracking.m:100
        calcLoopCoef(cfg->pll_noise_bandwidth, cfg->pll_damping_ratio, 0.25, &h->tau1carr, &h->tau2carr);   // tracking.m:109
        if (!h->hostCodes) {                                 // caller-supplied codes: replicas are built by the first gc_acquire
            int rc = build_replicas(h);
            if (rc != GC_OK) return rc;
        }
        laps.lap("replicas enqueued");
        GC_CUDA(h, cudaStreamSynchronize(h->stream));
        laps.lap("replicas done");
        return GC_OK;
    };
    int rc = setup();
    if (rc != GC_OK) return bail(rc);
    *out = h;
    return GC_OK;
}

void gc_destroy(gc_handle* h)
{
    if (!h) return;
    HostLaps laps("gc_destroy");
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->recOwned.release();
    h->twGen.release(); h->X.release(); h->T1.release(); h->T2.release(); h->T3.release();
    h->chipIdx.release(); h->twFused.release(); h->twCols.release();
    h->Cc.release(); h->W.release(); h->dphi.release(); h->fdphi.release(); h->codeTab.release(); h->chips.release(); h->fineSecondary.release(); h->slotSecondary.release();
    h->slotFreq0.release(); h->metricDev.release(); h->slotChipRow.release(); h->slotSv.release(); h->nAcqDev.release(); h->acqSlot.release(); h->fineChipRow.release();
    h->prnList.release(); h->slotGroup.release(); h->partIdx.release(); h->fineCodePhase.release(); h->fineBest.release(); h->fineSv.release(); h->partMax.release();
    h->peaks.release(); h->sigPower.release(); h->fineSums.release(); h->fineResult.release(); h->fineProd.release();
    h->slotResult.release(); h->vcSlot.release(); h->vbRows.release(); h->vbPeak.release(); h->vbIdx.release(); h->vbSeg.release();
    h->chans.release(); h->trackCodes.release(); h->trackPilot.release(); h->trackOut.release(); h->epochsDone.release();
    if (h->graph.exec) cudaGraphExecDestroy(h->graph.exec);
    if (h->pin) cudaFreeHost(h->pin);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) { if (h->evRows[i]) cudaEventDestroy(h->evRows[i]); if (h->evCols[i]) cudaEventDestroy(h->evCols[i]); }
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    laps.lap("buffers, events, streams freed");
    delete h;
}

int gc_set_code(gc_handle* h, int32_t sv, int32_t component, const int8_t* chips, int32_t nChips)
{
    if (!h) return GC_ERR_ARG;
    if (!h->hostCodes) return fail(h, GC_ERR_ARG, "gc_set_code: this signal generates its own codes");
    const bool secondary = (component == 2);
    if (secondary && h->cfg.signal != GC_SIG_GAL_E5A && h->cfg.signal != GC_SIG_BDS_B1C) return fail(h, GC_ERR_ARG, "gc_set_code: only GAL E5a takes a pilot secondary code");
    const bool l2c = h->cfg.signal == GC_SIG_GPS_L2C || h->cfg.signal == GC_SIG_BDS_B1C;   // 2*codeLength entries: the return-to-zero CM code
                                                                                            // (generateCMcode.m) / B1C BOC(1,1) sub-chips (generateDataBOC11.m)
    const bool isL2C = h->cfg.signal == GC_SIG_GPS_L2C, isB1C = h->cfg.signal == GC_SIG_BDS_B1C;
    if (isL2C && component == 1) {                            // the return-to-zero CL sequence, 75 CM periods (generateCLcode.m)
        if (sv < 1 || sv > h->resultLen || !chips || nChips != 150 * h->cfg.code_length)
            return fail(h, GC_ERR_ARG, "gc_set_code: the GPS L2C CL component takes 2*75*code_length entries");
        for (int i = 0; i < nChips; ++i)
            if (chips[i] < -1 || chips[i] > 1) return fail(h, GC_ERR_ARG, "gc_set_code: CL entries must be +-1 or 0");
        h->hostCode[1][sv - 1].assign(chips, chips + nChips);
        return GC_OK;
    }
    if (isB1C && component == 2) {                            // pilot BOC(6,1) sequence, 12 entries per chip (generatePilotBOC61.m)
        if (sv < 1 || sv > h->resultLen || !chips || nChips != 12 * h->cfg.code_length)
            return fail(h, GC_ERR_ARG, "gc_set_code: the BDS B1C BOC(6,1) component takes 12*code_length entries");
        for (int i = 0; i < nChips; ++i)
            if (chips[i] != 1 && chips[i] != -1) return fail(h, GC_ERR_ARG, "gc_set_code: chips must be +-1");
        h->hostCode[2][sv - 1].assign(chips, chips + nChips);
        return GC_OK;
    }
    if (sv < 1 || sv > h->resultLen || component < 0 || component > 2 || !chips || (h->varB && component != 0) ||
        nChips != (secondary ? 100 : l2c ? 2 * h->cfg.code_length : h->cfg.code_length))
        return fail(h, GC_ERR_ARG, "gc_set_code: bad argument (PRN in range, component 0/1 with code_length chips, or 2 with 100)");
    for (int i = 0; i < nChips; ++i)
        if (chips[i] != 1 && chips[i] != -1 && !(h->cfg.signal == GC_SIG_GPS_L2C && chips[i] == 0)) return fail(h, GC_ERR_ARG, "gc_set_code: chips must be +-1");
    std::vector<int8_t>& c = h->hostCode[component][sv - 1];
    if (h->e1c) {
        c.resize((size_t)2 * nChips);
        boc11(chips, nChips, c.data());                       // generateE1Bcode.m:58-64
    } else {
        c.assign(chips, chips + nChips);
    }
    if (!secondary) h->replicasReady = false;
    return GC_OK;
}

// sync = false: the copy stays in flight on the handle's stream (gc_acquire_host: the search that follows is stream ordered
// behind it and synchronises before it returns, so the caller's buffer is free again when the call is)
static int set_record_host(gc_handle* h, const void* bytes, size_t nbytes, bool sync)
{
    if (!h || !bytes || nbytes == 0) return fail(h, GC_ERR_ARG, "gc_set_record_host: bad argument");
    cudaSetDevice(h->cfg.device);
    const size_t cap = ((nbytes + 15) & ~(size_t)15) + 256;
    const bool grown = cap > h->recOwned.cap;
    GC_CUDA(h, h->recOwned.reserve(cap));
    GC_CUDA(h, cudaMemcpyAsync(h->recOwned.p, bytes, nbytes, cudaMemcpyHostToDevice, h->stream));
    if (grown || nbytes != h->recBytes || h->rec != h->recOwned.p)     // (the zero tail of an unchanged layout is still there)
        GC_CUDA(h, cudaMemsetAsync(h->recOwned.p + nbytes, 0, cap - nbytes, h->stream));
    if (sync) GC_CUDA(h, cudaStreamSynchronize(h->stream));
    h->rec = h->recOwned.p;
    h->recBytes = nbytes;
    return GC_OK;
}

int gc_set_record_host(gc_handle* h, const void* bytes, size_t nbytes) { return set_record_host(h, bytes, nbytes, true); }

int gc_set_record_device(gc_handle* h, const void* dptr, size_t nbytes)
{
    if (!h || !dptr || nbytes == 0) return fail(h, GC_ERR_ARG, "gc_set_record_device: bad argument");
    if (((uintptr_t)dptr & 15) != 0) return fail(h, GC_ERR_ARG, "gc_set_record_device: pointer must be 16-byte aligned");
    h->rec = static_cast<const int8_t*>(dptr);
    h->recBytes = nbytes;
    return GC_OK;
}

// Codes the caller has not supplied through gc_set_code are generated here, ON THE DEVICE (codegen.cu: one thread per SV and
// component running the reference's generators as bit-packed registers), and then take the same path as caller-supplied ones.
static int autofill_codes(gc_handle* h, int nSv, const int32_t* svList)
{
    if (!h->hostCodes) return GC_OK;
    const gc_config& c = h->cfg;
    std::vector<CodeJob> jobs;
    std::vector<std::pair<int, int>> what;                     // (sv, component) of each job
    for (int i = 0; i < nSv; ++i) {
        const int sv = svList[i];
        if (sv < 1 || sv > h->resultLen) continue;
        for (int comp = 0; comp < 3; ++comp) {
            bool need = comp <= 1;
            if (c.signal == GC_SIG_BDS_B1I) need = comp == 0;
            if (c.signal == GC_SIG_GPS_L2C) need = comp == 0 || (comp == 1 && c.pilot_trk_flag == 1);
            if (c.signal == GC_SIG_BDS_B1C) need = comp <= 1 || (comp == 2 && c.pilot_trk_flag == 2);
            if (c.signal == GC_SIG_GAL_E5A && comp == 2) need = true;
            if (!need || !h->hostCode[comp][sv - 1].empty()) continue;
            bool dup = false;
            for (auto& w : what) dup |= (w.first == sv && w.second == comp);
            CodeJob j;
            if (dup || !make_code_job(c.signal, sv, comp, &j)) continue;
            jobs.push_back(j);
            what.push_back({sv, comp});
        }
    }
    if (jobs.empty()) return GC_OK;
    cudaSetDevice(c.device);
    std::vector<int8_t> out;
    GC_CUDA(h, run_code_jobs_device(jobs, out, h->stream));
    for (size_t i = 0; i < jobs.size(); ++i) {
        const int rc = gc_set_code(h, what[i].first, what[i].second, out.data() + jobs[i].outOff, jobs[i].n);
        if (rc != GC_OK) return rc;
    }
    return GC_OK;
}

// ---- acquisition variant B (BDS/B1I/include/acquisition.m:42-176, GPS/GPS_L2C/include/acquisition.m:26-118) -------
static int varb_build_replicas(gc_handle* h)
{
    const gc_config& c = h->cfg;
    const int Lb = h->vb.Lb, S = h->vb.tabLen, nRep = h->resultLen;
    const bool b1i = c.signal == GC_SIG_BDS_B1I;
    std::vector<int8_t> tab((size_t)nRep * S, 0);
    const double ts = 1 / c.sampling_freq;
    for (int prn = 1; prn <= nRep; ++prn) {
        const std::vector<int8_t>& code = h->hostCode[0][prn - 1];
        if (code.empty()) continue;
        int8_t* t = tab.data() + (size_t)(prn - 1) * S;
        if (b1i) {                                            // makeCaTableDMA.m:13-19: [caCode caCode] sampled over two periods
            const double tc = 1 / c.code_freq_basis;
            for (int n = 1; n <= S; ++n) {
                int idx = (int)std::ceil((ts * (double)n) / tc);
                if (n == S) idx = 2 * c.code_length;
                t[n - 1] = code[(idx - 1) % c.code_length];
            }
        } else {                                              // makeCMTable.m:8-15 (return-to-zero CM code, 2*codeLength entries)
            const double tc = 1 / (c.code_freq_basis * 2);
            for (int n = 0; n < S; ++n) {
                int idx = (int)std::ceil((ts * (double)n) / tc);
                if (n == S - 1) idx = c.code_length * 2;
                if (n == 0) idx = 1;
                t[n] = code[idx - 1];
            }
        }
    }
    cudaStream_t st = h->stream;
    GC_CUDA(h, upload(h->codeTab, tab, st));
    GC_CUDA(h, h->Cc.reserve((size_t)nRep * Lb));
    if (h->fused) {                                          // [table zeros] (B1I :58, L2C :47) -> conj(fft(.))/L in the plan's layout
        FwdColsParams fp{};
        fp.N = S; fp.codeTab = h->codeTab.p; fp.out = h->Cc.p; fp.tw = h->twFused.p;
        GC_CUDA(h, launch_fwd_cols(Lb, fp, nRep, true, st));
        RowsParams rp{};
        rp.X = h->Cc.p; rp.nRows = (long long)nRep * h->fp.C;
        GC_CUDA(h, launch_fwd_rows(Lb, rp, st));
        GC_CUDA(h, launch_finish_replica(h->Cc.p, (size_t)nRep * Lb, Lb, st));
        h->replicasReady = true;
        return GC_OK;
    }
    GC_CUDA(h, h->T1.reserve((size_t)nRep * Lb));
    GC_CUDA(h, h->T2.reserve((size_t)nRep * Lb));
    GC_CUDA(h, launch_varb_pad(h->codeTab.p, S, nRep, h->T1.p, Lb, st));           // [table zeros] (B1I :58, L2C :47)
    float2 *src = h->T1.p, *dst = h->T2.p;
    int n = Lb, sd = 1;
    for (int f = 0; f < h->plan.nf; ++f) {
        GC_CUDA(h, launch_generic_stage(h->plan, f, n, sd, false, src, dst, nRep, st));
        n /= h->plan.fac[f]; sd *= h->plan.fac[f];
        std::swap(src, dst);
    }
    GC_CUDA(h, cudaMemcpyAsync(h->Cc.p, src, (size_t)nRep * Lb * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    GC_CUDA(h, launch_generic_conj_scale(h->Cc.p, (size_t)nRep * Lb, 1.0f / (float)Lb, st));   // conj(fft(.)) with ifft's 1/L
    h->replicasReady = true;
    return GC_OK;
}

static int acquire_varb(gc_handle* h, long long winStart, int32_t nSv, const int32_t* svList,
                        double* carrFreq, double* codePhase, double* peakMetric, int32_t* coarseBin, int32_t* coarseCodePhase)
{
    const gc_config& c = h->cfg;
    const int Lb = h->vb.Lb, nSig = h->vb.nSig, nShifts = h->vb.nShifts, nBins = h->vb.nBins;
    const bool b1i = c.signal == GC_SIG_BDS_B1I;
    cudaStream_t st = h->stream;
    const long long recSamples = rec_samples(h);
    if (winStart < 0 || winStart + (long long)nSig * Lb > recSamples)
        return fail(h, GC_ERR_SHORT_RECORD, "gc_acquire: record shorter than the acquisition blocks");
    for (int i = 0; i < h->resultLen; ++i) {
        carrFreq[i] = codePhase[i] = peakMetric[i] = 0;
        if (coarseBin) coarseBin[i] = 0;
        if (coarseCodePhase) coarseCodePhase[i] = 0;
    }
    cudaEventRecord(h->ev[0], st);
    int launches = 0;
    // wiped-off spectra of every sub-bin shift and block (B1I :62-76, L2C :60-68)
    std::vector<uint64_t> dphi(nShifts);
    for (int b = 0; b < nShifts; ++b)
        dphi[b] = turns_to_fix((h->vb.initFreq + h->vb.sign * b * (h->vb.freqRes / nShifts)) / c.sampling_freq);
    GC_CUDA(h, h->dphi.reserve(nShifts));
    GC_CUDA(h, cudaMemcpyAsync(h->dphi.p, dphi.data(), nShifts * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const int nX = nShifts * nSig;
    const int nRows = nShifts * nBins * nSig;
    std::vector<float> peaks((size_t)nSv * nRows);
    std::vector<VarbRow> win(nSv);
    std::vector<int> winShift(nSv), winBin(nSv);
    std::vector<float> maxPeak(nSv), second(nSv);
    std::vector<int> cp0(nSv);
    // the reference's running comparison, in its loop order (B1I :60-123, L2C :58-86)
    auto pick_winners = [&]() {
        for (int s = 0; s < nSv; ++s) {
            float prevmax = 0.f;
            int fbin = 0, fshift = 0, fsig = 0;
            for (int b = 0; b < nShifts; ++b)
                for (int k = 0; k < nBins; ++k) {
                    if (k == nBins - 1 && b > 0) continue;                                     // B1I :79-81
                    const float* pk = peaks.data() + (size_t)s * nRows + (size_t)(b * nBins + k) * nSig;
                    if (nSig == 2) {
                        if (pk[0] > prevmax || pk[1] > prevmax) {                              // B1I :101-114
                            if (pk[0] > pk[1]) { prevmax = pk[0]; fsig = 0; } else { prevmax = pk[1]; fsig = 1; }
                            fshift = b; fbin = k;
                        }
                    } else if (pk[0] > prevmax) { prevmax = pk[0]; fshift = b; fbin = k; fsig = 0; }   // L2C :78-83
                }
            win[s] = VarbRow{fshift * nSig + fsig, svList[s] - 1, fbin, 0};
            winShift[s] = fshift; winBin[s] = fbin;
        }
    };
    // code-phase ranges of the second-peak search outside +-1 chip of the peak (:127-139), 0-based inclusive
    auto second_peak_ranges = [&](std::vector<int4>& seg) {
        for (int s = 0; s < nSv; ++s) {
            const int cp = cp0[s] + 1;                                                         // 1-based codePhase (:125)
            const int e1 = cp - h->vb.chipSamples, e2 = cp + h->vb.chipSamples, N1 = h->vb.N1; // :127-128
            int a1, b1, a2 = 1, b2 = 0;                                                        // 1-based inclusive ranges; second one empty by default
            if (e1 < 2) { a1 = e2; b1 = N1 + e1; }                                             // :131-133
            else if (e2 >= N1) { a1 = e2 - N1 + 1; b1 = e1; }                                  // :134-136
            else { a1 = 1; b1 = e1; a2 = e2; b2 = N1; }                                        // :137-139
            seg[s] = make_int4(std::max(a1, 1) - 1, std::min(b1, Lb) - 1, std::max(a2, 1) - 1, std::min(b2, Lb) - 1);
        }
    };
    if (h->fused) {
        // the plan's kernels: forward spectra of every (shift, block); per SV every (shift, bin, block) row as X(j - bin) .* C,
        // inverse rows -> work buffer -> inverse columns + |.| + tile maxima; the winning rows once more with corrVec written out
        const int parts = h->fp.parts, C = h->fp.C;
        GC_CUDA(h, h->X.reserve((size_t)nX * Lb));
        FwdColsParams fp{};
        fp.rec = rec_of(h); fp.winStart = winStart; fp.N = Lb; fp.nonCoh = nSig; fp.swapIQ = 0;
        fp.dphi = h->dphi.p; fp.out = h->X.p; fp.tw = h->twFused.p;
        GC_CUDA(h, launch_fwd_cols(Lb, fp, nX, false, st)); ++launches;
        RowsParams rp{};
        rp.X = h->X.p; rp.nRows = (long long)nX * C;
        GC_CUDA(h, launch_fwd_rows(Lb, rp, st)); ++launches;
        cudaEventRecord(h->ev[2], st);
        std::vector<int2> map((size_t)nRows + nSv);
        for (int b = 0; b < nShifts; ++b)
            for (int k = 0; k < nBins; ++k)
                for (int g = 0; g < nSig; ++g) map[(b * nBins + k) * nSig + g] = make_int2(b * nSig + g, k);
        GC_CUDA(h, h->vbMap.reserve(map.size()));
        GC_CUDA(h, cudaMemcpyAsync(h->vbMap.p, map.data(), nRows * sizeof(int2), cudaMemcpyHostToDevice, st));
        std::vector<int> slotRep(nSv);
        for (int s = 0; s < nSv; ++s) slotRep[s] = svList[s] - 1;
        GC_CUDA(h, upload(h->prnList, slotRep, st));
        GC_CUDA(h, h->partMax.reserve((size_t)nSv * nRows * parts));
        GC_CUDA(h, h->partIdx.reserve((size_t)nSv * nRows * parts));
        GC_CUDA(h, h->peaks.reserve((size_t)nSv * nRows));
        int chunk = (int)std::max<long long>(1, (long long)(kWorkBytes / ((double)nRows * Lb * sizeof(float2))));
        chunk = std::min(chunk, (int)nSv);
        GC_CUDA(h, h->W.reserve(std::max((size_t)chunk * nRows, (size_t)nSv) * Lb));   // (the winner pass holds one row per SV)
        auto correlate = [&](int s0, int nc, int nB, const int2* bm, int bmStride, float* magOut, int b0 = 0, int nBTotal = 0) -> int {
            RowsParams ip{};
            ip.X = h->X.p; ip.Cc = h->Cc.p; ip.W = h->W.p; ip.tw = h->twFused.p;
            ip.nonCoh = 1; ip.nBins = nB; ip.bin0 = b0; ip.nRep = 1; ip.repStride = 1;
            ip.prnPerCta = 1; ip.mPerCta = 1; ip.binPerCta = 5; ip.binMap = bm; ip.binMapSlotStride = bmStride;
            if (nB == 1) { ip.prnPerCta = 5; ip.binPerCta = 1; }       // one row per SV: fill the CTA with SVs instead of bins
            ip.nPrnChunk = nc; ip.prnSlot0 = s0; ip.prnList = h->prnList.p;
            ip.zLoop = rows_zloop(Lb);
            GC_CUDA(h, launch_inv_rows(Lb, ip, st)); ++launches;
            InvColsParams cp{};
            cp.W = h->W.p; cp.colTw = h->twCols.p; cp.nBins = nB; cp.bin0 = b0; cp.nBinsTotal = nBTotal; cp.nonCoh = 1; cp.nPrnChunk = nc; cp.prnSlot0 = s0;
            cp.partMax = h->partMax.p; cp.partIdx = h->partIdx.p; cp.magOut = magOut;
            GC_CUDA(h, launch_inv_cols(Lb, cp, st)); ++launches;
            return GC_OK;
        };
        // GC_BIG_L2_MB = n: rows of one SV in groups whose work buffer is about n MB, so that the column pass reads what the row
        // pass wrote from the 126 MB L2 instead of HBM (every element of W is written once and read once here)
        const int l2mb = getenv("GC_BIG_L2_MB") ? atoi(getenv("GC_BIG_L2_MB")) : 0;
        if (l2mb > 0) {
            const int binChunk = std::max(5, (int)((double)l2mb * 1e6 / ((double)Lb * sizeof(float2))) / 5 * 5);
            for (int s0 = 0; s0 < nSv; ++s0)
                for (int b0 = 0; b0 < nRows; b0 += binChunk) {
                    const int rc = correlate(s0, 1, std::min(binChunk, nRows - b0), h->vbMap.p, 0, nullptr, b0, nRows);
                    if (rc != GC_OK) return rc;
                }
        } else
        for (int s0 = 0; s0 < nSv; s0 += chunk) {
            const int rc = correlate(s0, std::min(chunk, (int)nSv - s0), nRows, h->vbMap.p, 0, nullptr);
            if (rc != GC_OK) return rc;
        }
        GC_CUDA(h, launch_peak_select(h->partMax.p, h->partIdx.p, nSv * nRows, 1, parts, h->peaks.p, st)); ++launches;
        std::vector<PeakOut> po((size_t)nSv * nRows);
        GC_CUDA(h, cudaMemcpyAsync(po.data(), h->peaks.p, po.size() * sizeof(PeakOut), cudaMemcpyDeviceToHost, st));
        GC_CUDA(h, cudaStreamSynchronize(st));
        for (size_t i = 0; i < po.size(); ++i) peaks[i] = (float)po[i].peak;
        pick_winners();
        for (int s = 0; s < nSv; ++s) map[nRows + s] = make_int2(win[s].src, win[s].shift);
        GC_CUDA(h, cudaMemcpyAsync(h->vbMap.p + nRows, map.data() + nRows, nSv * sizeof(int2), cudaMemcpyHostToDevice, st));
        GC_CUDA(h, h->vbMag.reserve((size_t)nSv * Lb));
        {                                                     // corrVec of the winning row of every SV in one pass: slot s reads its own map
            const int rc = correlate(0, nSv, 1, h->vbMap.p + nRows, 1, h->vbMag.p);   // entry (partial maxima land in slot s, one bin)
            if (rc != GC_OK) return rc;
        }
        GC_CUDA(h, launch_peak_select(h->partMax.p, h->partIdx.p, nSv, 1, parts, h->peaks.p, st)); ++launches;
        GC_CUDA(h, cudaMemcpyAsync(po.data(), h->peaks.p, nSv * sizeof(PeakOut), cudaMemcpyDeviceToHost, st));
        GC_CUDA(h, cudaStreamSynchronize(st));
        for (int s = 0; s < nSv; ++s) { maxPeak[s] = (float)po[s].peak; cp0[s] = po[s].codePhase - 1; }
        std::vector<int4> seg(nSv);
        second_peak_ranges(seg);
        GC_CUDA(h, h->vbSeg.reserve(nSv));
        GC_CUDA(h, h->vbPeak.reserve(nSv));
        GC_CUDA(h, cudaMemcpyAsync(h->vbSeg.p, seg.data(), nSv * sizeof(int4), cudaMemcpyHostToDevice, st));
        GC_CUDA(h, launch_varb_segmax_mag(h->vbMag.p, nSv, Lb, h->vbSeg.p, h->vbPeak.p, st)); ++launches;
        GC_CUDA(h, cudaMemcpyAsync(second.data(), h->vbPeak.p, nSv * sizeof(float), cudaMemcpyDeviceToHost, st));
        cudaEventRecord(h->ev[1], st);
        GC_CUDA(h, cudaStreamSynchronize(st));
    } else {
        const size_t bufRows = (size_t)std::max(std::max(nRows, nX), std::max((int)nSv, h->resultLen));
        GC_CUDA(h, h->X.reserve((size_t)nX * Lb));
        GC_CUDA(h, h->T1.reserve(bufRows * Lb));
        GC_CUDA(h, h->T2.reserve(bufRows * Lb));
        GC_CUDA(h, launch_generic_wipe(rec_of(h), winStart, Lb, nSig, nShifts, 0, h->dphi.p, h->T1.p, Lb, st)); ++launches;
        float2 *src = h->T1.p, *dst = h->T2.p;
        int n = Lb, sd = 1;
        for (int f = 0; f < h->plan.nf; ++f) {
            GC_CUDA(h, launch_generic_stage(h->plan, f, n, sd, false, src, dst, nX, st)); ++launches;
            n /= h->plan.fac[f]; sd *= h->plan.fac[f];
            std::swap(src, dst);
        }
        GC_CUDA(h, cudaMemcpyAsync(h->X.p, src, (size_t)nX * Lb * sizeof(float2), cudaMemcpyDeviceToDevice, st));
        cudaEventRecord(h->ev[2], st);

        // every (shift, bin, block) row of every SV: peak of abs(ifft(circshift(X) .* codeFreqDom))
        GC_CUDA(h, h->vbRows.reserve((size_t)std::max(nRows, (int)nSv)));
        GC_CUDA(h, h->vbPeak.reserve((size_t)nSv * nRows));
        GC_CUDA(h, h->vbIdx.reserve((size_t)std::max(nRows, (int)nSv)));
        std::vector<VarbRow> rows(nRows);
        auto inverse = [&](int batch, float2** result) -> int {
            float2 *a = h->T1.p, *b = h->T2.p;
            int nn = Lb, ss = 1;
            for (int f = 0; f < h->plan.nf; ++f) {
                GC_CUDA(h, launch_generic_stage(h->plan, f, nn, ss, true, a, b, batch, st)); ++launches;
                nn /= h->plan.fac[f]; ss *= h->plan.fac[f];
                std::swap(a, b);
            }
            *result = a;
            return GC_OK;
        };
        for (int s = 0; s < nSv; ++s) {
            for (int b = 0; b < nShifts; ++b)
                for (int k = 0; k < nBins; ++k)
                    for (int g = 0; g < nSig; ++g)
                        rows[(b * nBins + k) * nSig + g] = VarbRow{b * nSig + g, svList[s] - 1, k, 0};
            GC_CUDA(h, cudaMemcpyAsync(h->vbRows.p, rows.data(), nRows * sizeof(VarbRow), cudaMemcpyHostToDevice, st));
            GC_CUDA(h, launch_varb_mulshift(h->X.p, h->Cc.p, h->vbRows.p, nRows, h->T1.p, Lb, st)); ++launches;
            float2* W = nullptr;
            int rc = inverse(nRows, &W);
            if (rc != GC_OK) return rc;
            GC_CUDA(h, launch_varb_rowpeak(W, nRows, Lb, h->vbPeak.p + (size_t)s * nRows, h->vbIdx.p, st)); ++launches;
            GC_CUDA(h, cudaStreamSynchronize(st));                // rows (pageable host vector) is rewritten for the next SV
        }
        GC_CUDA(h, cudaMemcpyAsync(peaks.data(), h->vbPeak.p, peaks.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
        GC_CUDA(h, cudaStreamSynchronize(st));

        pick_winners();
        // corrVec of the winning rows again, then its first maximum and the second peak outside +-1 chip
        GC_CUDA(h, cudaMemcpyAsync(h->vbRows.p, win.data(), nSv * sizeof(VarbRow), cudaMemcpyHostToDevice, st));
        GC_CUDA(h, launch_varb_mulshift(h->X.p, h->Cc.p, h->vbRows.p, nSv, h->T1.p, Lb, st)); ++launches;
        float2* W = nullptr;
        int rc = inverse(nSv, &W);
        if (rc != GC_OK) return rc;
        GC_CUDA(h, launch_varb_rowpeak(W, nSv, Lb, h->vbPeak.p, h->vbIdx.p, st)); ++launches;
        GC_CUDA(h, cudaMemcpyAsync(maxPeak.data(), h->vbPeak.p, nSv * sizeof(float), cudaMemcpyDeviceToHost, st));
        GC_CUDA(h, cudaMemcpyAsync(cp0.data(), h->vbIdx.p, nSv * sizeof(int), cudaMemcpyDeviceToHost, st));
        GC_CUDA(h, cudaStreamSynchronize(st));
        std::vector<int4> seg(nSv);
        second_peak_ranges(seg);
        GC_CUDA(h, h->vbSeg.reserve(nSv));
        GC_CUDA(h, cudaMemcpyAsync(h->vbSeg.p, seg.data(), nSv * sizeof(int4), cudaMemcpyHostToDevice, st));
        GC_CUDA(h, launch_varb_segmax(W, nSv, Lb, h->vbSeg.p, h->vbPeak.p, st)); ++launches;
        GC_CUDA(h, cudaMemcpyAsync(second.data(), h->vbPeak.p, nSv * sizeof(float), cudaMemcpyDeviceToHost, st));
        cudaEventRecord(h->ev[1], st);
        GC_CUDA(h, cudaStreamSynchronize(st));
    }
    int nAcq = 0;
    for (int s = 0; s < nSv; ++s) {
        const int ri = svList[s] - 1;
        peakMetric[ri] = (double)maxPeak[s] / (double)second[s];                           // :142
        if (coarseBin) coarseBin[ri] = winBin[s] + 1;
        if (coarseCodePhase) coarseCodePhase[ri] = cp0[s] + 1;
        if (peakMetric[ri] > c.acq_threshold) {                                            // :145
            ++nAcq;
            codePhase[ri] = cp0[s] + 1;
            carrFreq[ri] = h->vb.initFreq - h->vb.freqRes * winBin[s] + h->vb.sign * (h->vb.freqRes / nShifts) * winShift[s];   // B1I :150 ; L2C :95
        }
    }
    std::fill(h->clPhaseOut, h->clPhaseOut + 32, 0);
    if (!b1i && c.pilot_trk_flag == 1 && nAcq > 0) {
        // L2CL code phase from the detected CM code phase (GPS_L2C acquisition.m:100-137)
        const int N = h->N, segLen = 2 * c.code_length;
        std::vector<int> idx(N);
        const double ts = 1.0 / c.sampling_freq, tc = 1.0 / (c.code_freq_basis * 2);          // :117
        for (int i = 0; i < N; ++i) idx[i] = (int)std::ceil((ts * (double)i) / tc);          // :124
        idx[0] = 1;                                                                           // :127
        idx[N - 1] = (c.acq_coh_t > 0 && c.acq_coh_t <= 10) ? c.code_length : 2 * c.code_length;   // :128-133
        GC_CUDA(h, upload(h->clIdx, idx, st));
        GC_CUDA(h, h->clPower.reserve(75));
        for (int s = 0; s < nSv; ++s) {
            const int ri = svList[s] - 1;
            if (carrFreq[ri] == 0) continue;
            const std::vector<int8_t>& cl = h->hostCode[1][ri];
            if ((int)cl.size() != 75 * segLen) return fail(h, GC_ERR_ARG, "gc_acquire: pilotTRKflag == 1 needs the CL code of every SV (gc_set_code component 1)");
            const long long start = winStart + (long long)codePhase[ri] - 1;
            if (start + N > recSamples) return fail(h, GC_ERR_SHORT_RECORD, "gc_acquire: record too short for the CL phase search");
            GC_CUDA(h, upload(h->clDev, cl, st));
            GC_CUDA(h, launch_l2c_clphase(rec_of(h), start, N, h->clDev.p, segLen, h->clIdx.p, turns_to_fix(carrFreq[ri] * ts), h->clPower.p, st)); ++launches;
            double pw[75];
            GC_CUDA(h, cudaMemcpyAsync(pw, h->clPower.p, sizeof(pw), cudaMemcpyDeviceToHost, st));
            GC_CUDA(h, cudaStreamSynchronize(st));
            int best = 0;
            for (int i = 1; i < 75; ++i) if (pw[i] > pw[best]) best = i;                      // [~, CLCodePhase] = max(powerArray), :136
            h->clPhaseOut[ri] = best + 1;
        }
    }
    float total = 0, fwd = 0;
    cudaEventElapsedTime(&total, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&fwd, h->ev[0], h->ev[2]);
    h->stats.n_acquired = nAcq;
    h->stats.acq_fwd_ms = fwd; h->stats.acq_corr_ms = total - fwd; h->stats.acq_fine_ms = 0; h->stats.acq_total_ms = total;
    h->stats.corr_rows_ms = total - fwd; h->stats.corr_cols_ms = 0; h->stats.corr_row_launches = nSv;
    h->stats.acq_launches = launches;
    return GC_OK;
}

// ---- acquisition variant C (BDS/B1C/include/acquisition.m:128-276) -------------------------------------------------
static int varc_build_replicas(gc_handle* h)
{
    const gc_config& c = h->cfg;
    const int N = h->N, Lc = h->vc.Lc, xLen = h->vc.xLen, nSvMax = h->resultLen, nRep = h->nRep;
    // sampled BOC(1,1) tables (makeDataTable.m / makePilotTable.m): [sv][rep][N]; kept on the device for the fine search
    std::vector<int8_t> tab((size_t)nSvMax * 2 * N, 0);
    for (int prn = 1; prn <= nSvMax; ++prn)
        for (int r = 0; r < nRep; ++r)
            if (!h->hostCode[r][prn - 1].empty())
                make_boc_table(h->hostCode[r][prn - 1].data(), c.sampling_freq, c.code_freq_basis, c.code_length, N,
                               tab.data() + ((size_t)(prn - 1) * 2 + r) * N);
    cudaStream_t st = h->stream;
    GC_CUDA(h, upload(h->codeTab, tab, st));
    // local replicas [table(1:samplesXmsLen) zeros] -> conj(fft(.))/L   (:176-185); rows (prn-1)*2 + r
    const int rows = nSvMax * 2;
    std::vector<int8_t> head((size_t)rows * xLen);
    for (int r = 0; r < rows; ++r) std::copy(tab.begin() + (size_t)r * N, tab.begin() + (size_t)r * N + xLen, head.begin() + (size_t)r * xLen);
    GC_CUDA(h, upload(h->chips, head, st));
    GC_CUDA(h, h->Cc.reserve((size_t)rows * Lc));
    if (h->fused) {
        FwdColsParams fp{};
        fp.N = xLen; fp.codeTab = h->chips.p; fp.out = h->Cc.p; fp.tw = h->twFused.p;
        GC_CUDA(h, launch_fwd_cols(Lc, fp, rows, true, st));
        RowsParams rp{};
        rp.X = h->Cc.p; rp.nRows = (long long)rows * h->fp.C;
        GC_CUDA(h, launch_fwd_rows(Lc, rp, st));
        GC_CUDA(h, launch_finish_replica(h->Cc.p, (size_t)rows * Lc, Lc, st));
        GC_CUDA(h, cudaStreamSynchronize(st));
        h->replicasReady = true;
        return GC_OK;
    }
    const int chunk = 32;                                       // transform the replicas in chunks to bound the scratch buffers
    GC_CUDA(h, h->T1.reserve((size_t)chunk * Lc));
    GC_CUDA(h, h->T2.reserve((size_t)chunk * Lc));
    for (int r0 = 0; r0 < rows; r0 += chunk) {
        const int nr = std::min(chunk, rows - r0);
        GC_CUDA(h, launch_varb_pad(h->chips.p + (size_t)r0 * xLen, xLen, nr, h->T1.p, Lc, st));
        float2 *src = h->T1.p, *dst = h->T2.p;
        int n = Lc, sd = 1;
        for (int f = 0; f < h->plan.nf; ++f) {
            GC_CUDA(h, launch_generic_stage(h->plan, f, n, sd, false, src, dst, nr, st));
            n /= h->plan.fac[f]; sd *= h->plan.fac[f];
            std::swap(src, dst);
        }
        GC_CUDA(h, cudaMemcpyAsync(h->Cc.p + (size_t)r0 * Lc, src, (size_t)nr * Lc * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    }
    GC_CUDA(h, launch_generic_conj_scale(h->Cc.p, (size_t)rows * Lc, 1.0f / (float)Lc, st));
    GC_CUDA(h, cudaStreamSynchronize(st));
    h->replicasReady = true;
    return GC_OK;
}

static int acquire_varc(gc_handle* h, long long winStart, long long longLen, int32_t nSv, const int32_t* svList,
                        double* carrFreq, double* codePhase, double* peakMetric, int32_t* coarseBin, int32_t* coarseCodePhase)
{
    const gc_config& c = h->cfg;
    const int N = h->N, Lc = h->vc.Lc, nBins = h->nBins, nRep = h->nRep, nFine = h->vc.nFine;
    cudaStream_t st = h->stream;
    if (longLen <= 0) longLen = 2LL * N;                     // postProcessing.m:33 reads two code periods
    const long long recSamples = rec_samples(h);
    if (winStart < 0 || winStart + std::max<long long>(Lc, longLen) > recSamples || longLen < Lc)
        return fail(h, GC_ERR_SHORT_RECORD, "gc_acquire: record shorter than (10 + acqCohT) ms");
    for (int i = 0; i < h->resultLen; ++i) {
        carrFreq[i] = codePhase[i] = peakMetric[i] = 0;
        if (coarseBin) coarseBin[i] = 0;
        if (coarseCodePhase) coarseCodePhase[i] = 0;
    }
    int launches = 0;
    cudaEventRecord(h->ev[0], st);
    GC_CUDA(h, h->sigPower.reserve(1));
    GC_CUDA(h, launch_sig_power(rec_of(h), winStart, h->vc.xLen, h->sigPower.p, st)); ++launches;   // :163
    const uint64_t dphi0 = turns_to_fix(h->vc.initFreq / c.sampling_freq);
    GC_CUDA(h, h->dphi.reserve(1));
    GC_CUDA(h, cudaMemcpyAsync(h->dphi.p, &dphi0, sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    int parts = 1;
    if (h->fused) {
        // the plan's kernels: one forward spectrum; per SV the (bin, replica) rows as X(j - bin) .* C, inverse rows -> work buffer ->
        // inverse columns with the data / pilot magnitudes weighted sqrt(11) : sqrt(29) and scaled by 1/sqrt(40) (:213-214)
        parts = h->fp.parts;
        GC_CUDA(h, h->X.reserve((size_t)Lc));
        FwdColsParams fp{};
        fp.rec = rec_of(h); fp.winStart = winStart; fp.N = Lc; fp.nonCoh = 1; fp.swapIQ = 0;
        fp.dphi = h->dphi.p; fp.out = h->X.p; fp.tw = h->twFused.p;
        GC_CUDA(h, launch_fwd_cols(Lc, fp, 1, false, st)); ++launches;                      // :168-179
        RowsParams rp{};
        rp.X = h->X.p; rp.nRows = h->fp.C;
        GC_CUDA(h, launch_fwd_rows(Lc, rp, st)); ++launches;
        cudaEventRecord(h->ev[2], st);
        std::vector<int2> map(nBins);
        for (int k = 0; k < nBins; ++k) map[k] = make_int2(0, k);                          // circshift(IQfreqDom, k) (:203)
        GC_CUDA(h, upload_cached(h, h->vbMap, map, st));
        std::vector<int> slotRep(nSv);
        for (int s = 0; s < nSv; ++s) slotRep[s] = (svList[s] - 1) * 2;
        GC_CUDA(h, upload(h->prnList, slotRep, st));
        GC_CUDA(h, h->partMax.reserve((size_t)nSv * nBins * parts));
        GC_CUDA(h, h->partIdx.reserve((size_t)nSv * nBins * parts));
        GC_CUDA(h, h->peaks.reserve(nSv));
        int chunk = (int)std::max<long long>(1, (long long)(kWorkBytes / ((double)nBins * nRep * Lc * sizeof(float2))));
        chunk = std::min(chunk, (int)nSv);
        GC_CUDA(h, h->W.reserve((size_t)chunk * nBins * nRep * Lc));
        const int l2mb = getenv("GC_BIG_L2_MB") ? atoi(getenv("GC_BIG_L2_MB")) : 0;     // (see acquire_varb)
        const int binChunk = l2mb > 0 ? std::max(5, (int)((double)l2mb * 1e6 / ((double)nRep * Lc * sizeof(float2))) / 5 * 5) : nBins;
        if (l2mb > 0) chunk = 1;
        for (int s0 = 0; s0 < nSv; s0 += chunk)
        for (int b0 = 0; b0 < nBins; b0 += binChunk) {
            const int nc = std::min(chunk, (int)nSv - s0), nb = std::min(binChunk, nBins - b0);
            RowsParams ip{};
            ip.X = h->X.p; ip.Cc = h->Cc.p; ip.W = h->W.p; ip.tw = h->twFused.p;
            ip.nonCoh = 1; ip.nBins = nb; ip.bin0 = b0; ip.nRep = nRep; ip.repStride = 1;
            ip.prnPerCta = 1; ip.mPerCta = 1; ip.binPerCta = 5; ip.binMap = h->vbMap.p;
            ip.nPrnChunk = nc; ip.prnSlot0 = s0; ip.prnList = h->prnList.p;
            ip.zLoop = rows_zloop(Lc);
            GC_CUDA(h, launch_inv_rows(Lc, ip, st)); ++launches;
            InvColsParams cp{};
            cp.W = h->W.p; cp.colTw = h->twCols.p; cp.nBins = nb; cp.bin0 = b0; cp.nBinsTotal = nBins; cp.nonCoh = nRep; cp.nPrnChunk = nc; cp.prnSlot0 = s0;
            cp.partMax = h->partMax.p; cp.partIdx = h->partIdx.p;
            if (nRep == 2) { cp.weighted = 1; cp.w0 = 3.3166247903554f; cp.w1 = 5.385164807134504f; cp.wScale = 1.0f / 6.324555320336759f; }
            GC_CUDA(h, launch_inv_cols(Lc, cp, st)); ++launches;
        }
    } else {
        const int nRows = nBins * nRep;
        GC_CUDA(h, h->X.reserve((size_t)Lc));
        GC_CUDA(h, h->T1.reserve((size_t)std::max(nRows, 32) * Lc));
        GC_CUDA(h, h->T2.reserve((size_t)std::max(nRows, 32) * Lc));
        GC_CUDA(h, launch_generic_wipe(rec_of(h), winStart, Lc, 1, 1, 0, h->dphi.p, h->T1.p, Lc, st)); ++launches;   // :168-172
        {
            float2 *src = h->T1.p, *dst = h->T2.p;
            int n = Lc, sd = 1;
            for (int f = 0; f < h->plan.nf; ++f) {
                GC_CUDA(h, launch_generic_stage(h->plan, f, n, sd, false, src, dst, 1, st)); ++launches;
                n /= h->plan.fac[f]; sd *= h->plan.fac[f];
                std::swap(src, dst);
            }
            GC_CUDA(h, cudaMemcpyAsync(h->X.p, src, (size_t)Lc * sizeof(float2), cudaMemcpyDeviceToDevice, st));
        }
        cudaEventRecord(h->ev[2], st);
        GC_CUDA(h, h->vbRows.reserve(nRows));
        GC_CUDA(h, h->partMax.reserve((size_t)nSv * nBins));
        GC_CUDA(h, h->partIdx.reserve((size_t)nSv * nBins));
        GC_CUDA(h, h->peaks.reserve(nSv));
        std::vector<VarbRow> rows(nRows);
        for (int s = 0; s < nSv; ++s) {
            for (int k = 0; k < nBins; ++k)
                for (int r = 0; r < nRep; ++r) rows[k * nRep + r] = VarbRow{0, (svList[s] - 1) * 2 + r, k, 0};   // circshift(IQfreqDom, k) (:203)
            GC_CUDA(h, cudaMemcpyAsync(h->vbRows.p, rows.data(), nRows * sizeof(VarbRow), cudaMemcpyHostToDevice, st));
            GC_CUDA(h, launch_varb_mulshift(h->X.p, h->Cc.p, h->vbRows.p, nRows, h->T1.p, Lc, st)); ++launches;
            float2 *a = h->T1.p, *b = h->T2.p;
            int nn = Lc, ss = 1;
            for (int f = 0; f < h->plan.nf; ++f) {
                GC_CUDA(h, launch_generic_stage(h->plan, f, nn, ss, true, a, b, nRows, st)); ++launches;
                nn /= h->plan.fac[f]; ss *= h->plan.fac[f];
                std::swap(a, b);
            }
            GC_CUDA(h, launch_varc_combine(a, nBins, Lc, nRep, h->partMax.p, h->partIdx.p, (size_t)s * nBins, st)); ++launches;
            GC_CUDA(h, cudaStreamSynchronize(st));
        }
    }
    GC_CUDA(h, launch_peak_select(h->partMax.p, h->partIdx.p, nSv, nBins, parts, h->peaks.p, st)); ++launches;   // :221-225
    std::vector<PeakOut> peaks(nSv);
    double sigPower = 0;
    GC_CUDA(h, cudaMemcpyAsync(peaks.data(), h->peaks.p, nSv * sizeof(PeakOut), cudaMemcpyDeviceToHost, st));
    GC_CUDA(h, cudaMemcpyAsync(&sigPower, h->sigPower.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    cudaEventRecord(h->ev[1], st);
    GC_CUDA(h, cudaStreamSynchronize(st));
    std::vector<int> acq;
    std::vector<int> cps, slots;
    std::vector<double> selFreq;
    for (int s = 0; s < nSv; ++s) {
        const int ri = svList[s] - 1;
        peakMetric[ri] = peaks[s].peak / sigPower;                                         // :227
        int cp = peaks[s].codePhase;
        if ((long long)cp + N - 1 > longLen) cp -= N;                                      // :229-231
        if (coarseBin) coarseBin[ri] = peaks[s].bin;
        if (coarseCodePhase) coarseCodePhase[ri] = cp;
        if (peakMetric[ri] > c.acq_threshold) {                                            // :234
            acq.push_back(s); cps.push_back(cp); slots.push_back((svList[s] - 1) * 2);
            selFreq.push_back(h->vc.initFreq - (peaks[s].bin - 1) * c.acq_search_step);    // :223
        }
    }
    const int nAcq = (int)acq.size();
    float fineMs = 0;
    if (nAcq > 0) {
        std::vector<uint64_t> fd((size_t)nAcq * nFine);
        std::vector<double> ff((size_t)nAcq * nFine);
        for (int a = 0; a < nAcq; ++a)
            for (int j = 0; j < nFine; ++j) {
                ff[(size_t)a * nFine + j] = selFreq[a] + c.acq_search_step - 25.0 * j;     // :244
                fd[(size_t)a * nFine + j] = turns_to_fix(ff[(size_t)a * nFine + j] / c.sampling_freq);
            }
        GC_CUDA(h, upload(h->fdphi, fd, st));
        GC_CUDA(h, upload(h->fineCodePhase, cps, st));
        GC_CUDA(h, upload(h->vcSlot, slots, st));
        GC_CUDA(h, h->fineResult.reserve((size_t)nAcq * nFine));
        cudaEventRecord(h->ev[3], st);
        GC_CUDA(h, launch_varc_fine(rec_of(h), winStart, N, nRep, h->codeTab.p, h->vcSlot.p, h->fineCodePhase.p, h->fdphi.p, nFine, nAcq,
                                    h->fineResult.p, st)); ++launches;
        std::vector<double> fr((size_t)nAcq * nFine);
        GC_CUDA(h, cudaMemcpyAsync(fr.data(), h->fineResult.p, fr.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        cudaEventRecord(h->ev[4], st);
        GC_CUDA(h, cudaStreamSynchronize(st));
        cudaEventElapsedTime(&fineMs, h->ev[3], h->ev[4]);
        for (int a = 0; a < nAcq; ++a) {
            int best = 0;
            for (int j = 1; j < nFine; ++j) if (fr[(size_t)a * nFine + j] > fr[(size_t)a * nFine + best]) best = j;   // :252
            const int ri = svList[acq[a]] - 1;
            carrFreq[ri] = ff[(size_t)a * nFine + best];                                   // :253
            if (carrFreq[ri] == 0) carrFreq[ri] = 1;
            codePhase[ri] = cps[a];                                                        // :258
        }
    }
    float total = 0, fwd = 0;
    cudaEventElapsedTime(&total, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&fwd, h->ev[0], h->ev[2]);
    h->stats.n_acquired = nAcq;
    h->stats.acq_fwd_ms = fwd; h->stats.acq_corr_ms = total - fwd; h->stats.acq_fine_ms = fineMs; h->stats.acq_total_ms = total + fineMs;
    h->stats.corr_rows_ms = total - fwd; h->stats.corr_cols_ms = 0; h->stats.corr_row_launches = nSv;
    h->stats.acq_launches = launches;
    return GC_OK;
}

// device-side timings of the last variant A enqueue from its events (after the stream has been synchronised)
static void resolve_acq_stats(gc_handle* h)
{
    if (!h->statsPending) return;
    h->statsPending = false;
    const AcqEnq& info = h->pendingInfo;
    float rowsMs = 0, colsMs = 0, fwdMs = 0, fineMs = 0, coarseMs = 0, ms = 0;
    for (auto& pr : info.rowEv) if (cudaEventElapsedTime(&ms, h->ev[pr.first], h->ev[pr.second]) == cudaSuccess) rowsMs += ms;
    for (auto& pr : info.colEv) if (cudaEventElapsedTime(&ms, h->ev[pr.first], h->ev[pr.second]) == cudaSuccess) colsMs += ms;
    for (auto& pr : info.fwdEv) if (cudaEventElapsedTime(&ms, h->ev[pr.first], h->ev[pr.second]) == cudaSuccess) fwdMs += ms;
    if (info.fa >= 0 && cudaEventElapsedTime(&ms, h->ev[info.fa], h->ev[info.fb]) == cudaSuccess) fineMs = ms;
    if (info.e0 >= 0 && cudaEventElapsedTime(&ms, h->ev[info.e0], h->ev[info.e1]) == cudaSuccess) coarseMs = ms;
    cudaGetLastError();
    h->stats.acq_fwd_ms = fwdMs; h->stats.acq_corr_ms = coarseMs - fwdMs; h->stats.acq_fine_ms = fineMs; h->stats.acq_total_ms = coarseMs + fineMs;
    h->stats.corr_rows_ms = rowsMs; h->stats.corr_cols_ms = colsMs; h->stats.corr_row_launches = info.nRowLaunches; h->stats.acq_launches = info.launches;
}

// Variant A on a fused plan with the split correlation stage - the default path of every variant A signal (GPS L1CA, GLONASS,
// B3I, E1, L5C, E5a, E5b, B2a).  All host tables and buffer reservations first, then ONE enqueue of the kernels and the small
// result copies: issued directly the first time a call comes in, captured into a CUDA graph the second time the same call
// (window, SV list, record, result buffer, device buffers) arrives, one cudaGraphLaunch from the third on.  The ~12 launches,
// their event records and the four result copies cost more host time than the grid of a sharded search takes on the device.
// Several carrier grids (GLONASS: one per frequency number) are transformed and searched in ONE pass each.
static int acquire_plain(gc_handle* h, long long winStart, int32_t nSv, const int32_t* svList, const std::vector<int>& order,
                         const std::vector<int>& groupStart /* nGroups + 1 entries */, const std::vector<int>& slotGroup,
                         double* carrFreq, double* codePhase, double* peakMetric, int32_t* coarseBin, int32_t* coarseCodePhase,
                         double* dOut, bool async)
{
    HostLaps laps("gc_acquire");
    const gc_config& c = h->cfg;
    const int N = h->N, L = h->L, nBins = h->nBins, nonCoh = h->nonCoh, nKm = nBins * nonCoh;
    const int nGroups = (int)groupStart.size() - 1;
    const int codeLen = c.code_length;
    cudaStream_t st = h->stream;

    // ---- host tables and buffers (nothing here is part of the graph) -------------------------------------------------------
    std::vector<std::vector<double>> coarseFreqOf(nSv);   // per list slot: the bin frequencies it was searched on
    std::vector<uint64_t> dphi((size_t)nGroups * nBins);
    for (int gi = 0; gi < nGroups; ++gi) {
        const double off = sv_freq_offset(h, svList[order[groupStart[gi]]]);
        std::vector<double> coarseFreq(nBins);
        for (int k = 0; k < nBins; ++k) {
            coarseFreq[k] = (c.IF + off) + c.acq_search_band - c.acq_search_step * k;   // :169 (GLO :181-182)
            dphi[(size_t)gi * nBins + k] = turns_to_fix(coarseFreq[k] * h->ts);
        }
        for (int s = groupStart[gi]; s < groupStart[gi + 1]; ++s) coarseFreqOf[s] = coarseFreq;
    }
    const bool shifted = h->binShift > 0;                    // only the first binQ bins of every grid are transformed, the others are shifts of them
    const int nq = shifted ? std::min(h->binQ, nBins) : nBins;
    const int fwdRowsPerGroup = nq * nonCoh;
    GC_CUDA(h, h->X.reserve((size_t)nGroups * fwdRowsPerGroup * L));
    if (shifted) {
        std::vector<uint64_t> d0((size_t)nGroups * nq);
        for (int gi = 0; gi < nGroups; ++gi)
            for (int b = 0; b < nq; ++b) d0[(size_t)gi * nq + b] = dphi[(size_t)gi * nBins + b];
        GC_CUDA(h, upload_cached(h, h->dphi, d0, st));
        std::vector<int2> map(nBins);
        for (int k = 0; k < nBins; ++k) map[k] = make_int2(k % nq, (k / nq) * h->binShift);
        GC_CUDA(h, upload_cached(h, h->vbMap, map, st));
    } else {
        GC_CUDA(h, upload_cached(h, h->dphi, dphi, st));
    }
    GC_CUDA(h, upload_cached(h, h->slotGroup, slotGroup, st));
    int chunk = (int)std::max<long long>(1, (long long)(kWorkBytes / ((double)nKm * h->nRep * L * sizeof(float2))));
    chunk = std::min(chunk, (int)nSv);
    const int nChunks = (nSv + chunk - 1) / chunk;
    GC_CUDA(h, h->W.reserve((size_t)chunk * nKm * h->nRep * L));
    GC_CUDA(h, h->partMax.reserve((size_t)nSv * nBins * h->parts));
    GC_CUDA(h, h->partIdx.reserve((size_t)nSv * nBins * h->parts));
    GC_CUDA(h, h->peaks.reserve(nSv));
    GC_CUDA(h, h->sigPower.reserve(1));
    const int nCodes = h->fineTwoCodes ? 2 : 1;
    const int nPeriods = h->nFinePeriods;                                            // :146-148 (B3I :131-133)
    const int tabLen = codeLen * h->sub;                                             // chips, or BOC sub-chips, per code period
    const int maxEnt = nSv * nCodes;
    std::vector<double> slotFreq0(nSv);
    for (int s = 0; s < nSv; ++s) slotFreq0[s] = (c.IF + sv_freq_offset(h, svList[order[s]])) + c.acq_search_band;   // coarseFreqBin(1), :169 (GLO :181-182)
    bool haveSecondary = false;
    if (!h->noFine) {
        std::vector<int> slotChipRow(nSv), slotSv(nSv);
        std::vector<int8_t> slotSec;
        if (h->fineCombine == 4) slotSec.resize((size_t)nSv * nPeriods);
        for (int s = 0; s < nSv; ++s) {
            const int sv = svList[order[s]];
            slotChipRow[s] = h->glo ? 0 : sv_result_index(h, sv) * 2;
            slotSv[s] = sv;
            if (h->fineCombine == 4) {
                static const int8_t NH20[20] = {1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1};   // GPS_L5C acquisition.m:134
                for (int q = 0; q < nPeriods; ++q)
                    slotSec[(size_t)s * nPeriods + q] = (c.signal == GC_SIG_GAL_E5A) ? h->hostCode[2][sv - 1][q]   // generateE5aQ_secondary
                                                                                       : NH20[q % 20];
            }
        }
        haveSecondary = !slotSec.empty();
        GC_CUDA(h, upload_cached(h, h->slotFreq0, slotFreq0, st));
        GC_CUDA(h, upload_cached(h, h->slotChipRow, slotChipRow, st));
        GC_CUDA(h, upload_cached(h, h->slotSv, slotSv, st));
        if (haveSecondary) GC_CUDA(h, upload_cached(h, h->slotSecondary, slotSec, st));
        GC_CUDA(h, h->metricDev.reserve(nSv));
        GC_CUDA(h, h->nAcqDev.reserve(1));
        GC_CUDA(h, h->acqSlot.reserve(nSv));
        GC_CUDA(h, h->fineChipRow.reserve(maxEnt));
        GC_CUDA(h, h->fineCodePhase.reserve(maxEnt));
        GC_CUDA(h, h->fdphi.reserve((size_t)maxEnt * h->nFine));
        GC_CUDA(h, h->fineSv.reserve(nSv));
        GC_CUDA(h, h->fineSecondary.reserve((size_t)nSv * nPeriods));
        GC_CUDA(h, h->fineProd.reserve((size_t)maxEnt * nPeriods * N));
        GC_CUDA(h, h->fineSums.reserve((size_t)maxEnt * h->nFine * nPeriods * 2));
        GC_CUDA(h, h->fineResult.reserve((size_t)nSv * h->nFine));
        GC_CUDA(h, h->fineBest.reserve(nSv));
    }
    if (dOut) {                                                   // acqResults assembled on the device for the collective that follows
        std::vector<int> slotResult(nSv);
        for (int s = 0; s < nSv; ++s) slotResult[s] = sv_result_index(h, svList[order[s]]);
        GC_CUDA(h, upload_cached(h, h->slotResult, slotResult, st));
        if (h->noFine) GC_CUDA(h, upload_cached(h, h->slotFreq0, slotFreq0, st));
    }
    // pinned staging of the result copies: [PeakOut peaks[nSv] | double sigPower | int best[nSv] | int nAcq]
    const size_t offSig = (size_t)nSv * sizeof(PeakOut), offBest = offSig + sizeof(double), offN = offBest + (size_t)nSv * sizeof(int);
    const size_t needPin = offN + sizeof(int);
    if (needPin > h->pinBytes) {
        if (h->pin) cudaFreeHost(h->pin);
        h->pin = nullptr; h->pinBytes = 0;
        GC_CUDA(h, cudaHostAlloc((void**)&h->pin, needPin + 256, cudaHostAllocDefault));
        h->pinBytes = needPin + 256;
    }
    PeakOut* pinPeaks = reinterpret_cast<PeakOut*>(h->pin);
    double* pinSig = reinterpret_cast<double*>(h->pin + offSig);
    int* pinBest = reinterpret_cast<int*>(h->pin + offBest);
    int* pinN = reinterpret_cast<int*>(h->pin + offN);
    *pinN = -1;
    laps.lap("host tables, buffers");

    // ---- the enqueue ---------------------------------------------------------------------------------------------------------
    const bool perChunkEvents = 2 * nChunks + 8 <= kEvents;
    auto enqueue = [&](AcqEnq& q, bool capturing) -> int {
        int evn = 0;
        auto mark = [&]() {
            if (capturing) cudaEventRecordWithFlags(h->ev[evn], st, cudaEventRecordExternal);
            else cudaEventRecord(h->ev[evn], st);
            return evn++;
        };
        q = AcqEnq{};
        q.e0 = mark();
        GC_CUDA(h, launch_sig_power(rec_of(h), winStart, N, h->sigPower.p, st)); ++q.launches;   // :151
        int prev = mark();
        FwdColsParams fp{};
        fp.rec = rec_of(h); fp.winStart = winStart; fp.N = N; fp.nonCoh = nonCoh; fp.swapIQ = (h->glo && h->fmt != 2 && h->fmt != 3) ? 1 : 0;
        fp.dphi = h->dphi.p; fp.out = h->X.p; fp.tw = h->twFused.p;          // row (grid, bin, block): the phase table is indexed grid*nBins + bin
        GC_CUDA(h, launch_fwd_cols(L, fp, nGroups * fwdRowsPerGroup, false, st)); ++q.launches;
        RowsParams rp{};
        rp.X = h->X.p; rp.nRows = (long long)nGroups * fwdRowsPerGroup * h->fp.C;
        GC_CUDA(h, launch_fwd_rows(L, rp, st)); ++q.launches;
        { const int m = mark(); q.fwdEv.push_back({prev, m}); prev = m; }
        for (int s0 = 0; s0 < nSv; s0 += chunk) {
            const int nc = std::min(chunk, (int)nSv - s0);
            RowsParams ip{};
            ip.X = h->X.p; ip.Cc = h->Cc.p; ip.W = h->W.p; ip.tw = h->twFused.p;
            ip.nonCoh = nonCoh; ip.nBins = nBins; ip.prnPerCta = 1; ip.mPerCta = 5;   // 5 warps, 96 registers, 20 warps/SM
            ip.nRep = h->nRep; ip.repStride = 1;
            if (nonCoh * h->nRep < 5) { ip.prnPerCta = 5; ip.mPerCta = 1; }           // few transforms per cell (Galileo E1: 2): fill the CTA with SVs
            ip.zLoop = rows_zloop(L);
            ip.nPrnChunk = nc; ip.prnSlot0 = s0; ip.prnList = h->prnList.p;
            ip.slotGroup = h->slotGroup.p; ip.groupRows = fwdRowsPerGroup;
            if (shifted) ip.binMap = h->vbMap.p;
            GC_CUDA(h, launch_inv_rows(L, ip, st)); ++q.launches; ++q.nRowLaunches;
            if (perChunkEvents) { const int m = mark(); q.rowEv.push_back({prev, m}); prev = m; }
            InvColsParams cp{};
            cp.W = h->W.p; cp.colTw = h->twCols.p; cp.nBins = nBins; cp.nonCoh = nonCoh * h->nRep; cp.nPrnChunk = nc; cp.prnSlot0 = s0;
            cp.partMax = h->partMax.p; cp.partIdx = h->partIdx.p;
            GC_CUDA(h, launch_inv_cols(L, cp, st)); ++q.launches;
            if (perChunkEvents) { const int m = mark(); q.colEv.push_back({prev, m}); prev = m; }
        }
        GC_CUDA(h, launch_peak_select(h->partMax.p, h->partIdx.p, nSv, nBins, h->parts, h->peaks.p, st)); ++q.launches;
        q.e1 = mark();
        // threshold (:200-206) and fine search (:211-253) follow on the device without a host round trip: fine_setup_kernel
        // decides which list slots are above the threshold and prepares their fine-search entries, the fine kernels are
        // launched for the worst case (every slot acquired) and idle blocks leave at once
        if (!h->noFine) {
            FineSetup fs{};
            fs.peaks = h->peaks.p; fs.sigPower = h->sigPower.p; fs.nSv = nSv; fs.nonCoh = nonCoh; fs.nFine = h->nFine; fs.nCodes = nCodes;
            fs.nPeriods = nPeriods; fs.pilotComp = h->nRep - 1; fs.threshold = c.acq_threshold; fs.step = c.acq_search_step;
            fs.fineStep = h->fineStep; fs.ts = h->ts; fs.slotFreq0 = h->slotFreq0.p; fs.slotChipRow = h->slotChipRow.p; fs.slotSv = h->slotSv.p;
            fs.slotSecondary = haveSecondary ? h->slotSecondary.p : nullptr;
            fs.metric = h->metricDev.p; fs.nAcq = h->nAcqDev.p; fs.acqSlot = h->acqSlot.p; fs.chipRow = h->fineChipRow.p;
            fs.codePhase = h->fineCodePhase.p; fs.dphi = h->fdphi.p; fs.svId = h->fineSv.p; fs.secondary = h->fineSecondary.p;
            q.fa = q.e1;
            GC_CUDA(h, launch_fine_setup(fs, st)); ++q.launches;
            FineParams fpp{};
            fpp.rec = rec_of(h); fpp.winStart = winStart; fpp.N = N; fpp.nPeriods = nPeriods; fpp.nFine = h->nFine; fpp.codeLen = tabLen;
            fpp.swapIQ = (h->glo && h->fmt != 2 && h->fmt != 3) ? 1 : 0; fpp.combine = h->fineCombine; fpp.chipIdx = h->chipIdx.p; fpp.svId = h->fineSv.p;
            fpp.chips = h->chips.p; fpp.chipRow = h->fineChipRow.p; fpp.codePhase = h->fineCodePhase.p; fpp.dphi = h->fdphi.p; fpp.prod = h->fineProd.p;
            fpp.sums = h->fineSums.p; fpp.best = h->fineBest.p; fpp.fineResult = h->fineResult.p;
            fpp.nAcqDev = h->nAcqDev.p; fpp.nCodes = nCodes; fpp.secondary = h->fineSecondary.p;
            fpp.moments = fine_moments(h->nFine, h->fineStep, h->ts);
            GC_CUDA(h, launch_fine(fpp, maxEnt, nSv, st)); q.launches += 3;
            q.fb = mark();
        }
        if (dOut) {
            PackParams pk{};
            pk.peaks = h->peaks.p; pk.sigPower = h->sigPower.p; pk.slotFreq0 = h->slotFreq0.p; pk.slotResult = h->slotResult.p;
            pk.best = h->noFine ? nullptr : h->fineBest.p; pk.nSv = nSv; pk.nonCoh = nonCoh; pk.resultLen = h->resultLen; pk.noFine = h->noFine ? 1 : 0;
            pk.threshold = c.acq_threshold; pk.step = c.acq_search_step; pk.fineStep = h->fineStep; pk.out = dOut;
            GC_CUDA(h, launch_pack_results(pk, st)); ++q.launches;
        } else {
            GC_CUDA(h, cudaMemcpyAsync(pinPeaks, h->peaks.p, nSv * sizeof(PeakOut), cudaMemcpyDeviceToHost, st));
            GC_CUDA(h, cudaMemcpyAsync(pinSig, h->sigPower.p, sizeof(double), cudaMemcpyDeviceToHost, st));
            if (!h->noFine) {
                GC_CUDA(h, cudaMemcpyAsync(pinBest, h->fineBest.p, nSv * sizeof(int), cudaMemcpyDeviceToHost, st));
                GC_CUDA(h, cudaMemcpyAsync(pinN, h->nAcqDev.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            }
        }
        return GC_OK;
    };

    // ---- direct, capture or replay ----------------------------------------------------------------------------------------------
    std::vector<char> key;
    auto put = [&](const void* p_, size_t n_) { const char* b = static_cast<const char*>(p_); key.insert(key.end(), b, b + n_); };
    {
        // everything the enqueue bakes into kernel arguments: the call's own parameters and EVERY device pointer it passes (a buffer
        // that was reallocated since the capture changes the key, so a stale graph is never replayed; other handles' allocations do
        // not disturb it)
        const void* ptrs[] = {h->rec, dOut, h->pin, h->sigPower.p, h->dphi.p, h->X.p, h->twFused.p, h->Cc.p, h->W.p, h->prnList.p, h->slotGroup.p,
                              h->vbMap.p, h->partMax.p, h->partIdx.p, h->peaks.p, h->slotFreq0.p, h->slotChipRow.p, h->slotSv.p,
                              h->slotSecondary.p, h->metricDev.p, h->nAcqDev.p, h->acqSlot.p, h->fineChipRow.p, h->fineCodePhase.p, h->fdphi.p,
                              h->fineSv.p, h->fineSecondary.p, h->chipIdx.p, h->chips.p, h->fineProd.p, h->fineSums.p, h->fineBest.p,
                              h->fineResult.p, h->slotResult.p};
        put(&winStart, sizeof winStart); put(&nSv, sizeof nSv); put(svList, sizeof(int32_t) * nSv);
        put(&h->recBytes, sizeof h->recBytes); put(ptrs, sizeof ptrs);
    }
    AcqGraph& g = h->graph;
    AcqEnq info;
    int mode = 0;                                             // 0 direct, 1 capture, 2 replay
    if (!h->graphOff) {
        if (g.key == key) mode = g.exec ? 2 : (g.hits >= 1 ? 1 : 0);
        else {
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            g.key = key; g.hits = 0;
        }
    }
    if (mode == 1) {
        cudaGraph_t graph = nullptr;
        int rc = GC_OK;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) { h->graphOff = true; mode = 0; cudaGetLastError(); }
        else {
            rc = enqueue(info, true);
            const cudaError_t ee = cudaStreamEndCapture(st, &graph);
            if (rc == GC_OK && ee == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                g.info = info;
                mode = 2;
            } else {                                          // never again on this handle; run this call directly
                h->graphOff = true; g.exec = nullptr; mode = 0;
                cudaGetLastError();
            }
            if (graph) cudaGraphDestroy(graph);
        }
        laps.lap("graph capture + instantiate");
    }
    if (mode == 2) {
        info = g.info;
        GC_CUDA(h, cudaGraphLaunch(g.exec, st));
    } else {
        const int rc = enqueue(info, false);
        if (rc != GC_OK) return rc;
        ++g.hits;
    }
    laps.lap(mode == 2 ? "graph launch" : "direct enqueue");
    h->pendingInfo = info;
    h->statsPending = true;
    if (dOut && async) {                                      // gc_acquire_device_async: the caller orders its work behind the handle's stream
        h->stats.n_acquired = -1;
        return GC_OK;
    }
    GC_CUDA(h, cudaStreamSynchronize(st));
    laps.lap("synchronize");
    resolve_acq_stats(h);
    if (dOut) {
        h->stats.n_acquired = -1;                             // (in the device buffer: carrFreq != 0)
        return GC_OK;
    }
    const double sigPower = *pinSig;
    std::vector<int> acq;   // list slots above the threshold, in list order (what fine_setup_kernel found too)
    for (int s = 0; s < nSv; ++s) {
        const int ri = sv_result_index(h, svList[order[s]]);
        peakMetric[ri] = pinPeaks[s].peak / sigPower / nonCoh;                       // :200
        if (coarseBin) coarseBin[ri] = pinPeaks[s].bin;
        if (coarseCodePhase) coarseCodePhase[ri] = pinPeaks[s].codePhase;
        if (peakMetric[ri] > c.acq_threshold) acq.push_back(s);                      // :206
    }
    const int nAcq = (int)acq.size();
    h->stats.n_acquired = nAcq;
    if (!h->noFine && *pinN != nAcq) return fail(h, GC_ERR_CUDA, "gc_acquire: device and host disagree on the acquired set");
    for (int a = 0; a < nAcq; ++a) {
        const int s = acq[a];
        const int ri = sv_result_index(h, svList[order[s]]);
        const double coarse = coarseFreqOf[s][pinPeaks[s].bin - 1];
        carrFreq[ri] = h->noFine ? coarse                                            // GAL_E5b acquisition.m:203-205
                                 : coarse + c.acq_search_step / 2 - h->fineStep * pinBest[a];   // :227, :254
        codePhase[ri] = pinPeaks[s].codePhase;                                       // :256
        if (!h->noFine && carrFreq[ri] == 0) carrFreq[ri] = 1;                       // :258
    }
    laps.lap("results");
    return GC_OK;
}


static int acquire_impl(gc_handle* h, long long winStart, int32_t nSv, const int32_t* svList,
                        double* carrFreq, double* codePhase, double* peakMetric,
                        int32_t* coarseBin, int32_t* coarseCodePhase, long long longLen = 0 /* length(longSignal), B1C only */,
                        double* dOut = nullptr /* gc_acquire_device: [4][resultLen] on the device */, bool async = false)
{
    const gc_config& c = h->cfg;
    const int N = h->N, L = h->L, nBins = h->nBins, nonCoh = h->nonCoh, nKm = nBins * nonCoh;
    const int codeLen = c.code_length;
    if (!h->rec) return fail(h, GC_ERR_NO_RECORD, "gc_acquire: no record resident");
    if (nSv < 1 || nSv > kMaxSv || !svList || !carrFreq || !codePhase || !peakMetric)
        return fail(h, GC_ERR_ARG, "gc_acquire: bad argument");
    for (int i = 0; i < nSv; ++i)
        if (!sv_ok(h, svList[i])) return fail(h, GC_ERR_ARG, h->glo ? "gc_acquire: frequency number out of range -7..13" : "gc_acquire: PRN out of range");
    {
        const int rc = autofill_codes(h, nSv, svList);           // codes not supplied by the caller: generated on the device
        if (rc != GC_OK) return rc;
    }
    for (int i = 0; i < nSv; ++i)
        if (!sv_has_code(h, svList[i])) return fail(h, GC_ERR_ARG, "gc_acquire: no code set for an SV of the list (gc_set_code)");
    if (!h->replicasReady) {
        cudaSetDevice(c.device);
        int rc = h->varB ? varb_build_replicas(h) : h->varC ? varc_build_replicas(h) : build_replicas(h);
        if (rc != GC_OK) return rc;
    }
    if (h->varB || h->varC) {
        cudaSetDevice(c.device);
        std::vector<int32_t> cb(h->resultLen, 0);
        if (dOut && !coarseBin) coarseBin = cb.data();
        const int rc = h->varB ? acquire_varb(h, winStart, nSv, svList, carrFreq, codePhase, peakMetric, coarseBin, coarseCodePhase)
                               : acquire_varc(h, winStart, longLen, nSv, svList, carrFreq, codePhase, peakMetric, coarseBin, coarseCodePhase);
        if (rc == GC_OK && dOut) {   // variants B and C finish on the host (the reference's running comparison): hand the result vectors over
            const int n = h->resultLen;
            std::vector<double> pack(4 * (size_t)n);
            for (int i = 0; i < n; ++i) { pack[i] = peakMetric[i]; pack[n + i] = codePhase[i]; pack[2 * n + i] = carrFreq[i]; pack[3 * n + i] = (double)coarseBin[i]; }
            GC_CUDA(h, cudaMemcpyAsync(dOut, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            GC_CUDA(h, cudaStreamSynchronize(h->stream));
        }
        return rc;
    }
    // postProcessing.m:86 reads max(42, nonCoh+2) code periods (B3I: max(22, nonCoh+1), BDS/B3I/include/postProcessing.m:86)
    const int nPeriodsAcq = std::max(h->acqMinPeriods, nonCoh + h->acqExtraPeriods);
    const long long recSamples = rec_samples(h);
    if (winStart < 0 || winStart + (long long)nPeriodsAcq * N > recSamples)
        return fail(h, GC_ERR_SHORT_RECORD, "gc_acquire: record shorter than the acquisition window (max(42, acqNonCohTime+2) code periods; B3I max(22, acqNonCohTime+1))");
    cudaSetDevice(c.device);
    cudaStream_t st = h->stream;
    int launches = 0, evn = 0, nRowLaunches = 0;
    float rowsMs = 0, colsMs = 0, fwdMs = 0;
    auto mark = [&]() { cudaEventRecord(h->ev[evn], st); return evn++; };
    std::vector<std::pair<int, int>> rowEv, colEv, fwdEv;
    auto drain_events = [&]() {
        cudaStreamSynchronize(st);
        float ms;
        for (auto& pr : rowEv) { cudaEventElapsedTime(&ms, h->ev[pr.first], h->ev[pr.second]); rowsMs += ms; }
        for (auto& pr : colEv) { cudaEventElapsedTime(&ms, h->ev[pr.first], h->ev[pr.second]); colsMs += ms; }
        for (auto& pr : fwdEv) { cudaEventElapsedTime(&ms, h->ev[pr.first], h->ev[pr.second]); fwdMs += ms; }
        rowEv.clear(); colEv.clear(); fwdEv.clear();
        evn = 2;                                      // events 0 and 1 bracket the whole coarse search
    };

    for (int i = 0; i < h->resultLen; ++i) {
        carrFreq[i] = codePhase[i] = peakMetric[i] = 0;                              // acquisition.m:130-134
        if (coarseBin) coarseBin[i] = 0;
        if (coarseCodePhase) coarseCodePhase[i] = 0;
    }
    // SVs that share a carrier grid share the wiped-off spectra: GPS = one group of all PRNs,
    // GLONASS = one group per frequency number (grid shifted by -freqSpacing*K, GLO acquisition.m:181-182).
    std::vector<int> order(nSv);
    for (int i = 0; i < nSv; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sv_freq_offset(h, svList[a]) < sv_freq_offset(h, svList[b]); });
    std::vector<int> slotReplica(nSv);                // device list slot -> replica spectrum
    for (int s = 0; s < nSv; ++s) slotReplica[s] = sv_replica(h, svList[order[s]]);
    GC_CUDA(h, upload_cached(h, h->prnList, slotReplica, st));
    GC_CUDA(h, h->partMax.reserve((size_t)nSv * nBins * h->parts));
    GC_CUDA(h, h->partIdx.reserve((size_t)nSv * nBins * h->parts));
    GC_CUDA(h, h->peaks.reserve(nSv));
    GC_CUDA(h, h->sigPower.reserve(1));
    // carrier grids: runs of list slots with the same offset
    std::vector<int> groupStart;
    std::vector<int> slotGroup(nSv);
    for (int s = 0; s < nSv; ++s) {
        if (s == 0 || sv_freq_offset(h, svList[order[s]]) != sv_freq_offset(h, svList[order[s - 1]])) groupStart.push_back(s);
        slotGroup[s] = (int)groupStart.size() - 1;
    }
    const int nGroups = (int)groupStart.size();
    groupStart.push_back(nSv);
    // the default path: split correlation stage on a fused plan, every carrier grid in one pass, one enqueue (graph) - acquire_plain
    const bool plain = h->fused && !h->cluster && !h->queue && !h->overlap && !getenv("GC_ACQ_CHUNK_BINS") && !getenv("GC_ACQ_CHUNK_PRNS") &&
                       !getenv("GC_ROWS_VARIANT") && !getenv("GC_ACQ_LEGACY") && (double)nGroups * nKm * L * sizeof(float2) <= kWorkBytes;
    if (plain)
        return acquire_plain(h, winStart, nSv, svList, order, groupStart, slotGroup, carrFreq, codePhase, peakMetric, coarseBin, coarseCodePhase, dOut, async);
    GC_CUDA(h, h->X.reserve((size_t)(h->cluster ? nGroups : 1) * nKm * L));
    GC_CUDA(h, h->dphi.reserve((size_t)nGroups * nBins));
    std::vector<std::vector<double>> coarseFreqOf(nSv);   // per list slot: the bin frequencies it was searched on

    const int e0 = mark();
    GC_CUDA(h, launch_sig_power(rec_of(h), winStart, N, h->sigPower.p, st)); ++launches;   // :151
    mark();                                           // event 1 (re-recorded at the end of the coarse search)
    if (h->cluster) {
        // forward spectra of every carrier grid, then the whole SV x bin grid in one cluster launch
        std::vector<uint64_t> dphi((size_t)nGroups * nBins);
        for (int gi = 0; gi < nGroups; ++gi) {
            const double off = sv_freq_offset(h, svList[order[groupStart[gi]]]);
            std::vector<double> coarseFreq(nBins);
            for (int k = 0; k < nBins; ++k) {
                coarseFreq[k] = (c.IF + off) + c.acq_search_band - c.acq_search_step * k;   // :169 (GLO :181-182)
                dphi[(size_t)gi * nBins + k] = turns_to_fix(coarseFreq[k] * h->ts);
            }
            for (int s = groupStart[gi]; s < groupStart[gi + 1]; ++s) coarseFreqOf[s] = coarseFreq;
        }
        GC_CUDA(h, cudaMemcpyAsync(h->dphi.p, dphi.data(), dphi.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        GC_CUDA(h, upload_cached(h, h->slotGroup, slotGroup, st));
        const int f0 = mark();
        for (int gi = 0; gi < nGroups; ++gi) {
            FwdColsParams fp{};
            fp.rec = rec_of(h); fp.winStart = winStart; fp.N = N; fp.nonCoh = nonCoh; fp.swapIQ = (h->glo && h->fmt != 2 && h->fmt != 3) ? 1 : 0;
            fp.dphi = h->dphi.p + (size_t)gi * nBins; fp.out = h->X.p + (size_t)gi * nKm * L; fp.tw = h->twFused.p;
            GC_CUDA(h, launch_fwd_cols(L, fp, nKm, false, st)); ++launches;
        }
        RowsParams rp{};
        rp.X = h->X.p; rp.nRows = (long long)nGroups * nKm * h->fp.C;
        GC_CUDA(h, launch_fwd_rows(L, rp, st)); ++launches;
        const int f1 = mark();
        fwdEv.push_back({f0, f1});
        CorrParams cp{};
        cp.X = h->X.p; cp.Cc = h->Cc.p; cp.tw = h->twFused.p;
        cp.nonCoh = nonCoh; cp.nBins = nBins; cp.nSlots = nSv; cp.nRep = h->nRep; cp.repStride = 1;
        cp.slotReplica = h->prnList.p; cp.slotGroup = h->slotGroup.p;
        cp.partMax = h->partMax.p; cp.partIdx = h->partIdx.p;
        GC_CUDA(h, launch_corr_cluster(L, cp, st)); ++launches;
        rowEv.push_back({f1, mark()}); ++nRowLaunches;
    }
    for (int g0 = 0; g0 < nSv && !h->cluster;) {
        int g1 = g0 + 1;
        const double off = sv_freq_offset(h, svList[order[g0]]);
        while (g1 < nSv && sv_freq_offset(h, svList[order[g1]]) == off) ++g1;
        // coarse bin frequencies (:169) and their per-sample phase increments
        std::vector<double> coarseFreq(nBins);
        std::vector<uint64_t> dphi(nBins);
        for (int k = 0; k < nBins; ++k) {
            coarseFreq[k] = (c.IF + off) + c.acq_search_band - c.acq_search_step * k;
            dphi[k] = turns_to_fix(coarseFreq[k] * h->ts);
        }
        for (int s = g0; s < g1; ++s) coarseFreqOf[s] = coarseFreq;
        if (evn > kEvents - 12) drain_events();
        GC_CUDA(h, cudaMemcpyAsync(h->dphi.p, dphi.data(), nBins * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        const int f0 = mark();
        if (h->fused) {
            const bool shifted = h->binShift > 0 && h->binQ == 1;   // only bin 0 is transformed (dphi[0]); bin k = its spectrum shifted by k * binShift
            if (shifted) {
                std::vector<int2> map(nBins);
                for (int k = 0; k < nBins; ++k) map[k] = make_int2(0, k * h->binShift);
                GC_CUDA(h, upload_cached(h, h->vbMap, map, st));
            }
            const int fwdRows = shifted ? nonCoh : nKm;
            FwdColsParams fp{};
            fp.rec = rec_of(h); fp.winStart = winStart; fp.N = N; fp.nonCoh = nonCoh; fp.swapIQ = (h->glo && h->fmt != 2 && h->fmt != 3) ? 1 : 0;
            fp.dphi = h->dphi.p; fp.out = h->X.p; fp.tw = h->twFused.p;
            GC_CUDA(h, launch_fwd_cols(L, fp, fwdRows, false, st)); ++launches;
            RowsParams rp{};
            rp.X = h->X.p; rp.nRows = (long long)fwdRows * h->fp.C;
            GC_CUDA(h, launch_fwd_rows(L, rp, st)); ++launches;
            fwdEv.push_back({f0, mark()});
            if (h->queue) {
                // one persistent kernel for the whole SV x bin grid of this carrier grid: W is a ring of nSlots cells in L2
                const int nCells = (g1 - g0) * nBins, Mq = nonCoh * h->nRep;
                int nSlots = 12, lag = 6;
                if (const char* e = getenv("GC_Q_SLOTS")) nSlots = std::max(2, atoi(e));
                if (const char* e = getenv("GC_Q_LAG")) lag = std::max(1, atoi(e));
                lag = std::min(lag, nSlots - 1);
                GC_CUDA(h, h->W.reserve((size_t)nSlots * Mq * L));
                const size_t nCtrl = 2 + (2 + (size_t)h->parts) * nCells;
                GC_CUDA(h, h->qctrl.reserve(nCtrl));
                GC_CUDA(h, cudaMemsetAsync(h->qctrl.p, 0, nCtrl * sizeof(int), st));
                GC_CUDA(h, h->vbMag.reserve((size_t)nSlots * ((Mq + 4) / 5) * L));
                QueueParams qp{};
                qp.X = h->X.p; qp.Cc = h->Cc.p; qp.W = h->W.p; qp.tw = h->twFused.p;
                qp.nonCoh = nonCoh; qp.nBins = nBins; qp.nRep = h->nRep; qp.repStride = 1;
                qp.nPrn = g1 - g0; qp.prnSlot0 = g0; qp.prnList = h->prnList.p;
                qp.nSlots = nSlots; qp.lag = lag;
                qp.partMax = h->partMax.p; qp.partIdx = h->partIdx.p; qp.parts = h->parts; qp.ctrl = h->qctrl.p; qp.partial = h->vbMag.p;
                const int a = mark();
                GC_CUDA(h, launch_corr_queue(L, qp, st)); ++launches;
                rowEv.push_back({a, mark()}); ++nRowLaunches;
                int flag = 0;
                GC_CUDA(h, cudaMemcpyAsync(&flag, h->qctrl.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
                GC_CUDA(h, cudaStreamSynchronize(st));
                if (flag) return fail(h, GC_ERR_CUDA, "gc_acquire: the work-queue correlation kernel gave up waiting (GC_ACQ_PATH=queue)");
                g0 = g1;
                continue;
            }
            // PRN chunks sized so the inverse work buffer stays below kWorkBytes.  With GC_ACQ_OVERLAP the chunks are halved and
            // pipelined over two work buffers: the column pass of chunk c (HBM bound) runs on a second stream while the row
            // pass of chunk c+1 (FP32 bound) runs on the first.
            const bool overlap = h->overlap;
            int chunk = (int)std::max<long long>(1, (long long)((overlap ? kWorkBytes / 2 : kWorkBytes) / ((double)nKm * h->nRep * L * sizeof(float2))));
            if (const char* e = getenv("GC_ACQ_CHUNK_PRNS")) chunk = std::max(1, atoi(e));
            chunk = std::min(chunk, g1 - g0);
            GC_CUDA(h, h->W.reserve((size_t)chunk * nKm * h->nRep * L));
            if (overlap) GC_CUDA(h, h->W2.reserve((size_t)chunk * nKm * h->nRep * L));
            // GC_ACQ_CHUNK_BINS: bins per launch (experiment: PRN x bin chunks small enough for the 126 MB L2; the launches get too
            // short to pay, profiles/r01_l2_overlap_experiments.md)
            int binChunk = nBins;
            if (const char* e = getenv("GC_ACQ_CHUNK_BINS")) binChunk = std::min(nBins, std::max(1, atoi(e)));
            int ci = 0;
            for (int b0 = 0; b0 < nBins; b0 += binChunk)
            for (int s0 = g0; s0 < g1; s0 += chunk, ++ci) {
                const int nb = std::min(binChunk, nBins - b0);
                const int nc = std::min(chunk, g1 - s0);
                float2* Wc = (overlap && (ci & 1)) ? h->W2.p : h->W.p;
                RowsParams ip{};
                ip.X = h->X.p; ip.Cc = h->Cc.p; ip.W = Wc; ip.tw = h->twFused.p;
                ip.nonCoh = nonCoh; ip.nBins = nb; ip.bin0 = b0; ip.prnPerCta = 1; ip.mPerCta = 5;   // 5 warps, 96 registers, 20 warps/SM
                ip.nRep = h->nRep; ip.repStride = 1;
                if (nonCoh * h->nRep < 5) { ip.prnPerCta = 5; ip.mPerCta = 1; }       // few transforms per cell (Galileo E1: 2): fill the CTA with SVs
                if (const char* e = getenv("GC_ROWS_VARIANT")) {   // "PxM" warps per CTA = P PRNs x M blocks
                    int P = 0, M = 0;
                    if (sscanf(e, "%dx%d", &P, &M) == 2 && (P * M == 5 || P * M == 8)) { ip.prnPerCta = P; ip.mPerCta = M; }
                }
                ip.zLoop = rows_zloop(L);
                ip.nPrnChunk = nc; ip.prnSlot0 = s0; ip.prnList = h->prnList.p;
                if (shifted) ip.binMap = h->vbMap.p;
                if (evn > kEvents - 12) drain_events();
                if (overlap && ci >= 2) GC_CUDA(h, cudaStreamWaitEvent(st, h->evCols[ci & 1], 0));   // this buffer's previous columns are done
                const int a = mark();
                GC_CUDA(h, launch_inv_rows(L, ip, st)); ++launches;
                const int b = mark();
                InvColsParams cp{};
                cp.W = Wc; cp.colTw = h->twCols.p; cp.nBins = nb; cp.bin0 = b0; cp.nBinsTotal = nBins; cp.nonCoh = nonCoh * h->nRep; cp.nPrnChunk = nc; cp.prnSlot0 = s0;
                cp.partMax = h->partMax.p; cp.partIdx = h->partIdx.p;
                if (overlap) {
                    if (const char* e = getenv("GC_COLS_PERSIST")) cp.persist = std::max(0, atoi(e)) * 148;   // CTAs per SM of the persistent column pass
                    GC_CUDA(h, cudaEventRecord(h->evRows[ci & 1], st));
                    GC_CUDA(h, cudaStreamWaitEvent(h->stream2, h->evRows[ci & 1], 0));
                    GC_CUDA(h, launch_inv_cols(L, cp, h->stream2)); ++launches;
                    GC_CUDA(h, cudaEventRecord(h->evCols[ci & 1], h->stream2));
                    rowEv.push_back({a, b}); ++nRowLaunches;
                } else {
                    GC_CUDA(h, launch_inv_cols(L, cp, st)); ++launches;
                    const int d = mark();
                    rowEv.push_back({a, b}); colEv.push_back({b, d}); ++nRowLaunches;
                }
            }
            if (overlap)                                       // the peak search waits for every column pass
                for (int i = 0; i < std::min(ci, 2); ++i) GC_CUDA(h, cudaStreamWaitEvent(st, h->evCols[i], 0));
        } else {
            GC_CUDA(h, h->T1.reserve((size_t)std::max(h->nReplicas, nKm) * L));
            GC_CUDA(h, h->T2.reserve((size_t)std::max(h->nReplicas, nKm) * L));
            GC_CUDA(h, launch_generic_wipe(rec_of(h), winStart, N, nonCoh, nBins, (h->glo && h->fmt != 2 && h->fmt != 3) ? 1 : 0, h->dphi.p, h->T1.p, L, st)); ++launches;
            float2 *src = h->T1.p, *dst = h->T2.p;
            int n = L, s = 1;
            for (int f = 0; f < h->plan.nf; ++f) {
                GC_CUDA(h, launch_generic_stage(h->plan, f, n, s, false, src, dst, nKm, st)); ++launches;
                n /= h->plan.fac[f]; s *= h->plan.fac[f];
                std::swap(src, dst);
            }
            GC_CUDA(h, cudaMemcpyAsync(h->X.p, src, (size_t)nKm * L * sizeof(float2), cudaMemcpyDeviceToDevice, st));
            fwdEv.push_back({f0, mark()});
            if (h->nRep > 1) GC_CUDA(h, h->T3.reserve((size_t)nKm * L));
            for (int sl = g0; sl < g1; ++sl) {
                if (evn > kEvents - 12) drain_events();
                const int a = mark();
                const float2* Wr[2] = {nullptr, nullptr};           // abs(ifft(X .* C_r)) inputs, one per replica
                for (int r = 0; r < h->nRep; ++r) {
                    GC_CUDA(h, launch_generic_mul(h->X.p, h->Cc.p + (size_t)(slotReplica[sl] + r) * L, h->T1.p, L, nKm, st)); ++launches;
                    src = h->T1.p; dst = h->T2.p; n = L; s = 1;
                    for (int f = 0; f < h->plan.nf; ++f) {
                        GC_CUDA(h, launch_generic_stage(h->plan, f, n, s, true, src, dst, nKm, st)); ++launches;
                        n /= h->plan.fac[f]; s *= h->plan.fac[f];
                        std::swap(src, dst);
                    }
                    if (r + 1 < h->nRep) {                          // keep the data-replica result while the pilot one is computed
                        GC_CUDA(h, cudaMemcpyAsync(h->T3.p, src, (size_t)nKm * L * sizeof(float2), cudaMemcpyDeviceToDevice, st));
                        Wr[r] = h->T3.p;
                    } else {
                        Wr[r] = src;
                    }
                }
                const int b = mark();
                GC_CUDA(h, launch_generic_absacc(Wr[0], Wr[1], L, nBins, nonCoh, h->parts, h->partMax.p, h->partIdx.p,
                                                 (size_t)sl * nBins * h->parts, st)); ++launches;
                const int d = mark();
                rowEv.push_back({a, b}); colEv.push_back({b, d}); ++nRowLaunches;
            }
        }
        g0 = g1;
    }
    GC_CUDA(h, launch_peak_select(h->partMax.p, h->partIdx.p, nSv, nBins, h->parts, h->peaks.p, st)); ++launches;
    cudaEventRecord(h->ev[1], st);
    // threshold (:200-206) and fine search (:211-253) follow on the device without a host round trip: fine_setup_kernel
    // decides which list slots are above the threshold and prepares their fine-search entries, the fine kernels are
    // launched for the worst case (every slot acquired) and idle blocks leave at once
    const int nCodes = h->fineTwoCodes ? 2 : 1;
    const int nPeriods = h->nFinePeriods;                                            // :146-148 (B3I :131-133)
    const int tabLen = codeLen * h->sub;                                             // chips, or BOC sub-chips, per code period
    int fa = -1, fb = -1;
    if (!h->noFine) {
        std::vector<double> slotFreq0(nSv);
        std::vector<int> slotChipRow(nSv), slotSv(nSv);
        std::vector<int8_t> slotSec;
        if (h->fineCombine == 4) slotSec.resize((size_t)nSv * nPeriods);
        for (int s = 0; s < nSv; ++s) {
            const int sv = svList[order[s]];
            slotFreq0[s] = (c.IF + sv_freq_offset(h, sv)) + c.acq_search_band;       // coarseFreqBin(1), :169 (GLO :181-182)
            slotChipRow[s] = h->glo ? 0 : sv_result_index(h, sv) * 2;
            slotSv[s] = sv;
            if (h->fineCombine == 4) {
                static const int8_t NH20[20] = {1, 1, 1, 1, 1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, -1, -1, -1, 1};   // GPS_L5C acquisition.m:134
                for (int q = 0; q < nPeriods; ++q)
                    slotSec[(size_t)s * nPeriods + q] = (c.signal == GC_SIG_GAL_E5A) ? h->hostCode[2][sv - 1][q]   // generateE5aQ_secondary
                                                                                       : NH20[q % 20];
            }
        }
        GC_CUDA(h, upload_cached(h, h->slotFreq0, slotFreq0, st));
        GC_CUDA(h, upload_cached(h, h->slotChipRow, slotChipRow, st));
        GC_CUDA(h, upload_cached(h, h->slotSv, slotSv, st));
        if (!slotSec.empty()) GC_CUDA(h, upload_cached(h, h->slotSecondary, slotSec, st));
        const int maxEnt = nSv * nCodes;
        GC_CUDA(h, h->metricDev.reserve(nSv));
        GC_CUDA(h, h->nAcqDev.reserve(1));
        GC_CUDA(h, h->acqSlot.reserve(nSv));
        GC_CUDA(h, h->fineChipRow.reserve(maxEnt));
        GC_CUDA(h, h->fineCodePhase.reserve(maxEnt));
        GC_CUDA(h, h->fdphi.reserve((size_t)maxEnt * h->nFine));
        GC_CUDA(h, h->fineSv.reserve(nSv));
        GC_CUDA(h, h->fineSecondary.reserve((size_t)nSv * nPeriods));
        GC_CUDA(h, h->fineProd.reserve((size_t)maxEnt * nPeriods * N));
        GC_CUDA(h, h->fineSums.reserve((size_t)maxEnt * h->nFine * nPeriods * 2));
        GC_CUDA(h, h->fineResult.reserve((size_t)nSv * h->nFine));
        GC_CUDA(h, h->fineBest.reserve(nSv));
        FineSetup fs{};
        fs.peaks = h->peaks.p; fs.sigPower = h->sigPower.p; fs.nSv = nSv; fs.nonCoh = nonCoh; fs.nFine = h->nFine; fs.nCodes = nCodes;
        fs.nPeriods = nPeriods; fs.pilotComp = h->nRep - 1; fs.threshold = c.acq_threshold; fs.step = c.acq_search_step;
        fs.fineStep = h->fineStep; fs.ts = h->ts; fs.slotFreq0 = h->slotFreq0.p; fs.slotChipRow = h->slotChipRow.p; fs.slotSv = h->slotSv.p;
        fs.slotSecondary = slotSec.empty() ? nullptr : h->slotSecondary.p;
        fs.metric = h->metricDev.p; fs.nAcq = h->nAcqDev.p; fs.acqSlot = h->acqSlot.p; fs.chipRow = h->fineChipRow.p;
        fs.codePhase = h->fineCodePhase.p; fs.dphi = h->fdphi.p; fs.svId = h->fineSv.p; fs.secondary = h->fineSecondary.p;
        fa = mark();
        GC_CUDA(h, launch_fine_setup(fs, st)); ++launches;
        FineParams fp{};
        fp.rec = rec_of(h); fp.winStart = winStart; fp.N = N; fp.nPeriods = nPeriods; fp.nFine = h->nFine; fp.codeLen = tabLen;
        fp.swapIQ = (h->glo && h->fmt != 2 && h->fmt != 3) ? 1 : 0; fp.combine = h->fineCombine; fp.chipIdx = h->chipIdx.p; fp.svId = h->fineSv.p;
        fp.chips = h->chips.p; fp.chipRow = h->fineChipRow.p; fp.codePhase = h->fineCodePhase.p; fp.dphi = h->fdphi.p; fp.prod = h->fineProd.p;
        fp.sums = h->fineSums.p; fp.best = h->fineBest.p; fp.fineResult = h->fineResult.p;
        fp.nAcqDev = h->nAcqDev.p; fp.nCodes = nCodes; fp.secondary = h->fineSecondary.p;
        fp.moments = fine_moments(h->nFine, h->fineStep, h->ts);
        GC_CUDA(h, launch_fine(fp, maxEnt, nSv, st)); launches += 3;
        fb = mark();
    }
    if (dOut) {                                                   // acqResults assembled on the device for the collective that follows
        std::vector<int> slotResult(nSv);
        std::vector<double> slotFreq0(nSv);
        for (int s = 0; s < nSv; ++s) {
            slotResult[s] = sv_result_index(h, svList[order[s]]);
            slotFreq0[s] = (c.IF + sv_freq_offset(h, svList[order[s]])) + c.acq_search_band;
        }
        GC_CUDA(h, upload_cached(h, h->slotResult, slotResult, st));
        if (h->noFine) GC_CUDA(h, upload_cached(h, h->slotFreq0, slotFreq0, st));
        PackParams pk{};
        pk.peaks = h->peaks.p; pk.sigPower = h->sigPower.p; pk.slotFreq0 = h->slotFreq0.p; pk.slotResult = h->slotResult.p;
        pk.best = h->noFine ? nullptr : h->fineBest.p; pk.nSv = nSv; pk.nonCoh = nonCoh; pk.resultLen = h->resultLen; pk.noFine = h->noFine ? 1 : 0;
        pk.threshold = c.acq_threshold; pk.step = c.acq_search_step; pk.fineStep = h->fineStep; pk.out = dOut;
        GC_CUDA(h, launch_pack_results(pk, st)); ++launches;
        // the results stay on the device: no D2H of the peak / fine-search arrays and no host assembly - one synchronise, the timings
        const int fa_ = fa, fb_ = fb;
        float fineMs = 0, coarseMs = 0;
        GC_CUDA(h, cudaStreamSynchronize(st));
        if (fa_ >= 0) cudaEventElapsedTime(&fineMs, h->ev[fa_], h->ev[fb_]);
        drain_events();
        cudaEventElapsedTime(&coarseMs, h->ev[e0], h->ev[1]);
        h->stats.n_acquired = -1;                             // (in the device buffer: carrFreq != 0)
        h->stats.acq_fwd_ms = fwdMs; h->stats.acq_corr_ms = coarseMs - fwdMs; h->stats.acq_fine_ms = fineMs; h->stats.acq_total_ms = coarseMs + fineMs;
        h->stats.corr_rows_ms = rowsMs; h->stats.corr_cols_ms = colsMs; h->stats.corr_row_launches = nRowLaunches; h->stats.acq_launches = launches;
        return GC_OK;
    }
    std::vector<PeakOut> peaks(nSv);
    std::vector<int> best(nSv, 0);
    double sigPower = 0;
    int nAcqDevice = -1;
    GC_CUDA(h, cudaMemcpyAsync(peaks.data(), h->peaks.p, nSv * sizeof(PeakOut), cudaMemcpyDeviceToHost, st));
    GC_CUDA(h, cudaMemcpyAsync(&sigPower, h->sigPower.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!h->noFine) {
        GC_CUDA(h, cudaMemcpyAsync(best.data(), h->fineBest.p, nSv * sizeof(int), cudaMemcpyDeviceToHost, st));
        GC_CUDA(h, cudaMemcpyAsync(&nAcqDevice, h->nAcqDev.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    float fineMs = 0;
    {
        const int fa_ = fa, fb_ = fb;
        cudaStreamSynchronize(st);
        if (fa_ >= 0) cudaEventElapsedTime(&fineMs, h->ev[fa_], h->ev[fb_]);
    }
    drain_events();
    float coarseMs = 0;
    cudaEventElapsedTime(&coarseMs, h->ev[e0], h->ev[1]);

    std::vector<int> acq;   // list slots above the threshold, in list order (what fine_setup_kernel found too)
    for (int s = 0; s < nSv; ++s) {
        const int ri = sv_result_index(h, svList[order[s]]);
        peakMetric[ri] = peaks[s].peak / sigPower / nonCoh;                          // :200
        if (coarseBin) coarseBin[ri] = peaks[s].bin;
        if (coarseCodePhase) coarseCodePhase[ri] = peaks[s].codePhase;
        if (peakMetric[ri] > c.acq_threshold) acq.push_back(s);                      // :206
    }
    const int nAcq = (int)acq.size();
    h->stats.n_acquired = nAcq;
    if (!h->noFine && nAcqDevice != nAcq) return fail(h, GC_ERR_CUDA, "gc_acquire: device and host disagree on the acquired set");
    for (int a = 0; a < nAcq; ++a) {
        const int s = acq[a];
        const int ri = sv_result_index(h, svList[order[s]]);
        const double coarse = coarseFreqOf[s][peaks[s].bin - 1];
        carrFreq[ri] = h->noFine ? coarse                                            // GAL_E5b acquisition.m:203-205
                                 : coarse + c.acq_search_step / 2 - h->fineStep * best[a];   // :227, :254
        codePhase[ri] = peaks[s].codePhase;                                          // :256
        if (!h->noFine && carrFreq[ri] == 0) carrFreq[ri] = 1;                       // :258
    }
    h->stats.acq_fwd_ms = fwdMs;
    h->stats.acq_corr_ms = coarseMs - fwdMs;
    h->stats.acq_fine_ms = fineMs;
    h->stats.acq_total_ms = coarseMs + fineMs;
    h->stats.corr_rows_ms = rowsMs;
    h->stats.corr_cols_ms = colsMs;
    h->stats.corr_row_launches = nRowLaunches;
    h->stats.acq_launches = launches;
    return GC_OK;
}

int gc_acquire(gc_handle* h, int32_t nSv, const int32_t* svList,
               double* carrFreq, double* codePhase, double* peakMetric,
               int32_t* coarseBin, int32_t* coarseCodePhase)
{
    if (!h) return GC_ERR_ARG;
    // fseek(fid, dataAdaptCoeff*skipNumberOfBytes) (postProcessing.m:74): skip counts complex samples
    return acquire_impl(h, skip_samples(h), nSv, svList, carrFreq, codePhase, peakMetric,
                        coarseBin, coarseCodePhase);
}

int gc_acquire_device(gc_handle* h, int32_t nSv, const int32_t* svList, double* dResults)
{
    if (!h) return GC_ERR_ARG;
    if (!dResults) return fail(h, GC_ERR_ARG, "gc_acquire_device: null result buffer");
    std::vector<double> cf(h->resultLen), cp(h->resultLen), pm(h->resultLen);
    return acquire_impl(h, skip_samples(h), nSv, svList, cf.data(), cp.data(), pm.data(), nullptr, nullptr, 0, dResults);
}

int gc_acquire_device_async(gc_handle* h, int32_t nSv, const int32_t* svList, double* dResults)
{
    if (!h) return GC_ERR_ARG;
    if (!dResults) return fail(h, GC_ERR_ARG, "gc_acquire_device_async: null result buffer");
    std::vector<double> cf(h->resultLen), cp(h->resultLen), pm(h->resultLen);
    return acquire_impl(h, skip_samples(h), nSv, svList, cf.data(), cp.data(), pm.data(), nullptr, nullptr, 0, dResults, true);
}

int gc_acquire_host(gc_handle* h, const int8_t* iq, size_t nSamples, int32_t nSv, const int32_t* svList,
                    double* carrFreq, double* codePhase, double* peakMetric,
                    int32_t* coarseBin, int32_t* coarseCodePhase)
{
    if (!h || !iq) return fail(h, GC_ERR_ARG, "gc_acquire_host: bad argument");
    int rc = set_record_host(h, iq, (size_t)rec_of(h).bytes_of((long long)nSamples), false);
    if (rc != GC_OK) return rc;
    rc = acquire_impl(h, 0, nSv, svList, carrFreq, codePhase, peakMetric, coarseBin, coarseCodePhase, (long long)nSamples);
    if (rc != GC_OK) cudaStreamSynchronize(h->stream);        // (an early return may have left the copy of `iq` in flight)
    return rc;
}

int gc_track_nfields(const gc_handle* h)
{
    if (!h) return GC_TRACK_NFIELDS;
    return h->pilotMode >= 4 ? GC_TRACK_NFIELDS_PILOT6 : h->pilotMode >= 2 ? GC_TRACK_NFIELDS_PILOT : GC_TRACK_NFIELDS;
}

int gc_set_param(gc_handle* h, int32_t key, double value)
{
    if (!h) return GC_ERR_ARG;
    if (key == GC_PARAM_B1C_WB_FACTOR) {
        if (!(value >= 0.0 && value <= 1.0)) return fail(h, GC_ERR_ARG, "gc_set_param: the B1C weighting factor lies in [0, 1]");
        h->wbFactor = value;
        return GC_OK;
    }
    if (key == GC_PARAM_TRACK_EXACT_SUMS) { h->trackExact = value != 0.0; return GC_OK; }
    if (key == GC_PARAM_TRACK_FAST_DISC) { h->trackFastDisc = value != 0.0; return GC_OK; }
    return fail(h, GC_ERR_ARG, "gc_set_param: unknown key");
}

int gc_get_cl_code_phase(const gc_handle* h, int32_t* clCodePhase)
{
    if (!h || !clCodePhase || h->cfg.signal != GC_SIG_GPS_L2C) return GC_ERR_ARG;
    std::copy(h->clPhaseOut, h->clPhaseOut + 32, clCodePhase);
    return GC_OK;
}

int gc_set_cl_code_phase(gc_handle* h, int32_t nCh, const int32_t* clCodePhase)
{
    if (!h) return GC_ERR_ARG;
    if (h->cfg.signal != GC_SIG_GPS_L2C || nCh < 1 || !clCodePhase) return fail(h, GC_ERR_ARG, "gc_set_cl_code_phase: GPS L2C only, nCh >= 1");
    h->clPhaseIn.assign(clCodePhase, clCodePhase + nCh);
    return GC_OK;
}

int gc_track(gc_handle* h, int32_t nCh, const int32_t* sv, const double* acqFreq, const double* codePhase,
             const double* codeFreq0, int32_t nEpochs, double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone)
{
    if (!h) return GC_ERR_ARG;
    const gc_config& c = h->cfg;
    if (!h->rec) return fail(h, GC_ERR_NO_RECORD, "gc_track: no record resident");
    if (c.signal == GC_SIG_BDS_B1C && c.pilot_trk_flag != 1 && c.pilot_trk_flag != 2)
        return fail(h, GC_ERR_UNSUPPORTED, "gc_track: BDS B1C tracking needs pilotTRKflag 1 (NB_tracking.m) or 2 (WB_tracking.m), B1C postProcessing.m:34-38");
    const bool l2c = c.signal == GC_SIG_GPS_L2C;
    if (h->pilotMode == 5 && h->wbFactor < 0)
        return fail(h, GC_ERR_ARG, "gc_track: B1C full-band tracking needs gc_set_param(GC_PARAM_B1C_WB_FACTOR) first");
    if (h->pilotMode == 4 && (int)h->clPhaseIn.size() != nCh)
        return fail(h, GC_ERR_ARG, "gc_track: GPS L2C CL pilot needs gc_set_cl_code_phase for these channels first");
    if (nCh < 1 || nEpochs < 1 || !sv || !acqFreq || !codePhase || !out || !epochsDone)
        return fail(h, GC_ERR_ARG, "gc_track: bad argument");
    cudaSetDevice(c.device);
    cudaStream_t st = h->stream;
    {
        std::vector<int32_t> need;
        for (int ch = 0; ch < nCh; ++ch)
            if (sv[ch] >= 1 && sv[ch] <= h->resultLen) need.push_back(sv[ch]);
        const int rc = autofill_codes(h, (int)need.size(), need.data());
        if (rc != GC_OK) return rc;
    }
    // table entries per code period (BOC: sub-chips).  GPS L2C works in half chips throughout: codeLength*2 entries of the
    // return-to-zero CM code, code NCO at 2*codeFreqBasis, spacing*2 (GPS_L2C/include/tracking.m:93-94, 171)
    const int codeLen = c.code_length * h->sub * (c.signal == GC_SIG_GPS_L2C ? 2 : 1);
    const int stride = (codeLen + 2 + 15) & ~15;
    const bool pilot = h->pilotMode != 0;                     // GAL_E1C tracking.m:127; GPS_L5C tracking.m:167
    std::vector<TrackChan> chans(nCh);
    const int clLen = 150 * c.code_length;                   // padded CL sequence [CL(end) CL CL(1)] per channel (GPS_L2C tracking.m:164-166)
    const int pstride = h->pilotMode == 4 ? ((clLen + 2 + 15) & ~15) : stride;
    const int p61stride = (codeLen * 6 + 2 + 15) & ~15;
    std::vector<int8_t> tabs((size_t)nCh * stride, 0), ptabs(pilot ? (size_t)nCh * pstride : 0, 0);
    std::vector<int8_t> p61(h->pilotMode == 5 ? (size_t)nCh * p61stride : 0, 0);
    std::vector<char> live(nCh, 0);
    for (int ch = 0; ch < nCh; ++ch) {
        const bool active = h->glo ? (sv[ch] != GC_SV_NONE) : (sv[ch] != 0);   // tracking.m:136; GLO tracking.m:137
        live[ch] = active;
        chans[ch].prn = sv[ch]; chans[ch].pad = active ? 1 : 0;
        chans[ch].acqFreq = acqFreq[ch];
        chans[ch].codeFreq0 = codeFreq0 ? codeFreq0[ch] : c.code_freq_basis * (l2c ? 2 : 1);   // channel.codeFreq (B3I tracking.m:57)
        // fseek(fid, dataAdaptCoeff*(skipNumberOfBytes + codePhase-1)) (tracking.m:150); L2C seeks to codePhase (GPS_L2C tracking.m:153)
        chans[ch].startSample = skip_samples(h) + (long long)codePhase[ch] - (l2c ? 0 : 1);
        if (active) {
            if (!sv_ok(h, sv[ch])) return fail(h, GC_ERR_ARG, "gc_track: SV id out of range");
            if (!sv_has_code(h, sv[ch])) return fail(h, GC_ERR_ARG, "gc_track: no code set for a channel's SV (gc_set_code)");
            if (chans[ch].startSample < 0) return fail(h, GC_ERR_ARG, "gc_track: codePhase must be >= 1");
            int8_t* t = tabs.data() + (size_t)ch * stride;
            sv_chips(h, sv[ch], t + 1);                    // tracking.m:156 (GLO tracking.m:88)
            t[0] = t[codeLen]; t[codeLen + 1] = t[1];      // [c(L) c c(1)]  :158
            if (h->pilotMode == 4) {
                const std::vector<int8_t>& cl = h->hostCode[1][sv[ch] - 1];
                if ((int)cl.size() != clLen) return fail(h, GC_ERR_ARG, "gc_track: no CL code set for a channel's SV (gc_set_code component 1)");
                if (h->clPhaseIn[ch] < 1 || h->clPhaseIn[ch] > 75) return fail(h, GC_ERR_ARG, "gc_track: CLCodePhase must be 1..75");
                chans[ch].clPhase = h->clPhaseIn[ch];
                int8_t* u = ptabs.data() + (size_t)ch * pstride;
                std::copy(cl.begin(), cl.end(), u + 1);
                u[0] = u[clLen]; u[clLen + 1] = u[1];
            } else if (pilot) {                            // GAL_E1C tracking.m:127-130
                if (h->hostCode[1][sv[ch] - 1].empty()) return fail(h, GC_ERR_ARG, "gc_track: no pilot code set for a channel's SV (gc_set_code component 1)");
                int8_t* u = ptabs.data() + (size_t)ch * stride;
                sv_chips(h, sv[ch], u + 1, 1);
                u[0] = u[codeLen]; u[codeLen + 1] = u[1];
            }
            if (h->pilotMode == 5) {                       // B1C WB_tracking.m:181-183
                const std::vector<int8_t>& b = h->hostCode[2][sv[ch] - 1];
                if ((int)b.size() != codeLen * 6) return fail(h, GC_ERR_ARG, "gc_track: no pilot BOC(6,1) code set for a channel's SV (gc_set_code component 2)");
                int8_t* u = p61.data() + (size_t)ch * p61stride;
                std::copy(b.begin(), b.end(), u + 1);
                u[0] = u[codeLen * 6]; u[codeLen * 6 + 1] = u[1];
            }
        }
    }
    TrackParams p{};
    p.rec = h->rec; p.fmt = h->fmt;
    p.recSamples = rec_samples(h);
    p.fs = c.sampling_freq; p.invFs = 1.0 / c.sampling_freq; p.codeFreqBasis = c.code_freq_basis * (l2c ? 2 : 1); p.codeLength = (double)c.code_length * (l2c ? 2 : 1);
    p.spc = c.dll_correlator_spacing * (l2c ? 2 : 1);
    p.cA = h->tau2code / h->tau1code; p.cB = c.int_time / h->tau1code;     // tracking.m:326
    p.pA = h->tau2carr / h->tau1carr; p.pB = c.int_time / h->tau1carr;     // tracking.m:308
    {   // Common/calcLoopCoefCarr.m:41-56 (GLONASS carrier filter)
        const double Wn = 1.2 * c.pll_noise_bandwidth;
        p.pf3 = std::pow(Wn, 3) * std::pow(c.int_time, 2);
        p.pf2 = 2 * std::pow(Wn, 2) * c.int_time;
        p.pf1 = 2 * Wn;
    }
    p.loopType = (h->glo || h->b3i || h->hostCodes) ? 1 : 0;
    p.swapIQ = (h->glo && h->fmt != 2 && h->fmt != 3) ? 1 : 0;
    p.nEpochs = nEpochs;
    // discriminators: float64 atan / sqrt / divide as the reference evaluates them (measured cost: +1.5 % at 12 channels, +1.7 % at
    // 592, profiles/r02_ncu_track_baseline.md); GC_PARAM_TRACK_FAST_DISC selects the fp32 forms
    p.exactDisc = h->trackFastDisc ? 0 : 1;
    p.exact = h->trackExact ? 1 : 0;                          // the float64 checking mode (GC_PARAM_TRACK_EXACT_SUMS)
    if (p.exact) p.exactDisc = 1;
    // CTAs per channel: spread few channels over the chip (thread-block clusters), 1 CTA per channel once
    // the channel count fills it
    int nLive = 0;
    for (int ch = 0; ch < nCh; ++ch) nLive += (sv[ch] != 0);
    int cluster = 1;
    for (int g = 8; g >= 2; g /= 2)
        if (nCh * g <= 148) { cluster = g; break; }   // (every B200 has 148 SMs; `sms` below is read from the device for the batch split)
    if (const char* e = getenv("GC_TRACK_CLUSTER")) { const int g = atoi(e); if (g == 1 || g == 2 || g == 4 || g == 8) cluster = g; }
    (void)nLive;
    p.codeLen = codeLen; p.codeStride = stride; p.pilotStride = pstride; p.p61Stride = p61stride; p.subChip = h->sub; p.pilot = h->pilotMode;
    p.wbFactor = h->wbFactor;
    if (h->pilotMode == 5) cluster = 8;                       // three tables + the window only fit with the block spread over 8 CTAs
    p.nRows = gc_track_nfields(h);
    // long code periods (Galileo E1: 4 ms = 72000+ samples): spread the block over enough CTAs for the
    // double-buffered window to fit in shared memory
    while (cluster < 8 && track_smem_bytes(track_buf_bytes(h->N + 64, cluster), codeLen, p.pilot) > 227 * 1024) cluster *= 2;
    p.bufBytes = track_buf_bytes(h->N + 64, cluster);
    p.singleBuf = track_smem_bytes(p.bufBytes, codeLen, p.pilot) > 227 * 1024 ? 1 : 0;   // B1C: 45 KB windows + two 82 KB tables
    if (track_smem_bytes(p.bufBytes, codeLen, p.pilot, p.singleBuf) > 227 * 1024)
        return fail(h, GC_ERR_UNSUPPORTED, "gc_track: one code period of samples does not fit in shared memory");
    // more channels than SMs: 256-thread CTAs, two per SM (each needs half of the shared memory)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
    int batch = (cluster == 1 && nCh > sms && h->pilotMode != 5 && !getenv("GC_TRACK_NO_BATCH")) ? 1 : 0;
    if (batch && 2 * (track_smem_bytes(p.bufBytes, codeLen, p.pilot, p.singleBuf, 0) + 1024) > 227 * 1024) batch = 0;
    {
        const size_t withSlots = track_smem_bytes(p.bufBytes, codeLen, p.pilot, p.singleBuf, track_threads(cluster, batch));
        p.preSlots = (h->fmt == 0 && !p.exact && (batch ? 2 * (withSlots + 1024) : withSlots) <= 227 * 1024) ? 1 : 0;
    }
    GC_CUDA(h, upload(h->chans, chans, st));
    GC_CUDA(h, upload(h->trackCodes, tabs, st));
    if (pilot) GC_CUDA(h, upload(h->trackPilot, ptabs, st));
    if (h->pilotMode == 5) GC_CUDA(h, upload(h->trackP61, p61, st));
    const int nRows = gc_track_nfields(h);
    const size_t nOut = (size_t)nCh * nRows * nEpochs;
    GC_CUDA(h, h->trackOut.reserve(nOut));
    GC_CUDA(h, h->epochsDone.reserve(nCh));
    p.codeTables = h->trackCodes.p; p.pilotTables = h->trackPilot.p; p.p61Tables = h->trackP61.p; p.chans = h->chans.p; p.out = h->trackOut.p; p.epochsDone = h->epochsDone.p;
    long long* dbg = nullptr;
    if (getenv("GC_TRACK_DEBUG")) { cudaMalloc(&dbg, 96 * sizeof(long long)); cudaMemset(dbg, 0, 96 * sizeof(long long)); }
    p.dbg = dbg;
    GC_CUDA(h, launch_track_fill(h->trackOut.p, nCh, nRows, nEpochs, st));
    cudaEventRecord(h->ev[0], st);
    GC_CUDA(h, launch_track(p, nCh, cluster, batch, st));
    cudaEventRecord(h->ev[1], st);
    const int vint = c.cno_vsm_interval, nV = nEpochs / vint;
    const bool wantVsm = vsmValue && vsmIndex && nV > 0;
    if (wantVsm) {                                               // C/N0 on the device, straight from the rows the kernel just wrote
        GC_CUDA(h, h->vsmDev.reserve(2 * (size_t)nCh * nV));
        GC_CUDA(h, launch_cno_vsm(h->trackOut.p, nCh, nRows, nEpochs, vint, c.cno_acc_time, h->epochsDone.p, h->vsmDev.p,
                                  h->vsmDev.p + (size_t)nCh * nV, st));
        GC_CUDA(h, cudaMemcpyAsync(vsmValue, h->vsmDev.p, (size_t)nCh * nV * sizeof(double), cudaMemcpyDeviceToHost, st));
        GC_CUDA(h, cudaMemcpyAsync(vsmIndex, h->vsmDev.p + (size_t)nCh * nV, (size_t)nCh * nV * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    const bool pld = (c.signal == GC_SIG_BDS_B2A || c.signal == GC_SIG_BDS_B1C) && nV > 0;
    h->cnoPldCh = h->cnoPldV = 0;
    if (pld) {   // Calc_CNo_PLD.m on the device (BDS/B2a/include/tracking.m:409-431); pilot rows: B2a / B1C narrow band swap roles, B1C full band as recorded
        const int pmode = h->pilotMode == 0 ? 0 : h->pilotMode == 5 ? 2 : 1;
        GC_CUDA(h, h->cnoPldDev.reserve((size_t)nCh * GC_CNO_PLD_ROWS * nV));
        GC_CUDA(h, launch_cno_pld(h->trackOut.p, nCh, nRows, nEpochs, vint, c.int_time, pmode, h->epochsDone.p, h->cnoPldDev.p, st));
        h->cnoPld.assign((size_t)nCh * GC_CNO_PLD_ROWS * nV, 0.0);
        GC_CUDA(h, cudaMemcpyAsync(h->cnoPld.data(), h->cnoPldDev.p, h->cnoPld.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        h->cnoPldCh = nCh; h->cnoPldV = nV;
    }
    GC_CUDA(h, cudaMemcpyAsync(out, h->trackOut.p, nOut * sizeof(double), cudaMemcpyDeviceToHost, st));
    GC_CUDA(h, cudaMemcpyAsync(epochsDone, h->epochsDone.p, nCh * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GC_CUDA(h, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    h->stats.track_kernel_ms = ms;
    h->stats.track_launches = 2;
    if (dbg) {
        long long hd[96];
        cudaMemcpy(hd, dbg, sizeof(hd), cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        const char* names[8] = {"top", "mbar", "samples", "sync1", "cluster", "control", "sync2", "xwait"};
        for (int w = 0; w < 4; ++w) {
            if (w == 2) continue;
            fprintf(stderr, "[gc_track timing] warp %d cycles/epoch:", w);
            for (int i = 0; i < 8; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)hd[w * 8 + i] / nEpochs);
            fprintf(stderr, "  (cluster=%d)\n", cluster);
        }
        for (int r = 0; r < cluster && cluster > 1; ++r) {
            fprintf(stderr, "[gc_track timing] rank %d warp0 cycles/epoch:", r);
            for (int i = 0; i < 7; ++i) fprintf(stderr, " %s=%.0f", names[i], (double)hd[32 + r * 8 + i] / nEpochs);
            fprintf(stderr, "\n");
        }
    }

    // A short read makes the reference `return` from tracking() (tracking.m:241-245): channels
    // after the first one that ran out of data stay as initialised.
    int failed = -1;
    for (int ch = 0; ch < nCh && failed < 0; ++ch)
        if (live[ch] && epochsDone[ch] < nEpochs) failed = ch;
    const double inf = std::numeric_limits<double>::infinity();
    for (int ch = failed + 1; failed >= 0 && ch < nCh; ++ch) {
        double* o = out + (size_t)ch * nRows * nEpochs;
        for (int f = 0; f < nRows; ++f) {
            const double fill = (f == GC_F_ABSOLUTE_SAMPLE || (f >= GC_F_I_P && f <= GC_F_Q_L) || f >= GC_TRACK_NFIELDS) ? 0.0 : inf;
            std::fill(o + (size_t)f * nEpochs, o + (size_t)(f + 1) * nEpochs, fill);
        }
        epochsDone[ch] = 0;
    }
    if (pld)
        for (int ch = 0; ch < nCh; ++ch)
            if (!live[ch] || (failed >= 0 && ch > failed))
                std::fill(h->cnoPld.begin() + (size_t)ch * GC_CNO_PLD_ROWS * nV, h->cnoPld.begin() + (size_t)(ch + 1) * GC_CNO_PLD_ROWS * nV, 0.0);
    // C/N0 (tracking.m:351-358) came from the device; channels after one that ran out of data stay as initialised
    if (wantVsm)
        for (int ch = 0; ch < nCh; ++ch)
            if (!live[ch] || (failed >= 0 && ch > failed)) {
                std::fill(vsmValue + (size_t)ch * nV, vsmValue + (size_t)(ch + 1) * nV, 0.0);
                std::fill(vsmIndex + (size_t)ch * nV, vsmIndex + (size_t)(ch + 1) * nV, 0.0);
            }
    return GC_OK;
}

int gc_get_cno_pld(const gc_handle* h, int32_t nCh, int32_t nIntervals, double* out)
{
    if (!h) return GC_ERR_ARG;
    if (nIntervals == 0) return GC_OK;                         // a run shorter than one CNoInterval: zeros(1, 0) in the reference (NB_tracking.m:88-92)
    gc_handle* hm = const_cast<gc_handle*>(h);
    if (!out) return fail(hm, GC_ERR_ARG, "gc_get_cno_pld: null output");
    if (h->cnoPldCh == 0 || nCh != h->cnoPldCh || nIntervals != h->cnoPldV)
        return fail(hm, GC_ERR_ARG, "gc_get_cno_pld: no BDS B2a / B1C gc_track of that shape (channels, intervals) before this call");
    std::copy(h->cnoPld.begin(), h->cnoPld.end(), out);
    return GC_OK;
}

int gc_acquire_track(gc_handle* h, int32_t nSv, const int32_t* svList, int32_t nChannels, int32_t nEpochs,
                     double* carrFreq, double* codePhase, double* peakMetric,
                     int32_t* chanSv, double* chanAcqFreq, double* chanCodePhase,
                     double* out, double* vsmValue, double* vsmIndex, int32_t* epochsDone)
{
    if (!h) return GC_ERR_ARG;
    const gc_config& c = h->cfg;
    if (nChannels < 1 || !chanSv || !chanAcqFreq || !chanCodePhase) return fail(h, GC_ERR_ARG, "gc_acquire_track: bad argument");
    const bool aided = h->b3i || h->fam5 || h->varC;           // channel.codeFreq from settings.carrFreqBasis (GPS_L5C preRun.m:69-71)
    if (aided && !(c.carr_freq_basis > 0)) return fail(h, GC_ERR_ARG, "gc_acquire_track: this signal's code NCO is carrier aided - set cfg.carr_freq_basis (settings.carrFreqBasis)");
    int rc = gc_acquire(h, nSv, svList, carrFreq, codePhase, peakMetric, nullptr, nullptr);
    if (rc != GC_OK) return rc;
    // preRun.m:44-72: [~, PRNindexes] = sort(peakMetric, 'descend') (stable), the first min(numberOfChannels, #acquired) of them
    const int n = h->resultLen;
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return peakMetric[a] > peakMetric[b]; });
    int nAcq = 0;
    for (int i = 0; i < n; ++i) nAcq += (carrFreq[i] != 0);
    std::vector<double> codeFreq0(nChannels, 0.0);
    std::vector<int32_t> cl(nChannels, 1);
    for (int ch = 0; ch < nChannels; ++ch) {
        const bool on = ch < std::min(nChannels, nAcq);
        chanSv[ch] = on ? (h->glo ? order[ch] - 7 : order[ch] + 1) : (h->glo ? GC_SV_NONE : 0);   // GLO preRun.m: Kindexes(ii) - 8
        chanAcqFreq[ch] = on ? carrFreq[order[ch]] : 0.0;
        chanCodePhase[ch] = on ? codePhase[order[ch]] : 0.0;
        if (on && aided)
            codeFreq0[ch] = c.code_freq_basis + (chanAcqFreq[ch] - c.IF) / c.carr_freq_basis * c.code_freq_basis;
        if (on && h->pilotMode == 4) cl[ch] = h->clPhaseOut[order[ch]];                            // GPS_L2C preRun.m: channel.CLCodePhase
    }
    if (h->pilotMode == 4) {
        rc = gc_set_cl_code_phase(h, nChannels, cl.data());
        if (rc != GC_OK) return rc;
    }
    return gc_track(h, nChannels, chanSv, chanAcqFreq, chanCodePhase, aided ? codeFreq0.data() : nullptr, nEpochs, out, vsmValue, vsmIndex, epochsDone);
}

int gc_track_file(gc_handle* h, const char* path, int32_t nCh, const int32_t* sv, const double* acqFreq,
                  const double* codePhase, const double* codeFreq0, int32_t nEpochs, double* out, double* vsmValue,
                  double* vsmIndex, int32_t* epochsDone)
{
    if (!h || !path) return fail(h, GC_ERR_ARG, "gc_track_file: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(h, GC_ERR_IO, std::string("gc_track_file: unable to read file ") + path);   // postProcessing.m:155-158
    fseek(f, 0, SEEK_END);
    const long long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz <= 0) { fclose(f); return fail(h, GC_ERR_IO, "gc_track_file: empty file"); }
    void* pinned = nullptr;
    cudaSetDevice(h->cfg.device);
    if (cudaHostAlloc(&pinned, (size_t)sz, cudaHostAllocDefault) != cudaSuccess) { fclose(f); return fail(h, GC_ERR_CUDA, "gc_track_file: cudaHostAlloc failed"); }
    const size_t got = fread(pinned, 1, (size_t)sz, f);
    fclose(f);
    int rc = (got == (size_t)sz) ? gc_set_record_host(h, pinned, (size_t)sz) : fail(h, GC_ERR_IO, "gc_track_file: short read");
    cudaFreeHost(pinned);
    if (rc != GC_OK) return rc;
    return gc_track(h, nCh, sv, acqFreq, codePhase, codeFreq0, nEpochs, out, vsmValue, vsmIndex, epochsDone);
}

int gc_nav_sync(gc_handle* h, int32_t nCh, int32_t nEpochs, const double* I_P, int32_t* subFrameStart, uint8_t* navBits, int32_t* bitsValid)
{
    if (!h) return GC_ERR_ARG;
    if (h->cfg.signal != GC_SIG_GPS_L1CA) return fail(h, GC_ERR_UNSUPPORTED, "gc_nav_sync: GPS L1 C/A only");
    if (nCh < 1 || nEpochs < 1 || !I_P || !subFrameStart || !navBits || !bitsValid) return fail(h, GC_ERR_ARG, "gc_nav_sync: bad argument");
    cudaSetDevice(h->cfg.device);
    cudaStream_t st = h->stream;
    const size_t n = (size_t)nCh * nEpochs;
    GC_CUDA(h, h->trackOut.reserve(n));                       // the I_P rows
    GC_CUDA(h, h->navCand.reserve(n));
    GC_CUDA(h, h->navBits.reserve((size_t)nCh * GC_NAV_BITS));
    GC_CUDA(h, h->navInt.reserve(2 * (size_t)nCh));
    GC_CUDA(h, cudaMemcpyAsync(h->trackOut.p, I_P, n * sizeof(double), cudaMemcpyHostToDevice, st));
    GC_CUDA(h, launch_nav_sync(h->trackOut.p, nCh, nEpochs, 0 /* searchStartOffset, NAVdecoding.m:66 */, nEpochs, h->navCand.p,
                               h->navInt.p, h->navBits.p, h->navInt.p + nCh, st));
    std::vector<int> hi(2 * (size_t)nCh);
    GC_CUDA(h, cudaMemcpyAsync(hi.data(), h->navInt.p, hi.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    GC_CUDA(h, cudaMemcpyAsync(navBits, h->navBits.p, (size_t)nCh * GC_NAV_BITS, cudaMemcpyDeviceToHost, st));
    GC_CUDA(h, cudaStreamSynchronize(st));
    for (int ch = 0; ch < nCh; ++ch) {
        subFrameStart[ch] = hi[ch] == 0x7fffffff ? 0 : hi[ch];
        bitsValid[ch] = hi[nCh + ch];
    }
    return GC_OK;
}

int gc_code_entries(int32_t signal, int32_t component)
{
    if (signal == GC_SIG_GPS_L1CA) return component == 0 ? 1023 : 0;
    if (signal == GC_SIG_GLO_G1G2) return component == 0 ? 511 : 0;
    if (signal == GC_SIG_BDS_B3I) return component == 0 ? 10230 : 0;
    return code_entries(signal, component);
}

int gc_generate_code(int32_t signal, int32_t sv, int32_t component, int8_t* out, int32_t nOut)
{
    const int n = gc_code_entries(signal, component);
    if (!out || n == 0 || nOut < n) return GC_ERR_ARG;
    if (signal == GC_SIG_GPS_L1CA) { if (sv < 1 || sv > 32) return GC_ERR_ARG; ca_code(sv, out); return n; }
    if (signal == GC_SIG_GLO_G1G2) { glo_code(out); return n; }
    if (signal == GC_SIG_BDS_B3I) { if (sv < 1 || sv > 63) return GC_ERR_ARG; b3i_code(sv, out); return n; }
    CodeJob j;
    if (!make_code_job(signal, sv, component, &j)) return GC_ERR_ARG;
    run_code_job_host(j, out);
    return n;
}

int gc_generate_code_device(int32_t device, int32_t signal, int32_t nSv, const int32_t* svList, int32_t component, int8_t* out)
{
    const int n = code_entries(signal, component);
    if (!out || !svList || nSv < 1 || n == 0) return GC_ERR_ARG;
    std::vector<CodeJob> jobs(nSv);
    for (int i = 0; i < nSv; ++i)
        if (!make_code_job(signal, svList[i], component, &jobs[i])) return GC_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return GC_ERR_CUDA;
    std::vector<int8_t> host;
    if (run_code_jobs_device(jobs, host, nullptr) != cudaSuccess) return GC_ERR_CUDA;
    std::copy(host.begin(), host.end(), out);
    return n;
}

void* gc_get_stream(const gc_handle* h) { return h ? (void*)h->stream : nullptr; }

int gc_get_stats(const gc_handle* h, gc_stats* out)
{
    if (!h || !out) return GC_ERR_ARG;
    if (h->statsPending) {                                    // an asynchronous acquisition: its events are read once its stream has drained
        gc_handle* hm = const_cast<gc_handle*>(h);
        cudaSetDevice(hm->cfg.device);
        cudaStreamSynchronize(hm->stream);
        resolve_acq_stats(hm);
    }
    *out = h->stats;
    return GC_OK;
}

}  // extern "C"
