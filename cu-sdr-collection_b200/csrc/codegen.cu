// Code generation for the signals whose primary codes the reference builds at run time: the per-SV job (which
// register lengths, taps and ICD constants), the device kernel that runs the generators of codegen.h - one thread per (SV, component) -
// and the host twin used by gc_generate_code.  The engine fills every code the caller has not supplied through gc_set_code from here.
#include <cstring>
#include <vector>

#include "../../include/gnsscorr.h"
#include "codegen.h"
#include "codegen_jobs.h"

namespace gc {

namespace {
#include "icd_tables.inc"
}

// entries of component `comp` in gc_set_code's layout, 0 = the signal has no such component
int code_entries(int signal, int comp)
{
    switch (signal) {
        case GC_SIG_GAL_E1C: return comp <= 1 ? 4092 : 0;
        case GC_SIG_GPS_L5C: return comp <= 1 ? 10230 : 0;
        case GC_SIG_GAL_E5A: case GC_SIG_GAL_E5B: return comp <= 1 ? 10230 : comp == 2 ? 100 : 0;
        case GC_SIG_BDS_B2A: return comp <= 1 ? 10230 : 0;
        case GC_SIG_BDS_B1I: return comp == 0 ? 2046 : 0;
        case GC_SIG_GPS_L2C: return comp == 0 ? 20460 : comp == 1 ? 1534500 : 0;
        case GC_SIG_BDS_B1C: return comp <= 1 ? 20460 : comp == 2 ? 122760 : 0;
        default: return 0;
    }
}

bool make_code_job(int signal, int sv, int comp, CodeJob* j)
{
    memset(j, 0, sizeof(*j));
    j->n = code_entries(signal, comp);
    if (j->n == 0 || sv < 1) return false;
    switch (signal) {
        case GC_SIG_GAL_E1C:
            if (sv > 50) return false;
            j->kind = CodeJob::E1; j->hex = comp == 0 ? kE1B_HEX[sv - 1] : kE1C_HEX[sv - 1]; j->hexLen = 1023;
            return true;
        case GC_SIG_GPS_L5C:
            if (sv > 210) return false;
            j->kind = CodeJob::L5; j->a = comp == 0 ? kL5I_ADVANCE[sv - 1] : kL5Q_ADVANCE[sv - 1];
            return true;
        case GC_SIG_GAL_E5A: case GC_SIG_GAL_E5B: {
            if (sv > 50) return false;
            const bool a = signal == GC_SIG_GAL_E5A;
            if (comp == 2) { j->kind = CodeJob::GALSEC; j->hex = a ? kE5AQ_SECONDARY[sv - 1] : kE5BQ_SECONDARY[sv - 1]; j->hexLen = 25; return true; }
            j->kind = CodeJob::GALE5;
            j->a = a ? (comp == 0 ? kE5AI_START[sv - 1] : kE5AQ_START[sv - 1]) : (comp == 0 ? kE5BI_START[sv - 1] : kE5BQ_START[sv - 1]);
            j->b = a ? 040503 : 064021;                                   // Feedback_Reg1 (generateE5aIcode.m:65, generateE5bIcode.m)
            j->c = a ? 050661 : (comp == 0 ? 051445 : 043143);           // Feedback_Reg2
            return true;
        }
        case GC_SIG_BDS_B2A:
            if (sv > 63) return false;
            j->kind = CodeJob::B2A; j->a = comp == 0 ? kB2AD_REG2[sv - 1] : kB2AP_REG2[sv - 1]; j->b = comp;
            return true;
        case GC_SIG_BDS_B1I:
            if (sv > 58) return false;
            j->kind = CodeJob::B1I; j->a = kB1I_G2S1[sv - 1]; j->b = kB1I_G2S2[sv - 1]; j->c = sv > 37 ? kB1I_G2S3[sv - 38] : 0;
            return true;
        case GC_SIG_GPS_L2C:
            if (sv > 63) return false;
            j->kind = CodeJob::L2C; j->a = comp == 0 ? kL2CM_INIT[sv - 1] : kL2CL_INIT[sv - 1]; j->b = comp; j->nChips = j->n / 2;
            return true;
        case GC_SIG_BDS_B1C:
            if (sv > 63) return false;
            j->kind = CodeJob::B1C; j->a = comp == 0 ? kB1CD_W[sv - 1] : kB1CP_W[sv - 1]; j->b = comp == 0 ? kB1CD_P[sv - 1] : kB1CP_P[sv - 1]; j->c = comp;
            return true;
        default: return false;
    }
}

namespace {
// one job; `hex` = the job's hex characters (host or device copy), `scratch` = 10243 bytes (B1C only)
__host__ __device__ void run_job(const CodeJob& j, const char* hex, int8_t* out, uint8_t* scratch)
{
    switch (j.kind) {
        case CodeJob::L5: codegen::gen_l5(j.a, out); break;
        case CodeJob::GALE5: codegen::gen_gal_e5(j.a, j.b, j.c, out); break;
        case CodeJob::GALSEC: codegen::gen_gal_secondary(hex, out); break;
        case CodeJob::B2A: codegen::gen_b2a(j.a, j.b, out); break;
        case CodeJob::B1I: codegen::gen_b1i(j.a, j.b, j.c, out); break;
        case CodeJob::L2C: codegen::gen_l2c((uint32_t)j.a, j.nChips, j.b, out); break;
        case CodeJob::B1C: codegen::legendre_sequence(10243, scratch); codegen::gen_b1c(j.a, j.b, j.c, scratch, out); break;
        case CodeJob::E1: codegen::gen_e1(hex, out); break;
        default: break;
    }
}

__global__ void codegen_kernel(const CodeJob* jobs, const char* hexPool, int8_t* outPool, uint8_t* scratchPool, int nJobs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nJobs) return;
    const CodeJob j = jobs[i];
    run_job(j, hexPool + j.hexOff, outPool + j.outOff, scratchPool + (size_t)i * 10243);
}
}  // namespace

void run_code_job_host(const CodeJob& j, int8_t* out)
{
    std::vector<uint8_t> scratch(j.kind == CodeJob::B1C ? 10243 : 1);
    run_job(j, j.hex, out, scratch.data());
}

// runs the jobs on the device (stream `st`) and returns their outputs concatenated in `out` (host); offsets are filled in here
cudaError_t run_code_jobs_device(std::vector<CodeJob>& jobs, std::vector<int8_t>& out, cudaStream_t st)
{
    std::vector<char> hexPool;
    size_t outBytes = 0;
    for (auto& j : jobs) {
        j.hexOff = (long long)hexPool.size();
        if (j.hex) { hexPool.insert(hexPool.end(), j.hex, j.hex + j.hexLen); hexPool.push_back(0); }
        j.outOff = (long long)outBytes;
        outBytes += (size_t)j.n;
    }
    if (hexPool.empty()) hexPool.push_back(0);
    CodeJob* dJobs = nullptr; char* dHex = nullptr; int8_t* dOut = nullptr; uint8_t* dScr = nullptr;
    cudaError_t e;
    auto cleanup = [&]() { cudaFree(dJobs); cudaFree(dHex); cudaFree(dOut); cudaFree(dScr); };
    if ((e = cudaMalloc(&dJobs, jobs.size() * sizeof(CodeJob))) != cudaSuccess || (e = cudaMalloc(&dHex, hexPool.size())) != cudaSuccess ||
        (e = cudaMalloc(&dOut, outBytes)) != cudaSuccess || (e = cudaMalloc(&dScr, jobs.size() * 10243)) != cudaSuccess) { cleanup(); return e; }
    cudaMemcpyAsync(dJobs, jobs.data(), jobs.size() * sizeof(CodeJob), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dHex, hexPool.data(), hexPool.size(), cudaMemcpyHostToDevice, st);
    const int n = (int)jobs.size();
    codegen_kernel<<<(n + 31) / 32, 32, 0, st>>>(dJobs, dHex, dOut, dScr, n);
    e = cudaGetLastError();
    out.resize(outBytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out.data(), dOut, outBytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cleanup();
    return e;
}

}  // namespace gc
