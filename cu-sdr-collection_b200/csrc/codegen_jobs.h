// One code-generation job = one (SV, component) of a signal (internal to libgnsscorr; codegen.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace gc {

struct CodeJob {
    enum Kind { NONE = 0, L5, GALE5, GALSEC, B2A, B1I, L2C, B1C, E1 };
    int kind;
    int a, b, c;              // the generator's ICD constants (advance / start value / taps / w, p ...), codegen.h
    int n;                    // entries written
    long long nChips;         // GPS L2C: chips before the return-to-zero interleave
    const char* hex;          // host pointer to the job's hex digits (memory codes, secondary codes) or nullptr
    int hexLen;
    long long hexOff, outOff; // offsets into the device pools (filled by run_code_jobs_device)
};

int code_entries(int signal, int comp);
bool make_code_job(int signal, int sv, int comp, CodeJob* job);
void run_code_job_host(const CodeJob& job, int8_t* out);
cudaError_t run_code_jobs_device(std::vector<CodeJob>& jobs, std::vector<int8_t>& out, cudaStream_t st);

}  // namespace gc
