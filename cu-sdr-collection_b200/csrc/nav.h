// Navigation-bit front end kernels (internal to libgnsscorr).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gc {

// ip [nCh][n] prompt in-phase outputs (trackResults.I_P); cand [nCh][n] scratch; first [nCh] subFrameStart (1-based,
// INT_MAX = none); bits [nCh][GC_NAV_BITS]; valid [nCh]
cudaError_t launch_nav_sync(const double* ip, int nCh, int n, int offset, int msToProcess, uint8_t* cand, int* first,
                            uint8_t* bits, int* valid, cudaStream_t st);

}  // namespace gc
