// Micro-benchmark: issue rate of scalar and packed (f32x2) FP32 instructions on sm_100a, in
// warp-instructions per cycle per SM sub-partition.  Guides the DFT codelet generator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_issue fp32_issue.cu && ./fp32_issue
#include <cstdio>
#include <cuda_runtime.h>

#define REP 64      // unrolled groups per loop trip
#define TRIPS 64
#define NACC 8

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cyc, float s1, float s2)
{
    float a[NACC];
    u64 p[NACC];
    const float b = s1 + threadIdx.x * 1e-9f, c = s2 + threadIdx.x * 1e-9f;
    const u64 pb = pack(b, c), pc = pack(c, b);
    const u64 ps = pack(s1, s1);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = pack(a[i], a[i] + 1.f); }
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < TRIPS; ++t) {
#pragma unroll
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, %1;" : "+f"(a[i]) : "f"(c));
                if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                if (MODE == 3) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                if (MODE == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ps), "l"(pc));
                if (MODE == 6) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 7) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 8) {   // FFMA + FADD alternating
                    if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                    else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                }
                if (MODE == 9) {   // FFMA2 + FFMA alternating
                    if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                    else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                }
                if (MODE == 10) {  // FFMA imm + FADD alternating
                    if (i & 1) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, %1;" : "+f"(a[i]) : "f"(c));
                    else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                }
                if (MODE == 11) {  // FFMA2 + FADD2 alternating
                    if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                    else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                }
                if (MODE == 12) {  // FFMA2 + IADD alternating (integer on the alu pipe)
                    if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                    else { int v = __float_as_int(a[i]); asm volatile("add.s32 %0, %0, %1;" : "+r"(v) : "r"(__float_as_int(b))); a[i] = __int_as_float(v); }
                }
                if (MODE == 13) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b), "f"(c));   // acc as addend
                if (MODE == 14) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(pb), "l"(ps));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += a[i] + lo(p[i]);
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc)
{
    bench<MODE><<<148, 1024>>>(out, cyc, 1.0001f, 0.9999f);
    bench<MODE><<<148, 1024>>>(out, cyc, 1.0001f, 0.9999f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double instr = (double)TRIPS * REP * NACC * 8;   // warp-instructions per SMSP (8 warps each)
    printf("%-28s %.3f warp-inst/clk/SMSP   (%.0f cycles)\n", name, instr / avg, avg);
}

int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA reg,reg,reg", out, cyc);
    run<13>("FFMA b*c+acc", out, cyc);
    run<1>("FFMA reg,imm,reg", out, cyc);
    run<2>("FADD", out, cyc);
    run<3>("FMUL", out, cyc);
    run<4>("FFMA2 reg", out, cyc);
    run<5>("FFMA2 bcast-pair", out, cyc);
    run<14>("FFMA2 b*s+acc", out, cyc);
    run<6>("FADD2", out, cyc);
    run<7>("FMUL2", out, cyc);
    run<8>("FFMA+FADD", out, cyc);
    run<9>("FFMA2+FFMA", out, cyc);
    run<10>("FFMAimm+FADD", out, cyc);
    run<11>("FFMA2+FADD2", out, cyc);
    run<12>("FFMA2+IADD", out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
