// Micro-benchmark: issue rate and dependent latency of FP64 and conversion instructions on sm_100a (B200), in
// warp-instructions per cycle per SM sub-partition.  Guides the float64 inner loop of the tracking kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

#define REP 32
#define TRIPS 64
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(double* out, long long* cyc, double s1, double s2)
{
    double a[NACC];
    float f[NACC];
    int v[NACC];
    const double b = s1 + threadIdx.x * 1e-12, c = s2 + threadIdx.x * 1e-12;
    const float fb = (float)b, fc = (float)c;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = threadIdx.x * 0.001 + i; f[i] = (float)a[i]; v[i] = threadIdx.x + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < TRIPS; ++t) {
#pragma unroll
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
                if (MODE == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b));
                if (MODE == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b));
                if (MODE == 3) asm volatile("add.rp.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b));
                // conversions: the result feeds an integer xor chain so that nothing is dead (the xor/add partner is measured alone in 12/13)
                if (MODE == 4) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f[i])); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(__int_as_float(__double2hiint(d) & 1))); }
                if (MODE == 5) { double d; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(v[i])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(__double2hiint(d))); }
                if (MODE == 12) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(v[(i + 1) % NACC])); }
                if (MODE == 13) { asm volatile("bfe.s32 %0, %0, 8, 8;" : "+r"(v[i])); }
                if (MODE == 14) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(v[(i + 1) % NACC]), "r"(v[(i + 2) % NACC])); }
                if (MODE == 6) {   // DFMA + FFMA alternating
                    if (i & 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
                    else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc));
                }
                if (MODE == 7) {   // DFMA + integer alternating
                    if (i & 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
                    else asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(v[(i + 1) % NACC]));
                }
                if (MODE == 8) {   // 1 DFMA : 3 FFMA
                    if ((i & 3) == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(b), "d"(c));
                    else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc));
                }
                if (MODE == 9) asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(v[i]) : "r"(0x4B000000));
                if (MODE == 10) { float g; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(g) : "r"(v[i])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(__float_as_int(g))); }
                if (MODE == 11) { int w; asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(w) : "d"(a[i])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(w)); a[i] = __hiloint2double(v[i] & 0x3fffffff, v[i]); }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += a[i] + f[i] + v[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// dependent-chain latency: one warp per SM, a chain of N dependent ops
template <int MODE>
__global__ void lat(double* out, long long* cyc, double s1, double s2)
{
    double a = s1 + threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < 64; ++t) {
#pragma unroll
        for (int r = 0; r < 64; ++r) {
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a) : "d"(s1), "d"(s2));
            if (MODE == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a) : "d"(s2));
            if (MODE == 2) { float x = (float)a; asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"((float)s1), "f"((float)s2)); a = x; }
        }
    }
    const long long t1 = clock64();
    if (a == 123.456) out[0] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double* out, long long* cyc)
{
    bench<MODE><<<148, 1024>>>(out, cyc, 1.0000001, 0.9999999);
    bench<MODE><<<148, 1024>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double instr = (double)TRIPS * REP * NACC * 8;
    printf("%-28s %.3f warp-inst/clk/SMSP   (%.0f cycles)\n", name, instr / avg, avg);
}
template <int MODE>
void runlat(const char* name, double* out, long long* cyc)
{
    lat<MODE><<<148, 32>>>(out, cyc, 1.0000001, 0.9999999);
    lat<MODE><<<148, 32>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-28s %.1f cycles per dependent op\n", name, avg / (64.0 * 64.0));
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
    run<0>("DFMA", out, cyc);
    run<1>("DADD", out, cyc);
    run<2>("DMUL", out, cyc);
    run<3>("DADD.RP", out, cyc);
    run<12>("XOR alone", out, cyc);
    run<13>("BFE.S32", out, cyc);
    run<14>("LOP3", out, cyc);
    run<4>("F2F.F64.F32 + FADD", out, cyc);
    run<5>("I2F.F64.S32 + XOR", out, cyc);
    run<10>("I2F.F32.S32 + XOR", out, cyc);
    run<6>("DFMA+FFMA 1:1", out, cyc);
    run<8>("DFMA+FFMA 1:3", out, cyc);
    run<7>("DFMA+XOR 1:1", out, cyc);
    run<9>("PRMT", out, cyc);
    runlat<0>("DFMA latency", out, cyc);
    runlat<1>("DADD latency", out, cyc);
    runlat<2>("FFMA(+cvt) latency", out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
