// host check of the generated codelets: stubs for the CUDA vector type and packed intrinsics
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#define __device__
#define __forceinline__ inline
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#define CODELET_HOST_CHECK
#include "fft_codelets.cuh"
template <int N> double check()
{
    double worst = 0;
    for (int inv = 0; inv < 2; ++inv) {
        float2 x[N]; std::vector<std::complex<double>> in(N), ref(N), got(N);
        for (int i = 0; i < N; ++i) { in[i] = {sin(1.0 + 0.37 * i * i), cos(0.5 + 1.7 * i)}; x[i] = make_float2((float)in[i].real(), (float)in[i].imag()); in[i] = {x[i].x, x[i].y}; }
        const double sg = inv ? 1.0 : -1.0;
        for (int k = 0; k < N; ++k) { std::complex<double> s = 0; for (int n = 0; n < N; ++n) s += in[n] * std::polar(1.0, sg * 2 * M_PI * ((long long)k * n % N) / N); ref[k] = s; }
        auto emit = [&](int k, float re, float im) { got[k] = {re, im}; };
        if (inv) gc::codelet::dft<N, true>(x, emit); else gc::codelet::dft<N, false>(x, emit);
        for (int k = 0; k < N; ++k) worst = std::max(worst, std::abs(got[k] - ref[k]) / (std::sqrt((double)N)));
    }
    printf("N=%d max err %.3g\n", N, worst);
    return worst;
}
int main() { double w = 0; w = std::max(w, check<4>()); w = std::max(w, check<8>()); w = std::max(w, check<9>()); w = std::max(w, check<10>()); w = std::max(w, check<18>()); w = std::max(w, check<20>()); w = std::max(w, check<16>()); w = std::max(w, check<25>()); w = std::max(w, check<30>()); w = std::max(w, check<31>()); w = std::max(w, check<32>());
  w = std::max(w, check<33>()); w = std::max(w, check<40>()); w = std::max(w, check<45>()); w = std::max(w, check<50>()); return w < 5e-6 ? 0 : 1; }
