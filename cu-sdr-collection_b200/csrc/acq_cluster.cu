// Acquisition, correlation stage as ONE kernel (opt-in: GC_ACQ_PATH=cluster): a thread-block cluster
// computes a whole (SV, Doppler bin) cell of GPS/GPS_L1CA/include/acquisition.m:175-198 -
//
//     for every non-coherent block m (and every replica of a data+pilot pair):
//         IQfreqDom .* caCodeFreqDom  ->  ifft  ->  abs  ->  results(bin, :) += ...     (:186-190)
//     then the maximum / first arg-max of results(bin, :)                                (:196-198)
//
// with the complete L-point inverse transform resident in the cluster's shared memory, so the
// [SV][bin][block][L] intermediate of the two-kernel version (acq_fused.cu inv_rows + inv_cols,
// 2 x 2.4 GB of HBM traffic per 16 PRNs) never exists and DRAM traffic drops to the spectra.
//
//   cluster of CL CTAs, L = C x R (R = 32 x RB), plan and index maps as in acq_fused.cu:
//     rows phase : CTA q owns rows k1 in [q*RPC, (q+1)*RPC), one warp per row: load X .* Cc from
//                  global (L2 resident), RB-point and 32-point inverse DFTs through the row's own
//                  shared-memory buffer, result left in place in that buffer
//     cols phase : CTA q owns row positions pp in [q*R/CL, (q+1)*R/CL), one thread per position:
//                  gathers the C elements of its column from every CTA's row buffers (DSMEM loads),
//                  C-point inverse DFT in registers, |.|, accumulate into a shared-memory
//                  results(bin, :) slice [C][R/CL]
//   two cluster barriers per block; the second one is split (arrive after the column loads, wait
//   only before the next rows phase writes shared memory).
//
// Measured on B200 (profiles/r01_cluster_experiments.md): 2.85 ms for the 32 x 29 x 20 grid against
// 2.05 ms for the two-kernel version - the serialised phases at 17 warps per SM leave the FMA pipe
// at 40 %, and the DSMEM gather runs at ~8 B/clk per SM.  A warp-specialised variant (row warps
// pushing into peer slices, col warps decoupled by the cluster barrier) measured 3.9 ms: two
// cluster-barrier phases per 5k-cycle block cost more than they hide.  Kept as the zero-W design
// point; the default path is the two-kernel version.
#include "acq.h"
#include "acq_plan.cuh"
#include "common.cuh"
#include "fft_codelets.cuh"

namespace gc {

namespace {

__device__ __forceinline__ uint32_t cl_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ uint32_t cl_map(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ float2 cl_ld(uint32_t caddr)
{
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(caddr));
    return v;
}

template <class P, int CL>
struct Geo {
    static constexpr int RPC = (P::C + CL - 1) / CL;          // rows per CTA
    static constexpr int COLS = P::R / CL;                    // row positions per CTA in the cols phase
    static_assert(P::R % CL == 0, "row length must split evenly over the cluster");
    static constexpr size_t kRowBytes = sizeof(float2) * P::R;
    static constexpr size_t kSmem = kRowBytes * RPC + sizeof(float) * P::C * COLS;
    // warps per CTA: rows split evenly over the fewest rounds of at most 17 warps, and enough
    // threads for one round of the cols phase
    static constexpr int kRounds = (RPC + 16) / 17;
    static constexpr int kRowWarps = (RPC + kRounds - 1) / kRounds;
    static constexpr int kColWarps = (COLS + 31) / 32;
    static constexpr int NW = kRowWarps > kColWarps ? kRowWarps : (kColWarps <= 17 ? kColWarps : kRowWarps);
};

template <class P, int CL, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
corr_cluster_kernel(CorrParams p)
{
    using G = Geo<P, CL>;
    constexpr int C = P::C, RA = P::RA, RB = P::RB, R = P::R, RPC = G::RPC, COLS = G::COLS, NT = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_rows = reinterpret_cast<float2*>(smem_raw);                  // [RPC][R]
    float* s_acc = reinterpret_cast<float*>(smem_raw + G::kRowBytes * RPC); // [C][COLS]
    __shared__ float s_b[NW];
    __shared__ int s_i[NW];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t rank = cl_rank();
    const int cell = blockIdx.x / CL;
    const int k = cell / p.nSlots, slot = cell % p.nSlots;                  // Doppler bin, SV list slot
    const float2* Xg = p.X + ((size_t)p.slotGroup[slot] * p.nBins + k) * p.nonCoh * P::L;
    const int rep0 = p.slotReplica[slot];

    // DSMEM base address of every CTA's row buffers
    uint32_t rowBase[CL];
#pragma unroll
    for (int q = 0; q < CL; ++q) rowBase[q] = cl_map(smem_u32(s_rows), (uint32_t)q);

    cl_arrive();                                                            // pairs with the first rows-phase wait
    bool first = true;
    for (int m = 0; m < p.nonCoh; ++m) {                                    // acquisition.m:175
        for (int r = 0; r < p.nRep; ++r) {                                  // data + pilot replicas (e.g. GAL_E1C acquisition.m:186-192)
            const float2* Cg = p.Cc + (size_t)(rep0 + r * p.repStride) * P::L;
            // ------------------------------------------------ rows phase
            bool waited = false;
            for (int i = warp; i < RPC; i += NW) {
                const int k1 = (int)rank * RPC + i;
                if (k1 >= C) break;
                const float2* src = Xg + ((size_t)m * C + k1) * R;
                const float2* mul = Cg + (size_t)k1 * R;
                float2* s_x = s_rows + i * R;
                {                                                           // lane = ka: DFT-RB over kb
                    float2 u[RB];
#pragma unroll
                    for (int kb = 0; kb < RB; ++kb)                         // IQfreqDom .* caCodeFreqDom (:186)
                        u[kb] = cmul(__ldg(src + kb * RA + lane), __ldg(mul + kb * RA + lane));
                    // nobody may still be reading the row buffers of the previous block
                    if (!waited) { cl_wait(); waited = true; }
                    codelet::dft<RB, true>(u, [&](int tb, float re, float im) { s_x[lane * RB + tb] = make_float2(re, im); });
                }
                __syncwarp();
                if (lane < RB) {                                            // lane = tb: DFT-RA over ka, in place
                    float2 v[RA];
#pragma unroll
                    for (int ka = 0; ka < RA; ++ka) v[ka] = s_x[ka * RB + lane];
                    const float2* tw = p.tw + (size_t)k1 * R + lane;
                    codelet::dft<RA, true>(v, [&](int ta, float re, float im) {
                        float2 t = make_float2(re, im);
                        if (!P::kPfa) t = cmul_conj(t, __ldg(tw + ta * RB));  // conj(w_L^(j1*tau2))
                        s_x[ta * RB + lane] = t;
                    });
                }
            }
            if (!waited) cl_wait();
            cl_arrive();                                                    // rows of this block are in shared memory
            cl_wait();
            // ------------------------------------------------ cols phase
            for (int c = threadIdx.x; c < COLS; c += NT) {
                const int pp = (int)rank * COLS + c;
                float2 x[C];
#pragma unroll
                for (int k1 = 0; k1 < C; ++k1)
                    x[k1] = cl_ld(rowBase[k1 / RPC] + (uint32_t)(((k1 % RPC) * R + pp) * sizeof(float2)));
                float* acc = s_acc + c;
                codelet::dft<C, true>(x, [&](int t1, float re, float im) {
                    const float prev = first ? 0.f : acc[t1 * COLS];
                    acc[t1 * COLS] = prev + cabs_fast(re, im);              // results(bin, :) += abs(ifft(.)) (:188-190)
                });
            }
            cl_arrive();                                                    // done reading the row buffers
            first = false;
        }
    }
    cl_wait();                                                              // pairs with the last arrive; nobody exits early
    __syncthreads();

    // maximum / first arg-max over this CTA's slice of results(bin, :)
    float best = -1.f;
    int bidx = 0x7fffffff;
    for (int c = threadIdx.x; c < COLS; c += NT) {
        const int rest = P::row_index((int)rank * COLS + c);
#pragma unroll 4
        for (int t1 = 0; t1 < C; ++t1) {
            const float v = s_acc[t1 * COLS + c];
            const int idx = P::index(t1, rest);                             // code phase (lag) of this output
            if (v > best || (v == best && idx < bidx)) { best = v; bidx = idx; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { s_b[warp] = best; s_i[warp] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < NW; ++w)
            if (s_b[w] > best || (s_b[w] == best && s_i[w] < bidx)) { best = s_b[w]; bidx = s_i[w]; }
        const size_t o = ((size_t)slot * p.nBins + k) * kCorrClusterParts + rank;
        p.partMax[o] = best;
        p.partIdx[o] = bidx;
    }
    if (CL < kCorrClusterParts && threadIdx.x < kCorrClusterParts - CL && rank == 0) {   // unused part slots
        const size_t o = ((size_t)slot * p.nBins + k) * kCorrClusterParts + CL + threadIdx.x;
        p.partMax[o] = -1.f;
        p.partIdx[o] = 0x7fffffff;
    }
}

template <class P, int CL>
cudaError_t launch_t(const CorrParams& p, cudaStream_t s)
{
    using G = Geo<P, CL>;
    auto kern = corr_cluster_kernel<P, CL, G::NW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.nSlots * p.nBins * CL));
    cfg.blockDim = dim3(G::NW * 32);
    cfg.dynamicSmemBytes = G::kSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

constexpr int kSmemMax = 227 * 1024;

// the smallest cluster whose CTAs hold their share of the transform (row buffers + results slice)
template <class P>
struct Launch {
    static cudaError_t corr(const CorrParams& p, cudaStream_t s)
    {
        if constexpr (P::kBig) {
            return cudaErrorNotSupported;                     // two-level columns: only the two-kernel path
        } else if constexpr (Geo<P, 2>::kSmem + 1024 <= kSmemMax) {
            return launch_t<P, 2>(p, s);
        } else {
            static_assert(Geo<P, 4>::kSmem + 1024 <= kSmemMax, "transform does not fit a 4-CTA cluster");
            return launch_t<P, 4>(p, s);
        }
    }
};

}  // namespace

cudaError_t launch_corr_cluster(int L, const CorrParams& p, cudaStream_t s) { GC_PLAN_DISPATCH(L, corr(p, s)) }

}  // namespace gc
