// ubench_ntt.cu — NTT-only kernel experiments on the product's own seb_ntt.cuh (occupancy variants, and the
// 16- vs 32-coefficients-per-thread plans at n >= 8192: keys 13/14 vs 29/30, NttCfg in seb_ntt.cuh).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../seal-embedded_b200/csrc -o ubench_ntt ubench_ntt.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "seb_ntt.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

struct LoadPlain
{
    const uint32_t *src;
    __device__ __forceinline__ uint32_t operator()(int, uint32_t pos) const { return seb_ldg_stream(src + pos); }
};

// the 2-CTA cluster form of plan 30 (n = 16384, 32 coefficients per thread): 256-thread CTAs
template <int LOGN, int MINB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NttCfg<LOGN>::T / 2, MINB)
    k_ntt_c2(uint32_t *__restrict__ polys, const seb_oct *__restrict__ roots, uint32_t q)
{
    constexpr int N = 1 << NttCfg<LOGN>::LOGN;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t        = seb_ntt_thread<LOGN, 2>();
    uint32_t *data     = polys + (size_t)(blockIdx.x >> 1) * N;
    const uint32_t two_q = 2 * q;
    uint32_t x[1][NttCfg<LOGN>::E];
    LoadPlain ld{data};
    seb_ntt_forward_cluster2<LOGN>(x, smem, t, roots, q, two_q, ld);
    using O = NttOut<LOGN>;
#pragma unroll
    for (int i = 0; i < O::GPL; i++)
    {
        seb_oct *dst = reinterpret_cast<seb_oct *>(data + O::pos(t, i));
#pragma unroll
        for (int k = 0; k < O::RUN / 8; k++)
        {
            seb_oct v;
#pragma unroll
            for (int c = 0; c < 8; c++) v.v[c] = seb_final_reduce(x[0][i * O::RUN + 8 * k + c], q, two_q);
            seb_stg256_stream(dst + k, v);
        }
    }
}

template <int LOGN, int MINB>
__global__ void __launch_bounds__(NttCfg<LOGN>::T, MINB)
    k_ntt(uint32_t *__restrict__ polys, const seb_oct *__restrict__ roots, uint32_t q)
{
    constexpr int N = 1 << NttCfg<LOGN>::LOGN;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t        = threadIdx.x;
    uint32_t *data     = polys + (size_t)blockIdx.x * N;
    const uint32_t two_q = 2 * q;
    uint32_t x[1][NttCfg<LOGN>::E];
    LoadPlain ld{data};
    seb_ntt_forward<LOGN, 1>(x, smem, t, roots, q, two_q, ld);
    using O = NttOut<LOGN>;
#pragma unroll
    for (int i = 0; i < O::GPL; i++)
    {
        seb_oct *dst = reinterpret_cast<seb_oct *>(data + O::pos(t, i));
#pragma unroll
        for (int k = 0; k < O::RUN / 8; k++)
        {
            seb_oct v;
#pragma unroll
            for (int c = 0; c < 8; c++) v.v[c] = seb_final_reduce(x[0][i * O::RUN + 8 * k + c], q, two_q);
            seb_stg256_stream(dst + k, v);
        }
    }
}

static uint32_t mulmod(uint32_t a, uint32_t b, uint32_t q) { return (uint32_t)((uint64_t)a * b % q); }

template <int LOGN, int MINB>
static void run(uint32_t *d_polys, const uint32_t *d_init, size_t npoly, const seb_oct *d_tw, uint32_t q, double peak)
{
    constexpr int N = 1 << NttCfg<LOGN>::LOGN;
    const size_t smem = 4 * NttSmem<LOGN>::WORDS;
    CK(cudaFuncSetAttribute(k_ntt<LOGN, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ntt<LOGN, MINB>, NttCfg<LOGN>::T, smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_ntt<LOGN, MINB>);
    CK(cudaMemcpy(d_polys, d_init, npoly * N * 4, cudaMemcpyDeviceToDevice));
    for (int i = 0; i < 2; i++) k_ntt<LOGN, MINB><<<(unsigned)npoly, NttCfg<LOGN>::T, smem>>>(d_polys, d_tw, q);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; i++) k_ntt<LOGN, MINB><<<(unsigned)npoly, NttCfg<LOGN>::T, smem>>>(d_polys, d_tw, q);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double gbs = 8.0 * N * npoly / (ms * 1e-3) / 1e9;
    // digest of the first polynomial after ONE transform of the initial data (equal across plans of one degree:
    // the made-up roots are the same table, only its layout differs)
    CK(cudaMemcpy(d_polys, d_init, (size_t)N * 4, cudaMemcpyDeviceToDevice));
    k_ntt<LOGN, MINB><<<1, NttCfg<LOGN>::T, smem>>>(d_polys, d_tw, q);
    std::vector<uint32_t> h(N);
    CK(cudaMemcpy(h.data(), d_polys, (size_t)N * 4, cudaMemcpyDeviceToHost));
    uint64_t dg = 1469598103934665603ULL;
    for (uint32_t v : h) dg = (dg ^ v) * 1099511628211ULL;
    printf("n=%5d E=%d minb=%d regs=%3d spill=%zu occ=%d CTA/SM: %.3f ms  %.0f GB/s  %.1f%% of %.0f  digest %016llx\n", N,
           NttCfg<LOGN>::E, MINB, fa.numRegs, (size_t)fa.localSizeBytes, occ, ms, gbs, 100 * gbs / peak, peak,
           (unsigned long long)dg);
}

template <int LOGN, int MINB>
static void run_c2(uint32_t *d_polys, const uint32_t *d_init, size_t npoly, const seb_oct *d_tw, uint32_t q, double peak)
{
    constexpr int N = 1 << NttCfg<LOGN>::LOGN;
    const size_t smem = 4 * NttSmemHalf<LOGN>::WORDS;
    CK(cudaFuncSetAttribute(k_ntt_c2<LOGN, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_ntt_c2<LOGN, MINB>);
    CK(cudaMemcpy(d_polys, d_init, npoly * N * 4, cudaMemcpyDeviceToDevice));
    for (int i = 0; i < 2; i++) k_ntt_c2<LOGN, MINB><<<(unsigned)(2 * npoly), NttCfg<LOGN>::T / 2, smem>>>(d_polys, d_tw, q);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; i++) k_ntt_c2<LOGN, MINB><<<(unsigned)(2 * npoly), NttCfg<LOGN>::T / 2, smem>>>(d_polys, d_tw, q);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double gbs = 8.0 * N * npoly / (ms * 1e-3) / 1e9;
    CK(cudaMemcpy(d_polys, d_init, (size_t)N * 4, cudaMemcpyDeviceToDevice));
    k_ntt_c2<LOGN, MINB><<<2, NttCfg<LOGN>::T / 2, smem>>>(d_polys, d_tw, q);
    std::vector<uint32_t> h(N);
    CK(cudaMemcpy(h.data(), d_polys, (size_t)N * 4, cudaMemcpyDeviceToHost));
    uint64_t dg = 1469598103934665603ULL;
    for (uint32_t v : h) dg = (dg ^ v) * 1099511628211ULL;
    printf("n=%5d E=%d 2-CTA cluster minb=%d regs=%3d spill=%zu: %.3f ms  %.0f GB/s  %.1f%% of %.0f  digest %016llx\n", N,
           NttCfg<LOGN>::E, MINB, fa.numRegs, (size_t)fa.localSizeBytes, ms, gbs, 100 * gbs / peak, peak, (unsigned long long)dg);
}

template <int LOGN>
static seb_oct *make_tw(uint32_t q)
{
    // timing only: any table of valid Shoup pairs exercises the same instructions
    constexpr int N = 1 << NttCfg<LOGN>::LOGN;
    std::vector<uint2> roots(N);
    uint64_t s = 0x9E3779B97F4A7C15ULL + NttCfg<LOGN>::LOGN;
    for (int i = 0; i < N; i++)
    {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const uint32_t w = (uint32_t)(s % q);
        roots[i] = make_uint2(w, (uint32_t)(((uint64_t)w << 32) / q));
    }
    std::vector<seb_oct> tab(NttTwSize<LOGN>::OCTS);
    memset(tab.data(), 0, tab.size() * sizeof(seb_oct));
    seb_build_tw<LOGN>(roots.data(), tab.data());
    seb_oct *d_tw;
    CK(cudaMalloc(&d_tw, tab.size() * sizeof(seb_oct)));
    CK(cudaMemcpy(d_tw, tab.data(), tab.size() * sizeof(seb_oct), cudaMemcpyHostToDevice));
    return d_tw;
}

int main(int argc, char **argv)
{
    const double peak = argc > 1 ? atof(argv[1]) : 6535.4;
    const uint32_t q = 1053818881u;
    const size_t bytes = (size_t)3 << 30;  // 3 GiB of polynomials, >> L2
    uint32_t *d_polys, *d_init;
    CK(cudaMalloc(&d_polys, bytes));
    CK(cudaMalloc(&d_init, bytes));
    std::vector<uint32_t> h(bytes / 4);
    uint64_t s = 88172645463325252ULL;
    for (auto &v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s % q); }
    CK(cudaMemcpy(d_init, h.data(), bytes, cudaMemcpyHostToDevice));
    const int only = argc > 2 ? atoi(argv[2]) : 0;  // run one plan key only (for ncu captures)
#define SWEEP(L, ...)                                                        \
    if (!only || only == L)                                                  \
    {                                                                        \
        seb_oct *d_tw = make_tw<L>(q);                                       \
        const size_t npoly = bytes / ((size_t)4 << (L & 15));                       \
        __VA_ARGS__                                                          \
        cudaFree(d_tw);                                                      \
    }
#define R(L, B) run<L, B>(d_polys, d_init, npoly, d_tw, q, peak);
    SWEEP(10, R(10, 1) R(10, 8) R(10, 12) R(10, 16) R(10, 20) R(10, 24) R(10, 32))
    SWEEP(11, R(11, 1) R(11, 4) R(11, 6) R(11, 8) R(11, 10) R(11, 12) R(11, 16))
    SWEEP(12, R(12, 1) R(12, 4) R(12, 5) R(12, 6) R(12, 8))
    SWEEP(13, R(13, 1) R(13, 2) R(13, 3) R(13, 4))
    SWEEP(29, R(29, 1) R(29, 2) R(29, 3) R(29, 4) R(29, 5))
    SWEEP(14, R(14, 1) R(14, 2))
    SWEEP(30, R(30, 1) R(30, 2) run_c2<30, 2>(d_polys, d_init, npoly, d_tw, q, peak); run_c2<30, 3>(d_polys, d_init, npoly, d_tw, q, peak);
              run_c2<30, 4>(d_polys, d_init, npoly, d_tw, q, peak);)
    return 0;
}
