// ubench_ntt_tma.cu — does staging through shared memory with the bulk-copy engine (cp.async.bulk + mbarrier, the
// non-tensor TMA path) help the NTT-only kernel?  The north star of this project prescribes "TMA-staged twiddle
// tiles in shared memory"; the product reads twiddles with 256-bit read-only global loads of an L1/L2-resident
// table and the polynomial with plain coalesced loads.  Variants, all persistent (one CTA loops over polynomials)
// and all built on the product's own seb_ntt.cuh:
//   TW_SMEM : the per-pass twiddle tables (35 KB at n = 4096) are bulk-copied into shared memory once per CTA and
//             the butterflies read them from there (2 x LDS.128 per oct instead of 1 x LDG.256)
//   IN_TMA  : the polynomial is bulk-copied into a staging buffer (the copy of polynomial k+1 is issued as soon
//             as pass 0 of polynomial k has consumed the buffer) and pass 0 reads shared memory
// Build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../seal-embedded_b200/csrc -o ubench_ntt_tma ubench_ntt_tma.cu
//        (and -DUB_TW_SMEM -o ubench_ntt_tma_twsmem for the twiddles-in-shared-memory flavour)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "seb_common.cuh"

// Twiddle fetch used by seb_ntt.cuh.  The address space has to be a compile-time property of the load, so the
// two flavours are two builds of this file: default = the product's LDG.E.256.CONSTANT of the global table;
// -DUB_TW_SMEM = 2 x LDS.128 of the table staged in shared memory.
#ifdef UB_TW_SMEM
struct seb_oct;
__device__ __forceinline__ seb_oct tw_load_shared(const seb_oct *p);
#define SEB_TW_LOAD(p) tw_load_shared(p)
#endif
#include "seb_ntt.cuh"
#ifdef UB_TW_SMEM
__device__ __forceinline__ seb_oct tw_load_shared(const seb_oct *p)
{
    seb_oct r;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "r"(a));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "r"(a));
    return r;
}
constexpr bool kTwSmem = true;
#else
constexpr bool kTwSmem = false;
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared::cta (bytes: multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct LoadGlobal
{
    const uint32_t *src;
    __device__ __forceinline__ uint32_t operator()(int, uint32_t pos) const { return seb_ldg_stream(src + pos); }
};
struct LoadShared
{
    const uint32_t *src;
    __device__ __forceinline__ uint32_t operator()(int, uint32_t pos) const { return src[pos]; }
};

template <int LOGN, bool TW_SMEM, bool IN_TMA>
struct Lay
{
    static constexpr size_t TW_BYTES   = TW_SMEM ? (size_t)NttTwSize<LOGN>::OCTS * 32 : 0;
    static constexpr size_t IN_BYTES   = IN_TMA ? ((size_t)4 << LOGN) : 0;
    static constexpr size_t WORK_BYTES = 4 * (size_t)NttSmem<LOGN>::WORDS;
    static constexpr size_t TOTAL      = 64 + TW_BYTES + IN_BYTES + WORK_BYTES;
};

template <int LOGN, int MINB, bool TW_SMEM, bool IN_TMA>
__global__ void __launch_bounds__((1 << LOGN) / SEB_E, MINB)
    k_ntt_p(uint32_t *__restrict__ polys, const seb_oct *__restrict__ roots, uint32_t q, size_t npoly)
{
    constexpr int N = 1 << LOGN;
    using L         = Lay<LOGN, TW_SMEM, IN_TMA>;
    extern __shared__ __align__(128) uint8_t sm[];
    uint64_t *bars  = reinterpret_cast<uint64_t *>(sm);  // [0]: twiddles, [1]: input
    seb_oct *tw_s   = reinterpret_cast<seb_oct *>(sm + 64);
    uint32_t *stage = reinterpret_cast<uint32_t *>(sm + 64 + L::TW_BYTES);
    uint32_t *work  = reinterpret_cast<uint32_t *>(sm + 64 + L::TW_BYTES + L::IN_BYTES);
    const int t          = threadIdx.x;
    const uint32_t two_q = 2 * q;
    size_t poly          = blockIdx.x;
    if (poly >= npoly) return;

    if (TW_SMEM || IN_TMA)
    {
        if (t == 0)
        {
            mbar_init(bars + 0, 1);
            mbar_init(bars + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (TW_SMEM)
            {
                mbar_expect_tx(bars + 0, (uint32_t)L::TW_BYTES);
                bulk_g2s(tw_s, roots, (uint32_t)L::TW_BYTES, bars + 0);
            }
            if (IN_TMA)
            {
                mbar_expect_tx(bars + 1, (uint32_t)L::IN_BYTES);
                bulk_g2s(stage, polys + poly * N, (uint32_t)L::IN_BYTES, bars + 1);
            }
        }
        __syncthreads();
        if (TW_SMEM) mbar_wait(bars + 0, 0);
    }
    const seb_oct *tw = TW_SMEM ? tw_s : roots;
    uint32_t phase    = 0;
    for (; poly < npoly; poly += gridDim.x)
    {
        uint32_t *data = polys + poly * N;
        uint32_t x[1][SEB_E];
        if (IN_TMA)
        {
            mbar_wait(bars + 1, phase);
            phase ^= 1;
            LoadShared ld{stage};
            seb_ntt_first<LOGN, 1>(x, work, t, tw, q, two_q, ld);  // ends with __syncthreads(): stage is consumed
            const size_t next = poly + gridDim.x;
            if (t == 0 && next < npoly)
            {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bars + 1, (uint32_t)L::IN_BYTES);
                bulk_g2s(stage, polys + next * N, (uint32_t)L::IN_BYTES, bars + 1);
            }
        }
        else
        {
            LoadGlobal ld{data};
            seb_ntt_first<LOGN, 1>(x, work, t, tw, q, two_q, ld);
        }
        seb_ntt_rest<LOGN, 1>(x, work, t, tw, q, two_q);
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
        __syncthreads();  // work is rewritten by the next polynomial's pass 0
    }
}

// the product's structure: one CTA per polynomial, not persistent (tools/ubench/ubench_ntt.cu)
template <int LOGN, int MINB>
__global__ void __launch_bounds__((1 << LOGN) / SEB_E, MINB)
    k_ntt_np(uint32_t *__restrict__ polys, const seb_oct *__restrict__ roots, uint32_t q, size_t npoly)
{
    constexpr int N = 1 << LOGN;
    extern __shared__ __align__(128) uint8_t sm[];
    uint32_t *work       = reinterpret_cast<uint32_t *>(sm);
    const int t          = threadIdx.x;
    const uint32_t two_q = 2 * q;
    uint32_t *data       = polys + (size_t)blockIdx.x * N;
    uint32_t x[1][SEB_E];
    LoadGlobal ld{data};
    seb_ntt_forward<LOGN, 1>(x, work, t, roots, q, two_q, ld);
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

static uint64_t checksum(const uint32_t *d, size_t words)
{
    std::vector<uint32_t> h(words);
    CK(cudaMemcpy(h.data(), d, words * 4, cudaMemcpyDeviceToHost));
    uint64_t s = 1469598103934665603ULL;
    for (uint32_t v : h) s = (s ^ v) * 1099511628211ULL;
    return s;
}

template <class K>
static void time_kernel(const char *label, K kern, unsigned grid, int threads, size_t smem, uint32_t *d_polys,
                        const uint32_t *d_init, size_t npoly, int n, const seb_oct *d_tw, uint32_t q, double peak)
{
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    CK(cudaMemcpy(d_polys, d_init, npoly * n * 4, cudaMemcpyDeviceToDevice));
    kern<<<grid, threads, smem>>>(d_polys, d_tw, q, npoly);
    CK(cudaDeviceSynchronize());
    const uint64_t sum = checksum(d_polys, (size_t)1024 * n);  // same input, same transform: same digest in every variant
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; i++) kern<<<grid, threads, smem>>>(d_polys, d_tw, q, npoly);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double gbs = 8.0 * n * npoly / (ms * 1e-3) / 1e9;
    printf("n=%5d %-58s regs=%3d smem=%6zu B occ=%d CTA/SM: %.3f ms  %.0f GB/s  %.1f%% of %.0f  digest=%016llx\n", n, label,
           fa.numRegs, smem, occ, ms, gbs, 100 * gbs / peak, peak, (unsigned long long)sum);
}

template <int LOGN, int MINB, bool IN_TMA>
static void run_p(uint32_t *d_polys, const uint32_t *d_init, size_t npoly, const seb_oct *d_tw, uint32_t q, double peak,
                  int sms)
{
    constexpr int N   = 1 << LOGN;
    const size_t smem = Lay<LOGN, kTwSmem, IN_TMA>::TOTAL;
    auto kern         = k_ntt_p<LOGN, MINB, kTwSmem, IN_TMA>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, N / SEB_E, smem));
    char label[128];
    snprintf(label, sizeof label, "persistent, twiddles %s, input %s, minb %d", kTwSmem ? "bulk->smem (LDS)" : "LDG.256",
             IN_TMA ? "bulk->smem prefetched" : "LDG", MINB);
    if (occ == 0) { printf("n=%5d %s: does not fit\n", N, label); return; }
    time_kernel(label, kern, (unsigned)(sms * occ), N / SEB_E, smem, d_polys, d_init, npoly, N, d_tw, q, peak);
}

template <int LOGN, int MINB>
static void run_np(uint32_t *d_polys, const uint32_t *d_init, size_t npoly, const seb_oct *d_tw, uint32_t q, double peak)
{
    constexpr int N = 1 << LOGN;
    char label[128];
    snprintf(label, sizeof label, "one CTA per polynomial (product), twiddles LDG.256, input LDG, minb %d", MINB);
    time_kernel(label, k_ntt_np<LOGN, MINB>, (unsigned)npoly, N / SEB_E, 4 * (size_t)NttSmem<LOGN>::WORDS, d_polys, d_init,
                npoly, N, d_tw, q, peak);
}

template <int LOGN>
static seb_oct *make_tw(uint32_t q)
{
    constexpr int N = 1 << LOGN;
    std::vector<uint2> roots(N);
    uint64_t s = 0x9E3779B97F4A7C15ULL + LOGN;
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
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms      = prop.multiProcessorCount;
    const uint32_t q   = 1053818881u;
    const size_t bytes = (size_t)3 << 30;
    uint32_t *d_polys, *d_init;
    CK(cudaMalloc(&d_polys, bytes));
    CK(cudaMalloc(&d_init, bytes));
    std::vector<uint32_t> h(bytes / 4);
    uint64_t s = 88172645463325252ULL;
    for (auto &v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s % q); }
    CK(cudaMemcpy(d_init, h.data(), bytes, cudaMemcpyHostToDevice));
    {
        seb_oct *d_tw      = make_tw<12>(q);
        const size_t npoly = bytes / ((size_t)4 << 12);
#ifndef UB_TW_SMEM
        run_np<12, 5>(d_polys, d_init, npoly, d_tw, q, peak);
        run_p<12, 5, false>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<12, 5, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<12, 4, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<12, 6, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
#else
        run_p<12, 4, false>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<12, 3, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<12, 2, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
#endif
        cudaFree(d_tw);
    }
#ifndef UB_TW_SMEM
    {
        seb_oct *d_tw      = make_tw<10>(q);
        const size_t npoly = bytes / ((size_t)4 << 10);
        run_np<10, 20>(d_polys, d_init, npoly, d_tw, q, peak);
        run_p<10, 20, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<10, 16, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        cudaFree(d_tw);
    }
    {
        seb_oct *d_tw      = make_tw<13>(q);
        const size_t npoly = bytes / ((size_t)4 << 13);
        run_np<13, 4>(d_polys, d_init, npoly, d_tw, q, peak);
        run_p<13, 4, false>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<13, 4, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<13, 3, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<13, 2, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        cudaFree(d_tw);
    }
    {
        seb_oct *d_tw      = make_tw<14>(q);
        const size_t npoly = bytes / ((size_t)4 << 14);
        run_np<14, 2>(d_polys, d_init, npoly, d_tw, q, peak);
        run_p<14, 2, false>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<14, 2, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        run_p<14, 1, true>(d_polys, d_init, npoly, d_tw, q, peak, sms);
        cudaFree(d_tw);
    }
#endif
    return 0;
}
