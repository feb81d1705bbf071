// ubench_bfly.cu — which formulation of the Harvey/Shoup lazy NTT butterfly is cheapest on sm_100a?
// All variants map X,Y in [0,4q) to X+WY, X-WY in [0,4q) (device/lib/ntt.c:94-105 semantics); they
// differ in how floor(Y*W/q) is approximated and on which pipe the additions run.
//   V0  __umulhi(Y, floor(W*2^32/q))                       (IMAD.HI)            — product kernels, round 1
//   V1  hi32 of mul.wide.u32                               (IMAD.WIDE)
//   V2  FP64: floor(Y * RD(W/q)) via DADD + DFMA.RM        (FP64 pipe)
//   V3  V2 + additions folded into multiply-adds           (FMA pipe)
//   V4  V0 + additions folded into multiply-adds
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench_bfly ubench_bfly.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

struct Tw
{
    uint32_t w, wq;  // Shoup pair
    double winv;     // RD(w / q)
};

__constant__ uint32_t c_one[2] = {1u, 2u};

template <int V>
__device__ __forceinline__ void bfly(uint32_t &x, uint32_t &y, const Tw &tw, const uint32_t q, const uint32_t two_q,
                                     const uint32_t neg_q)
{
    uint32_t h;
    if (V == 0 || V == 4) h = __umulhi(y, tw.wq);
    if (V == 1)
    {
        uint64_t p;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(y), "r"(tw.wq));
        h = (uint32_t)(p >> 32);
    }
    if (V == 2 || V == 3)
    {
        const double yd = __hiloint2double(0x43300000, (int)y) - 4503599627370496.0;  // exact
        const double s  = __fma_rd(yd, tw.winv, 4503599627370496.0);                  // 2^52 + floor(y*winv)
        h               = (uint32_t)__double2loint(s);
    }
    const uint32_t u = min(x, x - two_q);
    if (V <= 2)
    {
        const uint32_t t = y * tw.w - h * q;  // [0, 2q)
        x                = u + t;
        y                = u - t + two_q;
    }
    else
    {
        // x' = u + y*w - h*q as two multiply-adds; y' = (2u + 2q) - x'
        uint32_t m, xn, v;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(m) : "r"(y), "r"(tw.w), "r"(u));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(xn) : "r"(h), "r"(neg_q), "r"(m));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(u), "r"(c_one[1]), "r"(two_q));
        x = xn;
        y = v - xn;
    }
}

constexpr int ITER = 256;

template <int V>
__global__ void __launch_bounds__(256) k_bfly(const Tw *__restrict__ tws, uint32_t *__restrict__ data, uint32_t q)
{
    const size_t tid     = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t two_q = 2 * q, neg_q = 0u - q;
    uint32_t x[16];
    Tw tw[15];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = data[tid * 16 + i];
#pragma unroll
    for (int i = 0; i < 15; i++) tw[i] = tws[(threadIdx.x & 31) * 15 + i];
#pragma unroll 1
    for (int it = 0; it < ITER; it++)
    {
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const int half = 8 >> r;
#pragma unroll
            for (int m = 0; m < (1 << r); m++)
#pragma unroll
                for (int t = 0; t < half; t++)
                    bfly<V>(x[m * 2 * half + t], x[m * 2 * half + t + half], tw[(1 << r) - 1 + m], q, two_q, neg_q);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++)
    {
        uint32_t v = x[i];
        v          = min(v, v - two_q);
        v          = min(v, v - q);
        data[tid * 16 + i] = v;
    }
}

template <int V>
static void run(const Tw *d_tw, uint32_t *d_data, const uint32_t *h_in, uint32_t *h_out, size_t nthreads, uint32_t q,
                const uint32_t *ref, int sms)
{
    CK(cudaMemcpy(d_data, h_in, nthreads * 64, cudaMemcpyHostToDevice));
    k_bfly<V><<<(unsigned)(nthreads / 256), 256>>>(d_tw, d_data, q);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_data, nthreads * 64, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    if (ref)
        for (size_t i = 0; i < nthreads * 16; i++) bad += h_out[i] != ref[i];
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k_bfly<V><<<(unsigned)(nthreads / 256), 256>>>(d_tw, d_data, q);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bf = 5.0 * nthreads * ITER * 32;
    const double rate = bf / (ms * 1e-3);
    printf("bfly V%d: %7.3f T butterflies/s = %5.2f SMSP-cycles per warp-butterfly (1.965 GHz)  mismatches vs V0: %zu\n", V,
           rate / 1e12, (double)sms * 4 * 1.965e9 * 32 / rate, bad);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms       = prop.multiProcessorCount;
    const uint32_t q    = 1053818881u;
    const size_t nthr   = (size_t)sms * 8 * 256;
    Tw h_tw[32 * 15];
    uint64_t s = 88172645463325252ULL;
    auto rnd   = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int i = 0; i < 32 * 15; i++)
    {
        const uint32_t w = (uint32_t)(rnd() % q);
        h_tw[i].w        = w;
        h_tw[i].wq       = (uint32_t)(((uint64_t)w << 32) / q);
        // RD(w/q): long double division then step down if needed
        double d = (double)w / (double)q;
        while ((long double)d * q > (long double)w) d = nextafter(d, 0.0);
        h_tw[i].winv = d;
    }
    // adversarial twiddles: 0, 1, q-1
    h_tw[0].w = 0; h_tw[0].wq = 0; h_tw[0].winv = 0.0;
    h_tw[1].w = 1; h_tw[1].wq = (uint32_t)(((uint64_t)1 << 32) / q);
    { double d = 1.0 / q; while ((long double)d * q > 1.0L) d = nextafter(d, 0.0); h_tw[1].winv = d; }
    h_tw[2].w = q - 1; h_tw[2].wq = (uint32_t)(((uint64_t)(q - 1) << 32) / q);
    { double d = (double)(q - 1) / q; while ((long double)d * q > (long double)(q - 1)) d = nextafter(d, 0.0); h_tw[2].winv = d; }
    Tw *d_tw;
    CK(cudaMalloc(&d_tw, sizeof h_tw));
    CK(cudaMemcpy(d_tw, h_tw, sizeof h_tw, cudaMemcpyHostToDevice));
    uint32_t *h_in = (uint32_t *)malloc(nthr * 64), *h_ref = (uint32_t *)malloc(nthr * 64), *h_out = (uint32_t *)malloc(nthr * 64);
    for (size_t i = 0; i < nthr * 16; i++) h_in[i] = (uint32_t)(rnd() % (4ull * q));
    h_in[0] = 4u * q - 1; h_in[1] = 0; h_in[2] = 2 * q; h_in[3] = 2 * q - 1;
    uint32_t *d_data;
    CK(cudaMalloc(&d_data, nthr * 64));
    run<0>(d_tw, d_data, h_in, h_ref, nthr, q, nullptr, sms);
    run<1>(d_tw, d_data, h_in, h_out, nthr, q, h_ref, sms);
    run<2>(d_tw, d_data, h_in, h_out, nthr, q, h_ref, sms);
    run<3>(d_tw, d_data, h_in, h_out, nthr, q, h_ref, sms);
    run<4>(d_tw, d_data, h_in, h_out, nthr, q, h_ref, sms);
    return 0;
}
