// ubench_overlap.cu — do an ALU-bound kernel (Keccak-f) and an FMA-bound kernel (lazy NTT butterflies)
// overlap when they are co-resident on the same SMs (two streams, bounded grids)?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../seal-embedded_b200/csrc -o ubench_overlap ubench_overlap.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "seb_keccak.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(128) k_keccak(uint64_t *out, uint64_t seed, int reps)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = seed * (i + 1) + tid;
    uint64_t acc = 0;
#pragma unroll 1
    for (int r = 0; r < reps; r++)
    {
        seb_keccak_f1600(a);
        acc ^= a[3];
        a[7] ^= r;
    }
    out[tid] = acc ^ a[0];
}

__device__ __forceinline__ void bfly(uint32_t &x, uint32_t &y, uint32_t w, uint32_t wq, uint32_t q, uint32_t two_q)
{
    const uint32_t u = min(x, x - two_q);
    const uint32_t t = y * w - __umulhi(y, wq) * q;
    x                = u + t;
    y                = u - t + two_q;
}

__global__ void __launch_bounds__(256) k_bfly(const uint2 *__restrict__ tws, uint32_t *__restrict__ data, uint32_t q, int iters)
{
    const size_t tid     = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t two_q = 2 * q;
    uint32_t x[16];
    uint2 tw[15];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = data[tid * 16 + i];
#pragma unroll
    for (int i = 0; i < 15; i++) tw[i] = tws[(threadIdx.x & 31) * 15 + i];
#pragma unroll 1
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const int half = 8 >> r;
#pragma unroll
            for (int m = 0; m < (1 << r); m++)
#pragma unroll
                for (int t = 0; t < half; t++)
                    bfly(x[m * 2 * half + t], x[m * 2 * half + t + half], tw[(1 << r) - 1 + m].x, tw[(1 << r) - 1 + m].y, q, two_q);
        }
#pragma unroll
    for (int i = 0; i < 16; i++) data[tid * 16 + i] = x[i];
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms    = prop.multiProcessorCount;
    const uint32_t q = 1053818881u;
    uint2 h_tw[32 * 15];
    for (int i = 0; i < 32 * 15; i++)
    {
        const uint32_t w = (uint32_t)((i * 2654435761ull + 12345) % q);
        h_tw[i]          = make_uint2(w, (uint32_t)(((uint64_t)w << 32) / q));
    }
    uint2 *d_tw;
    CK(cudaMalloc(&d_tw, sizeof h_tw));
    CK(cudaMemcpy(d_tw, h_tw, sizeof h_tw, cudaMemcpyHostToDevice));
    uint32_t *d_data;
    uint64_t *d_k;
    CK(cudaMalloc(&d_data, (size_t)sms * 8 * 256 * 64));
    CK(cudaMemset(d_data, 1, (size_t)sms * 8 * 256 * 64));
    CK(cudaMalloc(&d_k, (size_t)sms * 16 * 128 * 8));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f0, f1;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);

    printf("%s: Keccak CTAs/SM (128 thr) x butterfly CTAs/SM (256 thr); times in ms\n", prop.name);
    const int kc[] = {2, 3, 4, 6, 8}, bc[] = {1, 2, 3, 4};
    for (int ki = 0; ki < 5; ki++)
        for (int bi = 0; bi < 4; bi++)
        {
            const int gk = sms * kc[ki], gb = sms * bc[bi];
            // work sized so that each kernel alone takes a few ms and the same TOTAL work in every config
            const int reps  = 4096 / kc[ki];   // keccak perms per thread
            const int iters = 8192 / bc[bi];   // radix-16 passes per thread
            float t_k, t_b, t_both;
            k_keccak<<<gk, 128, 0, s1>>>(d_k, 3, reps);
            k_bfly<<<gb, 256, 0, s2>>>(d_tw, d_data, q, iters);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0, s1);
            k_keccak<<<gk, 128, 0, s1>>>(d_k, 3, reps);
            cudaEventRecord(e1, s1);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&t_k, e0, e1);
            cudaEventRecord(e0, s2);
            k_bfly<<<gb, 256, 0, s2>>>(d_tw, d_data, q, iters);
            cudaEventRecord(e1, s2);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&t_b, e0, e1);
            // concurrent: both start together
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0, 0);
            k_keccak<<<gk, 128, 0, s1>>>(d_k, 3, reps);
            k_bfly<<<gb, 256, 0, s2>>>(d_tw, d_data, q, iters);
            cudaEventRecord(f1, s2);
            cudaStreamWaitEvent(s1, f1, 0);
            cudaEventRecord(e1, s1);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&t_both, e0, e1);
            printf("keccak %d CTA/SM + bfly %d CTA/SM: keccak alone %7.3f  bfly alone %7.3f  sum %7.3f  together %7.3f  (%.2fx vs serial)\n",
                   kc[ki], bc[bi], t_k, t_b, t_k + t_b, t_both, (t_k + t_b) / t_both);
        }
    return 0;
}
