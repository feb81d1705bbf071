// How fast does ONE warp per SM sub-partition run Keccak-f[1600], and does anything help it?
// (configuration D's uniform sampler: 16384 sequential sponges = 512 warps for 592 sub-partitions.)
//   full   : W warps per SM sub-partition, 32 sponges per warp (one per lane)
//   half   : W warps, lanes 16..31 idle (16 sponges per warp): does a half-empty warp cost half an ALU slot?
//   dual   : W warps, TWO independent sponges per lane interleaved in one instruction stream (ILP 2)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../seal-embedded_b200/csrc -o ubench_lonewarp ubench_lonewarp.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "seb_keccak.cuh"

template <int MODE>
__global__ void __launch_bounds__(32) k(uint32_t *out, int perms)
{
    const uint32_t t = blockIdx.x * 32 + threadIdx.x;
    if (MODE == 1 && (threadIdx.x & 16)) return;
    uint32_t lo[25], hi[25], lo2[25], hi2[25];
#pragma unroll
    for (int i = 0; i < 25; i++)
    {
        lo[i] = t * 0x9E3779B9u + i, hi[i] = t * 0x7F4A7C15u - i;
        lo2[i] = ~lo[i], hi2[i] = hi[i] ^ 0x55u;
    }
    for (int p = 0; p < perms; p++)
    {
#pragma unroll 1
        for (int r = 0; r < 24; r++)
        {
            seb_keccak_round<false>(lo, hi, r);
            if (MODE == 2) seb_keccak_round<false>(lo2, hi2, r);
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 25; i++) x ^= lo[i] ^ hi[i] ^ (MODE == 2 ? lo2[i] ^ hi2[i] : 0u);
    out[t] = x;
}

template <int MODE>
static void run(const char *name, int warps_per_smsp, int smsps, uint32_t *out)
{
    const int blocks = warps_per_smsp * smsps, perms = 400;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 4; rep++)
    {
        cudaEventRecord(a);
        k<MODE><<<blocks, 32>>>(out, perms);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    const double sponges = (double)blocks * (MODE == 1 ? 16 : MODE == 2 ? 64 : 32);
    printf("%-5s %d warp(s)/sub-partition: %8.3f ms for %d permutations in sequence = %6.2f us each, %7.3f G Keccak-f/s\n", name,
           warps_per_smsp, best, perms, best * 1e3 / perms, sponges * perms / (best * 1e-3) / 1e9);
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const int smsps = pr.multiProcessorCount * 4;
    uint32_t *out;
    cudaMalloc(&out, 64 << 20);
    for (int w : {1, 2, 4, 8})
    {
        run<0>("full", w, smsps, out);
        run<1>("half", w, smsps, out);
        run<2>("dual", w, smsps, out);
    }
    cudaFree(out);
    return cudaGetLastError() != cudaSuccess;
}
