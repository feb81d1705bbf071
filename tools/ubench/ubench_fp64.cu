// ubench_fp64.cu — FP64 pipe rates on sm_100a for the operations the bit-exact encode is made of: it may not use
// FMA (every product and sum individually rounded), so DADD and DMUL are what count, not the DFMA peak.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64 ubench_fp64.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int OP>
__global__ void __launch_bounds__(256) k(double *out, double a, double b, int iters)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a + threadIdx.x + i;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (OP == 0) x[i] = __dadd_rn(x[i], b);
            if (OP == 1) x[i] = __dmul_rn(x[i], b);
            if (OP == 2) x[i] = __fma_rn(x[i], b, a);
            if (OP == 3) x[i] = __dadd_rn(__dmul_rn(x[i], b), a);  // the encode's pattern: DMUL then DADD
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
static void run(const char *name, int sms, double ghz, int ops_per)
{
    double *d;
    const int blocks = sms * 8, iters = 4096;
    cudaMalloc(&d, (size_t)blocks * 256 * 8);
    k<OP><<<blocks, 256>>>(d, 1.0000001, 0.9999999, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(d, 1.0000001, 0.9999999, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)blocks * 256 * iters * 8 * ops_per;
    printf("%-28s %7.2f T lane-instr/s = %5.1f lanes/clk/SM (at %.3f GHz)\n", name, lane_ops / (ms * 1e-3) / 1e12,
           lane_ops / (ms * 1e-3) / sms / (ghz * 1e9), ghz);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = 1.965;
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("DADD", p.multiProcessorCount, ghz, 1);
    run<1>("DMUL", p.multiProcessorCount, ghz, 1);
    run<2>("DFMA", p.multiProcessorCount, ghz, 1);
    run<3>("DMUL+DADD (dependent pair)", p.multiProcessorCount, ghz, 2);
    return 0;
}
