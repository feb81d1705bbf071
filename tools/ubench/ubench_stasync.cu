// ubench_stasync.cu — checks the DSMEM exchange primitive considered for the 2-CTA cluster NTT: remote stores with
// st.async ... mbarrier::complete_tx::bytes into the partner CTA, which waits on its OWN mbarrier for the expected bytes
// (no cluster-wide barrier, no release fence after the stores).  Prints ok/mismatch.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_stasync ubench_stasync.cu
#include <cstdint>
#include <cstdio>
#include <vector>
__global__ void __cluster_dims__(2, 1, 1) k(uint32_t *out)
{
    extern __shared__ __align__(16) uint32_t sm[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + 1024);
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(mbar);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(blockDim.x * 16) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    uint32_t rmb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rmb) : "r"(mb), "r"(rank ^ 1u));
    for (int j = 0; j < 4; j++)
    {
        uint32_t dst;
        const uint32_t la = (uint32_t)__cvta_generic_to_shared(sm + j * blockDim.x + threadIdx.x);
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(la), "r"(rank ^ 1u));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(dst),
                     "r"(1000u * rank + 4u * threadIdx.x + j), "r"(rmb)
                     : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(mb), "r"(0)
                     : "memory");
    for (int j = 0; j < 4; j++) out[(blockIdx.x * 4 + j) * blockDim.x + threadIdx.x] = sm[j * blockDim.x + threadIdx.x];
}
int main()
{
    const int T = 256, clusters = 1000;
    uint32_t *d;
    cudaMalloc(&d, (size_t)clusters * 2 * 4 * T * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 + 16);
    k<<<clusters * 2, T, 4096 + 16>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint32_t> h((size_t)clusters * 2 * 4 * T);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int b = 0; b < clusters * 2; b++)
        for (int j = 0; j < 4; j++)
            for (int t = 0; t < T; t++) bad += h[((size_t)b * 4 + j) * T + t] != 1000u * ((b & 1) ^ 1) + 4u * t + j;
    printf("st.async + mbarrier exchange: %s (%zu mismatches), cuda: %s\n", bad ? "MISMATCH" : "ok", bad, cudaGetErrorString(e));
    return bad != 0;
}
