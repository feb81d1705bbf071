// seb_encode.cu — CKKS encode: fp32 slots -> scatter -> FP64 inverse FFT -> scale -> round -> int64.
//
// Restates ckks_encode_base (device/lib/ckks_common.c:105-215) with ifft_inpl
// (device/lib/fft.c:69-144) for a batch, one CTA (or, for n = 16384, one 2-CTA cluster) per
// ciphertext with the whole complex vector in shared memory.
//
// Bit-exactness rules (SURVEY.md §7 "Bit-exact FP64 encode"):
//   * every FP64 operation is individually rounded (__dadd_rn/__dsub_rn/__dmul_rn, never an FMA)
//     and the complex product is re = a*c - b*d, im = a*d + b*c, the order GCC emits for C99
//     complex multiplication on x86-64 without -ffast-math;
//   * twiddles are the host libm's cos/sin (table built once in seb_api.cu with the reference's own
//     expression, fft.c:27-45), not CUDA's;
//   * radix-16 passes are four fused radix-2 stages (radix-8 / three with ENC_E = 8), never a re-associated butterfly;
//   * C round() (half away from zero) and the x86 cvttsd2si result for out-of-range values.
#include <cooperative_groups.h>

#include "seb_kernels.h"

namespace cg = cooperative_groups;

#include "seb_encode.cuh"

// resident CTAs per SM the kernel is compiled for: 64 registers per thread at every degree (two
// 512-thread CTAs per SM for n = 4096)
#ifndef ENC_THREADS_PER_SM
#define ENC_THREADS_PER_SM (ENC_E == 16 ? 512 : 1024)
#endif
template <int LOGN, int CL>
struct EncOcc
{
    static constexpr int T    = (1 << LOGN) / CL / ENC_E;
    // 8 values per thread: 64 registers, 1024 threads per SM; 16 per thread: 128 registers, 512 threads per SM
    // (ENC_THREADS_PER_SM: A/B switch)
    static constexpr int TPS  = ENC_THREADS_PER_SM;
    static constexpr int MINB = T >= TPS ? 1 : TPS / T;
};

template <int LOGN, int CL>
__global__ void __launch_bounds__((1 << LOGN) / CL / ENC_E, EncOcc<LOGN, CL>::MINB)
    k_encode(const float *__restrict__ values, size_t v_stride, int vlen, const uint16_t *__restrict__ src_map,
             const double2 *__restrict__ tw, double n_inv, int64_t *__restrict__ pt, int *__restrict__ fail,
             uint32_t *__restrict__ mag)
{
    constexpr int N     = 1 << LOGN;
    constexpr int NL    = N / CL;
    constexpr int LOGNL = (CL == 2) ? LOGN - 1 : LOGN;
    constexpr int T     = NL / ENC_E;
    constexpr int RL    = enc_r(LOGNL, enc_npass(LOGNL) - 1);
    constexpr int LSL   = ENC_LR * (enc_npass(LOGNL) - 1);
    extern __shared__ __align__(16) double esm[];
    double *sre   = esm;
    double *sim   = esm + NL;
    float *svals  = reinterpret_cast<float *>(esm + 2 * NL);  // the message, zero padded, skewed (enc_vskew)

    const int t            = threadIdx.x;
    const size_t b         = blockIdx.x / CL;
    const uint32_t rank    = blockIdx.x % CL;
    const uint32_t pos0    = rank * NL;
    const float *vals      = values + b * v_stride;
    int64_t *dst           = pt + b * N;
    int bad                = 0;
    uint32_t mx            = 0;  // max |coefficient| this thread produced, clipped to 32 bits

    // stage the message with coalesced loads (all issued before the first store, so their latencies
    // overlap); the gather of pass 0 then reads shared memory
    {
        constexpr int PER = (N / 2) / T;
        float v[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) v[k] = (t + k * T) < vlen ? __ldg(vals + t + k * T) : 0.0f;
#pragma unroll
        for (int k = 0; k < PER; k++) svals[enc_vskew<LOGN>((uint32_t)(t + k * T))] = v[k];
    }
    __syncthreads();

    double xr[ENC_E], xi[ENC_E];
    EncRun<LOGN, LOGNL, 0>::run(xr, xi, sre, sim, t, pos0, svals, src_map, tw);

    if (CL == 1)
    {
#pragma unroll
        for (int i = 0; i < (ENC_E >> RL); i++)
        {
            const uint32_t g    = (uint32_t)t + (uint32_t)i * T;
            const uint32_t off  = g & ((1u << LSL) - 1u);
            const uint32_t base = ((g >> LSL) << (LSL + RL)) | off;
#pragma unroll
            for (int j = 0; j < (1 << RL); j++)
                dst[base | ((uint32_t)j << LSL)] = enc_finish(xr[i * (1 << RL) + j], n_inv, bad, mx);
        }
    }
    else
    {
        // n = 16384: each CTA of the pair holds one half after 13 local stages; the last stage
        // (tt = n/2, twiddle index 1) pairs element k of rank 0 with element k of rank 1 through
        // distributed shared memory.  Only real parts are needed afterwards.
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int i = 0; i < (ENC_E >> RL); i++)
        {
            const uint32_t g    = (uint32_t)t + (uint32_t)i * T;
            const uint32_t off  = g & ((1u << LSL) - 1u);
            const uint32_t base = ((g >> LSL) << (LSL + RL)) | off;
#pragma unroll
            for (int j = 0; j < (1 << RL); j++)
            {
                const uint32_t pos = base | ((uint32_t)j << LSL);
                enc_st(sre, sim, pos, xr[i * (1 << RL) + j], xi[i * (1 << RL) + j]);
            }
        }
        cluster.sync();
        const double *ore = cluster.map_shared_rank(sre, rank ^ 1u);
        const double *oim = cluster.map_shared_rank(sim, rank ^ 1u);
        const double2 s   = __ldg(tw + 1);
        // Each CTA finishes BOTH outputs (k and k + n/2) for half of the positions k - rank r takes k in [r NL/2,
        // (r+1) NL/2) - instead of one output for all of them: u is rank 0's element k, v is rank 1's, so a CTA reads
        // NL/2 partner elements instead of NL (a third less traffic through distributed shared memory, the slowest
        // access of the kernel).  One element per iteration: issuing eight elements' remote loads before the first use
        // measured slower (3.09 against 3.00 ms).
        for (uint32_t k = rank * (NL / 2) + t; k < (rank + 1) * (uint32_t)(NL / 2); k += T)
        {
            double ar, ai, br, bi;
            enc_ld(sre, sim, k, ar, ai);
            enc_ld(ore, oim, k, br, bi);
            // own = (ar, ai), partner = (br, bi); rank 0 holds u, rank 1 holds v
            const double ur = rank == 0 ? ar : br, ui = rank == 0 ? ai : bi;
            const double vr = rank == 0 ? br : ar, vi = rank == 0 ? bi : ai;
            dst[k]      = enc_finish(__dadd_rn(ur, vr), n_inv, bad, mx);  // Re(u + v)
            const double dr = __dsub_rn(ur, vr), di = __dsub_rn(ui, vi);  // u - v
            dst[NL + k] = enc_finish(__dsub_rn(__dmul_rn(dr, s.x), __dmul_rn(di, s.y)), n_inv, bad, mx);  // Re((u - v) s)
        }
        // the partner may still be reading our shared memory: an execution barrier is enough (its loads have returned by
        // the time it arrives - their values went into its stores), so no release/acquire fence here
        asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
    }
    if (bad) fail[b] = 1;
    if (mag)
    {
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if ((threadIdx.x & 31) == 0) atomicMax(mag + b, mx);
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
size_t seb_enc_tw_entries(size_t n) { return enc_tw_entries(n); }
void seb_host_build_enc_tw0(size_t n, double2 *tw) { enc_build_tw0(n, tw); }

template <int LOGN, int CL>
static cudaError_t encode_cfg()
{
    return cudaFuncSetAttribute(k_encode<LOGN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(double) * 2 * ((1 << LOGN) / CL) + sizeof(float) * EncVals<LOGN>::WORDS));
}

cudaError_t seb_encode_configure(int logn)
{
    switch (logn)
    {
        case 10: return encode_cfg<10, 1>();
        case 11: return encode_cfg<11, 1>();
        case 12: return encode_cfg<12, 1>();
        case 13: return encode_cfg<13, 1>();
        case 14: return encode_cfg<14, 2>();
        default: return cudaErrorInvalidValue;
    }
}

template <int LOGN, int CL>
static cudaError_t encode_launch(const float *values, size_t v_stride, int vlen, const uint16_t *src_map,
                                 const double2 *tw, double n_inv, int64_t *pt, int *fail, uint32_t *mag, int batch,
                                 cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3((unsigned)batch * CL);
    cfg.blockDim           = dim3((1 << LOGN) / CL / ENC_E);
    cfg.dynamicSmemBytes   = sizeof(double) * 2 * ((1 << LOGN) / CL) + sizeof(float) * EncVals<LOGN>::WORDS;
    cfg.stream             = st;
    cudaLaunchAttribute attr[1];
    attr[0].id               = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs                = attr;
    cfg.numAttrs             = 1;
    return cudaLaunchKernelEx(&cfg, k_encode<LOGN, CL>, values, v_stride, vlen, src_map, tw, n_inv, pt, fail, mag);
}

cudaError_t seb_launch_encode(int logn, const float *values, size_t v_stride, int vlen, const uint16_t *src_map,
                              const double2 *tw, double n_inv, int64_t *pt, int *fail, uint32_t *mag, int batch,
                              cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
    switch (logn)
    {
        case 10: return encode_launch<10, 1>(values, v_stride, vlen, src_map, tw, n_inv, pt, fail, mag, batch, st);
        case 11: return encode_launch<11, 1>(values, v_stride, vlen, src_map, tw, n_inv, pt, fail, mag, batch, st);
        case 12: return encode_launch<12, 1>(values, v_stride, vlen, src_map, tw, n_inv, pt, fail, mag, batch, st);
        case 13: return encode_launch<13, 1>(values, v_stride, vlen, src_map, tw, n_inv, pt, fail, mag, batch, st);
        case 14: return encode_launch<14, 2>(values, v_stride, vlen, src_map, tw, n_inv, pt, fail, mag, batch, st);
        default: return cudaErrorInvalidValue;
    }
}
