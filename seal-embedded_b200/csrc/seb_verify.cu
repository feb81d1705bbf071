// seb_verify.cu — the receiving side of the round trip, batched, so that parity and property tests can
// check EVERY ciphertext of a full-size batch on the GPU instead of a handful on the host:
//
//   k_intt             : inverse negacyclic NTT (the inverse of ntt_inpl, device/lib/ntt.c:124-189;
//                        the reference's own inverse is device/lib/intt.c:226-501), bit-reversed
//                        order in, natural order out, scaled by n^-1, canonical residues
//   k_decrypt_decode   : ckks_decrypt (device/test/ckks_tests_common.c:132-171): pt^ = c0 + c1 (.) s^,
//                        then intt, then ckks_decode (ckks_tests_common.c:59-118: centred lift, /scale,
//                        forward FFT fft_inpl device/lib/fft.c:146-213, gather through the index map)
//
// This is SURVEY.md 8(f) row 3 ("GPU INTT/decode verifier").  It is a verifier, not a hot path: one
// CTA per polynomial, radix-2 stages with a barrier each, exact Barrett arithmetic, FP64 FFT in an
// L2-resident global scratch slice per CTA.  Decoded values carry the reference's tolerance (0.1
// against the message, device/test/ckks_tests_common.c:228), not bit-exactness.
#include "seb_kernels.h"

#define SEB_VERIFY_THREADS 256

// x * w mod q, exact, for a Shoup pair (w, floor(w*2^32/q)) and x < 2^32
__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint2 w, uint32_t q)
{
    return seb_csub(seb_mul_shoup_lazy(x, w.x, w.y, q), q);
}

// In-place inverse transform of the n residues in `v` (shared memory) by the whole CTA.
// iroots[h + j] = (psi^bitrev(h + j))^-1 as Shoup pairs, i.e. the inverses of the forward table
// (device/lib/ntt.c:40-52) at the same indices; ninv = n^-1 mod q.
__device__ __forceinline__ void intt_smem(uint32_t *v, const int n, const uint2 *__restrict__ iroots, const uint2 ninv,
                                          const uint32_t q)
{
    // undo the forward stages last to first: (x, y) -> (x + y, (x - y) * w^-1)
    for (int lt = 0, h = n >> 1; h >= 1; lt++, h >>= 1)
    {
        const int tt = 1 << lt;
        __syncthreads();
        for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x)
        {
            const int j     = i >> lt;
            const int k     = (j << (lt + 1)) + (i - (j << lt));
            const uint32_t x = v[k], y = v[k + tt];
            v[k]            = seb_csub(x + y, q);
            v[k + tt]       = mul_shoup(x + q - y, __ldg(iroots + h + j), q);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = mul_shoup(v[i], ninv, q);
    __syncthreads();
}

// polys [batch][np][n], in place (polynomial k uses prime k % np): the inverse of k_ntt_forward
__global__ void __launch_bounds__(SEB_VERIFY_THREADS)
    k_intt(uint32_t *__restrict__ polys, const uint2 *__restrict__ iroots, const uint2 *__restrict__ ninvs,
           const __grid_constant__ SebModuli mods, int n, int np, size_t npolys)
{
    extern __shared__ __align__(16) uint32_t vs[];
    for (size_t poly = blockIdx.x; poly < npolys; poly += gridDim.x)
    {
        const int p     = (int)(poly % (size_t)np);
        const uint32_t q = mods.m[p].q;
        uint32_t *data  = polys + poly * (size_t)n;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) vs[i] = data[i];
        intt_smem(vs, n, iroots + (size_t)p * n, __ldg(ninvs + p), q);
        for (int i = threadIdx.x; i < n; i += blockDim.x) data[i] = vs[i];
    }
}

// ct [batch][np][2][n]; s_hat [np][n] = ntt(s) in the reference's (bit-reversed) order;
// tw = the IFFT twiddle table (cos, -sin)(2 pi bitrev(i) / 2n): the forward FFT uses its conjugate;
// index_map as ckks_calc_index_map (ckks_common.c:32-68); work: gridDim.x slices of n double2;
// values_out [batch][vlen]
__global__ void __launch_bounds__(SEB_VERIFY_THREADS)
    k_decrypt_decode(const uint32_t *__restrict__ ct, const uint32_t *__restrict__ s_hat,
                     const uint2 *__restrict__ iroots, const uint2 *__restrict__ ninvs,
                     const double2 *__restrict__ tw, const uint16_t *__restrict__ index_map,
                     const __grid_constant__ SebModuli mods, int n, int np, int prime, double scale,
                     double2 *__restrict__ work, int vlen, float *__restrict__ values_out, size_t batch)
{
    extern __shared__ __align__(16) uint32_t vs[];
    const SebModulus m = mods.m[prime];
    const uint32_t q   = m.q;
    double2 *x         = work + (size_t)blockIdx.x * n;
    for (size_t b = blockIdx.x; b < batch; b += gridDim.x)
    {
        const uint32_t *c0 = ct + ((b * np + prime) * 2) * (size_t)n;
        const uint32_t *c1 = c0 + n;
        __syncthreads();
        // ckks_decrypt: c0 + c1 (.) ntt(s)
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint32_t prod = seb_barrett64((uint64_t)c1[i] * __ldg(s_hat + (size_t)prime * n + i), m);
            vs[i]               = seb_csub(prod + c0[i], q);
        }
        intt_smem(vs, n, iroots + (size_t)prime * n, __ldg(ninvs + prime), q);
        // ckks_decode: representative in (-q/2, q/2], divided by the scale
        for (int i = threadIdx.x; i < n; i += blockDim.x)
        {
            const uint32_t v = vs[i];
            const double d   = (v > q / 2) ? -(double)(q - v) : (double)v;
            x[i]             = make_double2(d / scale, 0.0);
        }
        // fft_inpl (fft.c:146-213): h groups of 2*tt, root e^{+i theta_(h+j)} applied to the upper half
        for (int h = 1, lt = 31 - __clz(n >> 1); lt >= 0; h <<= 1, lt--)
        {
            const int tt = 1 << lt;
            __syncthreads();
            for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x)
            {
                const int j     = i >> lt;
                const int k     = (j << (lt + 1)) + (i - (j << lt));
                const double2 s = __ldg(tw + h + j);  // (cos, -sin)
                const double2 u = x[k], a = x[k + tt];
                const double vr = a.x * s.x + a.y * s.y;   // a * (cos + i sin)
                const double vi = a.y * s.x - a.x * s.y;
                x[k]            = make_double2(u.x + vr, u.y + vi);
                x[k + tt]       = make_double2(u.x - vr, u.y - vi);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < vlen; i += blockDim.x)
            values_out[b * (size_t)vlen + i] = (float)x[__ldg(index_map + i)].x;
    }
}

// ---------------------------------------------------------------------------------------------
// per-item 64-bit digest of a batch of word streams, so that a FULL batch can be compared with the compiled
// reference item by item without moving 6 GB to the host:
//   digest(item) = sum over i of mix64((i << 32) | word_i)  mod 2^64,   mix64 = the splitmix64 finaliser
// (position-keyed, so a changed, moved or swapped word changes the sum; oracle/ref_shim.c and oracle/se_oracle.c
// compute the same function on the CPU).  One CTA per item, 128-bit loads, warp-shuffle reduction.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t seb_mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_digest(const uint32_t *__restrict__ words, size_t words_per_item, size_t items,
                                                uint64_t *__restrict__ digests)
{
    __shared__ uint64_t s_part[8];
    const size_t b = blockIdx.x;
    if (b >= items) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(words + b * words_per_item);
    uint64_t acc     = 0;
    for (size_t i = threadIdx.x; i < words_per_item / 4; i += blockDim.x)
    {
        const uint4 v    = src[i];
        const uint64_t k = (uint64_t)(4 * i) << 32;
        acc += seb_mix64(k | v.x) + seb_mix64((k + (1ULL << 32)) | v.y) + seb_mix64((k + (2ULL << 32)) | v.z) +
               seb_mix64((k + (3ULL << 32)) | v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint64_t t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_part[w];
        digests[b] = t;
    }
}

cudaError_t seb_launch_digest(const uint32_t *words, size_t words_per_item, size_t items, uint64_t *digests, cudaStream_t st)
{
    if (items == 0) return cudaSuccess;
    if (words_per_item % 4 != 0 || items > 0x7FFFFFFFu) return cudaErrorInvalidValue;
    k_digest<<<(unsigned)items, 256, 0, st>>>(words, words_per_item, items, digests);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// optional packed wire form: residues are < q < 2^30, so 16 of them fit 15 words (30 bits each, little endian:
// residue i of a group occupies bits 30i .. 30i+29 of the group's 480 bits).  6.25 % fewer bytes on the PCIe link,
// which is what bounds the host-pointer path (profiles/README.md); the full form stays the default.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack30(const uint4 *__restrict__ in, uint32_t *__restrict__ out, size_t groups)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    uint32_t r[16];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const uint4 v = seb_ldg_stream(in + 4 * g + k);
        r[4 * k] = v.x, r[4 * k + 1] = v.y, r[4 * k + 2] = v.z, r[4 * k + 3] = v.w;
    }
    uint32_t *dst = out + 15 * g;
#pragma unroll
    for (int w = 0; w < 15; w++)
    {
        // word w holds bits 32w .. 32w+31: the top of residue i = 32w / 30 and the bottom of residue i + 1
        const int i = (32 * w) / 30, sh = 32 * w - 30 * i;  // residue i contributes its bits sh .. 29
        dst[w] = (r[i] >> sh) | (r[i + 1] << (30 - sh));
    }
}

__global__ void __launch_bounds__(256) k_unpack30(const uint32_t *__restrict__ in, uint4 *__restrict__ out, size_t groups)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 15; k++) w[k] = in[15 * g + k];
    w[15] = 0;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
    {
        const int k = (30 * i) / 32, sh = 30 * i - 32 * k;  // residue i starts at bit sh of word k
        const uint64_t two = ((uint64_t)w[k + (k < 15 ? 1 : 0)] << 32) | w[k];
        r[i] = (uint32_t)(two >> sh) & 0x3FFFFFFFu;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) out[4 * g + k] = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
}

cudaError_t seb_launch_pack30(const uint32_t *in, uint32_t *out, size_t words, cudaStream_t st)
{
    if (words == 0) return cudaSuccess;
    if (words % 16) return cudaErrorInvalidValue;
    const size_t groups = words / 16;
    k_pack30<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(in), out, groups);
    return cudaGetLastError();
}

cudaError_t seb_launch_unpack30(const uint32_t *in, uint32_t *out, size_t words, cudaStream_t st)
{
    if (words == 0) return cudaSuccess;
    if (words % 16) return cudaErrorInvalidValue;
    const size_t groups = words / 16;
    k_unpack30<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(in, reinterpret_cast<uint4 *>(out), groups);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// in-run integer-issue ceilings (bench.py's roofline for the kernels that are not HBM-bound)
// ---------------------------------------------------------------------------------------------
// Register-only loops of the two inner operations of the path, on enough resident warps to saturate the pipes:
//   k_ceiling_keccak : 24 full rounds of Keccak-f[1600] in the cheapest form the samplers use (seb_keccak.cuh, bit-interleaved
//                      state: 174 ALU-pipe operations per round; the plain form of the ternary / uniform kernels is 180)
//   k_ceiling_bfly   : Harvey/Shoup lazy butterflies exactly as the NTT passes run them (seb_ntt.cuh: seb_bfly, radix-16
//                      register groups: IMAD.HI + 2 IMAD on the FMA pipe, 3 operations on the ALU pipe)
// What they reach IS the ceiling the sampler / NTT kernels are measured against ("frac" of an ALU- or FMA-bound kernel).
#include "seb_keccak.cuh"
#include "seb_ntt.cuh"

__global__ void __launch_bounds__(128) k_ceiling_keccak(uint64_t *__restrict__ out, int iters)
{
    uint32_t e[25], o[25];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 25; i++) e[i] = t * 0x9E3779B9u + (uint32_t)i, o[i] = t * 0x7F4A7C15u - (uint32_t)i;
    for (int it = 0; it < iters; it++)
    {
#pragma unroll 1
        for (int round = 0; round < 24; round++) seb_keccak_round_il<25>(e, o, round);
    }
    uint32_t x = 0, y = 0;
#pragma unroll
    for (int i = 0; i < 25; i++) x ^= e[i], y ^= o[i];
    out[t] = ((uint64_t)y << 32) | x;
}

__global__ void __launch_bounds__(256) k_ceiling_bfly(uint32_t *__restrict__ out, const seb_oct *__restrict__ tw, uint32_t q,
                                                     int iters)
{
    uint32_t x[SEB_E];
    uint2 w[15];
    const uint32_t t     = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t two_q = 2 * q;
#pragma unroll
    for (int i = 0; i < SEB_E; i++) x[i] = (t * 2654435761u + (uint32_t)i * 40503u) % q;
    // 15 roots of a radix-16 register group, fetched once (heap slots 1..15 of the first group of the table)
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const seb_oct o = seb_ldg256(tw + k);
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (4 * k + c >= 1) w[4 * k + c - 1] = make_uint2(o.v[2 * c], o.v[2 * c + 1]);
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++)
    {
        // 32 butterflies: the four stages of a radix-16 pass on the thread's 16 registers
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            const int half = 8 >> r;
#pragma unroll
            for (int m = 0; m < (1 << r); m++)
#pragma unroll
                for (int k = 0; k < half; k++) seb_bfly(x[m * 2 * half + k], x[m * 2 * half + k + half], w[(1 << r) - 1 + m], q, two_q);
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < SEB_E; i++) acc ^= x[i];
    out[t] = acc;
}

// keccak_per_s / bfly_per_s: best of 3 timed launches each (CUDA events on `st`); scratch: >= ctas * 256 * 8 bytes
cudaError_t seb_measure_ceilings(int sms, const seb_oct *tw, uint32_t q, void *scratch, double *keccak_per_s, double *bfly_per_s,
                                 cudaStream_t st)
{
    cudaEvent_t e0, e1;
    cudaError_t err = cudaEventCreate(&e0);
    if (err != cudaSuccess) return err;
    if ((err = cudaEventCreate(&e1)) != cudaSuccess) return err;
    const int kc = sms * 16, kt = 128, kiters = 64;   // 16 CTAs x 128 threads per SM
    const int bc = sms * 8, bt = 256, biters = 2048;  // 8 CTAs x 256 threads per SM wanted (registers decide)
    double best_k = 0, best_b = 0;
    for (int rep = 0; rep < 4 && err == cudaSuccess; rep++)
    {
        float ms = 0;
        cudaEventRecord(e0, st);
        k_ceiling_keccak<<<kc, kt, 0, st>>>(static_cast<uint64_t *>(scratch), kiters);
        cudaEventRecord(e1, st);
        if ((err = cudaEventSynchronize(e1)) != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms > 0) best_k = fmax(best_k, (double)kc * kt * kiters / (ms * 1e-3));
        cudaEventRecord(e0, st);
        k_ceiling_bfly<<<bc, bt, 0, st>>>(static_cast<uint32_t *>(scratch), tw, q, biters);
        cudaEventRecord(e1, st);
        if ((err = cudaEventSynchronize(e1)) != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms > 0) best_b = fmax(best_b, (double)bc * bt * biters * 32.0 / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (err == cudaSuccess) err = cudaGetLastError();
    *keccak_per_s = best_k;
    *bfly_per_s   = best_b;
    return err;
}

cudaError_t seb_verify_configure(int n)
{
    // The attribute belongs to the kernel, not to a context: contexts of several degrees live in one
    // process, so it is always raised to the largest degree's need (a later, smaller context must not
    // lower it under an earlier one's launches).
    (void)n;
    const int max_bytes = 4 * 16384;
    cudaError_t e = cudaFuncSetAttribute(k_intt, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_decrypt_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes);
    return e;
}

cudaError_t seb_launch_intt(uint32_t *polys, const uint2 *iroots, const uint2 *ninvs, const SebModuli &mods, int n,
                            int np, size_t npolys, int max_ctas, cudaStream_t st)
{
    if (npolys == 0) return cudaSuccess;
    const size_t grid = npolys < (size_t)max_ctas ? npolys : (size_t)max_ctas;
    k_intt<<<(unsigned)grid, SEB_VERIFY_THREADS, 4 * (size_t)n, st>>>(polys, iroots, ninvs, mods, n, np, npolys);
    return cudaGetLastError();
}

cudaError_t seb_launch_decrypt_decode(const uint32_t *ct, const uint32_t *s_hat, const uint2 *iroots,
                                      const uint2 *ninvs, const double2 *tw, const uint16_t *index_map,
                                      const SebModuli &mods, int n, int np, int prime, double scale, double2 *work,
                                      int work_ctas, int vlen, float *values_out, size_t batch, cudaStream_t st)
{
    if (batch == 0) return cudaSuccess;
    const size_t grid = batch < (size_t)work_ctas ? batch : (size_t)work_ctas;
    k_decrypt_decode<<<(unsigned)grid, SEB_VERIFY_THREADS, 4 * (size_t)n, st>>>(
        ct, s_hat, iroots, ninvs, tw, index_map, mods, n, np, prime, scale, work, vlen, values_out, batch);
    return cudaGetLastError();
}
