// seb_encrypt.cu — batched RNS NTT and the fused per-(ciphertext, prime) encrypt kernels.
//
//   k_ntt_forward   : ntt_inpl (device/lib/ntt.c:168-189) over a batch of [np][n] polynomials
//   k_encrypt_asym  : one prime of ckks_encode_encrypt_asym (device/lib/ckks_asym.c:205-286):
//                     expand u / reduce e1 / reduce (m+e0) on load, three NTTs through shared
//                     memory sharing every twiddle fetch, then
//                       c1 = pk1 (.) ntt(u) + ntt(e1),  c0 = pk0 (.) ntt(u) + ntt(m+e0)
//                     straight from registers with 256-bit stores.
//   k_encrypt_sym   : one prime of ckks_encode_encrypt_sym (device/lib/ckks_sym.c:199-301) with
//                     ntt(s) precomputed at setup:  c0 = -(a (.) ntt(s)) + ntt(m+e)
//
// One CTA of n/16 threads per polynomial slot; grid = (prime, item low, item high) so the CTAs that
// re-read one ciphertext's m/e/u are scheduled together and hit L2, and no thread divides by np.
#include <stdlib.h>

#include "seb_kernels.h"
#include "seb_ntt.cuh"


// ---------------------------------------------------------------------------------------------
// on-load conversions.  Their results feed lazy butterflies, which accept ANY representative of
// the residue below 4q, so none of them reduces to the canonical range (the reference's functions
// are cited for the value being represented, not for the representative).
// ---------------------------------------------------------------------------------------------
// 2-bit field t in {0,1,2} standing for t - 1 (device/lib/sample.c:98-116) -> t + q - 1 in {q-1, q, q+1}
__device__ __forceinline__ uint32_t expand_ternary(const uint8_t *__restrict__ u, uint32_t pos, uint32_t q)
{
    const uint32_t byte = __ldg(u + (pos >> 2));
    return ((byte >> (6 - 2 * (pos & 3))) & 3u) + (q - 1u);
}
// small signed e (device/lib/ckks_common.c:259-265) -> e + q in (0, 2q)
__device__ __forceinline__ uint32_t reduce_small(const int8_t *__restrict__ e, uint32_t pos, uint32_t q)
{
    return (uint32_t)((int)__ldg(e + pos)) + q;
}
// (m + e) int64 (device/lib/ckks_common.c:224-245).
// small == true promises |m + e| < 2q for every coefficient of this ciphertext (decided per CTA from
// the encode kernel's max |m|), so the low words carry the whole value and
// min(x, x + 2q) over unsigned words maps it into [0, 2q): one VIADDMNMX.  Otherwise: exact 64-bit
// Barrett reduction of |x| and a conditional negation.
template <bool SMALL>
__device__ __forceinline__ uint32_t reduce_pte(const int64_t *__restrict__ pt, const int8_t *__restrict__ e,
                                               uint32_t pos, const SebModulus &m)
{
    if (SMALL)
    {
        const uint32_t lo = __ldg(reinterpret_cast<const uint32_t *>(pt + pos));  // little endian: low word
        const uint32_t x  = lo + (uint32_t)((int)__ldg(e + pos));
        return min(x, x + m.two_q);
    }
    const uint64_t x  = (uint64_t)__ldg(pt + pos) + (uint64_t)(int64_t)__ldg(e + pos);
    const bool neg    = (int64_t)x < 0;
    const uint64_t ax = neg ? (uint64_t)0 - x : x;
    const uint32_t r  = seb_barrett64(ax, m);
    return neg ? m.q - r : r;
}
// CBD samples lie in [-21, 21] (device/lib/sample.c:278-284)
#define SEB_E_BOUND 21u

// ---------------------------------------------------------------------------------------------
// NTT only
// ---------------------------------------------------------------------------------------------
struct LoadPlain
{
    const uint32_t *src;
    __device__ __forceinline__ uint32_t operator()(int, uint32_t pos) const { return seb_ldg_stream(src + pos); }
};

// item index of this CTA for a grid built by seb_grid()
__device__ __forceinline__ size_t seb_item() { return (size_t)blockIdx.z * gridDim.y + blockIdx.y; }

// resident CTAs per SM the NTT-only kernel is compiled for, per degree: the best of the sweep in
// profiles/r01_ubench_ntt_occupancy.txt (48 registers for n <= 4096, 32 above)
template <int LOGN>
struct NttOcc
{
    static constexpr int MINB = LOGN == 10 ? 20 : LOGN == 11 ? 10 : LOGN == 12 ? 5 : LOGN == 13 ? 4 : 2;
};

template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / SEB_E, NttOcc<LOGN>::MINB)
    k_ntt_forward(uint32_t *__restrict__ polys, const seb_oct *__restrict__ roots,
                  const __grid_constant__ SebModuli mods, int np, size_t items)
{
    constexpr int N = 1 << LOGN;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t         = threadIdx.x;
    const size_t b      = seb_item();
    if (b >= items) return;
    const int p         = (int)blockIdx.x;
    const SebModulus &m = mods.m[p];
    uint32_t *data      = polys + (b * np + p) * N;

    uint32_t x[1][SEB_E];
    LoadPlain ld{data};
    seb_ntt_forward<LOGN, 1>(x, smem, t, roots + (size_t)p * NttTwSize<LOGN>::OCTS, m.q, m.two_q, ld);

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
            for (int c = 0; c < 8; c++) v.v[c] = seb_final_reduce(x[0][i * O::RUN + 8 * k + c], m.q, m.two_q);
            seb_stg256_stream(dst + k, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// asymmetric encrypt, one (ciphertext, prime)
// ---------------------------------------------------------------------------------------------
template <bool SMALL>
struct LoadAsym
{
    const uint8_t *u;
    const int8_t *e0;
    const int8_t *e1;
    const int64_t *pt;
    SebModulus m;
    __device__ __forceinline__ uint32_t operator()(int p, uint32_t pos) const
    {
        if (p == 0) return expand_ternary(u, pos, m.q);
        if (p == 1) return reduce_small(e1, pos, m.q);
        return reduce_pte<SMALL>(pt, e0, pos, m);
    }
};

// x*w mod q in [0,q) for a Shoup pair, then + y (y lazy in [0,4q)) mod q
__device__ __forceinline__ uint32_t mul_add_final(uint32_t x, uint2 w, uint32_t y, uint32_t q, uint32_t two_q)
{
    const uint32_t prod = seb_csub(seb_mul_shoup_lazy(x, w.x, w.y, q), q);
    return seb_csub(prod + seb_final_reduce(y, q, two_q), q);
}

// c0/c1 for 8 consecutive coefficients starting at pos; xe/xp = ntt(e1)/ntt(m+e0) values (lazy);
// k0/k1 point at the two key octs (4 Shoup pairs each) covering these coefficients, `ks` apart.
__device__ __forceinline__ void asym_store8(const uint32_t (&xu)[8], const uint32_t (&xe)[8], const uint32_t (&xp)[8],
                                            const seb_oct *__restrict__ k0, const seb_oct *__restrict__ k1,
                                            const int ks, uint32_t *__restrict__ c0, uint32_t *__restrict__ c1,
                                            uint32_t pos, uint32_t q, uint32_t two_q)
{
    seb_oct v0, v1;
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const seb_oct a = seb_ldg256(k0 + (size_t)h * ks);
        const seb_oct b = seb_ldg256(k1 + (size_t)h * ks);
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            v0.v[4 * h + c] = mul_add_final(xu[4 * h + c], make_uint2(a.v[2 * c], a.v[2 * c + 1]), xp[4 * h + c], q, two_q);
            v1.v[4 * h + c] = mul_add_final(xu[4 * h + c], make_uint2(b.v[2 * c], b.v[2 * c + 1]), xe[4 * h + c], q, two_q);
        }
    }
    seb_stg256_stream(reinterpret_cast<seb_oct *>(c0 + pos), v0);
    seb_stg256_stream(reinterpret_cast<seb_oct *>(c1 + pos), v1);
}

// Three polynomials at once (registers permitting): every twiddle fetched once for all three.
template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / SEB_E,
                                  (LOGN == 10 ? 12 : LOGN == 11 ? 6 : LOGN == 12 ? 3 : LOGN == 13 ? 2 : 1))
    k_encrypt_asym(const int64_t *__restrict__ pt, const uint32_t *__restrict__ mag, const int8_t *__restrict__ e,
                   const uint8_t *__restrict__ u,
                   const seb_oct *__restrict__ roots, const seb_oct *__restrict__ pk0s,
                   const seb_oct *__restrict__ pk1s, const __grid_constant__ SebModuli mods, int np,
                   uint32_t *__restrict__ out, size_t batch)
{
    constexpr int N = 1 << LOGN;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t        = threadIdx.x;
    const size_t b     = seb_item();
    if (b >= batch) return;
    const int p        = (int)blockIdx.x;
    const SebModulus m = mods.m[p];

    const seb_oct *tw = roots + (size_t)p * NttTwSize<LOGN>::OCTS;
    uint32_t x[3][SEB_E];
    if (__ldg(mag + b) < m.two_q - SEB_E_BOUND)  // CTA-uniform: |m + e0| < 2q everywhere in this ciphertext
    {
        LoadAsym<true> ld{u + b * (N / 4), e + b * 2 * N, e + b * 2 * N + N, pt + b * N, m};
        seb_ntt_first<LOGN, 3>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    else
    {
        LoadAsym<false> ld{u + b * (N / 4), e + b * 2 * N, e + b * 2 * N + N, pt + b * N, m};
        seb_ntt_first<LOGN, 3>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    seb_ntt_rest<LOGN, 3>(x, smem, t, tw, m.q, m.two_q);

    using O           = NttOut<LOGN>;
    uint32_t *c0      = out + (b * np + p) * 2 * (size_t)N;
    uint32_t *c1      = c0 + N;
    const seb_oct *k0 = pk0s + (size_t)p * (N / 4);
    const seb_oct *k1 = pk1s + (size_t)p * (N / 4);
#pragma unroll
    for (int i = 0; i < O::GPL; i++)
#pragma unroll
        for (int k = 0; k < O::RUN / 8; k++)
        {
            const int r = i * O::RUN + 8 * k;
            uint32_t xu[8], xe[8], xp[8];
#pragma unroll
            for (int c = 0; c < 8; c++)
            {
                xu[c] = x[0][r + c];
                xe[c] = x[1][r + c];
                xp[c] = x[2][r + c];
            }
            const uint32_t ei = seb_epi_index<LOGN>(t, i, 2 * k);
            asym_store8(xu, xe, xp, k0 + ei, k1 + ei, O::T, c0, c1, O::pos(t, i) + 8 * k, m.q, m.two_q);
        }
}

// ---------------------------------------------------------------------------------------------
// symmetric encrypt, one (ciphertext, prime)
// ---------------------------------------------------------------------------------------------
template <bool SMALL>
struct LoadSym
{
    const int8_t *e;
    const int64_t *pt;
    SebModulus m;
    __device__ __forceinline__ uint32_t operator()(int, uint32_t pos) const
    {
        return reduce_pte<SMALL>(pt, e, pos, m);
    }
};

// a / c0 are addressed as base + b*ct_stride + p*p_stride (words): the full layout has a in the c1 slot
// of the output (a = out + n, c0 = out, strides 2*np*n and 2n); the seed-compressed layout keeps a in
// scratch and writes c0 only ([batch][np][n], strides np*n and n).
template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / SEB_E)
    k_encrypt_sym(const int64_t *__restrict__ pt, const uint32_t *__restrict__ mag, const int8_t *__restrict__ e,
                  const seb_oct *__restrict__ roots,
                  const seb_oct *__restrict__ ntt_s, const __grid_constant__ SebModuli mods, uint32_t *a_base,
                  uint32_t *c0_base, size_t ct_stride, size_t p_stride, int quirk, size_t batch)
{
    constexpr int N = 1 << LOGN;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t        = threadIdx.x;
    const size_t b     = seb_item();
    if (b >= batch) return;
    const int p        = (int)blockIdx.x;
    const SebModulus m = mods.m[p];

    const seb_oct *tw = roots + (size_t)p * NttTwSize<LOGN>::OCTS;
    uint32_t x[1][SEB_E];
    if (__ldg(mag + b) < m.two_q - SEB_E_BOUND)  // CTA-uniform, see k_encrypt_asym
    {
        LoadSym<true> ld{e + b * N, pt + b * N, m};
        seb_ntt_first<LOGN, 1>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    else
    {
        LoadSym<false> ld{e + b * N, pt + b * N, m};
        seb_ntt_first<LOGN, 1>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    seb_ntt_rest<LOGN, 1>(x, smem, t, tw, m.q, m.two_q);

    using O           = NttOut<LOGN>;
    uint32_t *c0      = c0_base + b * ct_stride + (size_t)p * p_stride;
    uint32_t *c1      = a_base + b * ct_stride + (size_t)p * p_stride;
    const seb_oct *sk = ntt_s + (size_t)p * (N / 4);
#pragma unroll
    for (int i = 0; i < O::GPL; i++)
#pragma unroll
        for (int k = 0; k < O::RUN / 8; k++)
        {
            const uint32_t pos = O::pos(t, i) + 8 * k;
            const int r        = i * O::RUN + 8 * k;
            const uint4 a0     = *reinterpret_cast<const uint4 *>(c1 + pos);
            const uint4 a1     = *reinterpret_cast<const uint4 *>(c1 + pos + 4);
            const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            seb_oct cv, mv;
#pragma unroll
            for (int h = 0; h < 2; h++)
            {
                const seb_oct s = seb_ldg256(sk + seb_epi_index<LOGN>(t, i, 2 * k + h));
#pragma unroll
                for (int c = 0; c < 4; c++)
                {
                    const uint32_t prod =
                        seb_csub(seb_mul_shoup_lazy(av[4 * h + c], s.v[2 * c], s.v[2 * c + 1], m.q), m.q);
                    const uint32_t neg = prod ? m.q - prod : 0u;  // poly_neg_mod (polymodarith.h:67-70)
                    mv.v[4 * h + c]    = seb_final_reduce(x[0][r + 4 * h + c], m.q, m.two_q);
                    cv.v[4 * h + c]    = seb_csub(neg + mv.v[4 * h + c], m.q);
                }
            }
            seb_stg256_stream(reinterpret_cast<seb_oct *>(c0 + pos), cv);
            if (quirk) seb_stg256_stream(reinterpret_cast<seb_oct *>(c1 + pos), mv);
        }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// grid (np, items) folded into three dimensions: x = prime (fastest, so one item's primes are
// neighbours), y/z = item
static inline dim3 seb_grid(int np, size_t items)
{
    const size_t gy = items < 32768 ? items : 32768;
    return dim3((unsigned)np, (unsigned)gy, (unsigned)((items + gy - 1) / gy));
}

#define SEB_DISPATCH_LOGN(logn, CALL)      \
    switch (logn)                          \
    {                                      \
        case 10: CALL(10); break;          \
        case 11: CALL(11); break;          \
        case 12: CALL(12); break;          \
        case 13: CALL(13); break;          \
        case 14: CALL(14); break;          \
        default: return cudaErrorInvalidValue; \
    }

size_t seb_table_octs(int logn)
{
    switch (logn)
    {
        case 10: return NttTwSize<10>::OCTS;
        case 11: return NttTwSize<11>::OCTS;
        case 12: return NttTwSize<12>::OCTS;
        case 13: return NttTwSize<13>::OCTS;
        case 14: return NttTwSize<14>::OCTS;
    }
    return 0;
}

void seb_host_build_tw(int logn, const uint2 *roots_bitrev, seb_oct *out)
{
    switch (logn)
    {
        case 10: seb_build_tw<10>(roots_bitrev, out); break;
        case 11: seb_build_tw<11>(roots_bitrev, out); break;
        case 12: seb_build_tw<12>(roots_bitrev, out); break;
        case 13: seb_build_tw<13>(roots_bitrev, out); break;
        case 14: seb_build_tw<14>(roots_bitrev, out); break;
    }
}

void seb_host_build_epi(int logn, const uint2 *natural, seb_oct *out)
{
    switch (logn)
    {
        case 10: seb_build_epi<10>(natural, out); break;
        case 11: seb_build_epi<11>(natural, out); break;
        case 12: seb_build_epi<12>(natural, out); break;
        case 13: seb_build_epi<13>(natural, out); break;
        case 14: seb_build_epi<14>(natural, out); break;
    }
}

cudaError_t seb_encrypt_configure(int logn)
{
    cudaError_t err = cudaSuccess;
#define CFG(L)                                                                                                    \
    {                                                                                                             \
        err = cudaFuncSetAttribute(k_ntt_forward<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * NttSmem<L>::WORDS);        \
        if (err == cudaSuccess)                                                                                   \
            err = cudaFuncSetAttribute(k_encrypt_asym<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * NttSmem<L>::WORDS);  \
        if (err == cudaSuccess)                                                                                   \
            err = cudaFuncSetAttribute(k_encrypt_sym<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * NttSmem<L>::WORDS);    \
    }
    SEB_DISPATCH_LOGN(logn, CFG)
#undef CFG
    return err;
}

cudaError_t seb_launch_ntt(int logn, uint32_t *polys, const seb_oct *roots, const SebModuli &mods, int np,
                           size_t npolys_total, cudaStream_t st)
{
    if (npolys_total == 0) return cudaSuccess;
    const size_t items = npolys_total / (size_t)np;
#define RUN(L) k_ntt_forward<L><<<seb_grid(np, items), (1 << L) / SEB_E, 4 * NttSmem<L>::WORDS, st>>>(polys, roots, mods, np, items)
    SEB_DISPATCH_LOGN(logn, RUN)
#undef RUN
    return cudaGetLastError();
}

cudaError_t seb_launch_encrypt_asym(int logn, const int64_t *pt, const uint32_t *mag, const int8_t *e, const uint8_t *u,
                                    const seb_oct *roots, const seb_oct *pk0s, const seb_oct *pk1s,
                                    const SebModuli &mods, int np, uint32_t *out, int batch, cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
#define RUN(L)                                                                                            \
    k_encrypt_asym<L><<<seb_grid(np, (size_t)batch), (1 << L) / SEB_E, 12 * NttSmem<L>::WORDS, st>>>( \
        pt, mag, e, u, roots, pk0s, pk1s, mods, np, out, (size_t)batch)
    SEB_DISPATCH_LOGN(logn, RUN)
#undef RUN
    return cudaGetLastError();
}

cudaError_t seb_launch_encrypt_sym(int logn, const int64_t *pt, const uint32_t *mag, const int8_t *e, const seb_oct *roots,
                                   const seb_oct *ntt_s, const SebModuli &mods, int np, uint32_t *a, uint32_t *c0,
                                   size_t ct_stride, size_t p_stride, int quirk, int batch, cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
#define RUN(L)                                                                                                       \
    k_encrypt_sym<L><<<seb_grid(np, (size_t)batch), (1 << L) / SEB_E, 4 * NttSmem<L>::WORDS, st>>>(                 \
        pt, mag, e, roots, ntt_s, mods, a, c0, ct_stride, p_stride, quirk, (size_t)batch)
    SEB_DISPATCH_LOGN(logn, RUN)
#undef RUN
    return cudaGetLastError();
}
