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

// Which plan the ONE-polynomial kernels (k_ntt_forward, k_encrypt_sym) use per degree (seb_ntt.cuh, NttCfg): 32
// coefficients per thread at n = 8192 (three passes, one CTA-wide barrier, 7.5 instructions per butterfly instead of
// 8.2: 59.9 % of the HBM peak against 56.8 %), 16 everywhere else — at n = 16384 the 32-coefficient plan leaves two
// 512-thread CTAs per SM and measures 46.4 % against 48.0 % (profiles/r02_ubench_ntt_plans.txt).  The
// three-polynomial asymmetric kernel keeps 16 everywhere (3 x 32 values do not fit the register file at any useful
// occupancy).
// The SYMMETRIC kernel at n = 16384 runs in the SPLIT form (SEB_SYM14_SPLIT, k_encrypt_sym_split): stage 0 is folded into
// the LOAD (each CTA of a pair reads the inputs at i and i + n/2 and keeps its own output of that butterfly, so the one
// multiplication of stage 0 is done twice), after which the two halves are independent 8192-point transforms on the plan
// of n = 8192 (key 29) with their own root tables: no remote stores, no release/acquire cluster barrier.  5.66 -> 5.50 ms
// for configuration D's shard.  The NTT-only kernel keeps the 2-CTA cluster form (plan key 30): split, it reads every
// input twice and measured 49.9 % of the HBM peak against 51.8 %.
#ifndef SEB_SYM14_SPLIT
#define SEB_SYM14_SPLIT 1
#endif
#define SEB_SPLIT1(logn) (SEB_SYM14_SPLIT && (logn) == 14)
#define SEB_KEY_SPLIT(logn) SEB_NTT_KEY32((logn) - 1)
#define SEB_KEY1(logn) ((logn) >= 13 ? SEB_NTT_KEY32(logn) : (logn))
// ... and at n = 16384 the 512 threads of that plan run as a CLUSTER of two 256-thread CTAs, each holding half of the
// polynomial in its shared memory (seb_ntt.cuh, "two-CTA cluster form"): four resident CTAs per SM instead of two.
#define SEB_CLUSTER1(logn) ((logn) == 14)

// resident CTAs per SM the NTT-only kernel is compiled for, per plan: the best of the sweeps in
// profiles/r01_ubench_ntt_occupancy.txt (48 registers for n <= 4096) and profiles/r02_ubench_ntt_plans.txt (64
// registers for the 32-coefficient plans: 4 x 256 threads at n = 8192, 2 x 512 at n = 16384)
template <int K>
struct NttOcc
{
    static constexpr int MINB = K == 10 ? 20 : K == 11 ? 10 : K == 12 ? 5 : K == 13 ? 4 : K == 14 ? 2
                                : K == SEB_NTT_KEY32(13) ? 4 : 2;
};

// CL = 2: a cluster of two CTAs per polynomial, grid.x = 2 * prime + cluster rank (launched with the cluster attribute)
template <int K, int CL>
__global__ void __launch_bounds__(NttCfg<K>::T / CL, CL == 2 ? 4 : NttOcc<K>::MINB)
    k_ntt_forward(uint32_t *__restrict__ polys, const seb_oct *__restrict__ roots,
                  const __grid_constant__ SebModuli mods, int np, size_t items)
{
    constexpr int N = 1 << NttCfg<K>::LOGN;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t         = seb_ntt_thread<K, CL>();
    const size_t b      = seb_item();
    if (b >= items)  // uniform over a cluster: both CTAs leave together
    {
        if (CL == 2) seb_cluster_wait();
        return;
    }
    const int p         = (int)(blockIdx.x / CL);
    const SebModulus &m = mods.m[p];
    uint32_t *data      = polys + (b * np + p) * N;

    uint32_t x[1][NttCfg<K>::E];
    LoadPlain ld{data};
    if constexpr (CL == 2)
        seb_ntt_forward_cluster2<K>(x, smem, t, roots + (size_t)p * NttTwSize<K>::OCTS, m.q, m.two_q, ld);
    else
        seb_ntt_forward<K, 1>(x, smem, t, roots + (size_t)p * NttTwSize<K>::OCTS, m.q, m.two_q, ld);

    using O = NttOut<K>;
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

// (x*w + y) mod q in [0,q) for a Shoup pair (w, floor(w*2^32/q)), any 32-bit x, y lazy in [0,4q).
// The addition rides in the multiply-add: x*w - hi32(x*w')*q lies in [0,2q) (uintmodarith.h:308-331), y is brought
// to [0,2q) with one conditional subtraction, their sum is below 4q < 2^32, and one final reduction follows:
// 6 instructions per output where reducing the product and y separately took 8.
__device__ __forceinline__ uint32_t mul_add_final(uint32_t x, uint2 w, uint32_t y, uint32_t q, uint32_t two_q)
{
    const uint32_t y2 = seb_csub(y, two_q);
    const uint32_t r  = x * w.x + y2 - __umulhi(x, w.y) * q;
    return seb_final_reduce(r, q, two_q);
}

// (y - x*w) mod q in [0,q), same operands: y2 + 2q - (x*w - hi*q) lies in (0,4q)
__device__ __forceinline__ uint32_t mul_sub_final(uint32_t x, uint2 w, uint32_t y, uint32_t q, uint32_t two_q)
{
    const uint32_t y2 = seb_csub(y, two_q) + two_q;
    const uint32_t r  = __umulhi(x, w.y) * q + y2 - x * w.x;
    return seb_final_reduce(r, q, two_q);
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

#ifndef SEB_ASYM12_MINB
#define SEB_ASYM12_MINB 3
#endif
// Three polynomials at once (registers permitting): every twiddle fetched once for all three.
template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / SEB_E,
                                  (LOGN == 10 ? 12 : LOGN == 11 ? 6 : LOGN == 12 ? SEB_ASYM12_MINB : LOGN == 13 ? 2 : 1))
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
template <int K, int CL>
__global__ void __launch_bounds__(NttCfg<K>::T / CL, (K == SEB_NTT_KEY32(13) ? 3 : CL == 2 ? 3 : 0))
    k_encrypt_sym(const int64_t *__restrict__ pt, const uint32_t *__restrict__ mag, const int8_t *__restrict__ e,
                  const seb_oct *__restrict__ roots,
                  const seb_oct *__restrict__ ntt_s, const __grid_constant__ SebModuli mods, uint32_t *a_base,
                  uint32_t *c0_base, size_t ct_stride, size_t p_stride, int quirk, size_t batch)
{
    constexpr int N = 1 << NttCfg<K>::LOGN;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t        = seb_ntt_thread<K, CL>();
    const size_t b     = seb_item();
    if (b >= batch)  // uniform over a cluster
    {
        if (CL == 2) seb_cluster_wait();
        return;
    }
    const int p        = (int)(blockIdx.x / CL);
    const SebModulus m = mods.m[p];

    const seb_oct *tw = roots + (size_t)p * NttTwSize<K>::OCTS;
    uint32_t x[1][NttCfg<K>::E];
    if (__ldg(mag + b) < m.two_q - SEB_E_BOUND)  // uniform over the ciphertext, see k_encrypt_asym
    {
        LoadSym<true> ld{e + b * N, pt + b * N, m};
        seb_ntt_first<K, 1, LoadSym<true>, CL>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    else
    {
        LoadSym<false> ld{e + b * N, pt + b * N, m};
        seb_ntt_first<K, 1, LoadSym<false>, CL>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    seb_ntt_rest<K, 1, CL>(x, smem, t, tw, m.q, m.two_q);

    using O           = NttOut<K>;
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
                const seb_oct s = seb_ldg256(sk + seb_epi_index<K>(t, i, 2 * k + h));
#pragma unroll
                for (int c = 0; c < 4; c++)
                {
                    // c0 = -(a (.) s^) + (m+e)^ (poly_neg_mod + poly_add_mod, polymodarith.h:39-70) as one lazy
                    // multiply-subtract and one final reduction
                    cv.v[4 * h + c] = mul_sub_final(av[4 * h + c], make_uint2(s.v[2 * c], s.v[2 * c + 1]), x[0][r + 4 * h + c],
                                                    m.q, m.two_q);
                    if (quirk) mv.v[4 * h + c] = seb_final_reduce(x[0][r + 4 * h + c], m.q, m.two_q);
                }
            }
            seb_stg256_stream(reinterpret_cast<seb_oct *>(c0 + pos), cv);
            if (quirk) seb_stg256_stream(reinterpret_cast<seb_oct *>(c1 + pos), mv);
        }
}

// ---------------------------------------------------------------------------------------------
// the SPLIT form of the one-polynomial kernels: a polynomial of 2 * 2^LOGN(K) coefficients as two CTAs
// ---------------------------------------------------------------------------------------------
// Stage 0 of the forward transform pairs x[i] with x[i + n/2] under ONE root (roots[1]) and leaves two independent
// half-size transforms; half r uses, at its local stage s and group j, the root roots[2^(s+1) + r 2^s + j] of the full
// table (ntt.c:124-166 with the bit-reversed table of ntt.c:40-52).  CTA r = blockIdx.x & 1 loads both inputs of every
// stage-0 butterfly of its half, keeps output r and runs plan K on it with the table of half r.
// Table of a prime: [octs(K) of half 0][octs(K) of half 1][one oct whose first two words are roots[1] and its Shoup word].
template <class Inner>
struct LoadSplit
{
    Inner in;  // the loader of the full polynomial (positions 0 .. 2 NH)
    uint32_t nh, rank, q, two_q;
    uint2 w0;
    __device__ __forceinline__ uint32_t operator()(int poly, uint32_t pos) const
    {
        uint32_t a = in(poly, pos), b = in(poly, pos + nh);
        seb_bfly(a, b, w0, q, two_q);
        return rank ? b : a;
    }
};

// symmetric encrypt in the split form: inputs (pt, e) and outputs (c0; a is read and, with the quirk, rewritten at the
// thread's own positions) are different buffers, so the two CTAs of a polynomial do not have to meet at all
template <int K>
__global__ void __launch_bounds__(NttCfg<K>::T, 3)
    k_encrypt_sym_split(const int64_t *__restrict__ pt, const uint32_t *__restrict__ mag, const int8_t *__restrict__ e,
                        const seb_oct *__restrict__ roots, const seb_oct *__restrict__ ntt_s,
                        const __grid_constant__ SebModuli mods, uint32_t *a_base, uint32_t *c0_base, size_t ct_stride,
                        size_t p_stride, int quirk, size_t batch)
{
    constexpr int NH = 1 << NttCfg<K>::LOGN, N = 2 * NH;
    extern __shared__ __align__(16) uint32_t smem[];
    const int t    = threadIdx.x;
    const size_t b = seb_item();
    if (b >= batch) return;
    const int p         = (int)(blockIdx.x >> 1);
    const uint32_t rank = blockIdx.x & 1u;
    const SebModulus m  = mods.m[p];
    const seb_oct *tab  = roots + (size_t)p * (2 * NttTwSize<K>::OCTS + 1);
    const seb_oct w0o   = seb_ldg256(tab + 2 * NttTwSize<K>::OCTS);
    const uint2 w0      = make_uint2(w0o.v[0], w0o.v[1]);
    const seb_oct *tw   = tab + (size_t)rank * NttTwSize<K>::OCTS;

    uint32_t x[1][NttCfg<K>::E];
    if (__ldg(mag + b) < m.two_q - SEB_E_BOUND)  // uniform over the ciphertext, see k_encrypt_asym
    {
        LoadSplit<LoadSym<true>> ld{LoadSym<true>{e + b * N, pt + b * N, m}, (uint32_t)NH, rank, m.q, m.two_q, w0};
        seb_ntt_first<K, 1, LoadSplit<LoadSym<true>>, 1>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    else
    {
        LoadSplit<LoadSym<false>> ld{LoadSym<false>{e + b * N, pt + b * N, m}, (uint32_t)NH, rank, m.q, m.two_q, w0};
        seb_ntt_first<K, 1, LoadSplit<LoadSym<false>>, 1>(x, smem, t, tw, m.q, m.two_q, ld);
    }
    seb_ntt_rest<K, 1, 1>(x, smem, t, tw, m.q, m.two_q);

    using O           = NttOut<K>;
    uint32_t *c0      = c0_base + b * ct_stride + (size_t)p * p_stride + rank * NH;
    uint32_t *c1      = a_base + b * ct_stride + (size_t)p * p_stride + rank * NH;
    const seb_oct *sk = ntt_s + (size_t)p * (N / 4) + (size_t)rank * (NH / 4);
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
                const seb_oct s = seb_ldg256(sk + seb_epi_index<K>(t, i, 2 * k + h));
#pragma unroll
                for (int c = 0; c < 4; c++)
                {
                    cv.v[4 * h + c] = mul_sub_final(av[4 * h + c], make_uint2(s.v[2 * c], s.v[2 * c + 1]), x[0][r + 4 * h + c],
                                                    m.q, m.two_q);
                    if (quirk) mv.v[4 * h + c] = seb_final_reduce(x[0][r + 4 * h + c], m.q, m.two_q);
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
// the same over plan keys: the five 16-coefficient plans and the two 32-coefficient ones
#define SEB_DISPATCH_KEY(key, CALL, DEFAULT)         \
    switch (key)                                     \
    {                                                \
        case 10: CALL(10); break;                    \
        case 11: CALL(11); break;                    \
        case 12: CALL(12); break;                    \
        case 13: CALL(13); break;                    \
        case 14: CALL(14); break;                    \
        case SEB_NTT_KEY32(13): CALL(SEB_NTT_KEY32(13)); break; \
        case SEB_NTT_KEY32(14): CALL(SEB_NTT_KEY32(14)); break; \
        default: DEFAULT;                            \
    }

int seb_ntt_key1(int logn) { return SEB_KEY1(logn); }
int seb_sym_split(int logn) { return SEB_SPLIT1(logn) ? 1 : 0; }

size_t seb_table_octs(int key)
{
#define OCTS(K) return NttTwSize<K>::OCTS
    SEB_DISPATCH_KEY(key, OCTS, return 0)
#undef OCTS
    return 0;
}

void seb_host_build_tw(int key, const uint2 *roots_bitrev, seb_oct *out)
{
#define BUILD(K) seb_build_tw<K>(roots_bitrev, out)
    SEB_DISPATCH_KEY(key, BUILD, return)
#undef BUILD
}

void seb_host_build_epi(int key, const uint2 *natural, seb_oct *out)
{
#define BUILD(K) seb_build_epi<K>(natural, out)
    SEB_DISPATCH_KEY(key, BUILD, return)
#undef BUILD
}

// The tables of the SYMMETRIC kernel of degree 2^logn: the per-pass root table of a prime from its bit-reversed roots, and
// ntt(s) in epilogue order.  They are those of plan seb_ntt_key1(logn), except in the split form (seb_sym_split), where
// both are two half-size tables on the plan of the half (see k_encrypt_sym_split).
size_t seb_table_octs_sym(int logn)
{
    return SEB_SPLIT1(logn) ? 2 * seb_table_octs(SEB_KEY_SPLIT(logn)) + 1 : seb_table_octs(SEB_KEY1(logn));
}
void seb_host_build_tw_sym(int logn, const uint2 *roots_bitrev, seb_oct *out)
{
    if (!SEB_SPLIT1(logn))
    {
        seb_host_build_tw(SEB_KEY1(logn), roots_bitrev, out);
        return;
    }
    const int key = SEB_KEY_SPLIT(logn), lh = logn - 1;
    const size_t nh = (size_t)1 << lh, o = seb_table_octs(key);
    uint2 *half = new uint2[nh];
    for (int r = 0; r < 2; r++)
    {
        half[0] = roots_bitrev[0];  // never read: a transform starts at roots[1]
        for (int s = 0; s < lh; s++)
            for (size_t j = 0; j < ((size_t)1 << s); j++)
                half[((size_t)1 << s) + j] = roots_bitrev[((size_t)2 << s) + ((size_t)r << s) + j];
        seb_host_build_tw(key, half, out + (size_t)r * o);
    }
    delete[] half;
    seb_oct w0 = {};
    w0.v[0] = roots_bitrev[1].x, w0.v[1] = roots_bitrev[1].y;
    out[2 * o] = w0;
}
void seb_host_build_epi_sym(int logn, const uint2 *natural, seb_oct *out)
{
    if (!SEB_SPLIT1(logn))
    {
        seb_host_build_epi(SEB_KEY1(logn), natural, out);
        return;
    }
    const size_t nh = (size_t)1 << (logn - 1);
    for (int r = 0; r < 2; r++)
        seb_host_build_epi(SEB_KEY_SPLIT(logn), natural + (size_t)r * nh, out + (size_t)r * (nh / 4));
}

// shared memory of one CTA of a one-polynomial kernel
template <int K, int CL>
static constexpr size_t seb_smem1() { return 4 * (size_t)(CL == 2 ? NttSmemHalf<K>::WORDS : NttSmem<K>::WORDS); }

cudaError_t seb_encrypt_configure(int logn)
{
    cudaError_t err = cudaSuccess;
#define CFG(L)                                                                                                    \
    {                                                                                                             \
        constexpr int K1 = SEB_KEY1(L), C1 = SEB_CLUSTER1(L) ? 2 : 1;                                             \
        err = cudaFuncSetAttribute(k_ntt_forward<K1, C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seb_smem1<K1, C1>()); \
        if (err == cudaSuccess)                                                                                   \
            err = cudaFuncSetAttribute(k_encrypt_asym<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * NttSmem<L>::WORDS);  \
        if (err == cudaSuccess)                                                                                   \
            err = cudaFuncSetAttribute(k_encrypt_sym<K1, C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seb_smem1<K1, C1>()); \
    }
    SEB_DISPATCH_LOGN(logn, CFG)
#undef CFG
    if (err == cudaSuccess && SEB_SPLIT1(logn))
        err = cudaFuncSetAttribute(k_encrypt_sym_split<SEB_KEY_SPLIT(14)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)seb_smem1<SEB_KEY_SPLIT(14), 1>());
    return err;
}

// launch of a one-polynomial kernel over (np primes x items): plain, or as 2-CTA clusters along x
template <int K, int CL, class Kern, class... Args>
static cudaError_t seb_launch1(Kern kern, int np, size_t items, cudaStream_t st, Args... args)
{
    dim3 grid = seb_grid(np, items);
    grid.x *= CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = grid;
    cfg.blockDim           = dim3(NttCfg<K>::T / CL);
    cfg.dynamicSmemBytes   = seb_smem1<K, CL>();
    cfg.stream             = st;
    cudaLaunchAttribute attr[1];
    attr[0].id               = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs                = attr;
    cfg.numAttrs             = CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// roots: the table of plan seb_ntt_key1(logn)
cudaError_t seb_launch_ntt(int logn, uint32_t *polys, const seb_oct *roots, const SebModuli &mods, int np,
                           size_t npolys_total, cudaStream_t st)
{
    if (npolys_total == 0) return cudaSuccess;
    const size_t items = npolys_total / (size_t)np;
    cudaError_t err    = cudaSuccess;
#define RUN(L)                                                                                              \
    {                                                                                                       \
        constexpr int K1 = SEB_KEY1(L), C1 = SEB_CLUSTER1(L) ? 2 : 1;                                       \
        err = seb_launch1<K1, C1>(k_ntt_forward<K1, C1>, np, items, st, polys, roots, mods, np, items);     \
    }
    SEB_DISPATCH_LOGN(logn, RUN)
#undef RUN
    return err;
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

// roots / ntt_s: the symmetric kernel's tables (seb_host_build_tw_sym / seb_host_build_epi_sym)
cudaError_t seb_launch_encrypt_sym(int logn, const int64_t *pt, const uint32_t *mag, const int8_t *e, const seb_oct *roots,
                                   const seb_oct *ntt_s, const SebModuli &mods, int np, uint32_t *a, uint32_t *c0,
                                   size_t ct_stride, size_t p_stride, int quirk, int batch, cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
    cudaError_t err = cudaSuccess;
    if (SEB_SPLIT1(logn))  // n = 16384: two independent CTAs per (ciphertext, prime), each on the plan of n = 8192
    {
        constexpr int KS = SEB_KEY_SPLIT(14);
        dim3 grid        = seb_grid(np, (size_t)batch);
        grid.x *= 2;
        k_encrypt_sym_split<KS><<<grid, NttCfg<KS>::T, seb_smem1<KS, 1>(), st>>>(pt, mag, e, roots, ntt_s, mods, a, c0, ct_stride,
                                                                                p_stride, quirk, (size_t)batch);
        return cudaGetLastError();
    }
#define RUN(L)                                                                                                       \
    {                                                                                                                \
        constexpr int K1 = SEB_KEY1(L), C1 = SEB_CLUSTER1(L) ? 2 : 1;                                                \
        err = seb_launch1<K1, C1>(k_encrypt_sym<K1, C1>, np, (size_t)batch, st, pt, mag, e, roots, ntt_s, mods, a, c0, ct_stride, \
                                  p_stride, quirk, (size_t)batch);                                                   \
    }
    SEB_DISPATCH_LOGN(logn, RUN)
#undef RUN
    return err;
}
