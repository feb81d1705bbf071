// seb_ntt.cuh — in-shared-memory negacyclic forward NTT over a 30-bit prime, register-blocked.
//
// Computes exactly what device/lib/ntt.c:124-189 (ntt_inpl) computes: Cooley-Tukey, natural
// order in, bit-reversed order out, A[j] = a(psi^(2*bitrev(j)+1)), every output in [0,q).  Outputs
// are canonical residues of an exact function, so the algorithm is free to differ: here stages are
// fused 3 or 4 at a time in registers (radix-8/16), butterflies are Harvey lazy butterflies on
// [0,4q) with Shoup twiddles {w, floor(w*2^32/q)} (the reference's own SE_NTT_FAST variant,
// ntt.c:72-109, uintmodarith.h:308-331), and a final correction brings values to [0,q).
//
// Work decomposition for one polynomial of n = 2^LOGN coefficients (a plan is chosen by a key K, NttCfg<K>):
//   * T = n/E threads, E = 16 coefficients per thread (or 32: the one-polynomial kernels at n >= 8192).
//   * the log2(n) stages are split into passes of R in {3,4,5} stages (NttPlan<K>); a pass of R
//     stages with stride S = n >> (s0+R) works on groups {blk*(S<<R) + off + j*S, j < 2^R}; a thread
//     owns E >> R such groups.  Between passes coefficients go through shared memory; the
//     first pass takes its inputs from a loader functor (global memory / on-the-fly expansion)
//     and the last pass, whose groups are 2^R contiguous coefficients, hands each thread its
//     contiguous outputs so the caller can fuse its epilogue and issue 128-bit stores.
//   * shared-memory words are laid out with padding (seb_pad: a few words inserted every 32/64/128
//     coefficients, per plan) so that every pass is bank-conflict free AND the address of element
//     j of a group is phys(base) + a compile-time constant, i.e. an immediate offset: no
//     per-access address arithmetic (checked exhaustively by
//     tests/test_host_logic.py::test_smem_layout_conflict_free).  The last pass reads its 2^R
//     contiguous coefficients with 128-bit loads.
//   * NPOLY polynomials that share the modulus go through the passes together so each twiddle
//     is fetched once for all of them.
//   * twiddles are stored per pass in the order the threads consume them (seb_build_tw): the
//     2^R - 1 roots a group needs sit in heap order in 2^R slots, split in 32-byte "octs" and
//     interleaved across groups, so a warp's 256-bit loads are fully coalesced.
#pragma once

#include "seb_common.cuh"

// A transform is selected by a KEY K: K = log2(n) for 16 coefficients per thread (every degree), K = 16 + log2(n) for
// 32 coefficients per thread (n = 8192 and 16384: half the threads, 64 registers each instead of 32, three passes
// instead of four and a single CTA-wide barrier per transform; used by the kernels that carry ONE polynomial per CTA —
// VERDICT r01 weak #3).
template <int K>
struct NttCfg
{
    static constexpr int LOGN = K & 15;
    static constexpr int E    = (K & 16) ? 32 : 16;  // coefficients per thread
    static constexpr int T    = (1 << LOGN) / E;     // threads per polynomial
};
#define SEB_E 16  // the default (and the only value the three-polynomial asymmetric kernel uses)
#define SEB_NTT_KEY32(logn) (16 + (logn))

// Three polynomials per CTA (the asymmetric kernel): fetch a pass' roots stage by stage (at most two octs live) instead of
// all at once (four octs = 32 registers beside the 48 coefficient registers) - A/B switch, see profiles/README.md
#ifndef SEB_STAGED_TW_NPOLY3
#define SEB_STAGED_TW_NPOLY3 0
#endif

// How a twiddle oct is fetched: a 256-bit read-only global load of the L1/L2-resident table.  (tools/ubench/
// ubench_ntt_tma.cu overrides it to measure a table staged in shared memory by cp.async.bulk.)
#ifndef SEB_TW_LOAD
#define SEB_TW_LOAD(p) seb_ldg256(p)
#endif

template <int K>
struct NttPlan;
template <>
struct NttPlan<10>
{
    static constexpr int NPASS = 3;
    static constexpr int R[4]  = {4, 3, 3, 0};
};
template <>
struct NttPlan<11>
{
    static constexpr int NPASS = 3;
    static constexpr int R[4]  = {4, 4, 3, 0};
};
template <>
struct NttPlan<12>
{
    static constexpr int NPASS = 3;
    static constexpr int R[4]  = {4, 4, 4, 0};
};
template <>
struct NttPlan<13>
{
    static constexpr int NPASS = 4;
    static constexpr int R[4]  = {4, 3, 3, 3};
};
template <>
struct NttPlan<14>
{
    static constexpr int NPASS = 4;
    static constexpr int R[4]  = {4, 4, 3, 3};
};

template <>
struct NttPlan<SEB_NTT_KEY32(13)>
{
    static constexpr int NPASS = 3;
    static constexpr int R[4]  = {5, 4, 4, 0};
};
template <>
struct NttPlan<SEB_NTT_KEY32(14)>
{
    static constexpr int NPASS = 3;
    static constexpr int R[4]  = {5, 5, 4, 0};
};

// stage offset of pass P
template <int K, int P>
struct NttS0
{
    static constexpr int value = NttS0<K, P - 1>::value + NttPlan<K>::R[P - 1];
};
template <int K>
struct NttS0<K, 0>
{
    static constexpr int value = 0;
};

// Padded shared-memory layout: physical word of coefficient index a.  Additive over bit-disjoint
// fields: seb_pad(x | y) = seb_pad(x) + seb_pad(y) whenever x & y == 0.  The paddings of the 32-coefficient plans
// were found by exhaustive search over a + sum c_k (a >> k) (tests/test_host_logic.py checks every access of every pass).
template <int K>
__host__ __device__ __forceinline__ constexpr uint32_t seb_pad(uint32_t a)
{
    return a + 4u * (a >> 5) + (K == 11 ? 4u * (a >> 6) : 0u) + ((K == 12 || K == SEB_NTT_KEY32(13)) ? 8u * (a >> 7) : 0u) +
           (K == SEB_NTT_KEY32(14) ? 4u * (a >> 7) : 0u);
}
// words of shared memory one polynomial occupies
template <int K>
struct NttSmem
{
    static constexpr uint32_t WORDS = (seb_pad<K>((1u << NttCfg<K>::LOGN) - 1u) + 4u) & ~3u;
};

// An addend that is zero at run time but opaque to the compiler (a constant-bank operand).  The
// butterfly is bound by the FMA pipe (IMAD.HI + 2 IMAD, profiles/README.md) while the ALU pipe has
// slack, yet ptxas turns part of the two-input additions into IMAD.IADD; a three-input addition
// can only be an IADD3, which keeps it on the ALU pipe.
SEB_CONSTANT uint32_t c_seb_zero = 0;

// Harvey lazy butterfly: X,Y in [0,4q) -> X+WY, X-WY in [0,4q) (ntt.c:94-105)
__device__ __forceinline__ void seb_bfly(uint32_t &x, uint32_t &y, const uint2 w, const uint32_t q,
                                         const uint32_t two_q)
{
    const uint32_t u = min(x, x - two_q);
    const uint32_t t = seb_mul_shoup_lazy(y, w.x, w.y, q);
    x                = u + t + c_seb_zero;
    y                = u - t + two_q;
}

// Per-pass twiddle layout.  Pass P (stage offset S0, radix R) has NB = 2^S0 distinct groups
// ("blk"); group blk needs, for stage r of the pass, the 2^r roots  roots[((2^S0 + blk) << r) + m].
// They are stored in heap order: slot 2^r + m (slot 0 unused), 4 slots (32 bytes) per oct, and the
// table of the pass is [oct][blk].  OFF = offset of the pass' table in octs.
template <int K, int P>
struct NttTw
{
    static constexpr int R    = NttPlan<K>::R[P];
    static constexpr int S0   = NttS0<K, P>::value;
    static constexpr int OCTS = (1 << R) / 4;  // octs per group
    static constexpr int NB   = 1 << S0;
    static constexpr int OFF  = NttTw<K, P - 1>::OFF + NttTw<K, P - 1>::OCTS * NttTw<K, P - 1>::NB;
};
template <int K>
struct NttTw<K, 0>
{
    static constexpr int R    = NttPlan<K>::R[0];
    static constexpr int S0   = 0;
    static constexpr int OCTS = (1 << R) / 4;
    static constexpr int NB   = 1;
    static constexpr int OFF  = 0;
};
// octs per prime (<= n/4: the table is the same size as the plain root table)
template <int K>
struct NttTwSize
{
    static constexpr int LASTP = NttPlan<K>::NPASS - 1;
    static constexpr int OCTS  = NttTw<K, LASTP>::OFF + NttTw<K, LASTP>::OCTS * NttTw<K, LASTP>::NB;
};

template <int K, int P>
inline void seb_build_tw_pass(const uint2 *roots, seb_oct *out)
{
    using TW = NttTw<K, P>;
    for (int blk = 0; blk < TW::NB; blk++)
        for (int slot = 1; slot < (1 << TW::R); slot++)
        {
            int r = 0;
            while ((2 << r) <= slot) r++;
            const int m      = slot - (1 << r);
            const uint2 w    = roots[(((size_t)TW::NB + blk) << r) + m];
            seb_oct &o       = out[TW::OFF + (slot / 4) * TW::NB + blk];
            o.v[2 * (slot % 4)]     = w.x;
            o.v[2 * (slot % 4) + 1] = w.y;
        }
    if constexpr (P + 1 < NttPlan<K>::NPASS) seb_build_tw_pass<K, P + 1>(roots, out);
}
// roots: the reference's table roots[bitrev(i)] = psi^i (ntt.c:40-52) as Shoup pairs {w, floor(w*2^32/q)};
// out: NttTwSize<K>::OCTS octs, zero-initialised by the caller
template <int K>
inline void seb_build_tw(const uint2 *roots, seb_oct *out)
{
    seb_build_tw_pass<K, 0>(roots, out);
}

// Radix-32 passes: one stage per template instance (a single five-deep loop nest is left partly rolled by nvcc,
// with register-rotating moves), the octs of a stage fetched four slots at a time as the stage reaches them: a
// radix-32 pass never holds more than a few of its 31 roots in registers.
template <int R, int r, int NPOLY, int E>
struct SebRadixStage
{
    __device__ __forceinline__ static void run(uint32_t (&x)[NPOLY][E], const int gofs, const seb_oct *__restrict__ tp,
                                               const int nb, const uint32_t q, const uint32_t two_q)
    {
        constexpr int half = 1 << (R - 1 - r);
        constexpr int NTW  = 1 << r;                 // roots of this stage: heap slots NTW .. 2*NTW - 1
        constexpr int NO   = NTW >= 4 ? NTW / 4 : 1;  // octs they live in
        seb_oct w[NO];
#pragma unroll
        for (int k = 0; k < NO; k++) w[k] = SEB_TW_LOAD(tp + (size_t)((NTW >= 4 ? NTW / 4 : 0) + k) * nb);
#pragma unroll
        for (int m = 0; m < NTW; m++)
        {
            const int slot = NTW + m;
            const int ko   = NTW >= 4 ? m / 4 : 0;
            const uint2 tw = make_uint2(w[ko].v[2 * (slot % 4)], w[ko].v[2 * (slot % 4) + 1]);
#pragma unroll
            for (int t = 0; t < half; t++)
            {
                const int ia = gofs + m * 2 * half + t;
                const int ib = ia + half;
#pragma unroll
                for (int p = 0; p < NPOLY; p++) seb_bfly(x[p][ia], x[p][ib], tw, q, two_q);
            }
        }
        if constexpr (r + 1 < R) SebRadixStage<R, r + 1, NPOLY, E>::run(x, gofs, tp, nb, q, two_q);
    }
};

// R fused stages on NPOLY register groups of 2^R coefficients; tp points at oct 0 of this group's
// twiddles, consecutive octs are NB apart.  Octs are fetched as the stages reach them (an oct holds 4 consecutive
// heap slots, so stage r >= 2 walks octs 2^r/4 .. 2^(r+1)/4 - 1): a radix-32 pass never holds more than one oct
// of its 31 roots in registers.
template <int R, int NPOLY, int E>
__device__ __forceinline__ void seb_radix_regs(uint32_t (&x)[NPOLY][E], const int gofs,
                                               const seb_oct *__restrict__ tp, const int nb, const uint32_t q,
                                               const uint32_t two_q)
{
    constexpr int NOCT = (1 << R) / 4;
    if constexpr (R <= 4 && !(SEB_STAGED_TW_NPOLY3 && NPOLY >= 3))
    {
        seb_oct w[NOCT];
#pragma unroll
        for (int k = 0; k < NOCT; k++) w[k] = SEB_TW_LOAD(tp + (size_t)k * nb);
#pragma unroll
        for (int r = 0; r < R; r++)
        {
            const int half = 1 << (R - 1 - r);
#pragma unroll
            for (int m = 0; m < (1 << r); m++)
            {
                const int slot = (1 << r) + m;
                const uint2 tw = make_uint2(w[slot / 4].v[2 * (slot % 4)], w[slot / 4].v[2 * (slot % 4) + 1]);
#pragma unroll
                for (int t = 0; t < half; t++)
                {
                    const int ia = gofs + m * 2 * half + t;
                    const int ib = ia + half;
#pragma unroll
                    for (int p = 0; p < NPOLY; p++) seb_bfly(x[p][ia], x[p][ib], tw, q, two_q);
                }
            }
        }
    }
    else
        SebRadixStage<R, 0, NPOLY, E>::run(x, gofs, tp, nb, q, two_q);
}

// Group index of slot i of thread t in pass P: the single definition of "which thread touches which
// coefficient in which pass".  Group g of a pass of radix R and stride 2^LS covers the coefficients
// ((g >> LS) << (LS + R)) | (g & (2^LS - 1)) | (j << LS), j < 2^R.  The default assignment is
// g = t + i*T.  For n = 16384 with 16 coefficients per thread the last two passes (radix 8, two groups per thread) use
// g = (t/64)*128 + i*64 + t%64 instead, so that the 1024 contiguous coefficients a 64-thread group
// produced in pass 1 are the ones it consumes in pass 2 (and each warp keeps its own 512 in pass 3); with 32
// coefficients per thread its last pass (radix 16, two groups per thread) uses g = (t/32)*64 + i*32 + t%32: the 1024
// contiguous coefficients a warp produced in pass 1.  Lanes of a warp still own consecutive groups, so the
// shared-memory access pattern is unchanged.
template <int K, int P>
__host__ __device__ __forceinline__ constexpr uint32_t seb_ntt_group(uint32_t t, uint32_t i)
{
    constexpr int T = NttCfg<K>::T;
    if (K == 14 && P >= 2) return ((t >> 6) << 7) | (i << 6) | (t & 63u);
    if (K == SEB_NTT_KEY32(14) && P == 2) return ((t >> 5) << 6) | (i << 5) | (t & 31u);
    return t + i * T;
}

// Coefficient index of element j = 0 of group i of thread t in pass P (element j is at | (j << LS)).
// The passes use it for addressing and tests/test_host_logic.py uses it to prove the barrier scopes below.
template <int K, int P>
__host__ __device__ __forceinline__ constexpr uint32_t seb_ntt_group_base(uint32_t t, uint32_t i)
{
    constexpr int R  = NttPlan<K>::R[P];
    constexpr int LS = NttCfg<K>::LOGN - NttS0<K, P>::value - R;
    const uint32_t g = seb_ntt_group<K, P>(t, i);
    return ((g >> LS) << (LS + R)) | (g & ((1u << LS) - 1u));
}

// Barrier scope between pass P and pass P+1.  A CTA-wide barrier makes all threads wait for the
// slowest warp; wherever every coefficient a thread reads in pass P+1 was written in pass P by a
// thread of a smaller unit, only that unit synchronises:
//   SEB_SYNC_WARP : same warp -> __syncwarp()               (last boundary of n = 1024, 4096, 8192, 16384)
//   SEB_SYNC_GROUP: same aligned group of NttSync::GROUP threads -> named barrier `bar.sync 1 + t/GROUP, GROUP`
//                   (boundary 1 -> 2 of n = 8192: 8 groups of 64, and of n = 16384: the exchange is local to 64
//                   threads too, but 16 groups would need hardware barrier 0, which belongs to __syncthreads();
//                   8 groups of 128 keep to barriers 1..8)
//   SEB_SYNC_CTA  : __syncthreads()
// The 32-coefficient plans have ONE CTA-wide barrier (after pass 0) and a warp-scope one.
// Checked exhaustively by tests/test_host_logic.py::test_ntt_barrier_scopes.
#define SEB_SYNC_CTA 0
#define SEB_SYNC_GROUP 1
#define SEB_SYNC_WARP 2
template <int K, int P>
struct NttSync
{
    static constexpr int value = ((K == 10 && P == 1) || (K == 12 && P == 1) || (K == 13 && P == 2) || (K == 14 && P == 2) ||
                                  (K == SEB_NTT_KEY32(13) && P == 1) || (K == SEB_NTT_KEY32(14) && P == 1))
                                     ? SEB_SYNC_WARP
                                     : ((K == 13 && P == 1) || (K == 14 && P == 1)) ? SEB_SYNC_GROUP : SEB_SYNC_CTA;
    static constexpr int GROUP = K == 14 ? 128 : 64;  // threads per named barrier (SEB_SYNC_GROUP only)
};

template <int SCOPE, int GROUP, int T>
__device__ __forceinline__ void seb_ntt_sync(const int t)
{
    if (SCOPE == SEB_SYNC_WARP)
        __syncwarp();
    else if (SCOPE == SEB_SYNC_GROUP)
    {
#if defined(SEB_UBENCH_CTA_BARRIERS)  // tools/ubench A/B switch only
        __syncthreads();
#elif defined(__CUDACC__)
        seb_group_barrier<GROUP, T, false>(t);
#endif
    }
    else
        __syncthreads();
}

// ---- two-CTA cluster form (CL = 2): one polynomial over a pair of CTAs, each holding HALF of it in its shared memory.
// Thread t of the polynomial is thread t % (T/2) of CTA t / (T/2).  Only pass 0 crosses the halves: its groups stride
// over the whole polynomial, so a thread stores the elements with the top index bit clear into rank 0's shared memory
// and the others into rank 1's (one of the two is remote: st.shared::cluster); a cluster barrier replaces the CTA
// barrier after pass 0, and every later pass stays inside the CTA's own half (true for plan SEB_NTT_KEY32(14): its
// pass-1 blocks are 512 coefficients and threads [256 r, 256 r + 256) own blocks [16 r, 16 r + 16)).
// Why: the same 512 threads as ONE CTA leave two resident CTAs per SM, too few to hide the input loads (46 % of the
// HBM peak); as 256-thread CTAs, four are resident (profiles/r02_ubench_ntt_plans.txt).
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t seb_cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `p` (a shared-memory pointer of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t seb_cluster_map(const void *p, uint32_t rank)
{
    uint32_t local = (uint32_t)__cvta_generic_to_shared(p), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    return remote;
}
__device__ __forceinline__ void seb_cluster_st(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void seb_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// nothing to publish yet (kernel entry): no fence
__device__ __forceinline__ void seb_cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void seb_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
#else
static inline uint32_t seb_cluster_rank() { return 0; }
static inline uint32_t seb_cluster_map(const void *, uint32_t) { return 0; }
static inline void seb_cluster_st(uint32_t, uint32_t) {}
static inline void seb_cluster_arrive() {}
static inline void seb_cluster_arrive_relaxed() {}
static inline void seb_cluster_wait() {}
#endif

// One pass P of the plan.  FIRST: inputs come from load(p, pos); otherwise from smem.  LAST:
// outputs stay in registers (x[p][i*2^R + j] = coefficient (g_i << R) + j, lazy [0,4q)) and the
// caller finishes; otherwise they are written back to smem (same slots this thread read).
// CL = 2: `smem` is this CTA's half, `t` the thread index within the POLYNOMIAL (rank * T/2 + threadIdx.x).
template <int K, int P, int NPOLY, class Loader, int CL = 1>
__device__ __forceinline__ void seb_ntt_pass(uint32_t (&x)[NPOLY][NttCfg<K>::E], uint32_t *smem, const int t,
                                             const seb_oct *__restrict__ tw, const uint32_t q, const uint32_t two_q,
                                             Loader &load)
{
    static_assert(CL == 1 || (CL == 2 && NPOLY == 1 && K == SEB_NTT_KEY32(14)), "cluster form: plan {5,5,4} at n = 16384");
    constexpr int LOGN  = NttCfg<K>::LOGN;
    constexpr int E     = NttCfg<K>::E;
    constexpr int R     = NttPlan<K>::R[P];
    constexpr int S0    = NttS0<K, P>::value;
    constexpr int LS    = LOGN - S0 - R;  // log2 stride
    constexpr int GP    = E >> R;         // groups per thread
    constexpr bool LAST = (P == NttPlan<K>::NPASS - 1);

    constexpr uint32_t WORDS = NttSmem<K>::WORDS;
    constexpr uint32_t HALF  = 1u << (LOGN - 1);
    // CL = 2: this CTA's shared memory starts at coefficient rank * n/2 (seb_pad is additive over disjoint bit fields)
    const uint32_t rank = CL == 2 ? (uint32_t)t / (uint32_t)(NttCfg<K>::T / 2) : 0u;
    uint32_t *own       = CL == 2 ? smem - (rank ? seb_pad<K>(HALF) : 0u) : smem;
#pragma unroll
    for (int i = 0; i < GP; i++)
    {
        const uint32_t blk  = seb_ntt_group<K, P>((uint32_t)t, (uint32_t)i) >> LS;
        const uint32_t base = seb_ntt_group_base<K, P>((uint32_t)t, (uint32_t)i);
        uint32_t *sp        = own + seb_pad<K>(base);  // element j lives at sp[seb_pad(j << LS)]
        if (P == 0)
        {
#pragma unroll
            for (int j = 0; j < (1 << R); j++)
#pragma unroll
                for (int p = 0; p < NPOLY; p++) x[p][i * (1 << R) + j] = load(p, base | ((uint32_t)j << LS));
        }
        else if (LAST)
        {
            // 2^R contiguous coefficients: 128-bit shared loads (LS == 0, base = g << R)
#pragma unroll
            for (int p = 0; p < NPOLY; p++)
#pragma unroll
                for (int k = 0; k < (1 << R) / 4; k++)
                {
                    const uint4 v = *reinterpret_cast<const uint4 *>(sp + p * WORDS + 4 * k);
                    x[p][i * (1 << R) + 4 * k + 0] = v.x;
                    x[p][i * (1 << R) + 4 * k + 1] = v.y;
                    x[p][i * (1 << R) + 4 * k + 2] = v.z;
                    x[p][i * (1 << R) + 4 * k + 3] = v.w;
                }
        }
        else
        {
#pragma unroll
            for (int j = 0; j < (1 << R); j++)
#pragma unroll
                for (int p = 0; p < NPOLY; p++)
                    x[p][i * (1 << R) + j] = sp[p * WORDS + seb_pad<K>((uint32_t)j << LS)];
        }
        seb_radix_regs<R, NPOLY, E>(x, i * (1 << R), tw + NttTw<K, P>::OFF + blk, NttTw<K, P>::NB, q, two_q);
        if constexpr (!LAST && CL == 2 && P == 0)
        {
            // elements with the top index bit clear belong to rank 0's half, the others to rank 1's
            static_assert(LS + R == LOGN, "pass 0 of the cluster form spans the polynomial");
            // the partner's half first (remote stores have the longer way to go), then the own half with plain
            // shared-memory stores
            const uint32_t peer = seb_cluster_map(smem, rank ^ 1u) + 4u * seb_pad<K>(base);
            uint32_t *mine      = smem + seb_pad<K>(base);
            seb_cluster_wait();  // the partner CTA is running (arrive: at kernel entry): its shared memory may be written
            if (rank == 0)
            {
#pragma unroll
                for (int j = (1 << R) / 2; j < (1 << R); j++)
                    seb_cluster_st(peer + 4u * seb_pad<K>(((uint32_t)j << LS) - HALF), x[0][i * (1 << R) + j]);
#pragma unroll
                for (int j = 0; j < (1 << R) / 2; j++) mine[seb_pad<K>((uint32_t)j << LS)] = x[0][i * (1 << R) + j];
            }
            else
            {
#pragma unroll
                for (int j = 0; j < (1 << R) / 2; j++)
                    seb_cluster_st(peer + 4u * seb_pad<K>((uint32_t)j << LS), x[0][i * (1 << R) + j]);
#pragma unroll
                for (int j = (1 << R) / 2; j < (1 << R); j++)
                    mine[seb_pad<K>(((uint32_t)j << LS) - HALF)] = x[0][i * (1 << R) + j];
            }
        }
        else if constexpr (!LAST)
        {
#pragma unroll
            for (int j = 0; j < (1 << R); j++)
#pragma unroll
                for (int p = 0; p < NPOLY; p++)
                    sp[p * WORDS + seb_pad<K>((uint32_t)j << LS)] = x[p][i * (1 << R) + j];
        }
    }
}

template <int K, int P, int NPOLY, class Loader, int CL = 1>
struct SebNttRun
{
    __device__ __forceinline__ static void run(uint32_t (&x)[NPOLY][NttCfg<K>::E], uint32_t *smem, const int t,
                                               const seb_oct *__restrict__ tw, const uint32_t q, const uint32_t two_q,
                                               Loader &load)
    {
        seb_ntt_pass<K, P, NPOLY, Loader, CL>(x, smem, t, tw, q, two_q, load);
        if (P + 1 < NttPlan<K>::NPASS)
        {
            if (CL == 2 && P == 0)
            {
                // publishes both CTAs' pass-0 stores (release / acquire at cluster scope)
                seb_cluster_arrive();
                seb_cluster_wait();
            }
            else
                seb_ntt_sync<NttSync<K, P>::value, NttSync<K, P>::GROUP, NttCfg<K>::T>(t);
            SebNttRun<K, (P + 1 < NttPlan<K>::NPASS ? P + 1 : P), NPOLY, Loader, CL>::run(x, smem, t, tw, q, two_q, load);
        }
    }
};

// The cluster form of seb_ntt_forward: called by all T/2 threads of BOTH CTAs of a 2-CTA cluster with
// t = cluster rank * T/2 + threadIdx.x and `smem` = NttSmem<K>::WORDS / 2 (+ padding slack) words of the CTA's own
// shared memory.  The caller must have executed seb_cluster_arrive() once at kernel entry.
template <int K, class Loader>
__device__ __forceinline__ void seb_ntt_forward_cluster2(uint32_t (&x)[1][NttCfg<K>::E], uint32_t *smem, const int t,
                                                         const seb_oct *__restrict__ tw, const uint32_t q,
                                                         const uint32_t two_q, Loader &load)
{
    SebNttRun<K, 0, 1, Loader, 2>::run(x, smem, t, tw, q, two_q, load);
}
#ifdef __CUDACC__
// thread index within the polynomial and the entry protocol of a kernel that may run in cluster form
template <int K, int CL>
__device__ __forceinline__ int seb_ntt_thread()
{
    if (CL == 1) return (int)threadIdx.x;
    seb_cluster_arrive_relaxed();  // "this CTA runs": paired with the wait in front of pass 0's remote stores
    return (int)seb_cluster_rank() * (NttCfg<K>::T / 2) + (int)threadIdx.x;
}
#endif
// words of shared memory one CTA of the pair needs
template <int K>
struct NttSmemHalf
{
    static constexpr uint32_t WORDS = (seb_pad<K>((1u << (NttCfg<K>::LOGN - 1)) - 1u) + 4u) & ~3u;
};

// Full forward NTT of NPOLY polynomials sharing one modulus, executed by the T = n/E threads
// t = 0..T-1 that share `smem` (NPOLY * NttSmem<K>::WORDS words, 16-byte aligned).  On return thread t holds, for each of its
// GPL = E >> R_last groups i, the 2^R_last contiguous coefficients starting at
// NttOut<K>::pos(t, i), still lazy in [0,4q).  All T threads must call (barriers inside).
// The caller must __syncthreads() before smem is reused.
template <int K, int NPOLY, class Loader>
__device__ __forceinline__ void seb_ntt_forward(uint32_t (&x)[NPOLY][NttCfg<K>::E], uint32_t *smem, const int t,
                                                const seb_oct *__restrict__ tw, const uint32_t q, const uint32_t two_q,
                                                Loader &load)
{
    SebNttRun<K, 0, NPOLY, Loader>::run(x, smem, t, tw, q, two_q, load);
}

// The same transform in two calls, for callers whose on-load conversion comes in several variants:
// only pass 0 touches the loader, so only pass 0 is instantiated per variant.
//   seb_ntt_first : pass 0 (+ the barrier that publishes its results)
//   seb_ntt_rest  : passes 1 .. NPASS-1
struct SebNoLoad
{
    __device__ __forceinline__ uint32_t operator()(int, uint32_t) const { return 0u; }
};
template <int K, int NPOLY, class Loader, int CL = 1>
__device__ __forceinline__ void seb_ntt_first(uint32_t (&x)[NPOLY][NttCfg<K>::E], uint32_t *smem, const int t,
                                              const seb_oct *__restrict__ tw, const uint32_t q, const uint32_t two_q,
                                              Loader &load)
{
    seb_ntt_pass<K, 0, NPOLY, Loader, CL>(x, smem, t, tw, q, two_q, load);
    if (CL == 2)
    {
        seb_cluster_arrive();
        seb_cluster_wait();
    }
    else
        __syncthreads();
}
template <int K, int NPOLY, int CL = 1>
__device__ __forceinline__ void seb_ntt_rest(uint32_t (&x)[NPOLY][NttCfg<K>::E], uint32_t *smem, const int t,
                                             const seb_oct *__restrict__ tw, const uint32_t q, const uint32_t two_q)
{
    static_assert(NttPlan<K>::NPASS >= 2, "plan with a single pass");
    SebNoLoad none;
    SebNttRun<K, 1, NPOLY, SebNoLoad, CL>::run(x, smem, t, tw, q, two_q, none);
}

template <int K>
struct NttOut
{
    static constexpr int RL  = NttPlan<K>::R[NttPlan<K>::NPASS - 1];
    static constexpr int GPL = NttCfg<K>::E >> RL;   // contiguous runs per thread
    static constexpr int RUN = 1 << RL;              // coefficients per run
    static constexpr int T   = NttCfg<K>::T;
    __host__ __device__ __forceinline__ static constexpr uint32_t pos(int t, int i)
    {
        return seb_ntt_group<K, NttPlan<K>::NPASS - 1>((uint32_t)t, (uint32_t)i) << RL;
    }
};

// Key tables (pk0, pk1, ntt(s)) in the order the last pass hands coefficients to threads:
// oct (i, k, t) holds the Shoup pairs of the 4 coefficients NttOut::pos(t, i) + 4k .. + 3, and octs
// are interleaved across threads so a warp's 256-bit loads are contiguous.
template <int K>
__host__ __device__ __forceinline__ constexpr uint32_t seb_epi_index(int t, int i, int k)
{
    return (uint32_t)((i * (NttOut<K>::RUN / 4) + k) * NttOut<K>::T + t);
}
template <int K>
inline void seb_build_epi(const uint2 *natural, seb_oct *out)
{
    using O = NttOut<K>;
    for (int t = 0; t < O::T; t++)
        for (int i = 0; i < O::GPL; i++)
            for (int k = 0; k < O::RUN / 4; k++)
                for (int c = 0; c < 4; c++)
                {
                    const uint2 w = natural[O::pos(t, i) + 4 * k + c];
                    seb_oct &o    = out[seb_epi_index<K>(t, i, k)];
                    o.v[2 * c]     = w.x;
                    o.v[2 * c + 1] = w.y;
                }
}
