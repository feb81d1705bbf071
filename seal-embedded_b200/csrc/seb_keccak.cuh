// seb_keccak.cuh — Keccak-f[1600] / SHAKE256 for the reference's PRNG, one sponge per thread.
//
// The reference PRNG is SHAKE256(seed[64] || LE64(counter)) with a fresh sponge per call
// (device/lib/rng.h:78-91, device/lib/shake256/fips202.c:105-128): the 72-byte input is shorter than
// the 136-byte rate, so a call is "init state, permute once per 136 output bytes".
// The whole 25-lane state stays in registers as 32-bit halves; every logic step is a 3-input LOP3 and
// rotations are funnel shifts.
#pragma once

#include "seb_common.cuh"


// 3-input logic ops as single LOP3s.  Spelled out in PTX so the compiler cannot re-factor the theta
// step back into "D = C ^ rot(C); A ^= D" (one more ALU op per column): Keccak on this machine is
// bound by the ALU pipe alone (profiles/README.md), so the op count per round is the run time.
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t seb_xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t seb_chi(uint32_t a, uint32_t b, uint32_t c)  // a ^ (~b & c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
#else
static inline uint32_t seb_xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
static inline uint32_t seb_chi(uint32_t a, uint32_t b, uint32_t c) { return a ^ (~b & c); }
#endif

// 64-bit rotate left on a lane held as two 32-bit halves: two funnel shifts (none for 0 and 32)
template <int R>
__device__ __forceinline__ void seb_rotl64(uint32_t lo, uint32_t hi, uint32_t &olo, uint32_t &ohi)
{
    if constexpr (R == 0)
    {
        olo = lo;
        ohi = hi;
    }
    else if constexpr (R == 32)
    {
        olo = hi;
        ohi = lo;
    }
    else if constexpr (R > 32)
        seb_rotl64<R - 32>(hi, lo, olo, ohi);
    else
    {
        ohi = __funnelshift_l(lo, hi, R);
        olo = __funnelshift_l(hi, lo, R);
    }
}

SEB_CONSTANT uint32_t c_keccak_rc_lo[24] = {
    0x00000001u, 0x00008082u, 0x0000808au, 0x80008000u, 0x0000808bu, 0x80000001u, 0x80008081u, 0x00008009u,
    0x0000008au, 0x00000088u, 0x80008009u, 0x8000000au, 0x8000808bu, 0x0000008bu, 0x00008089u, 0x00008003u,
    0x00008002u, 0x00000080u, 0x0000800au, 0x8000000au, 0x80008081u, 0x00008080u, 0x80000001u, 0x80008008u};
SEB_CONSTANT uint32_t c_keccak_rc_hi[24] = {
    0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u,
    0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u,
    0x80000000u, 0x80000000u, 0x00000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x00000000u, 0x80000000u};

// theta + rho + pi for lane SRC: B[DST] = rotl(A[SRC] ^ C[x-1] ^ rotl(C[x+1], 1), ROT)
#define SEB_KECCAK_RP(SRC, DST, ROT)                                                                  \
    if (!PRUNE || (DST) < 14)                                                                         \
    {                                                                                                 \
        const uint32_t tl_ = X3(lo[SRC], cl[((SRC) % 5 + 4) % 5], rl[((SRC) % 5 + 1) % 5]);          \
        const uint32_t th_ = X3(hi[SRC], ch[((SRC) % 5 + 4) % 5], rh[((SRC) % 5 + 1) % 5]);          \
        seb_rotl64<ROT>(tl_, th_, bl[DST], bh[DST]);                                                  \
    }

// One round on 32-bit halves, 180 ALU operations: theta parities 20 (3-input XORs), their rotations
// 10, theta-apply fused into a 3-input XOR 50, rho 48 funnel shifts, chi 50, iota 2.
// PRUNE: only output lanes 0..11 (the first 96 bytes of the rate) are wanted, which need B lanes
// 0..13 only: 112 operations.
// FIRST: round 0 of a freshly initialised sponge (seb_prng_init), inlined outside the round loop.  Thirteen of its 25
// lanes are zero and two are constants; written with plain operators instead of the opaque LOP3s, the compiler folds
// them (column parities of two inputs, one theta word for the three empty lanes of a column): ~110 operations.
template <bool PRUNE, bool FIRST = false>
__device__ __forceinline__ void seb_keccak_round(uint32_t (&lo)[25], uint32_t (&hi)[25], const int round)
{
    auto X3 = [](uint32_t a, uint32_t b, uint32_t c) { return FIRST ? (a ^ b ^ c) : seb_xor3(a, b, c); };
    uint32_t cl[5], ch[5], rl[5], rh[5], bl[25], bh[25];
#pragma unroll
    for (int x = 0; x < 5; x++)
    {
        cl[x] = X3(X3(lo[x], lo[x + 5], lo[x + 10]), lo[x + 15], lo[x + 20]);
        ch[x] = X3(X3(hi[x], hi[x + 5], hi[x + 10]), hi[x + 15], hi[x + 20]);
    }
#pragma unroll
    for (int x = 0; x < 5; x++) seb_rotl64<1>(cl[x], ch[x], rl[x], rh[x]);
    SEB_KECCAK_RP(0, 0, 0) SEB_KECCAK_RP(1, 10, 1) SEB_KECCAK_RP(2, 20, 62) SEB_KECCAK_RP(3, 5, 28)
    SEB_KECCAK_RP(4, 15, 27) SEB_KECCAK_RP(5, 16, 36) SEB_KECCAK_RP(6, 1, 44) SEB_KECCAK_RP(7, 11, 6)
    SEB_KECCAK_RP(8, 21, 55) SEB_KECCAK_RP(9, 6, 20) SEB_KECCAK_RP(10, 7, 3) SEB_KECCAK_RP(11, 17, 10)
    SEB_KECCAK_RP(12, 2, 43) SEB_KECCAK_RP(13, 12, 25) SEB_KECCAK_RP(14, 22, 39) SEB_KECCAK_RP(15, 23, 41)
    SEB_KECCAK_RP(16, 8, 45) SEB_KECCAK_RP(17, 18, 15) SEB_KECCAK_RP(18, 3, 21) SEB_KECCAK_RP(19, 13, 8)
    SEB_KECCAK_RP(20, 14, 18) SEB_KECCAK_RP(21, 24, 2) SEB_KECCAK_RP(22, 9, 61) SEB_KECCAK_RP(23, 19, 56)
    SEB_KECCAK_RP(24, 4, 14)
#pragma unroll
    for (int y = 0; y < 25; y += 5)
#pragma unroll
        for (int x = 0; x < 5; x++)
            if (!PRUNE || y + x < 12)
            {
                lo[y + x] = seb_chi(bl[y + x], bl[y + (x + 1) % 5], bl[y + (x + 2) % 5]);
                hi[y + x] = seb_chi(bh[y + x], bh[y + (x + 1) % 5], bh[y + (x + 2) % 5]);
            }
    lo[0] ^= c_keccak_rc_lo[round];
    hi[0] ^= c_keccak_rc_hi[round];
}
#undef SEB_KECCAK_RP

// Keccak-f[1600].  NOUT = how many leading lanes of the result the caller reads: 25 for a full
// permutation (sponges that keep squeezing), <= 12 when only the first 96 bytes are used — every
// call of the ternary and centered-binomial samplers (device/lib/sample.c:223-241,311-356) — in
// which case the last round is pruned and lanes >= NOUT of `a` are left unspecified.
// The round loop stays rolled (one round ~190 instructions) so the kernels stay inside the
// instruction cache; tools/ubench measured no gain from unrolling by 2.
// FRESH: `a` comes straight from seb_prng_init (round 0 is specialised on its constant lanes).
template <int NOUT = 25, bool FRESH = false>
__device__ __forceinline__ void seb_keccak_f1600(uint64_t (&a)[25])
{
    constexpr bool PRUNE = NOUT <= 12;
    uint32_t lo[25], hi[25];
#pragma unroll
    for (int i = 0; i < 25; i++)
    {
        lo[i] = (uint32_t)a[i];
        hi[i] = (uint32_t)(a[i] >> 32);
    }
    if (FRESH) seb_keccak_round<false, true>(lo, hi, 0);
#pragma unroll 1
    for (int round = FRESH ? 1 : 0; round < (PRUNE ? 23 : 24); round++) seb_keccak_round<false>(lo, hi, round);
    if (PRUNE) seb_keccak_round<true>(lo, hi, 23);
#pragma unroll
    for (int i = 0; i < (PRUNE ? 12 : 25); i++) a[i] = ((uint64_t)hi[i] << 32) | lo[i];
}

// ---------------------------------------------------------------------------------------------
// The same permutation on a BIT-INTERLEAVED state: lane = (e, o), e bit i = lane bit 2i, o bit i = lane bit 2i + 1.
// A 64-bit rotation by an even amount R is two 32-bit rotations by R/2; by an odd amount it swaps the halves and
// rotates them by (R+1)/2 and (R-1)/2 - so the rotation by 1 of theta's five column parities costs ONE funnel shift
// instead of two, and so does rho's offset 1: 174 ALU operations per round instead of 180 (3.3 % of a kernel that is
// bound by exactly that).  It pays where the squeezed bits can be consumed without de-interleaving them: the
// centered-binomial sampler only counts bits (seb_sample.cuh, seb_cbd_block_il).
// ---------------------------------------------------------------------------------------------
SEB_CONSTANT uint32_t c_keccak_rc_even[24] = {
    0x00000001u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0x00000001u, 0x00000001u, 0x00000001u,
    0x00000000u, 0x00000000u, 0x00000001u, 0x00000000u, 0x00000001u, 0x00000001u, 0x00000001u, 0x00000001u,
    0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0x00000000u, 0x00000001u, 0x00000000u};
SEB_CONSTANT uint32_t c_keccak_rc_odd[24] = {
    0x00000000u, 0x00000089u, 0x8000008bu, 0x80008080u, 0x0000008bu, 0x00008000u, 0x80008088u, 0x80000082u,
    0x0000000bu, 0x0000000au, 0x00008082u, 0x00008003u, 0x0000808bu, 0x8000000bu, 0x8000008au, 0x80000081u,
    0x80000081u, 0x80000008u, 0x00000083u, 0x80008003u, 0x80008088u, 0x80000088u, 0x00008000u, 0x80008082u};

__device__ __forceinline__ uint32_t seb_rotl32(uint32_t x, const int r)  // r: compile-time after inlining
{
    return (r & 31) ? __funnelshift_l(x, x, r & 31) : x;
}
template <int R>
__device__ __forceinline__ void seb_rotl64_il(uint32_t e, uint32_t o, uint32_t &oe, uint32_t &oo)
{
    if constexpr ((R & 1) == 0)
    {
        oe = seb_rotl32(e, R / 2);
        oo = seb_rotl32(o, R / 2);
    }
    else
    {
        oe = seb_rotl32(o, (R + 1) / 2);
        oo = seb_rotl32(e, (R - 1) / 2);
    }
}

#define SEB_KECCAK_RP_IL(SRC, DST, ROT)                                                               \
    if ((DST) < NB)                                                                                   \
    {                                                                                                 \
        const uint32_t te_ = X3(e[SRC], ce[((SRC) % 5 + 4) % 5], re[((SRC) % 5 + 1) % 5]);           \
        const uint32_t to_ = X3(o[SRC], co[((SRC) % 5 + 4) % 5], ce[((SRC) % 5 + 1) % 5]);           \
        seb_rotl64_il<ROT>(te_, to_, be[DST], bo[DST]);                                               \
    }

// FIRST: round 0 of a freshly initialised sponge, inlined outside the round loop.  Thirteen of its 25 lanes are zero and
// two are constants; written with plain operators instead of the opaque LOP3s, the compiler folds them (column parities
// of two inputs, one theta word for the three empty lanes of a column).
// KEEP: how many leading lanes of the result are wanted - 25 (a full round), 12 (the last round of a 96-byte call: B lanes
// 0..13 suffice, 106 operations) or 1 (the last round of a 4-byte call: B lanes 0..2, 37 operations).
template <int KEEP = 25, bool FIRST = false>
__device__ __forceinline__ void seb_keccak_round_il(uint32_t (&e)[25], uint32_t (&o)[25], const int round)
{
    static_assert(KEEP == 25 || KEEP == 12 || KEEP == 1, "supported prunings");
    constexpr int NB = KEEP == 25 ? 25 : KEEP == 12 ? 14 : 3;  // B lanes the wanted outputs read
    auto X3 = [](uint32_t a, uint32_t b, uint32_t c) { return FIRST ? (a ^ b ^ c) : seb_xor3(a, b, c); };
    uint32_t ce[5], co[5], re[5], be[25], bo[25];
#pragma unroll
    for (int x = 0; x < 5; x++)
    {
        ce[x] = X3(X3(e[x], e[x + 5], e[x + 10]), e[x + 15], e[x + 20]);
        co[x] = X3(X3(o[x], o[x + 5], o[x + 10]), o[x + 15], o[x + 20]);
    }
    // rotl64(C, 1): even half = rotl32(co, 1), odd half = ce (a rename: SEB_KECCAK_RP_IL reads ce there)
#pragma unroll
    for (int x = 0; x < 5; x++) re[x] = seb_rotl32(co[x], 1);
    SEB_KECCAK_RP_IL(0, 0, 0) SEB_KECCAK_RP_IL(1, 10, 1) SEB_KECCAK_RP_IL(2, 20, 62) SEB_KECCAK_RP_IL(3, 5, 28)
    SEB_KECCAK_RP_IL(4, 15, 27) SEB_KECCAK_RP_IL(5, 16, 36) SEB_KECCAK_RP_IL(6, 1, 44) SEB_KECCAK_RP_IL(7, 11, 6)
    SEB_KECCAK_RP_IL(8, 21, 55) SEB_KECCAK_RP_IL(9, 6, 20) SEB_KECCAK_RP_IL(10, 7, 3) SEB_KECCAK_RP_IL(11, 17, 10)
    SEB_KECCAK_RP_IL(12, 2, 43) SEB_KECCAK_RP_IL(13, 12, 25) SEB_KECCAK_RP_IL(14, 22, 39) SEB_KECCAK_RP_IL(15, 23, 41)
    SEB_KECCAK_RP_IL(16, 8, 45) SEB_KECCAK_RP_IL(17, 18, 15) SEB_KECCAK_RP_IL(18, 3, 21) SEB_KECCAK_RP_IL(19, 13, 8)
    SEB_KECCAK_RP_IL(20, 14, 18) SEB_KECCAK_RP_IL(21, 24, 2) SEB_KECCAK_RP_IL(22, 9, 61) SEB_KECCAK_RP_IL(23, 19, 56)
    SEB_KECCAK_RP_IL(24, 4, 14)
#pragma unroll
    for (int y = 0; y < 25; y += 5)
#pragma unroll
        for (int x = 0; x < 5; x++)
            if (y + x < KEEP)
            {
                e[y + x] = seb_chi(be[y + x], be[y + (x + 1) % 5], be[y + (x + 2) % 5]);
                o[y + x] = seb_chi(bo[y + x], bo[y + (x + 1) % 5], bo[y + (x + 2) % 5]);
            }
    e[0] ^= c_keccak_rc_even[round];
    o[0] ^= c_keccak_rc_odd[round];
}
#undef SEB_KECCAK_RP_IL

// Keccak-f[1600] on the interleaved state; the last round is pruned to output lanes 0..11 (96 bytes)
__device__ __forceinline__ void seb_keccak_f1600_il12(uint32_t (&e)[25], uint32_t (&o)[25])
{
    seb_keccak_round_il<25, true>(e, o, 0);
#pragma unroll 1
    for (int round = 1; round < 23; round++) seb_keccak_round_il<25>(e, o, round);
    seb_keccak_round_il<12>(e, o, 23);
}

// the even (odd = 0) or odd (odd = 1) bits of a 32-bit word, packed into 16
__device__ __forceinline__ uint32_t seb_half_bits32(uint32_t w, const int odd)
{
    uint32_t t = (w >> odd) & 0x55555555u;
    t          = (t | (t >> 1)) & 0x33333333u;
    t          = (t | (t >> 2)) & 0x0F0F0F0Fu;
    t          = (t | (t >> 4)) & 0x00FF00FFu;
    t          = (t | (t >> 8)) & 0x0000FFFFu;
    return t;
}
// ... of a 64-bit lane, packed into 32
__device__ __forceinline__ uint32_t seb_half_bits(uint64_t w, const int odd)
{
    return seb_half_bits32((uint32_t)w, odd) | (seb_half_bits32((uint32_t)(w >> 32), odd) << 16);
}

// interleaved absorb of (seed || LE64(counter)): se/so = the seed's eight lanes already split, counter < 2^33
__device__ __forceinline__ void seb_prng_init_il(uint32_t (&e)[25], uint32_t (&o)[25], const uint32_t (&se)[8],
                                                 const uint32_t (&so)[8], uint64_t counter)
{
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = se[i], o[i] = so[i];
    e[8] = seb_half_bits32((uint32_t)counter, 0);
    o[8] = seb_half_bits32((uint32_t)counter, 1);
    if (const uint32_t chi = (uint32_t)(counter >> 32))  // the samplers' counters are small: never taken in practice
    {
        e[8] |= seb_half_bits32(chi, 0) << 16;
        o[8] |= seb_half_bits32(chi, 1) << 16;
    }
    e[9] = 0x7u;  // 0x1F: bits 0, 2, 4 | bits 1, 3
    o[9] = 0x3u;
#pragma unroll
    for (int i = 10; i < 25; i++) e[i] = 0u, o[i] = 0u;
    o[16] = 0x80000000u;  // bit 63
}

// v = bytes [a0, b0, a1, b1] of two 16-bit values a, b  ->  bit 2i = a_i, bit 2i+1 = b_i: the last three steps of the
// 32-bit perfect shuffle (the first, a swap of the two middle bytes, is folded into the byte permute that builds v)
__device__ __forceinline__ uint32_t seb_interleave_tail(uint32_t v)
{
    uint32_t t;
    t = (v ^ (v >> 4)) & 0x00F000F0u;
    v ^= t ^ (t << 4);
    t = (v ^ (v >> 2)) & 0x0C0C0C0Cu;
    v ^= t ^ (t << 2);
    t = (v ^ (v >> 1)) & 0x22222222u;
    v ^= t ^ (t << 1);
    return v;
}

// LE32 of the first four bytes of SHAKE256(seed || LE64(counter)) - a redraw of the uniform sampler (sample.c:39-57) -
// from a fresh interleaved sponge: round 0 folded, 22 full rounds, the last one pruned to the one word that is read
__device__ __forceinline__ uint32_t seb_prng_word_il(const uint32_t (&se)[8], const uint32_t (&so)[8], uint64_t counter)
{
    uint32_t e[25], o[25];
    seb_prng_init_il(e, o, se, so, counter);
    seb_keccak_round_il<25, true>(e, o, 0);
#pragma unroll 1
    for (int round = 1; round < 23; round++) seb_keccak_round_il<25>(e, o, round);
    seb_keccak_round_il<1>(e, o, 23);
    return seb_interleave_tail(__byte_perm(e[0], o[0], 0x5140));
}
// the seed's eight lanes split into even and odd bits by one thread
__device__ __forceinline__ void seb_seed_split(const uint8_t *seeds, size_t b, uint32_t (&se)[8], uint32_t (&so)[8])
{
    const uint64_t *p = reinterpret_cast<const uint64_t *>(seeds + b * SEB_SEED_BYTES);
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        const uint64_t w = __ldg(p + i);
        se[i] = seb_half_bits(w, 0), so[i] = seb_half_bits(w, 1);
    }
}

// SHAKE256 absorb of (seed || LE64(counter)): 72 bytes, domain byte 0x1F at offset 72, final bit
// 0x80 at offset 135 (device/lib/shake256/fips202.c:46-66).  seed8 = the seed as 8 LE words.
__device__ __forceinline__ void seb_prng_init(uint64_t (&a)[25], const uint64_t (&seed8)[8], uint64_t counter)
{
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed8[i];
    a[8] = counter;
    a[9] = 0x1FULL;
#pragma unroll
    for (int i = 10; i < 25; i++) a[i] = 0;
    a[16] = 0x8000000000000000ULL;
}
