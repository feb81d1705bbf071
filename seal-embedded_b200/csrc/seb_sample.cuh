// seb_sample.cuh — per-thread pieces of the samplers (shared by the kernels in seb_sample.cu and
// by the g++ host emulation in tests/host_emul).
#pragma once

#include "seb_keccak.cuh"

// per-byte x mod 3 on four packed bytes (exact for all byte values; modulo.h:150-164 is the
// reference's constant-time equivalent for r < 0xFE)
__device__ __forceinline__ uint32_t seb_mod3_bytes(uint32_t x)
{
    uint32_t r = ((x >> 4) & 0x0F0F0F0Fu) + (x & 0x0F0F0F0Fu);  // <= 30
    r          = ((r >> 2) & 0x07070707u) + (r & 0x03030303u);  // <= 10
    r          = ((r >> 2) & 0x03030303u) + (r & 0x03030303u);  // <= 5
    r          = ((r >> 2) & 0x01010101u) + (r & 0x03030303u);  // <= 3
    const uint32_t three = ((r + 0x01010101u) >> 2) & 0x01010101u;
    return r - (three | (three << 1));
}

// One 96-byte PRNG block as a ternary block (sample.c:223-241): packed[6] = the 24 output bytes (MSB-first 2-bit
// fields), m0..m2 = bit i set when byte i >= 0xFE (rejected).  The fields of rejected bytes hold a DON'T-CARE value (their
// byte mod 3): the walk overwrites them with the redraws.
//
// The samplers are bound by the ALU pipe (LOP3/SHF/PRMT/IADD: Keccak-f keeps it 98 % busy) while the FMA pipe idles, so
// this is written to put its arithmetic into IMADs - 9 ALU + 9 IMAD instructions per four bytes instead of ~30 ALU:
//  * rejected flags: z = ~w & 0xFE..FE has a zero byte exactly where w has 0xFE/0xFF; z's bytes are even, so the borrow
//    of (z - 0x01..01) never produces a false positive and (z - 0x01..01) & ~z & 0x80..80 is exact; one multiplication
//    gathers the four flags (bits 7, 15, 23, 31) into bits 28..31 and a funnel shift appends them to the mask;
//  * x mod 3 for two bytes per word (16-bit fields): x * 171 = 512 floor(x / 3) + l (exact for x < 256), so
//    x * 512 - 3 * (x * 171 & 0xFE00) = 512 (x mod 3) - also when the fields' partial products overflow into each other,
//    the identity holds modulo 2^32;
//  * the four 2-bit results (bits 9-10 and 25-26 of two words) are moved to one byte by two more multiplications whose
//    partial products do not collide, and a byte permute drops it into the output word.
__device__ __forceinline__ void seb_ternary_block(const uint64_t (&a)[25], uint32_t (&packed)[6], uint32_t &m0,
                                                  uint32_t &m1, uint32_t &m2)
{
    m0 = m1 = m2 = 0;
#pragma unroll
    for (int g = 0; g < 6; g++) packed[g] = 0;
#pragma unroll
    for (int k = 23; k >= 0; k--)  // downwards: the funnel shift pushes earlier words towards the high nibbles
    {
        const uint32_t w    = (k & 1) ? (uint32_t)(a[k >> 1] >> 32) : (uint32_t)a[k >> 1];
        const uint32_t z    = ~w & 0xFEFEFEFEu;
        const uint32_t ge   = (z - 0x01010101u) & ~z & 0x80808080u;  // bit 8i+7: byte i rejected
        const uint32_t nib  = ge * 0x00204081u;                       // ... gathered in bits 28 + i
        if (k < 8)
            m0 = __funnelshift_l(nib, m0, 4);
        else if (k < 16)
            m1 = __funnelshift_l(nib, m1, 4);
        else
            m2 = __funnelshift_l(nib, m2, 4);
        const uint32_t e  = __byte_perm(w, 0u, 0x4240);  // byte 0 | byte 2 << 16
        const uint32_t o  = __byte_perm(w, 0u, 0x4143);  // byte 3 | byte 1 << 16
        const uint32_t re = e * 512u - 3u * ((e * 171u) & 0xFE00FE00u);  // v0 at bits 9-10, v2 at bits 25-26
        const uint32_t ro = o * 512u - 3u * ((o * 171u) & 0xFE00FE00u);  // v3 at bits 9-10, v1 at bits 25-26
        // re * (2^21 + 2): v0 -> bits 30-31, v2 -> 26-27 (and v0 -> 10-11); ro * (2^15 + 2^3): v3 -> 24-25, v1 -> 28-29
        // (and v3 -> 12-13): the top byte is v0 << 6 | v1 << 4 | v2 << 2 | v3
        const uint32_t sum = re * 0x00200002u + ro * 0x00008008u;
        packed[k >> 2] = __byte_perm(packed[k >> 2], sum, (0x3210u & ~(0xFu << (4 * (k & 3)))) | (7u << (4 * (k & 3))));
    }
}

// centered binomial, k = 21 (sample.c:263-284): 6 bytes per sample
__device__ __forceinline__ int cbd_one(uint64_t v48)
{
    const uint32_t pos = (uint32_t)v48 & 0x1FFFFFu;          // bytes 0,1 and low 5 bits of byte 2
    const uint32_t neg = (uint32_t)(v48 >> 24) & 0x1FFFFFu;  // bytes 3,4 and low 5 bits of byte 5
    return __popc(pos) - __popc(neg);
}

// One 96-byte PRNG block as 16 CBD samples packed as int8 (sample.c:311-321)
__device__ __forceinline__ void seb_cbd_block(const uint64_t (&a)[25], uint32_t (&o)[4])
{
#pragma unroll
    for (int g = 0; g < 4; g++)
    {
        const uint64_t l0 = a[3 * g], l1 = a[3 * g + 1], l2 = a[3 * g + 2];
        const int s0      = cbd_one(l0);
        const int s1      = cbd_one((l0 >> 48) | (l1 << 16));
        const int s2      = cbd_one((l1 >> 32) | (l2 << 32));
        const int s3      = cbd_one(l2 >> 16);
        o[g] = (uint32_t)(s0 & 0xFF) | ((uint32_t)(s1 & 0xFF) << 8) | ((uint32_t)(s2 & 0xFF) << 16) |
               ((uint32_t)(s3 & 0xFF) << 24);
    }
}

// The same 16 samples from a BIT-INTERLEAVED block (seb_keccak.cuh: e[l] bit i = lane l bit 2i, o[l] bit i = bit 2i + 1).
// A sample is popcount(21 bits) - popcount(21 bits), and a popcount does not care where its bits sit: the 21-bit ranges
// [48s, 48s + 21) and [48s + 24, 48s + 45) of the stream become 11 even + 10 odd positions each, in one or two lanes.
// With popc(x) - popc(y) = popc(x) + popc(~y & mask) - |mask| the positive and the negative bits that share a word are
// counted by ONE population count of (w ^ neg_mask) & (pos_mask | neg_mask): 12 counts per four samples.
__host__ __device__ constexpr uint32_t seb_il_mask(int lane, int odd, int b0, int len)  // positions [b0, b0+len) of the stream
{
    uint32_t m = 0;
    for (int i = 0; i < 32; i++)
    {
        const int pos = 64 * lane + 2 * i + odd;
        if (pos >= b0 && pos < b0 + len) m |= 1u << i;
    }
    return m;
}
// (x ^ n) & m with constant n and m as ONE LOP3: an instruction takes one immediate, so the compiler splits the
// expression into two ALU operations; with m forced into a register the second becomes a move, which goes to the idle
// FMA pipe.  n = 0 is a plain AND.
template <uint32_t N, uint32_t M>
__device__ __forceinline__ uint32_t seb_xor_and(uint32_t x)
{
#ifdef __CUDACC__
    if constexpr (N == 0u)
        return x & M;
    else
    {
        uint32_t r;
        asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(r) : "r"(x), "n"(N), "r"(M));
        return r;
    }
#else
    return (x ^ N) & M;
#endif
}
// Adds the counts of sample S found in lane L to byte S % 4 of acc.  The counts of one sample sum to s + 21 in [0, 42]:
// no carry between the bytes.  Spelled as multiply-adds so that they issue on the FMA pipe.
template <int S, int L>
__device__ __forceinline__ void seb_cbd_il_part(const uint32_t (&e)[25], const uint32_t (&o)[25], uint32_t &acc)
{
    constexpr uint32_t pe = seb_il_mask(L, 0, 48 * S, 21), ne = seb_il_mask(L, 0, 48 * S + 24, 21);
    constexpr uint32_t po = seb_il_mask(L, 1, 48 * S, 21), no = seb_il_mask(L, 1, 48 * S + 24, 21);
    constexpr uint32_t w  = 1u << (8 * (S & 3));
    if constexpr ((pe | ne) != 0u) acc = (uint32_t)__popc(seb_xor_and<ne, (pe | ne)>(e[L])) * w + acc;
    if constexpr ((po | no) != 0u) acc = (uint32_t)__popc(seb_xor_and<no, (po | no)>(o[L])) * w + acc;
}
template <int S>
__device__ __forceinline__ void seb_cbd_il_one(const uint32_t (&e)[25], const uint32_t (&o)[25], uint32_t &acc)
{
    constexpr int L0 = (48 * S) / 64, L1 = (48 * S + 44) / 64;
    seb_cbd_il_part<S, L0>(e, o, acc);
    if constexpr (L1 != L0) seb_cbd_il_part<S, L1>(e, o, acc);
}
// four samples as int8: byte j collects s_j + 21 on top of 0x80 - 21 = 0x6B (bit 7 keeps the per-byte subtraction of
// the bias from borrowing), and flipping bit 7 afterwards leaves s_j modulo 256
template <int G>
__device__ __forceinline__ uint32_t seb_cbd_il_word(const uint32_t (&e)[25], const uint32_t (&o)[25])
{
    uint32_t acc = 0x6B6B6B6Bu;
    seb_cbd_il_one<4 * G>(e, o, acc);
    seb_cbd_il_one<4 * G + 1>(e, o, acc);
    seb_cbd_il_one<4 * G + 2>(e, o, acc);
    seb_cbd_il_one<4 * G + 3>(e, o, acc);
    return acc ^ 0x80808080u;
}
__device__ __forceinline__ void seb_cbd_block_il(const uint32_t (&e)[25], const uint32_t (&o)[25], uint32_t (&out)[4])
{
    out[0] = seb_cbd_il_word<0>(e, o);
    out[1] = seb_cbd_il_word<1>(e, o);
    out[2] = seb_cbd_il_word<2>(e, o);
    out[3] = seb_cbd_il_word<3>(e, o);
}
