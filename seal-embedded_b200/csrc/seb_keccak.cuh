// seb_keccak.cuh — Keccak-f[1600] / SHAKE256 for the reference's PRNG, one sponge per thread.
//
// The reference PRNG is SHAKE256(seed[64] || LE64(counter)) with a fresh sponge per call
// (device/lib/rng.h:78-91, device/lib/shake256/fips202.c:105-128): the 72-byte input is shorter than
// the 136-byte rate, so a call is "init state, permute once per 136 output bytes".
// The whole 25-lane state stays in registers; rotations are funnel shifts on 32-bit halves and
// the chi step is one LOP3 per half-lane.
#pragma once

#include "seb_common.cuh"

SEB_CONSTANT uint64_t c_keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

template <int R>
__device__ __forceinline__ uint64_t seb_rotl64(uint64_t x)
{
    if (R == 0) return x;
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t nlo, nhi;
    if (R == 32)
    {
        nlo = hi;
        nhi = lo;
    }
    else if (R < 32)
    {
        nhi = __funnelshift_l(lo, hi, R);
        nlo = __funnelshift_l(hi, lo, R);
    }
    else
    {
        nhi = __funnelshift_l(hi, lo, R - 32);
        nlo = __funnelshift_l(lo, hi, R - 32);
    }
    return ((uint64_t)nhi << 32) | nlo;
}

// 24 rounds, state as 25 named registers (x + 5y indexing).  The round loop is kept rolled
// (one round body ~200 instructions) so the kernel stays inside the instruction cache.
__device__ __forceinline__ void seb_keccak_f1600(uint64_t (&a)[25])
{
#pragma unroll 1
    for (int round = 0; round < 24; round++)
    {
        uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20];
        uint64_t c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
        uint64_t c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22];
        uint64_t c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
        uint64_t c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
        uint64_t d0 = c4 ^ seb_rotl64<1>(c1);
        uint64_t d1 = c0 ^ seb_rotl64<1>(c2);
        uint64_t d2 = c1 ^ seb_rotl64<1>(c3);
        uint64_t d3 = c2 ^ seb_rotl64<1>(c4);
        uint64_t d4 = c3 ^ seb_rotl64<1>(c0);

        // theta + rho + pi: b[y + 5*((2x+3y)%5)] = rotl(a[x+5y] ^ d[x], r[x][y])
        uint64_t b0  = a[0] ^ d0;
        uint64_t b10 = seb_rotl64<1>(a[1] ^ d1);
        uint64_t b20 = seb_rotl64<62>(a[2] ^ d2);
        uint64_t b5  = seb_rotl64<28>(a[3] ^ d3);
        uint64_t b15 = seb_rotl64<27>(a[4] ^ d4);
        uint64_t b16 = seb_rotl64<36>(a[5] ^ d0);
        uint64_t b1  = seb_rotl64<44>(a[6] ^ d1);
        uint64_t b11 = seb_rotl64<6>(a[7] ^ d2);
        uint64_t b21 = seb_rotl64<55>(a[8] ^ d3);
        uint64_t b6  = seb_rotl64<20>(a[9] ^ d4);
        uint64_t b7  = seb_rotl64<3>(a[10] ^ d0);
        uint64_t b17 = seb_rotl64<10>(a[11] ^ d1);
        uint64_t b2  = seb_rotl64<43>(a[12] ^ d2);
        uint64_t b12 = seb_rotl64<25>(a[13] ^ d3);
        uint64_t b22 = seb_rotl64<39>(a[14] ^ d4);
        uint64_t b23 = seb_rotl64<41>(a[15] ^ d0);
        uint64_t b8  = seb_rotl64<45>(a[16] ^ d1);
        uint64_t b18 = seb_rotl64<15>(a[17] ^ d2);
        uint64_t b3  = seb_rotl64<21>(a[18] ^ d3);
        uint64_t b13 = seb_rotl64<8>(a[19] ^ d4);
        uint64_t b14 = seb_rotl64<18>(a[20] ^ d0);
        uint64_t b24 = seb_rotl64<2>(a[21] ^ d1);
        uint64_t b9  = seb_rotl64<61>(a[22] ^ d2);
        uint64_t b19 = seb_rotl64<56>(a[23] ^ d3);
        uint64_t b4  = seb_rotl64<14>(a[24] ^ d4);

        // chi (+ iota on lane 0)
        a[0]  = b0 ^ (~b1 & b2) ^ c_keccak_rc[round];
        a[1]  = b1 ^ (~b2 & b3);
        a[2]  = b2 ^ (~b3 & b4);
        a[3]  = b3 ^ (~b4 & b0);
        a[4]  = b4 ^ (~b0 & b1);
        a[5]  = b5 ^ (~b6 & b7);
        a[6]  = b6 ^ (~b7 & b8);
        a[7]  = b7 ^ (~b8 & b9);
        a[8]  = b8 ^ (~b9 & b5);
        a[9]  = b9 ^ (~b5 & b6);
        a[10] = b10 ^ (~b11 & b12);
        a[11] = b11 ^ (~b12 & b13);
        a[12] = b12 ^ (~b13 & b14);
        a[13] = b13 ^ (~b14 & b10);
        a[14] = b14 ^ (~b10 & b11);
        a[15] = b15 ^ (~b16 & b17);
        a[16] = b16 ^ (~b17 & b18);
        a[17] = b17 ^ (~b18 & b19);
        a[18] = b18 ^ (~b19 & b15);
        a[19] = b19 ^ (~b15 & b16);
        a[20] = b20 ^ (~b21 & b22);
        a[21] = b21 ^ (~b22 & b23);
        a[22] = b22 ^ (~b23 & b24);
        a[23] = b23 ^ (~b24 & b20);
        a[24] = b24 ^ (~b20 & b21);
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
