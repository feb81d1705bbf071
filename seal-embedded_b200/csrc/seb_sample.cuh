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

// One 96-byte PRNG block as a ternary block (sample.c:223-241): packed[6] = the 24 output bytes
// (MSB-first 2-bit fields, rejected slots left 0), m0..m2 = bit i set when byte i >= 0xFE.
__device__ __forceinline__ void seb_ternary_block(const uint64_t (&a)[25], uint32_t (&packed)[6], uint32_t &m0,
                                                  uint32_t &m1, uint32_t &m2)
{
    m0 = m1 = m2 = 0;
#pragma unroll
    for (int k = 0; k < 24; k++)
    {
        const uint32_t w  = (k & 1) ? (uint32_t)(a[k >> 1] >> 32) : (uint32_t)a[k >> 1];
        const uint32_t ge = __vcmpgeu4(w, 0xFEFEFEFEu);
        const uint32_t nb = ((ge & 0x01010101u) * 0x01020408u) >> 24;  // bit i = byte i rejected
        if (k < 8)
            m0 |= nb << (4 * k);
        else if (k < 16)
            m1 |= nb << (4 * (k - 8));
        else
            m2 |= nb << (4 * (k - 16));
        const uint32_t r    = seb_mod3_bytes(w) & ~ge;      // rejected slots stay 0 for now
        const uint32_t byte = (r * 0x40100401u) >> 24;      // v0<<6 | v1<<4 | v2<<2 | v3
        if ((k & 3) == 0)
            packed[k >> 2] = byte;
        else
            packed[k >> 2] |= byte << (8 * (k & 3));
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
