/*
 * se_oracle.c — plain-C CPU restatement of SEAL-Embedded's CKKS encode+encrypt path.
 *
 * TEST INFRASTRUCTURE ONLY (see se_oracle.h).  Parity: PINNED against oracle/_ref (the
 * reference library compiled from /root/reference/device/lib) and tests/golden fixtures.
 *
 * Build: gcc -O2 -std=gnu11 -ffp-contract=off -fPIC -shared se_oracle.c -lm   (no -march=native,
 * no -ffast-math: the FP64 encode must round every operation exactly like the reference build).
 *
 * References are to /root/reference/device/lib/<file>:<line>.
 */
#include "se_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------
 * SHAKE256 (FIPS 202).  Restates shake256/fips202.c:30-128 + keccakf1600.c (unrolled there;
 * loop form here).  Round constants come from the FIPS-202 LFSR, rotation offsets from the
 * (x,y) -> (y, 2x+3y) walk, so no table is transcribed.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t rotl64(uint64_t v, unsigned r)
{
    return r ? (v << r) | (v >> (64 - r)) : v;
}

void orc_keccak_f1600(uint64_t A[25])
{
    uint8_t lfsr = 1;
    for (int round = 0; round < 24; round++)
    {
        uint64_t C[5], B[5];
        /* theta */
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
        for (int x = 0; x < 5; x++)
        {
            uint64_t D = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
            for (int y = 0; y < 5; y++) A[x + 5 * y] ^= D;
        }
        /* rho + pi */
        {
            int x = 1, y = 0;
            uint64_t cur = A[1];
            for (int t = 0; t < 24; t++)
            {
                unsigned r   = (unsigned)(((t + 1) * (t + 2) / 2) % 64);
                int X        = y;
                int Y        = (2 * x + 3 * y) % 5;
                uint64_t tmp = A[X + 5 * Y];
                A[X + 5 * Y] = rotl64(cur, r);
                cur          = tmp;
                x            = X;
                y            = Y;
            }
        }
        /* chi */
        for (int y = 0; y < 5; y++)
        {
            for (int x = 0; x < 5; x++) B[x] = A[x + 5 * y];
            for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x] ^ (~B[(x + 1) % 5] & B[(x + 2) % 5]);
        }
        /* iota */
        for (int j = 0; j < 7; j++)
        {
            int bit = lfsr & 1;
            lfsr    = (uint8_t)((lfsr & 0x80) ? ((lfsr << 1) ^ 0x71) : (lfsr << 1));
            if (bit) A[0] ^= (uint64_t)1 << ((1u << j) - 1);
        }
    }
}

#define SHAKE256_RATE 136

void orc_shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen)
{
    uint64_t st[25];
    uint8_t blk[SHAKE256_RATE];
    memset(st, 0, sizeof st);
    /* absorb (fips202.c:46-66): full blocks, then pad10*1 with domain byte 0x1F */
    while (inlen >= SHAKE256_RATE)
    {
        for (int i = 0; i < SHAKE256_RATE / 8; i++)
        {
            uint64_t w;
            memcpy(&w, in + 8 * i, 8);
            st[i] ^= w;
        }
        orc_keccak_f1600(st);
        in += SHAKE256_RATE;
        inlen -= SHAKE256_RATE;
    }
    memset(blk, 0, sizeof blk);
    memcpy(blk, in, inlen);
    blk[inlen] = 0x1F;
    blk[SHAKE256_RATE - 1] |= 0x80;
    for (int i = 0; i < SHAKE256_RATE / 8; i++)
    {
        uint64_t w;
        memcpy(&w, blk + 8 * i, 8);
        st[i] ^= w;
    }
    /* squeeze (fips202.c:80-128): permute before every block */
    while (outlen)
    {
        size_t take = outlen < SHAKE256_RATE ? outlen : SHAKE256_RATE;
        orc_keccak_f1600(st);
        memcpy(blk, st, SHAKE256_RATE); /* little-endian host */
        memcpy(out, blk, take);
        out += take;
        outlen -= take;
    }
}

/* rng.h:78-91: SHAKE256(seed || LE64(counter)); the caller advances the counter. */
void orc_prng_fill(const uint8_t seed[ORC_SEED_BYTES], uint64_t counter, size_t nbytes,
                   uint8_t *out)
{
    uint8_t in[ORC_SEED_BYTES + 8];
    memcpy(in, seed, ORC_SEED_BYTES);
    memcpy(in + ORC_SEED_BYTES, &counter, 8);
    orc_shake256(out, nbytes, in, sizeof in);
}

/* ------------------------------------------------------------------------------------------
 * Parameter tables.  parameters.c:129-174 (prime chains), :191-227 (legal (n,np), scale),
 * ntt.c:213-289 (psi per (n,q)), modulus.c:30-47 (const_ratio = floor(2^64/q)).
 * ---------------------------------------------------------------------------------------- */
static const uint32_t k_primes27[3]  = {134012929u, 134111233u, 134176769u};
static const uint32_t k_primes30[13] = {1053818881u, 1054015489u, 1054212097u, 1055260673u,
                                        1056178177u, 1056440321u, 1058209793u, 1060175873u,
                                        1060700161u, 1060765697u, 1061093377u, 1062469633u,
                                        1062535169u};

int orc_default_primes(size_t n, size_t np, uint32_t *primes)
{
    const uint32_t *src;
    size_t maxp;
    switch (n)
    {
        case 1024:
        case 2048: src = k_primes27; maxp = 1; break;
        case 4096: src = k_primes30; maxp = 3; break;
        case 8192: src = k_primes30; maxp = 6; break;
        case 16384: src = k_primes30; maxp = 13; break;
        default: return 0;
    }
    if (np < 1 || np > maxp) return 0;
    for (size_t i = 0; i < np; i++) primes[i] = src[i];
    return 1;
}

double orc_default_scale(size_t n)
{
    return n == 1024 ? 1048576.0 : 33554432.0; /* 2^20 / 2^25, parameters.c:197-225 */
}

uint32_t orc_ntt_root(size_t n, uint32_t q)
{
    static const uint32_t psi4k27[3]  = {7470u, 3856u, 24149u};
    static const uint32_t psi4k30[3]  = {503422u, 16768u, 7305u};
    static const uint32_t psi8k[6]    = {374229u, 123363u, 79941u, 38869u, 162146u, 81884u};
    static const uint32_t psi16k[13]  = {13040u, 507u,   1595u,   68507u,  3073u,   6854u, 44467u,
                                         16117u, 27607u, 222391u, 105471u, 310222u, 2005u};
    if (n == 1024) return q == 134012929u ? 142143u : 0;
    if (n == 2048) return q == 134012929u ? 85250u : 0;
    if (n == 4096)
    {
        for (int i = 0; i < 3; i++)
        {
            if (q == k_primes27[i]) return psi4k27[i];
            if (q == k_primes30[i]) return psi4k30[i];
        }
        return 0;
    }
    if (n == 8192)
    {
        for (int i = 0; i < 6; i++)
            if (q == k_primes30[i]) return psi8k[i];
        return 0;
    }
    if (n == 16384)
    {
        for (int i = 0; i < 13; i++)
            if (q == k_primes30[i]) return psi16k[i];
        return 0;
    }
    return 0;
}

void orc_const_ratio(uint32_t q, uint32_t ratio[2])
{
    unsigned __int128 one = (unsigned __int128)1 << 64;
    uint64_t r            = (uint64_t)(one / q);
    ratio[0]              = (uint32_t)r;
    ratio[1]              = (uint32_t)(r >> 32);
}

/* ------------------------------------------------------------------------------------------
 * Modular arithmetic.
 * ---------------------------------------------------------------------------------------- */
/* modulo.h:21-32 shift_result: one conditional subtraction */
static inline uint32_t cond_sub(uint32_t x, uint32_t q)
{
    return x >= q ? x - q : x;
}

/* modulo.h:43-75: t = hi32(x * floor(2^64/q).hi); r = x - t*q; one correction */
uint32_t orc_barrett32(uint32_t x, uint32_t q)
{
    uint32_t ratio[2];
    orc_const_ratio(q, ratio);
    uint32_t t = (uint32_t)(((uint64_t)x * ratio[1]) >> 32);
    return cond_sub(x - t * q, q);
}

/* modulo.h:84-116: t = floor(x * floor(2^64/q) / 2^64) mod 2^32 (the word-by-word carry chain
 * there computes exactly this); r = lo32(x) - t*q; one correction */
uint32_t orc_barrett64(uint32_t lo, uint32_t hi, uint32_t q)
{
    uint32_t ratio[2];
    orc_const_ratio(q, ratio);
    uint64_t x            = ((uint64_t)hi << 32) | lo;
    uint64_t r64          = ((uint64_t)ratio[1] << 32) | ratio[0];
    unsigned __int128 big = (unsigned __int128)x * r64;
    uint32_t t            = (uint32_t)(uint64_t)(big >> 64);
    return cond_sub(lo - t * q, q);
}

uint32_t orc_add_mod(uint32_t a, uint32_t b, uint32_t q) /* uintmodarith.h:26-33 */
{
    return cond_sub(a + b, q);
}
uint32_t orc_neg_mod(uint32_t a, uint32_t q) /* uintmodarith.h:56-62 */
{
    return a ? q - a : 0;
}
uint32_t orc_sub_mod(uint32_t a, uint32_t b, uint32_t q) /* uintmodarith.h:83-88 */
{
    return orc_add_mod(a, orc_neg_mod(b, q), q);
}
uint32_t orc_mul_mod(uint32_t a, uint32_t b, uint32_t q) /* uintmodarith.h:123-128 */
{
    uint64_t p = (uint64_t)a * b;
    return orc_barrett64((uint32_t)p, (uint32_t)(p >> 32), q);
}
uint32_t orc_pow_mod(uint32_t a, uint64_t e, uint32_t q)
{
    uint32_t r = 1;
    while (e)
    {
        if (e & 1) r = orc_mul_mod(r, a, q);
        a = orc_mul_mod(a, a, q);
        e >>= 1;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Encode.
 * ---------------------------------------------------------------------------------------- */
static size_t ilog2(size_t n)
{
    size_t l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

size_t orc_bitrev(size_t x, size_t nbits) /* fft.h:48-55 */
{
    size_t r = 0;
    for (size_t i = 0; i < nbits; i++) r |= ((x >> i) & 1) << (nbits - 1 - i);
    return r;
}

/* ckks_common.c:32-68: generator 3 of the odd residues mod 2n; slot i and its conjugate slot */
void orc_index_map(size_t n, uint16_t *map)
{
    size_t logn  = ilog2(n);
    uint64_t m   = 2 * (uint64_t)n;
    uint64_t pos = 1;
    for (size_t i = 0; i < n / 2; i++)
    {
        size_t a       = (size_t)((pos - 1) / 2);
        size_t b       = n - 1 - a;
        map[i]         = (uint16_t)orc_bitrev(a, logn);
        map[i + n / 2] = (uint16_t)orc_bitrev(b, logn);
        pos            = (pos * 3) & (m - 1);
    }
}

/* fft.c:27-45 + :129: s(h+j) = conj(cos t + i sin t), t = 2*pi*k/(2n), k = bitrev(h+j, logn).
 * Evaluation order of the angle follows calc_angle: ((2*M_PI)*k)/m. */
void orc_ifft_twiddles(size_t n, double *tw)
{
    size_t logn = ilog2(n);
    size_t m    = 2 * n;
    tw[0] = 1.0;
    tw[1] = 0.0;
    for (size_t i = 1; i < n; i++)
    {
        size_t k     = orc_bitrev(i, logn) & (m - 1);
        double angle = 2 * M_PI * (double)k / (double)m;
        tw[2 * i]     = cos(angle);
        tw[2 * i + 1] = -sin(angle);
    }
}

/* ckks_common.c:105-215 (scatter, ifft_inpl, scale, round) with fft.c:69-144 inlined.
 * Complex product written out the way GCC expands it without -ffast-math on x86-64 (no FMA):
 * re = a*c - b*d, im = a*d + b*c, every operation individually rounded. */
int orc_encode(size_t n, double scale, const float *values, size_t vlen, int64_t *out)
{
    size_t logn   = ilog2(n);
    uint16_t *map = malloc(n * sizeof *map);
    double *x     = calloc(2 * n, sizeof *x); /* interleaved re,im */
    double *tw    = malloc(2 * n * sizeof *tw);
    int ok        = 1;
    orc_index_map(n, map);
    orc_ifft_twiddles(n, tw);
    if (vlen > n / 2) vlen = n / 2;
    for (size_t i = 0; i < n / 2; i++)
    {
        double v             = i < vlen ? (double)values[i] : 0.0;
        x[2 * map[i]]        = v;
        x[2 * map[i + n / 2]] = v;
    }
    size_t tt = 1, h = n / 2;
    for (size_t r = 0; r < logn; r++, tt *= 2, h /= 2)
    {
        for (size_t j = 0, k0 = 0; j < h; j++, k0 += 2 * tt)
        {
            double sr = tw[2 * (h + j)], si = tw[2 * (h + j) + 1];
            for (size_t k = k0; k < k0 + tt; k++)
            {
                double ur = x[2 * k], ui = x[2 * k + 1];
                double vr = x[2 * (k + tt)], vi = x[2 * (k + tt) + 1];
                double dr = ur - vr, di = ui - vi;
                x[2 * k]            = ur + vr;
                x[2 * k + 1]        = ui + vi;
                x[2 * (k + tt)]     = dr * sr - di * si;
                x[2 * (k + tt) + 1] = dr * si + di * sr;
            }
        }
    }
    double n_inv = scale / (double)n;
    for (size_t i = 0; i < n; i++)
    {
        double c = round(x[2 * i] * n_inv);
        if (fabs(c) > 9223372036854775808.0) /* (double)0x7FFF...F == 2^63, ckks_common.c:30 */
        {
            ok = 0;
            break;
        }
        out[i] = (int64_t)c;
    }
    free(map);
    free(x);
    free(tw);
    return ok;
}

/* device/test/ckks_tests_common.c:59-118 with fft.c:146-213 (roots on the fly, not conjugated) */
void orc_decode(size_t n, double scale, uint32_t q, const uint32_t *pt, size_t vlen, float *values)
{
    size_t logn   = ilog2(n);
    size_t m      = 2 * n;
    uint16_t *map = malloc(n * sizeof *map);
    double *x     = calloc(2 * n, sizeof *x);
    orc_index_map(n, map);
    for (size_t i = 0; i < n; i++)
    {
        uint32_t v = pt[i];
        double d   = (v > q / 2) ? -(double)(q - v) : (double)v;
        x[2 * i]   = d / scale;
    }
    size_t h = 1, tt = n / 2;
    for (size_t r = 0; r < logn; r++, h *= 2, tt /= 2)
    {
        for (size_t j = 0, k0 = 0; j < h; j++, k0 += 2 * tt)
        {
            size_t kk    = orc_bitrev(h + j, logn) & (m - 1);
            double angle = 2 * M_PI * (double)kk / (double)m;
            double sr = cos(angle), si = sin(angle);
            for (size_t k = k0; k < k0 + tt; k++)
            {
                double ur = x[2 * k], ui = x[2 * k + 1];
                double ar = x[2 * (k + tt)], ai = x[2 * (k + tt) + 1];
                double vr = ar * sr - ai * si, vi = ar * si + ai * sr;
                x[2 * k]            = ur + vr;
                x[2 * k + 1]        = ui + vi;
                x[2 * (k + tt)]     = ur - vr;
                x[2 * (k + tt) + 1] = ui - vi;
            }
        }
    }
    for (size_t i = 0; i < vlen; i++) values[i] = (float)x[2 * map[i]];
    free(map);
    free(x);
}

/* ------------------------------------------------------------------------------------------
 * Samplers.
 * ---------------------------------------------------------------------------------------- */
/* modulo.h:150-164 is an exact r % 3 for r < 0xFE */
/* sample.c:218-242 + :61-87: 96-byte block per 96 coefficients; a byte >= 0xFE is redrawn with
 * single-byte PRNG calls (each bumps the counter); value r%3 stored MSB-first, 4 per byte. */
void orc_sample_ternary_small(size_t n, const uint8_t *seed, uint64_t *counter, uint8_t *packed)
{
    memset(packed, 0, n / 4);
    for (size_t j = 0; j < n; j += 96)
    {
        uint8_t buf[96];
        orc_prng_fill(seed, (*counter)++, 96, buf);
        size_t stop = (j + 96 <= n) ? 96 : n - j;
        for (size_t i = 0; i < stop; i++)
        {
            uint8_t r = buf[i];
            while (r >= 0xFE) orc_prng_fill(seed, (*counter)++, 1, &r);
            size_t idx = j + i;
            packed[idx / 4] |= (uint8_t)((r % 3) << (6 - 2 * (idx % 4)));
        }
    }
}

/* sample.c:263-284, :311-321: 16 samples per 96-byte PRNG call; sample = popcnt over bytes
 * 0,1 and the low 5 bits of byte 2, minus the same over bytes 3,4,5 (k = 21). */
void orc_sample_cbd(size_t n, const uint8_t *seed, uint64_t *counter, int8_t *out)
{
    for (size_t j = 0; j < n; j += 16)
    {
        uint8_t buf[96];
        orc_prng_fill(seed, (*counter)++, 96, buf);
        for (size_t i = 0; i < 16; i++)
        {
            const uint8_t *x = buf + 6 * i;
            int pos = __builtin_popcount(x[0]) + __builtin_popcount(x[1]) +
                      __builtin_popcount(x[2] & 0x1F);
            int neg = __builtin_popcount(x[3]) + __builtin_popcount(x[4]) +
                      __builtin_popcount(x[5] & 0x1F);
            out[j + i] = (int8_t)(pos - neg);
        }
    }
}

/* sample.c:39-57: one 4n-byte PRNG call, then in index order every word >= max_multiple is
 * redrawn with 4-byte PRNG calls until accepted; result reduced mod q. */
void orc_sample_uniform(size_t n, uint32_t q, const uint8_t *seed, uint64_t *counter,
                        uint32_t *out)
{
    uint32_t max_multiple = 0xFFFFFFFFu - orc_barrett32(0xFFFFFFFFu, q) - 1;
    orc_prng_fill(seed, (*counter)++, 4 * n, (uint8_t *)out);
    for (size_t i = 0; i < n; i++)
    {
        uint32_t r = out[i];
        while (r >= max_multiple) orc_prng_fill(seed, (*counter)++, 4, (uint8_t *)&r);
        out[i] = orc_barrett32(r, q);
    }
}

/* sample.c:89-129: stored t in {0,1,2} means coefficient t-1, i.e. {q-1, 0, 1} */
void orc_expand_ternary(size_t n, uint32_t q, const uint8_t *packed, uint32_t *out)
{
    for (size_t i = 0; i < n; i++)
    {
        uint32_t t = (packed[i / 4] >> (6 - 2 * (i % 4))) & 3;
        out[i]     = t == 0 ? q - 1 : t - 1;
    }
}

void orc_reduce_small(size_t n, uint32_t q, const int8_t *e, uint32_t *out) /* ckks_common.c:259-265 */
{
    for (size_t i = 0; i < n; i++) out[i] = e[i] < 0 ? q + (uint32_t)(int32_t)e[i] : (uint32_t)e[i];
}

/* ckks_common.c:224-245: |x| mod q via the 64->32 Barrett, then q - r when x < 0 (which yields q,
 * not 0, for negative multiples of q — kept as the reference does it). */
void orc_reduce_pte(size_t n, uint32_t q, const int64_t *pte, uint32_t *out)
{
    for (size_t i = 0; i < n; i++)
    {
        int64_t x   = pte[i];
        uint64_t ax = x < 0 ? (uint64_t)0 - (uint64_t)x : (uint64_t)x;
        uint32_t r  = orc_barrett64((uint32_t)ax, (uint32_t)(ax >> 32), q);
        out[i]      = x < 0 ? q - r : r;
    }
}

/* ------------------------------------------------------------------------------------------
 * NTT.
 * ---------------------------------------------------------------------------------------- */
/* ntt.c:40-52: roots[bitrev(i)] = psi^i */
void orc_ntt_roots(size_t n, uint32_t q, uint32_t psi, uint32_t *roots)
{
    size_t logn = ilog2(n);
    uint32_t p  = psi;
    roots[0]    = 1;
    for (size_t i = 1; i < n; i++)
    {
        roots[orc_bitrev(i, logn)] = p;
        p                          = orc_mul_mod(p, psi, q);
    }
}

/* ntt.c:124-165: Cooley-Tukey, natural in, bit-reversed out, fully reduced every butterfly */
void orc_ntt(size_t n, uint32_t q, const uint32_t *roots, uint32_t *vec)
{
    size_t h = 1, tt = n / 2;
    for (; tt >= 1; h *= 2, tt /= 2)
    {
        for (size_t j = 0, k0 = 0; j < h; j++, k0 += 2 * tt)
        {
            uint32_t s = roots[h + j];
            for (size_t k = k0; k < k0 + tt; k++)
            {
                uint32_t u  = vec[k];
                uint32_t v  = orc_mul_mod(vec[k + tt], s, q);
                vec[k]      = orc_add_mod(u, v, q);
                vec[k + tt] = orc_sub_mod(u, v, q);
            }
        }
    }
}

void orc_ntt_psi(size_t n, uint32_t q, uint32_t psi, uint32_t *vec)
{
    uint32_t *roots = malloc(n * sizeof *roots);
    orc_ntt_roots(n, q, psi, roots);
    orc_ntt(n, q, roots, vec);
    free(roots);
}

void orc_ntt_default(size_t n, uint32_t q, uint32_t *vec)
{
    orc_ntt_psi(n, q, orc_ntt_root(n, q), vec);
}

/* Inverse of orc_ntt (test helper): undo the stages last to first, then scale by n^-1. */
void orc_intt(size_t n, uint32_t q, uint32_t psi, uint32_t *vec)
{
    uint32_t *roots = malloc(n * sizeof *roots);
    orc_ntt_roots(n, q, psi, roots);
    for (size_t tt = 1, h = n / 2; h >= 1; tt *= 2, h /= 2)
    {
        for (size_t j = 0, k0 = 0; j < h; j++, k0 += 2 * tt)
        {
            uint32_t sinv = orc_pow_mod(roots[h + j], (uint64_t)q - 2, q);
            for (size_t k = k0; k < k0 + tt; k++)
            {
                uint32_t x  = vec[k] % q, y = vec[k + tt] % q;
                vec[k]      = orc_add_mod(x, y, q);
                vec[k + tt] = orc_mul_mod(orc_sub_mod(x, y, q), sinv, q);
            }
        }
    }
    uint32_t ninv = orc_pow_mod((uint32_t)(n % q), (uint64_t)q - 2, q);
    for (size_t i = 0; i < n; i++) vec[i] = orc_mul_mod(vec[i], ninv, q);
    free(roots);
}

void orc_negacyclic_mul(size_t n, uint32_t q, const uint32_t *a, const uint32_t *b, uint32_t *c)
{
    for (size_t k = 0; k < n; k++) c[k] = 0;
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++)
        {
            uint32_t p = orc_mul_mod(a[i] % q, b[j] % q, q);
            size_t k   = i + j;
            if (k < n)
                c[k] = orc_add_mod(c[k], p, q);
            else
                c[k - n] = orc_sub_mod(c[k - n], p, q);
        }
}

/* ------------------------------------------------------------------------------------------
 * Full path.
 * ---------------------------------------------------------------------------------------- */
static void pointwise_mul(size_t n, uint32_t q, uint32_t *a, const uint32_t *b) /* polymodarith.h:83-87 */
{
    for (size_t i = 0; i < n; i++) a[i] = orc_mul_mod(a[i], b[i], q);
}
static void pointwise_add(size_t n, uint32_t q, uint32_t *a, const uint32_t *b) /* polymodarith.h:39-42 */
{
    for (size_t i = 0; i < n; i++) a[i] = orc_add_mod(a[i], b[i], q);
}

/* seal_embedded.c:98-215 (asymmetric branch), ckks_asym.c:173-203 and :205-286 */
/* The chain is explicit (primes, the 2n-th roots psis and the scale): the default entry point below passes the
 * reference's tables; a caller chain (set_custom_parms_ckks, parameters.c:232-249) passes its own. */
int orc_encrypt_asym_ex(size_t n, size_t np, const uint32_t *primes, const uint32_t *psis, double scale,
                        const float *values, size_t vlen, const uint8_t *seed, const uint32_t *pk0,
                        const uint32_t *pk1, uint32_t *out)
{
    int64_t *pt  = malloc(n * sizeof *pt);
    uint8_t *u   = malloc(n / 4);
    int8_t *e    = malloc(n);
    int8_t *e1   = malloc(n);
    uint32_t *t  = malloc(n * sizeof *t);
    uint32_t *nu = malloc(n * sizeof *nu);
    int ok       = orc_encode(n, scale, values, vlen, pt);
    if (ok)
    {
        uint64_t ctr = 0;
        orc_sample_ternary_small(n, seed, &ctr, u);
        orc_sample_cbd(n, seed, &ctr, e);
        for (size_t i = 0; i < n; i++) pt[i] = (int64_t)((uint64_t)pt[i] + (uint64_t)(int64_t)e[i]);
        orc_sample_cbd(n, seed, &ctr, e1);
        for (size_t p = 0; p < np; p++)
        {
            uint32_t q   = primes[p];
            uint32_t psi = psis[p];
            uint32_t *c0 = out + (2 * p) * n;
            uint32_t *c1 = out + (2 * p + 1) * n;
            orc_expand_ternary(n, q, u, nu);
            orc_ntt_psi(n, q, psi, nu);
            memcpy(c1, pk1 + p * n, n * sizeof *c1);
            memcpy(c0, pk0 + p * n, n * sizeof *c0);
            pointwise_mul(n, q, c1, nu);
            pointwise_mul(n, q, c0, nu);
            orc_reduce_small(n, q, e1, t);
            orc_ntt_psi(n, q, psi, t);
            pointwise_add(n, q, c1, t);
            orc_reduce_pte(n, q, pt, t);
            orc_ntt_psi(n, q, psi, t);
            pointwise_add(n, q, c0, t);
        }
    }
    free(pt);
    free(u);
    free(e);
    free(e1);
    free(t);
    free(nu);
    return ok;
}

static int default_chain(size_t n, size_t np, uint32_t *primes, uint32_t *psis)
{
    if (!orc_default_primes(n, np, primes)) return 0;
    for (size_t p = 0; p < np; p++) psis[p] = orc_ntt_root(n, primes[p]);
    return 1;
}

int orc_encrypt_asym(size_t n, size_t np, const float *values, size_t vlen, const uint8_t *seed,
                     const uint32_t *pk0, const uint32_t *pk1, uint32_t *out)
{
    uint32_t primes[16], psis[16];
    if (!default_chain(n, np, primes, psis)) return 0;
    return orc_encrypt_asym_ex(n, np, primes, psis, orc_default_scale(n), values, vlen, seed, pk0, pk1, out);
}

/* one prime of ckks_sym.c:199-301; a is left in c1, ntt(m+e) in ntt_pte */
static void sym_core(size_t n, uint32_t q, uint32_t psi, const uint8_t *share_seed, uint64_t *ctr_a,
                     const uint8_t *sk_packed, const int64_t *pt, const int8_t *ep, uint32_t *c0,
                     uint32_t *c1, uint32_t *ntt_pte)
{
    orc_sample_uniform(n, q, share_seed, ctr_a, c1);
    orc_expand_ternary(n, q, sk_packed, c0);
    orc_ntt_psi(n, q, psi, c0);
    pointwise_mul(n, q, c0, c1);
    for (size_t i = 0; i < n; i++) c0[i] = orc_neg_mod(c0[i], q);
    if (ep)
        orc_reduce_small(n, q, ep, ntt_pte);
    else
        orc_reduce_pte(n, q, pt, ntt_pte);
    orc_ntt_psi(n, q, psi, ntt_pte);
    pointwise_add(n, q, c0, ntt_pte);
}

/* seal_embedded.c:98-215 (symmetric branch), ckks_sym.c:181-197 and :199-301 */
int orc_encrypt_sym_ex(size_t n, size_t np, const uint32_t *primes, const uint32_t *psis, double scale,
                       const float *values, size_t vlen, const uint8_t *share_seed, const uint8_t *seed,
                       const uint8_t *sk_packed, int ref_quirk, uint32_t *out)
{
    int64_t *pt = malloc(n * sizeof *pt);
    int8_t *e   = malloc(n);
    uint32_t *t = malloc(n * sizeof *t);
    int ok      = orc_encode(n, scale, values, vlen, pt);
    if (ok)
    {
        uint64_t ctr_e = 0, ctr_a = 0;
        orc_sample_cbd(n, seed, &ctr_e, e);
        for (size_t i = 0; i < n; i++) pt[i] = (int64_t)((uint64_t)pt[i] + (uint64_t)(int64_t)e[i]);
        for (size_t p = 0; p < np; p++)
        {
            uint32_t *c0 = out + (2 * p) * n;
            uint32_t *c1 = out + (2 * p + 1) * n;
            sym_core(n, primes[p], psis[p], share_seed, &ctr_a, sk_packed, pt, NULL, c0, c1, t);
            if (ref_quirk) memcpy(c1, t, n * sizeof *c1);
        }
    }
    free(pt);
    free(e);
    free(t);
    return ok;
}

int orc_encrypt_sym(size_t n, size_t np, const float *values, size_t vlen,
                    const uint8_t *share_seed, const uint8_t *seed, const uint8_t *sk_packed,
                    int ref_quirk, uint32_t *out)
{
    uint32_t primes[16], psis[16];
    if (!default_chain(n, np, primes, psis)) return 0;
    return orc_encrypt_sym_ex(n, np, primes, psis, orc_default_scale(n), values, vlen, share_seed, seed, sk_packed,
                              ref_quirk, out);
}

void orc_gen_pk_prime_ex(size_t n, uint32_t q, uint32_t psi, const uint8_t *seed, const uint8_t *sk_packed,
                         const int8_t *ep, uint32_t *pk0, uint32_t *pk1)
{
    uint64_t ctr = 0;
    uint32_t *t  = malloc(n * sizeof *t);
    sym_core(n, q, psi, seed, &ctr, sk_packed, NULL, ep, pk0, pk1, t);
    free(t);
}

void orc_gen_pk_prime(size_t n, uint32_t q, const uint8_t *seed, const uint8_t *sk_packed,
                      const int8_t *ep, uint32_t *pk0, uint32_t *pk1)
{
    orc_gen_pk_prime_ex(n, q, orc_ntt_root(n, q), seed, sk_packed, ep, pk0, pk1);
}

void orc_decrypt_ntt_ex(size_t n, uint32_t q, uint32_t psi, const uint32_t *c0, const uint32_t *c1,
                        const uint8_t *sk_packed, uint32_t *pt_ntt)
{
    uint32_t *s = malloc(n * sizeof *s);
    orc_expand_ternary(n, q, sk_packed, s);
    orc_ntt_psi(n, q, psi, s);
    for (size_t i = 0; i < n; i++)
        pt_ntt[i] = orc_add_mod(orc_mul_mod(c1[i], s[i], q), c0[i], q);
    free(s);
}

void orc_decrypt_ntt(size_t n, uint32_t q, const uint32_t *c0, const uint32_t *c1,
                     const uint8_t *sk_packed, uint32_t *pt_ntt)
{
    uint32_t *s = malloc(n * sizeof *s);
    orc_expand_ternary(n, q, sk_packed, s);
    orc_ntt_default(n, q, s);
    for (size_t i = 0; i < n; i++)
        pt_ntt[i] = orc_add_mod(orc_mul_mod(c1[i], s[i], q), c0[i], q);
    free(s);
}

int orc_encrypt_asym_batch(size_t n, size_t np, size_t batch, const float *values, size_t vlen,
                           const uint8_t *seeds, const uint32_t *pk0, const uint32_t *pk1,
                           uint32_t *out)
{
    int ok = 1;
    for (size_t b = 0; b < batch; b++)
        ok &= orc_encrypt_asym(n, np, values + b * vlen, vlen, seeds + b * ORC_SEED_BYTES, pk0, pk1,
                               out + b * 2 * np * n);
    return ok;
}
