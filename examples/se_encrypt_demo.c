/*
 * se_encrypt_demo.c — a C application written against SEAL-Embedded's public API
 * (device/lib/seal_embedded.h:91-130), linked against libseal_embedded_b200.so instead of the
 * reference's static library.  Nothing here is specific to the GPU build except the header name.
 *
 *   gcc -std=c11 -O2 examples/se_encrypt_demo.c -Iinclude -Lseal-embedded_b200 -lseal_embedded_b200 \
 *       -Wl,-rpath,$PWD/seal-embedded_b200 -o se_encrypt_demo
 *   ./se_encrypt_demo <asym|sym> <degree> <nprimes> <out.bin> [seed_byte] [share_seed_byte]
 *
 * Reads key material from ./adapter_output_data/ like the reference (fileops.c:140-204), encrypts the
 * message v[i] = (i % 17) - 8 + i / 1024 (i < degree/2) with seeds made of one repeated byte (default 0x5A
 * / 0xA5; the seeded entry point makes the output reproducible) and writes the bytes handed to the send
 * callback — c0 then c1 per prime, degree * 4 bytes each (seal_embedded.c:180-204) — to <out.bin>.
 * Then does the same through the batch extension for 3 messages and appends those ciphertexts.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "seal_embedded_b200.h"

static FILE *g_out;
static size_t g_calls;

static size_t send_to_file(void *buf, size_t nbytes)
{
    g_calls++;
    return fwrite(buf, 1, nbytes, g_out);
}

int main(int argc, char **argv)
{
    if (argc < 5)
    {
        fprintf(stderr, "usage: %s <asym|sym> <degree> <nprimes> <out.bin> [seed_byte] [share_seed_byte]\n", argv[0]);
        return 2;
    }
    const EncryptType type = strcmp(argv[1], "asym") == 0 ? SE_ASYM_ENCR : SE_SYM_ENCR;
    const size_t degree = (size_t)atol(argv[2]), nprimes = (size_t)atol(argv[3]);
    uint8_t seed[SE_PRNG_SEED_BYTE_COUNT], share[SE_PRNG_SEED_BYTE_COUNT];
    memset(seed, argc > 5 ? atoi(argv[5]) : 0x5A, sizeof seed);
    memset(share, argc > 6 ? atoi(argv[6]) : 0xA5, sizeof share);

    SE_PARMS *se_parms = se_setup(degree, nprimes, 0.0, type);
    const size_t n = se_parms->parms->coeff_count; /* callers read the parameter block (api_tests.c:66-100) */
    if (n != degree || se_parms->parms->nprimes != nprimes) return 3;

    const size_t vlen = n / 2;
    flpt *v = malloc(3 * vlen * sizeof(flpt));
    for (size_t b = 0; b < 3; b++)
        for (size_t i = 0; i < vlen; i++) v[b * vlen + i] = (flpt)((int)((i + b) % 17) - 8) + (flpt)i / 1024.0f;

    g_out = fopen(argv[4], "wb");
    if (!g_out) return 4;
    bool ok = se_encrypt_seeded(share, seed, &send_to_file, v, vlen * sizeof(flpt), false, se_parms);
    if (!ok || g_calls != 2 * nprimes) return 5;

    /* batch extension: 3 messages, seeds = the same bytes + item index in byte 0 */
    uint8_t seeds[3 * SE_PRNG_SEED_BYTE_COUNT], shares[3 * SE_PRNG_SEED_BYTE_COUNT];
    for (size_t b = 0; b < 3; b++)
    {
        memcpy(seeds + b * SE_PRNG_SEED_BYTE_COUNT, seed, SE_PRNG_SEED_BYTE_COUNT);
        memcpy(shares + b * SE_PRNG_SEED_BYTE_COUNT, share, SE_PRNG_SEED_BYTE_COUNT);
        seeds[b * SE_PRNG_SEED_BYTE_COUNT]  = (uint8_t)b;
        shares[b * SE_PRNG_SEED_BYTE_COUNT] = (uint8_t)(b + 100);
    }
    ZZ *cts = malloc(3 * nprimes * 2 * n * sizeof(ZZ));
    ok = se_encrypt_batch_seeded(type == SE_SYM_ENCR ? shares : NULL, seeds, v, vlen, 3, cts, se_parms);
    if (!ok) return 6;
    fwrite(cts, sizeof(ZZ), 3 * nprimes * 2 * n, g_out);
    fclose(g_out);
    printf("wrote %zu + %zu bytes (%zu send callbacks)\n", 2 * nprimes * n * sizeof(ZZ), 3 * nprimes * 2 * n * sizeof(ZZ),
           g_calls);
    free(cts);
    free(v);
    se_cleanup(se_parms);
    return se_parms->parms == 0 ? 0 : 7; /* seal_embedded.c:234 */
}
