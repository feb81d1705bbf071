/*
 * se_reference_api_demo.c — an application that uses ONLY what SEAL-Embedded's own header declares
 * (device/lib/seal_embedded.h:52-130): it includes "seal_embedded.h" and nothing of this repository.
 *
 * It is compiled twice by the tests:
 *   (1) against the reference's own header:  gcc -I/root/reference/device/lib ...      (oracle/Makefile: refdemo)
 *   (2) against include/seal_embedded.h (the compat shim over seal_embedded_b200.h)
 * and linked against libseal_embedded_b200.so both times.  Build (1) is the literal drop-in claim: the
 * reference's struct layouts and prototypes, this library's code.
 *
 *   ./se_reference_api_demo <asym|sym> <degree> <nprimes> <out.bin> [q0 q1 ...]
 *
 * With moduli on the command line the context is made by se_setup_custom (caller chain, caller scale 2^24,
 * ratios = floor(2^64/q) high word first as seal_embedded.h:86-87 documents); otherwise by se_setup.
 * Key material comes from ./adapter_output_data/ (fileops.c:140-204).  Output file: the bytes handed to the send
 * callback (c0, c1 per prime: seal_embedded.c:180-204), then n uint16 index-map entries read through
 * se_parms->se_ptrs->index_map_ptr, then per prime {value, const_ratio[0], const_ratio[1]} read through
 * se_parms->parms->moduli, then the scale as a double.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "seal_embedded.h"

static FILE *g_out;
static size_t g_calls, g_bytes;

static size_t send_to_file(void *buf, size_t nbytes)
{
    g_calls++;
    g_bytes += nbytes;
    return fwrite(buf, 1, nbytes, g_out);
}

int main(int argc, char **argv)
{
    if (argc < 5)
    {
        fprintf(stderr, "usage: %s <asym|sym> <degree> <nprimes> <out.bin> [q0 q1 ...]\n", argv[0]);
        return 2;
    }
    const EncryptType type = strcmp(argv[1], "asym") == 0 ? SE_ASYM_ENCR : SE_SYM_ENCR;
    const size_t degree = (size_t)atol(argv[2]), nprimes = (size_t)atol(argv[3]);
    uint8_t seed[64], share[64];
    memset(seed, 0x3C, sizeof seed);
    memset(share, 0xC3, sizeof share);

    SE_PARMS *se_parms;
    if (argc >= 5 + (int)nprimes)
    {
        ZZ *q = malloc(nprimes * sizeof(ZZ)), *ratios = malloc(2 * nprimes * sizeof(ZZ));
        for (size_t i = 0; i < nprimes; i++)
        {
            q[i]             = (ZZ)strtoul(argv[5 + i], NULL, 10);
            unsigned __int128 r = ((unsigned __int128)1 << 64) / q[i];
            ratios[2 * i]     = (ZZ)(r >> 32); /* high word, followed by low word */
            ratios[2 * i + 1] = (ZZ)r;
        }
        se_parms = se_setup_custom(degree, nprimes, q, ratios, 16777216.0, type);
        free(q);
        free(ratios);
    }
    else
        se_parms = se_setup(degree, nprimes, 0.0, type);

    const Parms *parms = se_parms->parms;
    const size_t n     = parms->coeff_count;
    if (n != degree || parms->nprimes != nprimes || ((size_t)1 << parms->logn) != n) return 3;
    if (parms->is_asymmetric != (type == SE_ASYM_ENCR)) return 3;

    const size_t vlen = n / 2;
    flpt *v = malloc(vlen * sizeof(flpt));
    for (size_t i = 0; i < vlen; i++) v[i] = (flpt)((int)(i % 23) - 11) + (flpt)i / 2048.0f;

    g_out = fopen(argv[4], "wb");
    if (!g_out) return 4;
    bool ok = se_encrypt_seeded(share, seed, &send_to_file, v, vlen * sizeof(flpt), false, se_parms);
    if (!ok || g_calls != 2 * nprimes || g_bytes != 2 * nprimes * n * sizeof(ZZ)) return 5;
    /* after the call the parameter block points at the last prime, as in the reference's loop */
    if (parms->curr_modulus_idx != nprimes - 1 || parms->curr_modulus != &parms->moduli[nprimes - 1]) return 6;

    fwrite(se_parms->se_ptrs->index_map_ptr, sizeof(uint16_t), n, g_out);
    for (size_t i = 0; i < nprimes; i++)
    {
        ZZ m[3] = {parms->moduli[i].value, parms->moduli[i].const_ratio[0], parms->moduli[i].const_ratio[1]};
        fwrite(m, sizeof(ZZ), 3, g_out);
    }
    double scale = parms->scale;
    fwrite(&scale, sizeof scale, 1, g_out);
    fclose(g_out);
    free(v);
    se_cleanup(se_parms);
    return se_parms->parms == 0 ? 0 : 7; /* seal_embedded.c:234 */
}
