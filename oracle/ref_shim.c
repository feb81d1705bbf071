/*
 * ref_shim.c — thin C harness compiled TOGETHER WITH the reference's own sources
 * (/root/reference/device/lib/*.c, never copied into this repo) into oracle/_ref/libseref.so by
 * oracle/Makefile.  It only calls the reference's functions and exposes them with plain
 * pointer/size signatures so tests can load them through ctypes.
 *
 * TEST INFRASTRUCTURE ONLY: used to pin the oracle and the CUDA path, and as the
 * "reference" CPU baseline in bench.py.  Not part of the product.
 *
 * The reference keeps static global state (seal_embedded.c:18-22): one context per process.
 */
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "ckks_asym.h"
#include "ckks_common.h"
#include "ckks_sym.h"
#include "fft.h"
#include "modulo.h"
#include "ntt.h"
#include "parameters.h"
#include "sample.h"
#include "seal_embedded.h"
#include "uintmodarith.h"

static SE_PARMS *g_se   = NULL;
static uint8_t *g_sink  = NULL;
static size_t g_sinkpos = 0;
static int g_saved_fd   = -1;

/* the reference prints from se_setup and friends; keep test logs readable */
static void quiet_begin(void)
{
    fflush(stdout);
    g_saved_fd = dup(1);
    int nul    = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
}
static void quiet_end(void)
{
    fflush(stdout);
    if (g_saved_fd >= 0)
    {
        dup2(g_saved_fd, 1);
        close(g_saved_fd);
        g_saved_fd = -1;
    }
}

static size_t capture_send(void *data, size_t nbytes)
{
    memcpy(g_sink + g_sinkpos, data, nbytes);
    g_sinkpos += nbytes;
    return nbytes;
}

/* chdir(workdir) (it must hold adapter_output_data/), then the reference's se_setup */
int ref_setup(size_t n, size_t np, int asym, const char *workdir)
{
    if (workdir && chdir(workdir) != 0) return 0;
    quiet_begin();
    g_se = se_setup(n, np, 0.0 /* overridden by set_parms_ckks */, asym ? SE_ASYM_ENCR : SE_SYM_ENCR);
    quiet_end();
    return g_se != NULL;
}

void ref_cleanup(void)
{
    if (g_se) se_cleanup(g_se);
    g_se = NULL;
}

double ref_scale(void) { return g_se->parms->scale; }
size_t ref_nprimes(void) { return g_se->parms->nprimes; }
uint32_t ref_prime(size_t i) { return g_se->parms->moduli[i].value; }
uint32_t ref_ratio(size_t i, size_t w) { return g_se->parms->moduli[i].const_ratio[w]; }

/* se_encrypt_seeded through the public API; out receives the concatenated send() payloads */
int ref_encrypt_seeded(const uint8_t *share_seed, const uint8_t *seed, const float *v,
                       size_t vlen_bytes, uint32_t *out)
{
    g_sink    = (uint8_t *)out;
    g_sinkpos = 0;
    /* fresh-pool semantics for short inputs: clear the staged values first */
    memset(g_se->se_ptrs->values, 0, (g_se->parms->coeff_count / 2) * sizeof(flpt));
    bool ok = se_encrypt_seeded((uint8_t *)share_seed, (uint8_t *)seed, capture_send, (void *)v,
                                vlen_bytes, false, g_se);
    return ok ? 1 : 0;
}

/* timing loop for the CPU baseline: count encryptions of consecutive items, seconds returned */
double ref_encrypt_loop(size_t count, const uint8_t *share_seeds, const uint8_t *seeds,
                        const float *v, size_t vlen, uint32_t *out_last)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (size_t b = 0; b < count; b++)
    {
        g_sink    = (uint8_t *)out_last;
        g_sinkpos = 0;
        se_encrypt_seeded(share_seeds ? (uint8_t *)share_seeds + 64 * b : NULL,
                          (uint8_t *)seeds + 64 * b, capture_send, (void *)(v + b * vlen),
                          vlen * sizeof(float), false, g_se);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* The per-item digest of seb_digest_device (seal-embedded_b200/csrc/seb_verify.cu: k_digest):
 * sum_i mix64((i << 32) | word_i) mod 2^64, mix64 = the splitmix64 finaliser. */
static uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t ref_digest_words(const uint32_t *w, size_t count)
{
    uint64_t acc = 0;
    for (size_t i = 0; i < count; i++) acc += mix64(((uint64_t)i << 32) | w[i]);
    return acc;
}

/* `count` consecutive items through se_encrypt_seeded (the reference's own API, byte stream captured from the
 * send callback); digests[b] = digest of item b's [nprimes][2][n] words.  Returns how many calls returned false. */
size_t ref_encrypt_digests(size_t count, const uint8_t *share_seeds, const uint8_t *seeds, const float *v,
                           size_t vlen, size_t words_per_item, uint32_t *scratch, uint64_t *digests)
{
    size_t bad = 0;
    for (size_t b = 0; b < count; b++)
    {
        g_sink    = (uint8_t *)scratch;
        g_sinkpos = 0;
        bool ok = se_encrypt_seeded(share_seeds ? (uint8_t *)share_seeds + 64 * b : NULL, (uint8_t *)seeds + 64 * b,
                                    capture_send, (void *)(v + b * vlen), vlen * sizeof(float), false, g_se);
        bad += !ok || g_sinkpos != words_per_item * sizeof(uint32_t);
        digests[b] = ref_digest_words(scratch, words_per_item);
    }
    return bad;
}

/* ---- stage-level entry points (own scratch, independent of the API's static state) ---- */

static void local_parms(size_t n, size_t np, int asym, Parms *parms)
{
    memset(parms, 0, sizeof *parms);
    parms->is_asymmetric = asym != 0;
    parms->pk_from_file  = 0;
    parms->sample_s      = 0;
    parms->small_u       = 1;
    parms->small_s       = 1;
    set_parms_ckks(n, np, parms);
}

void ref_index_map(size_t n, uint16_t *map)
{
    Parms parms;
    local_parms(n, 1, 0, &parms);
    ckks_calc_index_map(&parms, map);
    delete_parameters(&parms);
}

/* ckks_encode_base; out = n int64.  scratch is 16n bytes allocated here. */
int ref_encode(size_t n, const float *v, size_t vlen, int64_t *out)
{
    Parms parms;
    local_parms(n, 1, 0, &parms);
    uint16_t *map        = calloc(n, sizeof *map);
    double complex *conj = calloc(n, sizeof *conj);
    float *vals          = calloc(n / 2, sizeof *vals);
    memcpy(vals, v, (vlen > n / 2 ? n / 2 : vlen) * sizeof *vals);
    ckks_calc_index_map(&parms, map);
    quiet_begin();
    bool ok = ckks_encode_base(&parms, vals, n / 2, map, NULL, conj);
    quiet_end();
    if (ok) memcpy(out, conj, n * sizeof *out);
    free(map);
    free(conj);
    free(vals);
    delete_parameters(&parms);
    return ok ? 1 : 0;
}

void ref_prng_fill(const uint8_t *seed, uint64_t counter, size_t nbytes, uint8_t *out)
{
    SE_PRNG prng;
    prng_randomize_reset(&prng, (uint8_t *)seed);
    prng.counter = counter;
    prng_fill_buffer(nbytes, &prng, out);
}

/* ckks_asym_init: u (packed, n/4 B), pt += e0, e1; returns the PRNG counter afterwards */
uint64_t ref_asym_init(size_t n, const uint8_t *seed, int64_t *pt_inout, uint8_t *u_packed,
                       int8_t *e1)
{
    Parms parms;
    SE_PRNG prng;
    local_parms(n, 1, 1, &parms);
    memset(u_packed, 0, n / 4);
    ckks_asym_init(&parms, (uint8_t *)seed, &prng, pt_inout, (ZZ *)u_packed, e1);
    delete_parameters(&parms);
    return prng.counter;
}

uint64_t ref_sample_ternary_small(size_t n, const uint8_t *seed, uint64_t counter, uint8_t *packed)
{
    SE_PRNG prng;
    prng_randomize_reset(&prng, (uint8_t *)seed);
    prng.counter = counter;
    memset(packed, 0, n / 4);
    sample_small_poly_ternary_prng_96(n, &prng, (ZZ *)packed);
    return prng.counter;
}

uint64_t ref_sample_cbd(size_t n, const uint8_t *seed, uint64_t counter, int8_t *out)
{
    SE_PRNG prng;
    prng_randomize_reset(&prng, (uint8_t *)seed);
    prng.counter = counter;
    sample_poly_cbd_generic_prng_16(n, &prng, out);
    return prng.counter;
}

uint64_t ref_sample_uniform(size_t n, size_t np, size_t prime_idx, const uint8_t *seed,
                            uint64_t counter, uint32_t *out)
{
    Parms parms;
    SE_PRNG prng;
    local_parms(n, np, 0, &parms);
    for (size_t i = 0; i < prime_idx; i++) next_modulus(&parms);
    prng_randomize_reset(&prng, (uint8_t *)seed);
    prng.counter = counter;
    sample_poly_uniform(&parms, &prng, out);
    delete_parameters(&parms);
    return prng.counter;
}

/* ntt_roots_initialize + ntt_inpl under prime prime_idx of the default (n,np) chain */
void ref_ntt(size_t n, size_t np, size_t prime_idx, uint32_t *vec)
{
    Parms parms;
    local_parms(n, np, 0, &parms);
    for (size_t i = 0; i < prime_idx; i++) next_modulus(&parms);
    ZZ *roots = calloc(2 * n, sizeof *roots);
    ntt_roots_initialize(&parms, roots);
    ntt_inpl(&parms, roots, vec);
    free(roots);
    delete_parameters(&parms);
}

/* timing loop for the NTT-only CPU baseline (SURVEY 8d "(ii) NTT-only loop"): the root table is built once,
 * then ntt_inpl (ntt.c:168-189) runs `reps` times on the same buffer; seconds returned */
double ref_ntt_loop(size_t n, size_t np, size_t prime_idx, uint32_t *vec, size_t reps)
{
    Parms parms;
    local_parms(n, np, 0, &parms);
    for (size_t i = 0; i < prime_idx; i++) next_modulus(&parms);
    ZZ *roots = calloc(2 * n, sizeof *roots);
    ntt_roots_initialize(&parms, roots);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (size_t r = 0; r < reps; r++) ntt_inpl(&parms, roots, vec);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(roots);
    delete_parameters(&parms);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void ref_reduce_pte(size_t n, size_t np, size_t prime_idx, const int64_t *pte, uint32_t *out)
{
    Parms parms;
    local_parms(n, np, 0, &parms);
    for (size_t i = 0; i < prime_idx; i++) next_modulus(&parms);
    reduce_set_pte(&parms, pte, out);
    delete_parameters(&parms);
}

/* symmetric encryption at the ckks_encode_encrypt_sym level with c1_save, so the true `a` is
 * returned as c1 (the API's c1 buffer is clobbered).  out = [np][2][n]. */
int ref_encrypt_sym_c1a(size_t n, size_t np, const uint8_t *share_seed, const uint8_t *seed,
                        const uint8_t *sk_packed, const float *v, size_t vlen, uint32_t *out)
{
    Parms parms;
    SE_PRNG prng, shareable;
    local_parms(n, np, 0, &parms);
    uint16_t *map        = calloc(n, sizeof *map);
    double complex *conj = calloc(n, sizeof *conj);
    float *vals          = calloc(n / 2, sizeof *vals);
    ZZ *s_small          = calloc(n, sizeof *s_small);
    ZZ *roots            = calloc(2 * n, sizeof *roots);
    ZZ *ntt_pte          = calloc(n, sizeof *ntt_pte);
    ZZ *c1               = calloc(n, sizeof *c1);
    memcpy(vals, v, (vlen > n / 2 ? n / 2 : vlen) * sizeof *vals);
    memcpy(s_small, sk_packed, n / 4);
    ckks_calc_index_map(&parms, map);
    quiet_begin();
    bool ok = ckks_encode_base(&parms, vals, n / 2, map, NULL, conj);
    quiet_end();
    if (ok)
    {
        ckks_sym_init(&parms, (uint8_t *)share_seed, (uint8_t *)seed, &shareable, &prng,
                      (int64_t *)conj);
        for (size_t p = 0; p < np; p++)
        {
            ckks_encode_encrypt_sym(&parms, (int64_t *)conj, NULL, &shareable, s_small, ntt_pte,
                                    roots, out + (2 * p) * n, c1, NULL, out + (2 * p + 1) * n);
            if (p + 1 < np) ckks_next_prime_sym(&parms, s_small);
        }
    }
    free(map);
    free(conj);
    free(vals);
    free(s_small);
    free(roots);
    free(ntt_pte);
    free(c1);
    delete_parameters(&parms);
    return ok ? 1 : 0;
}

/* gen_pk (ckks_asym.c:159-171) for every prime: seed_base with byte 0 replaced by the prime
 * index; ep shared by all primes.  pk0, pk1: [np][n]. */
void ref_gen_pk(size_t n, size_t np, const uint8_t *sk_packed, const int8_t *ep,
                const uint8_t *seed_base, uint32_t *pk0, uint32_t *pk1)
{
    Parms parms;
    SE_PRNG shareable;
    local_parms(n, np, 1, &parms);
    ZZ *s_small = calloc(n, sizeof *s_small);
    ZZ *roots   = calloc(2 * n, sizeof *roots);
    ZZ *ntt_ep  = calloc(n, sizeof *ntt_ep);
    memcpy(s_small, sk_packed, n / 4);
    for (size_t p = 0; p < np; p++)
    {
        uint8_t seed[64];
        memcpy(seed, seed_base, 64);
        seed[0] = (uint8_t)p;
        gen_pk(&parms, s_small, roots, seed, &shareable, NULL, (int8_t *)ep, ntt_ep, pk0 + p * n,
               pk1 + p * n);
        if (p + 1 < np) next_modulus(&parms);
    }
    free(s_small);
    free(roots);
    free(ntt_ep);
    delete_parameters(&parms);
}

/* ---- scalar KAT wrappers over the header-only arithmetic ---- */
static void mod_for(uint32_t q, Modulus *m)
{
    quiet_begin();
    set_modulus(q, m);
    quiet_end();
}
uint32_t ref_barrett32(uint32_t x, uint32_t q)
{
    Modulus m;
    mod_for(q, &m);
    return barrett_reduce_32input_32modulus(x, &m);
}
uint32_t ref_barrett64(uint32_t lo, uint32_t hi, uint32_t q)
{
    Modulus m;
    uint32_t in[2] = {lo, hi};
    mod_for(q, &m);
    return barrett_reduce_64input_32modulus(in, &m);
}
uint32_t ref_mul_mod(uint32_t a, uint32_t b, uint32_t q)
{
    Modulus m;
    mod_for(q, &m);
    return mul_mod(a, b, &m);
}
uint32_t ref_add_mod(uint32_t a, uint32_t b, uint32_t q)
{
    Modulus m;
    mod_for(q, &m);
    return add_mod(a, b, &m);
}
uint32_t ref_sub_mod(uint32_t a, uint32_t b, uint32_t q)
{
    Modulus m;
    mod_for(q, &m);
    return sub_mod(a, b, &m);
}
