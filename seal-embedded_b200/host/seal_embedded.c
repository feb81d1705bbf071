/*
 * seal_embedded.c — the reference's public API (device/lib/seal_embedded.{h,c}) implemented in C on
 * top of the B200 kernels' C ABI (seb_* in include/seal_embedded_b200.h).
 *
 * Same symbols, argument meaning, key-file conventions, callback protocol and error behaviour as
 * the reference, so an application written against SEAL-Embedded's device library links against
 * this one unchanged:
 *   - se_setup_custom / se_setup / se_setup_default      seal_embedded.c:24-96
 *   - se_encrypt_seeded / se_encrypt                     seal_embedded.c:98-221
 *   - se_cleanup                                         seal_embedded.c:223-235
 * A single static context per process, like the reference (seal_embedded.c:18-22).
 */
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/random.h>
#include <unistd.h>

#include "../../include/seal_embedded_b200.h"

#ifndef SE_DATA_PATH
#define SE_DATA_PATH "adapter_output_data" /* device/CMakeLists.txt:285 */
#endif

static Parms g_parms;
static SE_PTRS g_ptrs;
static SE_PARMS g_se_parms;
static seb_ctx *g_ctx   = NULL;
static uint32_t *g_ct   = NULL; /* [nprimes][2][n] host copy of the last ciphertext */
static int g_ref_quirk  = 0;
static int g_print_full = 0;
static int g_sym_seed_ct = 0; /* se_encrypt sends (seed, c0) per prime in symmetric mode */
static int g_pk_loaded  = 0;

/* fileops.c:60-138 read_from_image + check_ret: a missing or short key file is fatal */
static void read_key_file(const char *fpath, size_t bytes_expected, void *vec)
{
    int fd = open(fpath, O_RDONLY);
    if (fd < 0)
    {
        printf("Error: problem with opening or closing file\n");
        printf("errno value: %d\n", errno);
        printf("errno message: %s\n", strerror(errno));
        printf("file path: %s\n", fpath);
        exit(1);
    }
    size_t got = 0;
    while (got < bytes_expected)
    {
        ssize_t r = read(fd, (char *)vec + got, bytes_expected - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    if (got != bytes_expected)
    {
        printf("Error: problem with reading from file\n");
        printf("bytes read     : %zu bytes\n", got);
        printf("bytes expected : %zu bytes\n", bytes_expected);
        printf("file path: %s\n", fpath);
        exit(1);
    }
}

/* free() of a buffer that held secret material */
static void wipe_free(void *p, size_t bytes)
{
    if (!p) return;
    explicit_bzero(p, bytes);
    free(p);
}

static void die_on(int rc, const char *what)
{
    if (rc == 0) return;
    printf("Error! %s failed: %s\n", what, seb_last_error());
    exit(1);
}

/* load_pki (fileops.c:172-204) for every prime, once, instead of on every encryption */
static void load_public_key(void)
{
    size_t n = g_parms.coeff_count, np = g_parms.nprimes;
    ZZ *pk0 = malloc(np * n * sizeof(ZZ)), *pk1 = malloc(np * n * sizeof(ZZ));
    char fpath[512];
    if (!pk0 || !pk1)
    {
        printf("Error! Allocation failed. Exiting...\n");
        exit(1);
    }
    for (size_t p = 0; p < np; p++)
    {
        snprintf(fpath, sizeof fpath, "%s/pk0_ntt_%zu_%u.dat", SE_DATA_PATH, n, g_parms.moduli[p].value);
        read_key_file(fpath, n * sizeof(ZZ), pk0 + p * n);
        snprintf(fpath, sizeof fpath, "%s/pk1_ntt_%zu_%u.dat", SE_DATA_PATH, n, g_parms.moduli[p].value);
        read_key_file(fpath, n * sizeof(ZZ), pk1 + p * n);
    }
    die_on(seb_set_public_key(g_ctx, pk0, pk1), "seb_set_public_key");
    free(pk0);
    free(pk1);
    g_pk_loaded = 1;
}

static void release_all(void)
{
    if (g_ctx) seb_destroy(g_ctx);
    g_ctx = NULL;
    free(g_parms.moduli);
    if (g_parms.coeff_count)
    {
        wipe_free(g_ptrs.values, (g_parms.coeff_count / 2) * sizeof(flpt)); /* the last message */
        wipe_free(g_ptrs.ternary, g_parms.coeff_count / 4);                 /* the secret key */
    }
    free(g_ptrs.index_map_ptr);
    free(g_ct);
    g_ct = NULL;
    memset(&g_parms, 0, sizeof g_parms);
    memset(&g_ptrs, 0, sizeof g_ptrs);
    g_pk_loaded = 0;
}

SE_PARMS *se_setup_custom(size_t degree, size_t nprimes, const ZZ *modulus_vals, const ZZ *ratios,
                          double scale, EncryptType encrypt_type)
{
    if (g_ctx) release_all();
    int asym   = encrypt_type == SE_ASYM_ENCR;
    int device = -1;
    const char *dev_env = getenv("SE_B200_DEVICE");
    if (dev_env && *dev_env) device = atoi(dev_env);

    /* With default parameters the requested scale is overridden by the degree's default exactly
     * as set_parms_ckks does (parameters.c:197-225).  Custom moduli keep the caller's
     * scale (parameters.c:232-249).  Their 2n-th roots come from the reference's table where it has
     * one (ntt.c:199-291) and are computed otherwise (seb_minimal_psi); seb_create rejects a modulus
     * that is not a prime = 1 mod 2n below 2^30.
     * `ratios` only selects the custom path (non-NULL, as in parameters.c:235): its VALUES are not
     * read.  The header documents "high word, low word" per modulus (seal_embedded.h:86-87) while
     * set_custom_parms_ckks reads ratios[i], ratios[i+1] (parameters.c:246) - two layouts that agree
     * for one prime only - so const_ratio is always floor(2^64/q) computed here (modulus.c:23-56),
     * which is what either layout is meant to carry. */
    int custom = modulus_vals && ratios;
    g_ctx      = seb_create(degree, nprimes, custom ? modulus_vals : NULL, NULL, custom ? scale : 0.0, asym, device);
    if (!g_ctx)
    {
        printf("Error! se_setup failed: %s\n", seb_last_error());
        exit(1);
    }

    size_t n              = degree;
    g_parms.coeff_count   = n;
    g_parms.logn          = (size_t)log2((double)n);
    g_parms.nprimes       = nprimes;
    g_parms.scale         = seb_scale(g_ctx);
    g_parms.is_asymmetric = asym;
    g_parms.pk_from_file  = 1; /* seal_embedded.c:37-41 */
    g_parms.sample_s      = 0;
    g_parms.small_u       = 1;
    g_parms.small_s       = 1;
    g_parms.moduli        = calloc(nprimes, sizeof(Modulus));
    g_ptrs.values         = calloc(n / 2, sizeof(flpt));
    g_ptrs.index_map_ptr  = calloc(n, sizeof(uint16_t));
    g_ptrs.ternary        = calloc(n / 4, 1);
    g_ct                  = calloc(2 * nprimes * n, sizeof(ZZ));
    if (!g_parms.moduli || !g_ptrs.values || !g_ptrs.index_map_ptr || !g_ptrs.ternary || !g_ct)
    {
        printf("Error! Allocation failed. Exiting...\n");
        exit(1);
    }
    for (size_t i = 0; i < nprimes; i++)
    {
        uint32_t q = seb_prime(g_ctx, i);
        /* floor(2^64/q), low then high word (modulus.c:23-56) */
        unsigned __int128 one     = (unsigned __int128)1 << 64;
        uint64_t ratio            = (uint64_t)(one / q);
        g_parms.moduli[i].value          = q;
        g_parms.moduli[i].const_ratio[0] = (ZZ)ratio;
        g_parms.moduli[i].const_ratio[1] = (ZZ)(ratio >> 32);
    }
    g_parms.curr_modulus_idx = 0;
    g_parms.curr_modulus     = &g_parms.moduli[0];

    /* ckks_calc_index_map (ckks_common.c:32-68), kept host-visible like SE_INDEX_MAP_PERSIST */
    {
        uint64_t m = 2 * (uint64_t)n, pos = 1;
        size_t logn = g_parms.logn;
        for (size_t i = 0; i < n / 2; i++)
        {
            size_t a = (size_t)((pos - 1) / 2), b = n - 1 - a, ra = 0, rb = 0;
            for (size_t k = 0; k < logn; k++)
            {
                ra |= ((a >> k) & 1) << (logn - 1 - k);
                rb |= ((b >> k) & 1) << (logn - 1 - k);
            }
            g_ptrs.index_map_ptr[i]         = (uint16_t)ra;
            g_ptrs.index_map_ptr[i + n / 2] = (uint16_t)rb;
            pos                             = (pos * 3) & (m - 1);
        }
    }

    if (!asym)
    {
        /* ckks_setup_s -> load_sk (ckks_sym.c:162-179, fileops.c:140-170) */
        char fpath[512];
        snprintf(fpath, sizeof fpath, "%s/sk_%zu.dat", SE_DATA_PATH, n);
        read_key_file(fpath, n / 4, g_ptrs.ternary);
        die_on(seb_set_secret_key(g_ctx, (const uint8_t *)g_ptrs.ternary), "seb_set_secret_key");
    }
    g_ptrs.c0_ptr      = g_ct;
    g_ptrs.c1_ptr      = g_ct + n;
    g_se_parms.parms   = &g_parms;
    g_se_parms.se_ptrs = &g_ptrs;
    return &g_se_parms;
}

SE_PARMS *se_setup(size_t degree, size_t nprimes, double scale, EncryptType encrypt_type)
{
    return se_setup_custom(degree, nprimes, NULL, NULL, scale, encrypt_type);
}

SE_PARMS *se_setup_default(EncryptType encrypt_type)
{
    return se_setup(4096, 3, pow(2, 25), encrypt_type); /* seal_embedded.c:90-96 */
}

void se_b200_set_reference_quirk(int on) { g_ref_quirk = on != 0; }

void se_b200_set_print_full(int on) { g_print_full = on != 0; }

void se_b200_set_sym_seed_ct(int on) { g_sym_seed_ct = on != 0; }

/* print_poly (device/lib/util_print.h:478-489) in the reference's two build flavours: the default
 * SE_PRINT_SMALL build prints PRINT_LEN_SMALL = 8 values and "... }" (defines.h:47-50); without it
 * the whole polynomial is printed, which is the text the adapter's ct_string_file_load /
 * poly_string_file_load parse (adapter/fileops.h:221-301, adapter/fileops.cpp:492-538). */
static void print_poly_like_reference(const char *name, const ZZ *a, size_t len)
{
    size_t print_len = (!g_print_full && len > 8) ? 8 : len;
    printf("%s : { ", name);
    for (size_t i = 0; i < print_len; i++)
    {
        printf("%u", a[i]);
        printf(i < len - 1 ? ", " : " "); /* print_comma, util_print.h:86-92 */
    }
    printf("%s", len == print_len ? "}\n" : "... }\n"); /* print_end_string, util_print.h:99-103 */
}

seb_ctx *se_b200_context(SE_PARMS *se_parms)
{
    return (se_parms == &g_se_parms) ? g_ctx : NULL;
}

/* rng.h:45-53: a NULL seed means a fresh random one */
static void fill_seeds(uint8_t *dst, const uint8_t *src, size_t count)
{
    if (src)
    {
        memcpy(dst, src, count * SE_PRNG_SEED_BYTE_COUNT);
        return;
    }
    size_t want = count * SE_PRNG_SEED_BYTE_COUNT, got = 0;
    while (got < want)
    {
        ssize_t r = getrandom(dst + got, want - got, 0);
        if (r < 0)
        {
            printf("Error: getrandom failed\n");
            exit(1);
        }
        got += (size_t)r;
    }
}

bool se_encrypt_batch_seeded(const uint8_t *shareable_seeds, const uint8_t *seeds, const flpt *v, size_t vlen,
                             size_t batch, ZZ *out, SE_PARMS *se_parms)
{
    if (!se_parms || se_parms != &g_se_parms || !g_ctx || !v || !out) return false;
    if (batch == 0) return true;
    /* vlen is also the row stride of v: clamping it would read items b >= 1 from the wrong offsets */
    if (vlen > g_parms.coeff_count / 2)
    {
        printf("Error! vlen %zu exceeds n/2 = %zu values per message.\n", vlen, g_parms.coeff_count / 2);
        return false;
    }
    uint8_t *sd = malloc(batch * SE_PRNG_SEED_BYTE_COUNT);
    uint8_t *ss = g_parms.is_asymmetric ? NULL : malloc(batch * SE_PRNG_SEED_BYTE_COUNT);
    if (!sd || (!g_parms.is_asymmetric && !ss))
    {
        printf("Error! Allocation failed. Exiting...\n");
        exit(1);
    }
    fill_seeds(sd, seeds, batch);
    int rc;
    if (g_parms.is_asymmetric)
    {
        if (!g_pk_loaded) load_public_key();
        rc = seb_encrypt_asym_host(g_ctx, v, vlen, sd, batch, out);
    }
    else
    {
        fill_seeds(ss, shareable_seeds, batch);
        rc = seb_encrypt_sym_host(g_ctx, v, vlen, ss, sd, batch, out, g_ref_quirk);
    }
    wipe_free(sd, batch * SE_PRNG_SEED_BYTE_COUNT); /* the private seeds are secrets */
    free(ss);
    if (rc == SE_ERR_ENCODE_RANGE)
    {
        printf("Error! Value is possibly too large.\n"); /* ckks_common.c:197 */
        return false;
    }
    die_on(rc, "se_encrypt");
    return true;
}

/* Seed-compressed symmetric batch: c0 only, [batch][nprimes][n] words; the receiver rebuilds c1 = a from
 * shareable_seeds (seb_expand_seedct_device).  The shareable seeds are the caller's because they ARE
 * the second half of each ciphertext. */
bool se_encrypt_batch_seedct(const uint8_t *shareable_seeds, const uint8_t *seeds, const flpt *v, size_t vlen,
                             size_t batch, ZZ *c0_out, SE_PARMS *se_parms)
{
    if (!se_parms || se_parms != &g_se_parms || !g_ctx || !v || !c0_out || !shareable_seeds) return false;
    if (g_parms.is_asymmetric) return false;
    if (batch == 0) return true;
    if (vlen > g_parms.coeff_count / 2)
    {
        printf("Error! vlen %zu exceeds n/2 = %zu values per message.\n", vlen, g_parms.coeff_count / 2);
        return false;
    }
    uint8_t *sd = malloc(batch * SE_PRNG_SEED_BYTE_COUNT);
    if (!sd)
    {
        printf("Error! Allocation failed. Exiting...\n");
        exit(1);
    }
    fill_seeds(sd, seeds, batch);
    int rc = seb_encrypt_sym_seedct_host(g_ctx, v, vlen, shareable_seeds, sd, batch, c0_out);
    wipe_free(sd, batch * SE_PRNG_SEED_BYTE_COUNT);
    if (rc == SE_ERR_ENCODE_RANGE)
    {
        printf("Error! Value is possibly too large.\n"); /* ckks_common.c:197 */
        return false;
    }
    die_on(rc, "se_encrypt");
    return true;
}

bool se_encrypt_seeded(uint8_t *shareable_seed, uint8_t *seed, SEND_FNCT_PTR network_send_function, void *v,
                       size_t vlen_bytes, bool print, SE_PARMS *se_parms)
{
    if (!se_parms || se_parms != &g_se_parms || !g_ctx || !v) return false;
    size_t n = g_parms.coeff_count, np = g_parms.nprimes;
    /* seed-compressed mode needs the shareable seed in hand: draw it here when the caller gave none */
    uint8_t own_sseed[SE_PRNG_SEED_BYTE_COUNT];
    const bool seed_ct = g_sym_seed_ct && !g_parms.is_asymmetric;
    if (seed_ct && !shareable_seed)
    {
        fill_seeds(own_sseed, NULL, 1);
        shareable_seed = own_sseed;
    }

    /* seal_embedded.c:108-111: at most n/2 values are taken; slots past the input keep what the
     * previous call staged there (zero after setup), as in the reference */
    size_t copy_size_bytes = (n / 2) * sizeof(ZZ);
    if (vlen_bytes < copy_size_bytes) copy_size_bytes = vlen_bytes;
    memset(g_ptrs.values, 0, copy_size_bytes);
    memcpy(g_ptrs.values, v, copy_size_bytes);

    bool ok = se_encrypt_batch_seeded(shareable_seed, seed, g_ptrs.values, n / 2, 1, g_ct, se_parms);
    if (!ok) return false;

    /* seal_embedded.c:145-213: per prime, c0 then c1, n words each; the callback must consume
     * exactly what it is given */
    for (size_t i = 0; i < np; i++)
    {
        g_parms.curr_modulus_idx = i;
        g_parms.curr_modulus     = &g_parms.moduli[i];
        g_ptrs.c0_ptr            = g_ct + (2 * i) * n;
        g_ptrs.c1_ptr            = g_ct + (2 * i + 1) * n;
        if (print) /* seal_embedded.c:161-165 */
        {
            print_poly_like_reference("c0: ", g_ptrs.c0_ptr, n);
            print_poly_like_reference("c1: ", g_ptrs.c1_ptr, n);
        }
        if (network_send_function)
        {
            size_t nbytes_send = n * sizeof(ZZ);
            if (seed_ct)
            {
                /* the reference's unfinished SE_ENABLE_SYM_SEED_CT (seal_embedded.c:184-194) sends the
                 * 64-byte shareable seed in place of one component; here the pair is (seed, c0), which
                 * is what a receiver needs to rebuild (c0, c1 = a) */
                size_t nb = network_send_function(shareable_seed, SE_PRNG_SEED_BYTE_COUNT);
                if (nb != SE_PRNG_SEED_BYTE_COUNT) return false;
                nb = network_send_function(g_ptrs.c0_ptr, nbytes_send);
                if (nb != nbytes_send) return false;
                continue;
            }
            size_t nbytes_recv = network_send_function(g_ptrs.c0_ptr, nbytes_send);
            if (nbytes_recv != nbytes_send) return false;
            nbytes_recv = network_send_function(g_ptrs.c1_ptr, nbytes_send);
            if (nbytes_recv != nbytes_send) return false;
        }
    }
    return true;
}

bool se_encrypt(SEND_FNCT_PTR network_send_function, void *v, size_t vlen_bytes, bool print, SE_PARMS *se_parms)
{
    return se_encrypt_seeded(NULL, NULL, network_send_function, v, vlen_bytes, print, se_parms);
}

/* SEAL-side layout.  The device library's stream is, per ciphertext, [nprimes][2][n] 32-bit words
 * (c0_p0, c1_p0, c0_p1, ...: seal_embedded.c:196-203).  A seal::Ciphertext of size 2 holds [2][nprimes][n] 64-bit
 * coefficients, which is how the adapter fills it from that stream: ct_ptr[i + j*n] = c0 of prime j,
 * ct_ptr[i + j*n + nprimes*n] = c1 of prime j (adapter/fileops.cpp:518-527).  Host-side, no GPU involved. */
int seb_ct_to_seal_layout(const ZZ *ct, size_t batch, size_t nprimes, size_t n, uint64_t *seal)
{
    if (!ct || !seal || !nprimes || !n) return SE_ERR_INVALD_ARGUMENT;
    for (size_t b = 0; b < batch; b++)
    {
        const ZZ *src = ct + b * 2 * nprimes * n;
        uint64_t *dst = seal + b * 2 * nprimes * n;
        for (size_t j = 0; j < nprimes; j++)
            for (size_t k = 0; k < 2; k++)
            {
                const ZZ *s = src + (2 * j + k) * n;
                uint64_t *d = dst + (k * nprimes + j) * n;
                for (size_t i = 0; i < n; i++) d[i] = s[i];
            }
    }
    return SE_SUCCESS;
}

/* the inverse; a coefficient that does not fit 32 bits (a SEAL ciphertext over wider primes) is an error */
int seb_ct_from_seal_layout(const uint64_t *seal, size_t batch, size_t nprimes, size_t n, ZZ *ct)
{
    if (!ct || !seal || !nprimes || !n) return SE_ERR_INVALD_ARGUMENT;
    for (size_t b = 0; b < batch; b++)
    {
        const uint64_t *src = seal + b * 2 * nprimes * n;
        ZZ *dst             = ct + b * 2 * nprimes * n;
        for (size_t j = 0; j < nprimes; j++)
            for (size_t k = 0; k < 2; k++)
            {
                const uint64_t *s = src + (k * nprimes + j) * n;
                ZZ *d             = dst + (2 * j + k) * n;
                for (size_t i = 0; i < n; i++)
                {
                    if (s[i] >> 32) return SE_ERR_INVALD_ARGUMENT;
                    d[i] = (ZZ)s[i];
                }
            }
    }
    return SE_SUCCESS;
}

void se_cleanup(SE_PARMS *se_parms)
{
    if (!se_parms || se_parms != &g_se_parms) return;
    release_all();
    se_parms->parms = 0; /* seal_embedded.c:234 */
}
