/*
 * seal_embedded_b200.h — C ABI of the B200-native CKKS encode+encrypt library.
 *
 * Two layers, both `extern "C"`, plain pointers and sizes only:
 *
 *  (1) The reference's own public API, symbol for symbol, so the library drops in for
 *      SEAL-Embedded's device library on the encrypt path
 *      (reference: device/lib/seal_embedded.h:91-130):
 *          se_setup_custom / se_setup / se_setup_default   seal_embedded.h:91-118
 *          se_encrypt_seeded / se_encrypt                  seal_embedded.h:120-124
 *          se_cleanup                                      seal_embedded.h:126-130
 *      with the caller-visible structs Parms (parameters.h:43-67), Modulus (modulus.h:22-30),
 *      SE_PTRS (ckks_common.h:36-52) and SE_PARMS (seal_embedded.h:52-56) laid out as in the
 *      reference's default (SE_USE_MALLOC, 32-bit ZZ) configuration.
 *
 *  (2) The batch extension the reference lacks (one call = many independent ciphertexts), in a
 *      host-pointer and a device-pointer flavour, plus stage-level entry points used by the parity
 *      tests and the NTT micro-benchmark.  These are the `seb_*` functions.
 *
 * All functions return 0 (SE_SUCCESS) or a negative SE_ERR_* code unless stated otherwise;
 * seb_last_error() gives the text of the last failure on the calling thread.
 * One process drives one GPU (select it with seb_create's `device` or SE_B200_DEVICE before
 * se_setup).  A context is not thread-safe, like the reference (seal_embedded.c:18-22).
 */
#ifndef SEAL_EMBEDDED_B200_H
#define SEAL_EMBEDDED_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* (1) reference-compatible API                                                               */
/* ------------------------------------------------------------------------------------------ */
typedef uint32_t ZZ; /* defines.h:373 */
typedef float flpt;  /* defines.h:376 */

#define SE_PRNG_SEED_BYTE_COUNT 64 /* defines.h:67 */

#define SE_SUCCESS 0 /* seal_embedded.h:35-40 */
#define SE_ERR_NO_MEMORY -12
#define SE_ERR_INVALD_ARGUMENT -22
#define SE_ERR_UNKNOWN -1000
#define SE_ERR_MINIMUM -9999
/* extension codes (inside the reference's reserved range) */
#define SE_ERR_CUDA -1001        /* a CUDA call failed */
#define SE_ERR_ENCODE_RANGE -1002 /* ckks_encode_base would have returned false */
#define SE_ERR_NO_KEY -1003      /* key material missing for the requested encryption type */

typedef struct Modulus /* modulus.h:22-30 */
{
    ZZ value;
    ZZ const_ratio[2]; /* floor(2^64/q): [0] low word, [1] high word */
} Modulus;

typedef struct /* parameters.h:43-67 (SE_USE_MALLOC, no SE_REVERSE_CT_GEN_ENABLED) */
{
    size_t coeff_count;
    size_t logn;
    Modulus *moduli;
    Modulus *curr_modulus;
    size_t curr_modulus_idx;
    size_t nprimes;
    double scale;
    bool is_asymmetric;
    bool pk_from_file;
    bool sample_s;
    bool small_s;
    bool small_u;
} Parms;

/* ckks_common.h:36-52.  Host-side views; see INTEGRATION.md for which are populated:
 * values, c0_ptr, c1_ptr (valid during the send callback), index_map_ptr and ternary (the packed
 * secret key in symmetric mode) are; the MCU scratch pointers are NULL. */
typedef struct SE_PTRS
{
    void *conj_vals;  /* `double complex *` in the reference */
    void *ifft_roots; /* `double complex *` in the reference */
    flpt *values;
    ZZ *ternary;
    int64_t *conj_vals_int_ptr;
    ZZ *c0_ptr;
    ZZ *c1_ptr;
    uint16_t *index_map_ptr;
    ZZ *ntt_roots_ptr;
    ZZ *ntt_pte_ptr;
    int8_t *e1_ptr;
} SE_PTRS;

typedef struct /* seal_embedded.h:52-56 */
{
    Parms *parms;
    SE_PTRS *se_ptrs;
} SE_PARMS;

typedef enum { SE_SYM_ENCR, SE_ASYM_ENCR } EncryptType; /* seal_embedded.h:58 */

typedef size_t (*SEND_FNCT_PTR)(void *, size_t);                    /* seal_embedded.h:65 */
typedef ssize_t (*RND_FNCT_PTR)(void *, size_t, unsigned int flags); /* seal_embedded.h:73 */

SE_PARMS *se_setup_custom(size_t degree, size_t nprimes, const ZZ *modulus_vals, const ZZ *ratios,
                          double scale, EncryptType encrypt_type);
SE_PARMS *se_setup(size_t degree, size_t nprimes, double scale, EncryptType encrypt_type);
SE_PARMS *se_setup_default(EncryptType encrypt_type);
bool se_encrypt_seeded(uint8_t *shareable_seed, uint8_t *seed, SEND_FNCT_PTR network_send_function,
                       void *v, size_t vlen_bytes, bool print, SE_PARMS *se_parms);
bool se_encrypt(SEND_FNCT_PTR network_send_function, void *v, size_t vlen_bytes, bool print,
                SE_PARMS *se_parms);
void se_cleanup(SE_PARMS *se_parms);

/* Batch extension at the se_* level (host buffers).  Each item b encrypts v[b*vlen .. +vlen)
 * (vlen <= n/2 floats, zero padded) with seeds[b*64 .. +64) (and shareable_seeds in symmetric
 * mode; NULL seeds are drawn from getrandom()).  out receives, per item, the byte stream
 * se_encrypt would have sent: [nprimes][2][n] words = c0_p0, c1_p0, c0_p1, ...
 * (seal_embedded.c:196-203).  Symmetric c1 is `a` unless se_b200_set_reference_quirk(1).
 * Returns false if any item failed to encode (reference: ckks_common.c:195-204). */
bool se_encrypt_batch_seeded(const uint8_t *shareable_seeds, const uint8_t *seeds, const flpt *v,
                             size_t vlen, size_t batch, ZZ *out, SE_PARMS *se_parms);
/* Reproduce se_encrypt's symmetric byte stream exactly, where the c1 buffer handed to the send
 * callback holds ntt(m+e) (ckks_sym.c:86-88 aliasing).  Default 0: c1 = a. */
void se_b200_set_reference_quirk(int on);
/* se_encrypt(print = true) prints each component with the reference's print_poly: 8 values and "... }"
 * as in the default SE_PRINT_SMALL build (0), or the whole polynomial as in a build without it (1) —
 * the text format the adapter's ct_string_file_load parses (adapter/fileops.cpp:492-538). */
void se_b200_set_print_full(int on);
/* Seed-compressed symmetric ciphertexts (the reference's unfinished SE_ENABLE_SYM_SEED_CT,
 * seal_embedded.c:184-194): with the switch on, symmetric se_encrypt* calls send, per prime, the 64-byte
 * shareable seed and then c0 (n words) instead of (c0, c1); c1 = a is a function of the seed alone
 * (sample.c:39-57) and the receiver rebuilds it (seb_expand_seedct_device).  Default 0. */
void se_b200_set_sym_seed_ct(int on);
/* Batch form: c0_out [batch][nprimes][n] words, half the bytes of se_encrypt_batch_seeded's output.
 * shareable_seeds [batch][64] are required (they are the other half of each ciphertext); seeds may be NULL. */
bool se_encrypt_batch_seedct(const uint8_t *shareable_seeds, const uint8_t *seeds, const flpt *v, size_t vlen,
                             size_t batch, ZZ *c0_out, SE_PARMS *se_parms);
/* SEAL-side ciphertext layout (adapter/fileops.cpp:518-527): the stream above is [nprimes][2][n] 32-bit words per
 * ciphertext; a seal::Ciphertext of size 2 holds [2][nprimes][n] 64-bit coefficients (c0 of every prime, then c1 of
 * every prime).  Host-side conversions for `batch` ciphertexts, both directions; `from` fails with
 * SE_ERR_INVALD_ARGUMENT when a coefficient does not fit 32 bits. */
int seb_ct_to_seal_layout(const ZZ *ct, size_t batch, size_t nprimes, size_t n, uint64_t *seal);
int seb_ct_from_seal_layout(const uint64_t *seal, size_t batch, size_t nprimes, size_t n, ZZ *ct);
/* The context behind the static SE_PARMS (for the seb_* calls below); NULL before se_setup. */
struct seb_ctx *se_b200_context(SE_PARMS *se_parms);

/* ------------------------------------------------------------------------------------------ */
/* (2) batch / device-pointer extension                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct seb_ctx seb_ctx;

const char *seb_last_error(void);

/* primes == NULL: the reference's default chain for (n, nprimes) (parameters.c:129-230) with its
 * tabulated 2n-th roots (ntt.c:199-291) and default scale when scale <= 0.  With explicit primes
 * (distinct, each a PRIME < 2^30 and = 1 mod 2n: checked, deterministic Miller-Rabin), psis[i] must be a
 * primitive 2n-th root of unity mod primes[i]; psis == NULL takes the reference's tabulated root where there is
 * one and seb_minimal_psi otherwise.  device < 0: keep the current CUDA device.
 * Test / A-B switches are read from the environment here, once (SEB_UNIFORM_COOP, SEB_UNIFORM_FIX_WIDE,
 * SEB_UNIFORM_SPEC, SEB_UNIFORM_PAIR, SEB_HOST_CHUNK, SEB_UNIFORM_LIST_CAP, SEB_UNIFORM_SPEC_SIGMAS); afterwards
 * seb_set_option changes them. */
seb_ctx *seb_create(size_t n, size_t nprimes, const uint32_t *primes, const uint32_t *psis,
                    double scale, int asym, int device);
void seb_destroy(seb_ctx *ctx);
/* Smallest primitive 2n-th root of unity mod q, 0 if there is none (host arithmetic, no GPU needed).
 * Equals the reference's table get_ntt_root (ntt.c:199-291) wherever that is defined. */
uint32_t seb_minimal_psi(size_t n, uint32_t q);
/* name in {"uniform_coop", "uniform_fix_wide", "uniform_spec", "uniform_pair"}: 0 / 1 force a code path, a negative
 * value restores the automatic choice; "uniform_fix_lanes": 4, 8 or 32 lanes per ciphertext in the uniform sampler's
 * fix-up; "uniform_fix_stream": 2 / 4 / 8 ciphertexts per warp in its streamed form (0: off); "sym_partition": 0 / 1 the
 * symmetric path with its sampler chain and its encode / CBD work on disjoint SM partitions (CUDA green contexts; by
 * itself for batches of ~16k items at n >= 8192), "sym_side_percent": the share of the batch whose encode / CBD run on
 * the side partition; "host_chunk": items per chunk of the host-pointer pipeline (<= 0: automatic). */
int seb_set_option(seb_ctx *ctx, const char *name, long value);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the context's own */
int seb_set_stream(seb_ctx *ctx, void *cuda_stream);
/* pk0, pk1: host [nprimes][n], NTT form (files pk{0,1}_ntt_<n>_<q>.dat, fileops.c:172-204) */
int seb_set_public_key(seb_ctx *ctx, const uint32_t *pk0, const uint32_t *pk1);
/* sk: host n/4 bytes, 2 bits per coefficient (file sk_<n>.dat, fileops.c:140-170) */
int seb_set_secret_key(seb_ctx *ctx, const uint8_t *sk_packed);
/* ckks_setup_s with sample_s (ckks_sym.c:162-173): s = sample_small_poly_ternary_prng_96(PRNG(seed), counter 0)
 * (sample.c:218-242) on the GPU; sk_out receives the n/4 packed bytes (sk_<n>.dat format) and the key is installed */
int seb_gen_secret_key(seb_ctx *ctx, const uint8_t *seed, uint8_t *sk_out);
/* gen_pk (ckks_asym.c:159-171) on the GPU: loads sk (as seb_set_secret_key), samples ep = CBD(PRNG(ep_seed))
 * and, per prime p, a = uniform(PRNG(a_seed_base with byte 0 := p)); writes pk0 = -(a (.) ntt(s)) + ntt(ep)
 * and pk1 = a to host [nprimes][n] and installs them when the context is asymmetric */
int seb_gen_public_key(seb_ctx *ctx, const uint8_t *sk_packed, const uint8_t *ep_seed, const uint8_t *a_seed_base,
                       uint32_t *pk0, uint32_t *pk1);
/* pre-size the per-batch scratch (otherwise grown on demand) */
int seb_reserve(seb_ctx *ctx, size_t batch);

size_t seb_degree(const seb_ctx *ctx);
size_t seb_nprimes(const seb_ctx *ctx);
double seb_scale(const seb_ctx *ctx);
uint32_t seb_prime(const seb_ctx *ctx, size_t i);
/* number of kernels this library has launched so far on this context */
uint64_t seb_launch_count(const seb_ctx *ctx);

/* ---- full path, device pointers, asynchronous on the context's stream ---- */
/* d_values [batch][vlen] fp32, d_seeds [batch][64], d_out [batch][nprimes][2][n] u32.
 * All pointers 16-byte aligned. */
int seb_encrypt_asym_device(seb_ctx *ctx, const float *d_values, size_t vlen, const uint8_t *d_seeds,
                            size_t batch, uint32_t *d_out);
int seb_encrypt_sym_device(seb_ctx *ctx, const float *d_values, size_t vlen,
                           const uint8_t *d_shareable_seeds, const uint8_t *d_seeds, size_t batch,
                           uint32_t *d_out, int ref_quirk);
/* Seed-compressed symmetric form: d_c0_out [batch][nprimes][n] receives c0 only (a stays in scratch). */
int seb_encrypt_sym_seedct_device(seb_ctx *ctx, const float *d_values, size_t vlen,
                                  const uint8_t *d_shareable_seeds, const uint8_t *d_seeds, size_t batch,
                                  uint32_t *d_c0_out);
/* Receiver side: d_out [batch][nprimes][2][n] gets c0 from d_c0 [batch][nprimes][n] (NULL: c0 slots left
 * untouched) and c1 = a regenerated from the shareable seeds, bit-identical to what the full-size
 * call would have produced. */
int seb_expand_seedct_device(seb_ctx *ctx, const uint8_t *d_shareable_seeds, const uint32_t *d_c0, size_t batch,
                             uint32_t *d_out);
/* Small symmetric calls (up to 4 ciphertexts at n = 16384 x 6 primes, 16 at n = 4096 x 3) run every prime's uniform squeeze at once on speculated PRNG counters
 * (the counter of prime p depends on the redraws of the primes before it; all counters within 5 sigma of the mean
 * are squeezed in parallel and the right one is selected afterwards — same bytes, a third to a sixth of the
 * dependent work).  Returns how many squeezes fell outside their window and were redone on the spot
 * (diagnostic; synchronises the stream). */
long seb_uniform_spec_misses(seb_ctx *ctx);
/* synchronises the stream; returns how many items of the last *_device call failed to encode
 * (>= 0) or a negative error */
int seb_encode_failures(seb_ctx *ctx);

/* Per-kernel CUDA-event timing of the next max_steps *_device full-path calls (events recorded on
 * the context's stream around each kernel).  seb_profile_end synchronises, writes ms[step][4]
 * (asym: encode, sample_ternary, sample_cbd, encrypt; sym: encode, sample_cbd, sample_uniform,
 * encrypt) and returns the number of steps recorded. */
int seb_profile_begin(seb_ctx *ctx, int max_steps);
int seb_profile_end(seb_ctx *ctx, float *ms);

/* ---- full path, host pointers: pinned staging, chunked, H2D/compute/D2H overlapped ---- */
int seb_encrypt_asym_host(seb_ctx *ctx, const float *values, size_t vlen, const uint8_t *seeds,
                          size_t batch, uint32_t *out);
int seb_encrypt_sym_host(seb_ctx *ctx, const float *values, size_t vlen,
                         const uint8_t *shareable_seeds, const uint8_t *seeds, size_t batch,
                         uint32_t *out, int ref_quirk);

int seb_encrypt_sym_seedct_host(seb_ctx *ctx, const float *values, size_t vlen,
                                const uint8_t *shareable_seeds, const uint8_t *seeds, size_t batch,
                                uint32_t *c0_out);

/* Optional packed wire form of the host-pointer calls: every residue is below 2^30, so a ciphertext's [nprimes][2][n]
 * words travel as 30 bits per residue — seb_packed30_words(ctx) = 2 * nprimes * n * 15 / 16 words per ciphertext,
 * residue i of each group of 16 in bits 30i .. 30i+29 (little endian) of the group's 15 words.  The host-pointer path
 * is bound by the PCIe link (6.25 % fewer bytes = 6.25 % more ciphertexts/s); the full form stays the default.
 * seb_unpack30 (host, no GPU) / seb_unpack30_device restore the full form bit for bit: `words` = full-form words, a
 * multiple of 16. */
size_t seb_packed30_words(const seb_ctx *ctx);
int seb_encrypt_asym_host_packed30(seb_ctx *ctx, const float *values, size_t vlen, const uint8_t *seeds,
                                   size_t batch, uint32_t *out_packed);
int seb_encrypt_sym_host_packed30(seb_ctx *ctx, const float *values, size_t vlen, const uint8_t *shareable_seeds,
                                  const uint8_t *seeds, size_t batch, uint32_t *out_packed, int ref_quirk);
int seb_unpack30(const uint32_t *packed, size_t words, uint32_t *out);
int seb_unpack30_device(seb_ctx *ctx, const uint32_t *d_packed, size_t words, uint32_t *d_out);

/* ---- stage level (device pointers, asynchronous) ---- */
/* ckks_encode_base: d_pt [batch][n] int64 */
int seb_encode_device(seb_ctx *ctx, const float *d_values, size_t vlen, size_t batch, int64_t *d_pt);
/* ckks_asym_init samplers: d_u [batch][n/4], d_e [batch][2][n] int8 (e0 then e1), d_ctr [batch]
 * = PRNG counter after u (e0 uses the next n/16 counters, e1 the n/16 after those) */
int seb_sample_asym_device(seb_ctx *ctx, const uint8_t *d_seeds, size_t batch, uint8_t *d_u, int8_t *d_e,
                           uint32_t *d_ctr);
/* CBD polynomials from counter d_ctr[b] (NULL: 0): d_e [batch][npoly][n] */
int seb_sample_cbd_device(seb_ctx *ctx, const uint8_t *d_seeds, const uint32_t *d_ctr, size_t npoly,
                          size_t batch, int8_t *d_e);
/* sample_poly_uniform under prime prime_idx: row b written at d_out + b*ct_stride (n words);
 * d_ctr [batch] is read and advanced */
int seb_sample_uniform_device(seb_ctx *ctx, const uint8_t *d_seeds, uint32_t *d_ctr, size_t prime_idx,
                              size_t batch, uint32_t *d_out, size_t ct_stride);
/* ntt_inpl, in place: d_polys [batch][nprimes][n] (polynomial k uses prime k % nprimes) */
int seb_ntt_device(seb_ctx *ctx, uint32_t *d_polys, size_t batch);
/* ---- verifier (the receiving side, so whole batches can be round-tripped on the GPU) ---- */
/* inverse of seb_ntt_device: intt (device/lib/intt.c:226-501), in place, canonical residues */
int seb_intt_device(seb_ctx *ctx, uint32_t *d_polys, size_t batch);
/* ckks_decrypt + intt + ckks_decode under prime prime_idx (device/test/ckks_tests_common.c:59-171):
 * d_ct [batch][nprimes][2][n] -> d_values_out [batch][vlen] fp32.  Needs seb_set_secret_key (also on
 * an asymmetric context).  Symmetric ciphertexts must carry c1 = a (ref_quirk off). */
int seb_decrypt_decode_device(seb_ctx *ctx, const uint32_t *d_ct, size_t batch, size_t prime_idx, size_t vlen,
                              float *d_values_out);
/* per-item digest of d_words [items][words_per_item] (words_per_item % 4 == 0):
 * d_digests[b] = sum_i mix64((i << 32) | word_i) mod 2^64 with the splitmix64 finaliser — 8 bytes per ciphertext
 * to compare a full-size batch with the reference run on the host (tests/, oracle/ref_shim.c: ref_encrypt_digests) */
int seb_digest_device(seb_ctx *ctx, const uint32_t *d_words, size_t words_per_item, size_t items, uint64_t *d_digests);
/* The integer-issue ceilings of this device, measured on the spot with register-only loops of the path's two inner
 * operations: Keccak-f[1600] permutations/s (what bounds the samplers) and lazy NTT butterflies/s (what bounds the
 * NTT / encrypt kernels beside HBM).  bench.py's roofline denominators for the kernels that are not HBM-bound. */
int seb_measure_ceilings(seb_ctx *ctx, double *keccak_f_per_s, double *butterflies_per_s);
/* first 136-byte block of SHAKE256(seed_i || LE64(counter_i)): d_out [count][17] u64 */
int seb_prng_blocks_device(seb_ctx *ctx, const uint8_t *d_seeds, const uint64_t *d_counters, size_t count,
                           uint64_t *d_out);

#ifdef __cplusplus
}
#endif
#endif
