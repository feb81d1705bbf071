/*
 * se_oracle.h — CPU restatement of SEAL-Embedded's CKKS encode+encrypt path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (seal-embedded_b200/csrc + host/) never links or calls this code.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the reference
 * library itself (oracle/_ref/libseref.so, compiled from /root/reference/device/lib by
 * oracle/Makefile) in tests/test_oracle_vs_ref.py, against the reference's scalar KATs
 * (device/test/modulo_tests.c, uintmodarith_tests.c) and against the committed fixtures in
 * tests/golden/ that were generated from that library (tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/device/lib unless noted).
 */
#ifndef SE_ORACLE_H
#define SE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_SEED_BYTES 64 /* defines.h:67 SE_PRNG_SEED_BYTE_COUNT */

/* ---- SHAKE256 / PRNG (shake256/fips202.c:105-128, rng.h:78-91) ---- */
void orc_keccak_f1600(uint64_t st[25]);
void orc_shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen);
void orc_prng_fill(const uint8_t seed[ORC_SEED_BYTES], uint64_t counter, size_t nbytes,
                   uint8_t *out);

/* ---- parameter tables (parameters.c:129-230, modulus.c:23-56, ntt.c:199-291) ---- */
/* Fills primes[0..np) for the default chain of degree n; returns 0 on an illegal (n,np). */
int orc_default_primes(size_t n, size_t np, uint32_t *primes);
double orc_default_scale(size_t n);
/* psi = first power of the 2n-th root the reference tabulates for (n,q); 0 if not tabulated. */
uint32_t orc_ntt_root(size_t n, uint32_t q);
/* floor(2^64/q) as {lo,hi} 32-bit words (Modulus.const_ratio[0], [1]). */
void orc_const_ratio(uint32_t q, uint32_t ratio[2]);

/* ---- modular arithmetic (modulo.h:43-116, uintmodarith.h:26-168) ---- */
uint32_t orc_barrett32(uint32_t x, uint32_t q);
uint32_t orc_barrett64(uint32_t lo, uint32_t hi, uint32_t q);
uint32_t orc_add_mod(uint32_t a, uint32_t b, uint32_t q);
uint32_t orc_neg_mod(uint32_t a, uint32_t q);
uint32_t orc_sub_mod(uint32_t a, uint32_t b, uint32_t q);
uint32_t orc_mul_mod(uint32_t a, uint32_t b, uint32_t q);
uint32_t orc_pow_mod(uint32_t a, uint64_t e, uint32_t q);

/* ---- CKKS encode (ckks_common.c:32-68, 105-215; fft.c:27-45, 69-144) ---- */
size_t orc_bitrev(size_t x, size_t nbits);
void orc_index_map(size_t n, uint16_t *map);
/* twiddle table used by the IFFT: tw[2*i], tw[2*i+1] = conj(e^{2 pi i bitrev(i,logn)/2n}), i in [1,n) */
void orc_ifft_twiddles(size_t n, double *tw);
/* values[0..vlen) zero-padded to n/2; writes n int64 coefficients; returns 0 when the
 * reference would return false (|coeff| > 2^63). */
int orc_encode(size_t n, double scale, const float *values, size_t vlen, int64_t *out);
/* inverse direction, test helper (device/test/ckks_tests_common.c:59-118, fft.c:146-213) */
void orc_decode(size_t n, double scale, uint32_t q, const uint32_t *pt, size_t vlen, float *values);

/* ---- samplers (sample.c:39-57, 218-242, 263-356; Appendix C of SURVEY.md) ---- */
void orc_sample_ternary_small(size_t n, const uint8_t *seed, uint64_t *counter, uint8_t *packed);
void orc_sample_cbd(size_t n, const uint8_t *seed, uint64_t *counter, int8_t *out);
void orc_sample_uniform(size_t n, uint32_t q, const uint8_t *seed, uint64_t *counter,
                        uint32_t *out);
void orc_expand_ternary(size_t n, uint32_t q, const uint8_t *packed, uint32_t *out);
void orc_reduce_small(size_t n, uint32_t q, const int8_t *e, uint32_t *out);
void orc_reduce_pte(size_t n, uint32_t q, const int64_t *pte, uint32_t *out);

/* ---- NTT (ntt.c:24-60, 124-189) ---- */
void orc_ntt_roots(size_t n, uint32_t q, uint32_t psi, uint32_t *roots);
void orc_ntt(size_t n, uint32_t q, const uint32_t *roots, uint32_t *vec);
void orc_intt(size_t n, uint32_t q, uint32_t psi, uint32_t *vec); /* test helper */
void orc_ntt_default(size_t n, uint32_t q, uint32_t *vec);        /* roots from orc_ntt_root */
void orc_ntt_psi(size_t n, uint32_t q, uint32_t psi, uint32_t *vec); /* roots from an explicit psi */
/* O(n^2) negacyclic product, test helper (polymodmult.c:37-101) */
void orc_negacyclic_mul(size_t n, uint32_t q, const uint32_t *a, const uint32_t *b, uint32_t *c);

/* ---- full path (seal_embedded.c:98-215, ckks_asym.c:173-286, ckks_sym.c:181-301) ---- */
/* out layout: [np][2][n] = c0_p0, c1_p0, c0_p1, c1_p1 ... (wire order, seal_embedded.c:196-203).
 * pk0/pk1: [np][n] NTT form.  Returns 0 when encode fails. */
int orc_encrypt_asym(size_t n, size_t np, const float *values, size_t vlen, const uint8_t *seed,
                     const uint32_t *pk0, const uint32_t *pk1, uint32_t *out);
/* sk_packed: n/4 bytes, 2 bits per coefficient.  ref_quirk != 0 reproduces the byte stream of
 * the reference's se_encrypt in its default memory layout, where the c1 buffer handed to the send
 * callback holds ntt(m+e) instead of a (ckks_sym.c:86-88 aliasing). */
int orc_encrypt_sym(size_t n, size_t np, const float *values, size_t vlen,
                    const uint8_t *share_seed, const uint8_t *seed, const uint8_t *sk_packed,
                    int ref_quirk, uint32_t *out);
/* gen_pk (ckks_asym.c:159-171): pk = sym encryption of 0 with error ep under prime index p,
 * a drawn from PRNG(seed) starting at counter 0.  pk0,pk1: n words each. */
void orc_gen_pk_prime(size_t n, uint32_t q, const uint8_t *seed, const uint8_t *sk_packed,
                      const int8_t *ep, uint32_t *pk0, uint32_t *pk1);
/* c0 + c1*s in NTT form (device/test/ckks_tests_common.c:142-171) */
void orc_decrypt_ntt(size_t n, uint32_t q, const uint32_t *c0, const uint32_t *c1,
                     const uint8_t *sk_packed, uint32_t *pt_ntt);

/* batch helpers used as the CPU baseline ("port" kind): sequential loop over items */
/* The same with an explicit chain: primes[np], their primitive 2n-th roots psis[np] and the scale (a caller
 * chain as se_setup_custom / set_custom_parms_ckks would install, parameters.c:232-249).  The reference's own
 * custom path does not terminate, so these are "parity unpinned" against the reference beyond
 * the default chains, where they coincide with the functions above. */
int orc_encrypt_asym_ex(size_t n, size_t np, const uint32_t *primes, const uint32_t *psis, double scale,
                        const float *values, size_t vlen, const uint8_t *seed, const uint32_t *pk0,
                        const uint32_t *pk1, uint32_t *out);
int orc_encrypt_sym_ex(size_t n, size_t np, const uint32_t *primes, const uint32_t *psis, double scale,
                       const float *values, size_t vlen, const uint8_t *share_seed, const uint8_t *seed,
                       const uint8_t *sk_packed, int ref_quirk, uint32_t *out);
void orc_gen_pk_prime_ex(size_t n, uint32_t q, uint32_t psi, const uint8_t *seed, const uint8_t *sk_packed,
                         const int8_t *ep, uint32_t *pk0, uint32_t *pk1);
void orc_decrypt_ntt_ex(size_t n, uint32_t q, uint32_t psi, const uint32_t *c0, const uint32_t *c1,
                        const uint8_t *sk_packed, uint32_t *pt_ntt);
int orc_encrypt_asym_batch(size_t n, size_t np, size_t batch, const float *values, size_t vlen,
                           const uint8_t *seeds, const uint32_t *pk0, const uint32_t *pk1,
                           uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
