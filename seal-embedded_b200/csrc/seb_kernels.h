// seb_kernels.h — launchers of the sm_100a kernels (internal; the public C ABI is
// include/seal_embedded_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "seb_common.cuh"

struct SebModuli
{
    SebModulus m[SEB_MAX_PRIMES];
};

// Test / A-B switches of a context.  -1 = automatic (the library decides from the batch size), 0 / 1 force a path.
// Initialised from the environment ONCE in seb_create (SEB_UNIFORM_COOP, SEB_UNIFORM_FIX_WIDE, SEB_UNIFORM_SPEC,
// SEB_UNIFORM_PAIR, SEB_HOST_CHUNK) and changed at run time with seb_set_option; never read from the environment
// on the call path.
struct SebKnobs
{
    int uniform_coop     = -1;  // bulk squeeze by a warp per ciphertext (25 lanes per sponge)
    int uniform_fix_wide = -1;  // fix-up by a CTA per ciphertext
    int uniform_spec     = -1;  // speculative prime chain for lone calls
    int uniform_pair     = -1;  // bulk squeeze by two lanes per sponge (bit-interleaved halves)
    int uniform_fix_lanes = -1; // fix-up lanes per ciphertext: 4, 8 or 32 (-1: by the expected number of rejections)
    int uniform_fix_stream = -1; // the 32-lane fix-up as a stream over 2 / 4 / 8 ciphertexts per warp (2, 4, other > 0; 0: off; -1: by batch size)
    int sym_partition    = -1;  // symmetric path: sampler chain and encode / CBD on disjoint SM partitions (-1: by shape)
    int sym_side_percent = -1;  // ... share of the batch whose encode / CBD run on the side partition (-1: from the kernels' rates)
    long host_chunk      = 0;   // items per chunk of the host-pointer pipeline (0 = automatic)
    int uniform_mix      = -1;  // bulk squeeze: the sponges beyond the last full layer of thread-kernel warps in the two-lane kernel
    int sms              = 0;   // SM count of the context's device (kernel selection by machine fill; not an option)
    // a second stream and two events of the context (not options): the two kernels of the mixed squeeze run side by side
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t aux_ev[2]   = {nullptr, nullptr};
};

// ---- samplers (seb_sample.cu) ----
void seb_launch_prng_blocks(const uint8_t *seeds, const uint64_t *counters, uint64_t *out, int count,
                            cudaStream_t st);
void seb_launch_sample_ternary(const uint8_t *seeds, uint8_t *u_out, uint32_t *ctr_out, int n, int batch,
                               cudaStream_t st);
void seb_launch_sample_cbd(const uint8_t *seeds, const uint32_t *ctr_base, int8_t *e_out, int n, int npoly,
                           int batch, cudaStream_t st);
// rej_idx [batch][rej_cap] / rej_cnt [batch]: scratch for the per-ciphertext lists of rejected words
void seb_launch_uniform(const uint8_t *seeds, uint32_t *ctr, uint32_t *out, size_t ct_stride, int n,
                        const SebModulus &mod, int batch, uint16_t *rej_idx, uint32_t *rej_cnt, uint32_t rej_cap,
                        const SebKnobs &knobs, cudaStream_t st);

// Lone calls: all primes' squeezes at once, speculating on the chained counters (seb_sample.cu).
struct SebSpecPrime
{
    uint32_t lo, width, first;
};
struct SebSpecPlan
{
    SebSpecPrime p[SEB_MAX_PRIMES];
    uint32_t total;
};
void seb_uniform_spec_plan(int n, const SebModuli &mods, int np, double sigmas, SebSpecPlan *plan);
void seb_launch_uniform_chain_spec(const uint8_t *seeds, uint32_t *ctr, uint32_t *out_p0, size_t ct_stride, size_t p_stride,
                                   int n, const SebModuli &mods, int np, const SebSpecPlan &plan, int batch,
                                   uint32_t *cand_rows, uint16_t *cand_list, uint32_t *cand_cnt, uint16_t *rej_idx,
                                   uint32_t *rej_cnt, uint32_t rej_cap, uint32_t *misses, const SebKnobs &knobs,
                                   cudaStream_t st);

// ---- encode (seb_encode.cu) ----
// values: [batch][v_stride] floats, the first vlen of each row are used (zero padded to n/2);
// src_map[pos] = slot whose value lands on position pos; tw[i] = IFFT twiddle (re, im), i in [1,n), followed by
// the pass-0 copies seb_host_build_enc_tw0 appends (seb_enc_tw_entries(n) entries in all)
// fail[b] is set when item b overflows int64; mag[b] (may be NULL; zeroed by the caller) receives
// max |coefficient| of item b clipped to 2^32 - 1
cudaError_t seb_launch_encode(int logn, const float *values, size_t v_stride, int vlen, const uint16_t *src_map,
                              const double2 *tw, double n_inv, int64_t *pt, int *fail, uint32_t *mag, int batch,
                              cudaStream_t st);
cudaError_t seb_encode_configure(int logn);
size_t seb_enc_tw_entries(size_t n);
void seb_host_build_enc_tw0(size_t n, double2 *tw);  // tw[0..n) filled on entry

// ---- NTT + encrypt (seb_encrypt.cu) ----
// Tables are laid out per NTT PLAN, named by a key (seb_ntt.cuh: NttCfg): key = logn for the 16-coefficients-per-thread
// plans (the asymmetric kernel, every degree) and seb_ntt_key1(logn) for the plan of the one-polynomial kernels
// (seb_launch_ntt, seb_launch_encrypt_sym), which differs from logn at n >= 8192.
// roots: per prime, the per-pass twiddle tables of seb_build_tw (seb_table_octs(key) octs each);
// pk0s/pk1s (key = logn) and ntt_s (key = seb_ntt_key1(logn)): per prime, n/4 octs in the epilogue order of seb_build_epi
int seb_ntt_key1(int logn);
// the symmetric kernel's own tables: those of plan seb_ntt_key1(logn), except where it runs as two half-size transforms
// per polynomial (seb_sym_split: n = 16384)
int seb_sym_split(int logn);
size_t seb_table_octs_sym(int logn);  // octs per prime of its root table
void seb_host_build_tw_sym(int logn, const uint2 *roots_bitrev, seb_oct *out);
void seb_host_build_epi_sym(int logn, const uint2 *natural, seb_oct *out);  // n coefficients (ntt(s)) in its epilogue order
size_t seb_table_octs(int key);
void seb_host_build_tw(int key, const uint2 *roots_bitrev, seb_oct *out);
void seb_host_build_epi(int key, const uint2 *natural, seb_oct *out);
cudaError_t seb_launch_ntt(int logn, uint32_t *polys, const seb_oct *roots, const SebModuli &mods, int np,
                           size_t npolys_total, cudaStream_t st);
// asym: out[b][p][0] = pk0 (.) ntt(u) + ntt(m+e0), out[b][p][1] = pk1 (.) ntt(u) + ntt(e1)
// mag: per item max |pt| clipped to 32 bits (from seb_launch_encode); selects the 32-bit reduction path
cudaError_t seb_launch_encrypt_asym(int logn, const int64_t *pt, const uint32_t *mag, const int8_t *e, const uint8_t *u,
                                    const seb_oct *roots, const seb_oct *pk0s, const seb_oct *pk1s,
                                    const SebModuli &mods, int np, uint32_t *out, int batch, cudaStream_t st);
// sym: a (already sampled) and c0 live at base + b*ct_stride + p*p_stride words; c0 = -(a (.) ntt(s)) + ntt(m+e).
// Full layout: a = out + n, c0 = out, ct_stride = 2*np*n, p_stride = 2n.  Seed-compressed layout
// (SURVEY 8f-2): a in scratch, c0 = the [batch][np][n] output, ct_stride = np*n, p_stride = n.
// quirk != 0 additionally overwrites a's slot with ntt(m+e) (reference byte stream)
cudaError_t seb_launch_encrypt_sym(int logn, const int64_t *pt, const uint32_t *mag, const int8_t *e, const seb_oct *roots,
                                   const seb_oct *ntt_s, const SebModuli &mods, int np, uint32_t *a, uint32_t *c0,
                                   size_t ct_stride, size_t p_stride, int quirk, int batch, cudaStream_t st);
cudaError_t seb_encrypt_configure(int logn);

// ---- verifier: inverse NTT, decrypt + decode, per-item digests (seb_verify.cu) ----
// digests[b] = sum_i mix64((i << 32) | words[b][i]) mod 2^64 (words_per_item a multiple of 4, rows 16-byte aligned)
cudaError_t seb_launch_digest(const uint32_t *words, size_t words_per_item, size_t items, uint64_t *digests, cudaStream_t st);
// packed wire form: `words` full-form words (a multiple of 16, 16-byte aligned) <-> words * 15 / 16 packed words
cudaError_t seb_launch_pack30(const uint32_t *in, uint32_t *out, size_t words, cudaStream_t st);
cudaError_t seb_launch_unpack30(const uint32_t *in, uint32_t *out, size_t words, cudaStream_t st);
// register-only Keccak-f[1600] and lazy-butterfly loops: the integer-issue ceilings of this device, measured now
// (tw: any valid twiddle table of a 16-coefficient plan; scratch >= sms * 16 * 128 * 8 bytes)
cudaError_t seb_measure_ceilings(int sms, const seb_oct *tw, uint32_t q, void *scratch, double *keccak_per_s, double *bfly_per_s,
                                 cudaStream_t st);
// iroots: per prime n Shoup pairs, iroots[i] = inverse of the forward table's root i; ninvs: n^-1 per prime
cudaError_t seb_verify_configure(int n);
cudaError_t seb_launch_intt(uint32_t *polys, const uint2 *iroots, const uint2 *ninvs, const SebModuli &mods, int n,
                            int np, size_t npolys, int max_ctas, cudaStream_t st);
// work: work_ctas slices of n double2
cudaError_t seb_launch_decrypt_decode(const uint32_t *ct, const uint32_t *s_hat, const uint2 *iroots,
                                      const uint2 *ninvs, const double2 *tw, const uint16_t *index_map,
                                      const SebModuli &mods, int n, int np, int prime, double scale, double2 *work,
                                      int work_ctas, int vlen, float *values_out, size_t batch, cudaStream_t st);
