// seb_api.cu — context, setup-time tables and the `seb_*` C ABI (include/seal_embedded_b200.h, part 2).
//
// Host code here is the GPU-side counterpart of the reference's L0 layer: parameter chains
// (device/lib/parameters.c:129-230), const_ratio (modulus.c:23-56), psi per (n,q)
// (ntt.c:199-291), root tables (ntt.c:40-52), index map (ckks_common.c:32-68) and IFFT twiddles
// (fft.c:27-45) are rebuilt once per context and kept resident in HBM.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cuda.h>  // driver types for the SM partition (green contexts); its functions are fetched at run time

#include "../../include/seal_embedded_b200.h"
#include "seb_kernels.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *seb_last_error(void) { return g_err; }

#define CU(call)                                                                                         \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(SE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,   \
                        __LINE__);                                                                       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// parameter tables (facts of the reference's default parameter sets)
// ---------------------------------------------------------------------------------------------
static const uint32_t k_primes27[3]  = {134012929u, 134111233u, 134176769u};
static const uint32_t k_primes30[13] = {1053818881u, 1054015489u, 1054212097u, 1055260673u, 1056178177u,
                                        1056440321u, 1058209793u, 1060175873u, 1060700161u, 1060765697u,
                                        1061093377u, 1062469633u, 1062535169u};

// parameters.c:191-227: legal (n, nprimes) and the chain used
static bool default_primes(size_t n, size_t np, uint32_t *out)
{
    const uint32_t *src = k_primes30;
    size_t maxp         = 0;
    switch (n)
    {
        case 1024:
        case 2048: src = k_primes27; maxp = 1; break;
        case 4096: maxp = 3; break;
        case 8192: maxp = 6; break;
        case 16384: maxp = 13; break;
        default: return false;
    }
    if (np < 1 || np > maxp) return false;
    for (size_t i = 0; i < np; i++) out[i] = src[i];
    return true;
}

// ntt.c:213-289
static uint32_t default_psi(size_t n, uint32_t q)
{
    static const uint32_t psi4k27[3]  = {7470u, 3856u, 24149u};
    static const uint32_t psi4k30[3]  = {503422u, 16768u, 7305u};
    static const uint32_t psi8k[6]    = {374229u, 123363u, 79941u, 38869u, 162146u, 81884u};
    static const uint32_t psi16k[13]  = {13040u, 507u,   1595u,   68507u,  3073u,   6854u, 44467u,
                                         16117u, 27607u, 222391u, 105471u, 310222u, 2005u};
    if (n == 1024) return q == 134012929u ? 142143u : 0u;
    if (n == 2048) return q == 134012929u ? 85250u : 0u;
    if (n == 4096)
        for (int i = 0; i < 3; i++)
        {
            if (q == k_primes27[i]) return psi4k27[i];
            if (q == k_primes30[i]) return psi4k30[i];
        }
    if (n == 8192)
        for (int i = 0; i < 6; i++)
            if (q == k_primes30[i]) return psi8k[i];
    if (n == 16384)
        for (int i = 0; i < 13; i++)
            if (q == k_primes30[i]) return psi16k[i];
    return 0u;
}

static inline uint32_t mulmod(uint32_t a, uint32_t b, uint32_t q) { return (uint32_t)(((uint64_t)a * b) % q); }
static uint32_t powmod(uint32_t a, uint64_t e, uint32_t q)
{
    uint32_t r = 1;
    while (e)
    {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
static inline uint32_t shoup(uint32_t w, uint32_t q) { return (uint32_t)(((uint64_t)w << 32) / q); }
static inline size_t bitrev(size_t x, int bits)
{
    size_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// deterministic Miller-Rabin for q < 2^32 (bases 2, 7, 61)
static bool is_prime32(uint32_t q)
{
    if (q < 2) return false;
    for (uint32_t p : {2u, 3u, 5u, 7u, 11u, 13u, 17u, 19u, 23u, 29u, 31u, 37u})
        if (q % p == 0) return q == p;
    uint32_t d = q - 1;
    int r      = 0;
    while (!(d & 1)) d >>= 1, r++;
    for (uint32_t a : {2u, 7u, 61u})
    {
        uint32_t x = powmod(a % q, d, q);
        if (x == 1 || x == q - 1) continue;
        bool comp = true;
        for (int i = 1; i < r && comp; i++)
        {
            x = mulmod(x, x, q);
            if (x == q - 1) comp = false;
        }
        if (comp) return false;
    }
    return true;
}

// Smallest primitive 2n-th root of unity mod q (0 when q - 1 is not a multiple of 2n or no root is found):
// what SEAL's try_minimal_primitive_root computes and — checked for all 27 pairs by the CPU suite — what
// the reference tabulates per (n, q) in get_ntt_root (ntt.c:199-291).  Lets custom prime chains
// (se_setup_custom) and the NTT sweep's "1..8 primes at every degree" work without a table.
extern "C" uint32_t seb_minimal_psi(size_t n, uint32_t q)
{
    if (n == 0 || q < 3 || (q - 1) % (2 * n) != 0) return 0;
    const uint64_t e = (q - 1) / (2 * n);
    uint32_t r       = 0;
    for (uint32_t g = 2; g < 1000 && g < q; g++)
    {
        const uint32_t c = powmod(g, e, q);
        if (powmod(c, n, q) == q - 1)  // order exactly 2n (n is a power of two)
        {
            r = c;
            break;
        }
    }
    if (!r) return 0;
    // the primitive 2n-th roots are the odd powers of r
    const uint32_t r2 = mulmod(r, r, q);
    uint32_t best = r, cur = r;
    for (size_t i = 1; i < n; i++)
    {
        cur = mulmod(cur, r2, q);
        if (cur < best) best = cur;
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct Scratch  // per in-flight chunk
{
    size_t cap       = 0;  // ciphertexts
    int64_t *pt      = nullptr;
    int8_t *e        = nullptr;
    uint8_t *u       = nullptr;
    uint32_t *ctr    = nullptr;
    uint32_t *ctr_a  = nullptr;
    int *fail        = nullptr;
    uint32_t *mag    = nullptr;  // max |plaintext coefficient| per item, clipped to 32 bits
    uint16_t *rej_idx = nullptr;  // [cap][n/8] uniform sampler: indices of rejected words (symmetric mode)
    uint32_t *rej_cnt = nullptr;  // [cap]
    uint32_t *a_buf   = nullptr;  // [a_cap][np][n] `a` of the seed-compressed symmetric path (allocated on first use)
    size_t a_cap      = 0;
    // lone-call speculation (seb_launch_uniform_chain_spec), allocated on first use for SEB_SPEC_MAX_BATCH items
    uint32_t *cand_rows = nullptr;
    uint16_t *cand_list = nullptr;
    uint32_t *cand_cnt  = nullptr;
    size_t cand_cap     = 0;  // ciphertexts the candidate buffers hold
    // host-API staging
    size_t io_cap    = 0;
    size_t io_out_words = 0;  // words of d_out / h_out per ciphertext
    float *d_values  = nullptr;
    uint8_t *d_seeds = nullptr, *d_sseeds = nullptr;
    uint32_t *d_out  = nullptr;
    uint32_t *d_pack = nullptr;  // packed wire form of d_out (allocated on first use)
    size_t pack_words = 0;
    int *h_fail      = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done    = nullptr;
};

struct seb_ctx
{
    int device = 0;
    size_t n = 0, np = 0;
    int logn = 0;
    bool asym = false;
    double scale = 0;
    uint32_t primes[SEB_MAX_PRIMES] = {0};
    uint32_t psis[SEB_MAX_PRIMES]   = {0};
    SebModuli mods;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // SM partition for the symmetric path's latency-bound regime (seb_partition_*): two green contexts, the sampler
    // chain's and the side work's, each with a stream; created on first use, `part_state` -1 = unavailable
    int part_state = 0;
    CUgreenCtx part_ctx[2] = {nullptr, nullptr};
    cudaStream_t part_stream[2] = {nullptr, nullptr};
    cudaEvent_t part_ev[3] = {nullptr, nullptr, nullptr};
    int part_sms[2] = {0, 0};
    uint32_t rej_cap = 0;  // capacity of the uniform sampler's per-ciphertext reject lists (n/8)
    // resident tables
    seb_oct *d_roots    = nullptr;  // [np][seb_table_octs(logn)]: per-pass twiddle tables, 16-coefficient plan (asymmetric kernel)
    seb_oct *d_roots1   = nullptr;  // the same roots in the plan of the one-polynomial kernels (= d_roots below n = 8192)
    seb_oct *d_roots_sym = nullptr;  // ... in the form the symmetric kernel reads (= d_roots1 except at n = 16384: split form)
    int key1 = 0;                   // seb_ntt_key1(logn)
    double2 *d_tw       = nullptr;  // [n] natural order + [7][n/8] pass-0 copies (seb_encode.cuh)
    uint16_t *d_src_map = nullptr;  // [n]
    seb_oct *d_pk0 = nullptr, *d_pk1 = nullptr;  // [np][n/4] Shoup pairs, epilogue order
    seb_oct *d_ntt_s = nullptr;                  // [np][n/4] Shoup pairs, epilogue order
    // verifier (seb_verify.cu)
    uint32_t *d_ntt_s_nat = nullptr;  // [np][n] ntt(s), reference order
    uint2 *d_iroots       = nullptr;  // [np][n] inverse roots, Shoup pairs
    uint2 *d_ninv         = nullptr;  // [np] n^-1
    uint16_t *d_index_map = nullptr;  // [n] ckks_calc_index_map
    double2 *d_work       = nullptr;  // [verify_ctas][n] FFT scratch
    int sms = 0, verify_ctas = 0;
    bool have_pk = false, have_sk = false;
    SebSpecPlan spec_plan;            // counter windows for the lone-call uniform chain (symmetric, np >= 2)
    uint32_t *d_spec_misses = nullptr;  // how often a true counter fell outside its window (diagnostic)
    // The asynchronous *_device entry points own `dev` (used on c->stream); the host-pointer pipeline owns
    // `hslot` (two chunks in flight, each on its own stream).  They share nothing that is written, so a host call
    // may follow an un-synchronised device call (and the other way round), and seb_encode_failures() always
    // reports the last *_device call.
    Scratch dev;
    Scratch hslot[2];
    SebKnobs knobs;  // test / A-B switches: environment read ONCE in seb_create, then seb_set_option
    size_t last_batch = 0;
    uint64_t launches = 0;
    // per-kernel CUDA-event timing of the *_device full-path calls (seb_profile_begin/end)
    std::vector<cudaEvent_t> prof_ev;
    int prof_max = 0, prof_step = 0;
    bool prof_on = false;
};

#define SEB_PROF_SEGMENTS 4
static inline void prof_mark(seb_ctx *c, cudaStream_t st, int k)
{
    if (c->prof_on && st == c->stream && c->prof_step < c->prof_max)
        cudaEventRecord(c->prof_ev[(size_t)c->prof_step * (SEB_PROF_SEGMENTS + 1) + k], st);
}
static inline void prof_next(seb_ctx *c, cudaStream_t st)
{
    if (c->prof_on && st == c->stream && c->prof_step < c->prof_max) c->prof_step++;
}

// cudaFree of a buffer that held secrets (plaintexts, error polynomials, seeds, key material): wiped first.
// The memset is queued on the legacy default stream and cudaFree synchronises the device, so the order holds.
static void seb_wipe(void *p, size_t bytes)
{
    volatile unsigned char *v = static_cast<volatile unsigned char *>(p);
    for (size_t i = 0; i < bytes; i++) v[i] = 0;
}

static void wipe_free(void *p, size_t bytes)
{
    if (!p) return;
    if (bytes) cudaMemset(p, 0, bytes);
    cudaFree(p);
}

// every cudaMalloc of a group succeeds or none is kept: a failed grow leaves the old buffers in place
struct AllocGroup
{
    std::vector<void **> slots;
    std::vector<void *> fresh;
    cudaError_t err = cudaSuccess;
    void want(void **slot, size_t bytes)
    {
        if (err != cudaSuccess) return;
        void *p = nullptr;
        err     = cudaMalloc(&p, bytes ? bytes : 1);
        if (err != cudaSuccess) return;
        slots.push_back(slot);
        fresh.push_back(p);
    }
    // on failure: release what was obtained, report; on success the caller frees the old buffers and commits
    bool failed()
    {
        if (err == cudaSuccess) return false;
        for (void *p : fresh) cudaFree(p);
        fresh.clear();
        cudaGetLastError();
        return true;
    }
    void commit()
    {
        for (size_t i = 0; i < slots.size(); i++) *slots[i] = fresh[i];
    }
};

static int ensure_scratch(seb_ctx *c, Scratch &s, size_t batch)
{
    if (batch <= s.cap) return 0;
    const size_t n = c->n, rc = (size_t)(c->rej_cap ? c->rej_cap : 1);
    Scratch f;  // the new buffers
    AllocGroup g;
    g.want((void **)&f.pt, batch * n * sizeof(int64_t));
    g.want((void **)&f.e, batch * 2 * n);
    g.want((void **)&f.u, batch * (n / 4));
    g.want((void **)&f.ctr, batch * sizeof(uint32_t));
    g.want((void **)&f.ctr_a, batch * sizeof(uint32_t));
    g.want((void **)&f.fail, batch * sizeof(int));
    g.want((void **)&f.mag, batch * sizeof(uint32_t));
    if (!c->asym)
    {
        g.want((void **)&f.rej_idx, batch * rc * sizeof(uint16_t));
        g.want((void **)&f.rej_cnt, batch * sizeof(uint32_t));
    }
    if (g.failed())
        return fail(SE_ERR_NO_MEMORY, "scratch for %zu ciphertexts: %s (the context keeps its %zu-item scratch)", batch,
                    cudaGetErrorString(g.err), s.cap);
    g.commit();
    wipe_free(s.pt, s.cap * n * sizeof(int64_t));
    wipe_free(s.e, s.cap * 2 * n);
    wipe_free(s.u, s.cap * (n / 4));
    cudaFree(s.ctr);
    cudaFree(s.ctr_a);
    cudaFree(s.fail);
    cudaFree(s.mag);
    cudaFree(s.rej_idx);
    cudaFree(s.rej_cnt);
    s.pt = f.pt, s.e = f.e, s.u = f.u, s.ctr = f.ctr, s.ctr_a = f.ctr_a, s.fail = f.fail, s.mag = f.mag;
    s.rej_idx = f.rej_idx, s.rej_cnt = f.rej_cnt;
    s.cap     = batch;
    return 0;
}

static void free_scratch(Scratch &s, size_t n)
{
    wipe_free(s.pt, s.cap * n * sizeof(int64_t));
    wipe_free(s.e, s.cap * 2 * n);
    wipe_free(s.u, s.cap * (n / 4));
    cudaFree(s.ctr);
    cudaFree(s.ctr_a);
    cudaFree(s.fail);
    cudaFree(s.mag);
    cudaFree(s.rej_idx);
    cudaFree(s.rej_cnt);
    cudaFree(s.a_buf);
    cudaFree(s.cand_rows);
    cudaFree(s.cand_list);
    cudaFree(s.cand_cnt);
    wipe_free(s.d_values, s.io_cap * (n / 2) * sizeof(float));
    wipe_free(s.d_seeds, s.io_cap * SEB_SEED_BYTES);
    cudaFree(s.d_sseeds);
    cudaFree(s.d_out);
    cudaFree(s.d_pack);
    cudaFreeHost(s.h_fail);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
    s = Scratch();
}

static int build_tables(seb_ctx *c)
{
    const size_t n = c->n;
    // NTT roots: bit-reversed powers of psi in Shoup form (ntt.c:40-52, uintmodarith.h:293-297),
    // re-ordered per pass into the layout the kernels read with coalesced 256-bit loads
    c->key1 = seb_ntt_key1(c->logn);
    const size_t octs = seb_table_octs(c->logn), octs1 = seb_table_octs(c->key1);
    const size_t octs_sym = seb_table_octs_sym(c->logn);
    std::vector<seb_oct> tabs_sym(seb_sym_split(c->logn) ? c->np * octs_sym : 0);
    if (!tabs_sym.empty()) memset(tabs_sym.data(), 0, tabs_sym.size() * sizeof(seb_oct));
    std::vector<seb_oct> tabs(c->np * octs), tabs1(c->key1 != c->logn ? c->np * octs1 : 0);
    memset(tabs.data(), 0, tabs.size() * sizeof(seb_oct));
    if (!tabs1.empty()) memset(tabs1.data(), 0, tabs1.size() * sizeof(seb_oct));
    std::vector<uint2> roots(n);
    for (size_t p = 0; p < c->np; p++)
    {
        const uint32_t q = c->primes[p], psi = c->psis[p];
        uint32_t pw      = 1;
        for (size_t i = 0; i < n; i++)
        {
            roots[bitrev(i, c->logn)] = make_uint2(pw, shoup(pw, q));
            pw                        = mulmod(pw, psi, q);
        }
        seb_host_build_tw(c->logn, roots.data(), tabs.data() + p * octs);
        if (!tabs1.empty()) seb_host_build_tw(c->key1, roots.data(), tabs1.data() + p * octs1);
        if (!tabs_sym.empty()) seb_host_build_tw_sym(c->logn, roots.data(), tabs_sym.data() + p * octs_sym);
    }
    CU(cudaMalloc(&c->d_roots, tabs.size() * sizeof(seb_oct)));
    CU(cudaMemcpy(c->d_roots, tabs.data(), tabs.size() * sizeof(seb_oct), cudaMemcpyHostToDevice));
    if (tabs1.empty())
        c->d_roots1 = c->d_roots;
    else
    {
        CU(cudaMalloc(&c->d_roots1, tabs1.size() * sizeof(seb_oct)));
        CU(cudaMemcpy(c->d_roots1, tabs1.data(), tabs1.size() * sizeof(seb_oct), cudaMemcpyHostToDevice));
    }
    if (tabs_sym.empty())
        c->d_roots_sym = c->d_roots1;
    else
    {
        CU(cudaMalloc(&c->d_roots_sym, tabs_sym.size() * sizeof(seb_oct)));
        CU(cudaMemcpy(c->d_roots_sym, tabs_sym.data(), tabs_sym.size() * sizeof(seb_oct), cudaMemcpyHostToDevice));
    }

    // inverse roots for the verifier's INTT: iroots[bitrev(i)] = psi^-i, and n^-1
    {
        std::vector<uint2> inv(c->np * n), ninv(c->np);
        for (size_t p = 0; p < c->np; p++)
        {
            const uint32_t q = c->primes[p], ipsi = powmod(c->psis[p], 2 * n - 1, q);
            uint32_t pw      = 1;
            for (size_t i = 0; i < n; i++)
            {
                inv[p * n + bitrev(i, c->logn)] = make_uint2(pw, shoup(pw, q));
                pw                              = mulmod(pw, ipsi, q);
            }
            const uint32_t ni = powmod((uint32_t)(n % q), (uint64_t)q - 2, q);
            ninv[p]           = make_uint2(ni, shoup(ni, q));
        }
        CU(cudaMalloc(&c->d_iroots, inv.size() * sizeof(uint2)));
        CU(cudaMemcpy(c->d_iroots, inv.data(), inv.size() * sizeof(uint2), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c->d_ninv, ninv.size() * sizeof(uint2)));
        CU(cudaMemcpy(c->d_ninv, ninv.data(), ninv.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    }

    // IFFT twiddles from the host libm, same expression as fft.c:27-45 (+ conj at fft.c:129)
    std::vector<double2> tw(seb_enc_tw_entries(n));
    const size_t m = 2 * n;
    tw[0]          = make_double2(1.0, 0.0);
    for (size_t i = 1; i < n; i++)
    {
        const size_t k     = bitrev(i, c->logn) & (m - 1);
        const double angle = 2 * M_PI * (double)k / (double)m;
        tw[i]              = make_double2(cos(angle), -sin(angle));
    }
    seb_host_build_enc_tw0(n, tw.data());  // pass-0 roots again, in the order the encode kernel's lanes read them
    CU(cudaMalloc(&c->d_tw, tw.size() * sizeof(double2)));
    CU(cudaMemcpy(c->d_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));

    // index map (ckks_common.c:32-68) inverted: position -> slot
    std::vector<uint16_t> src(n), fwd(n);
    uint64_t pos = 1;
    for (size_t i = 0; i < n / 2; i++)
    {
        const size_t a            = (size_t)((pos - 1) / 2);
        const size_t b            = n - 1 - a;
        src[bitrev(a, c->logn)]   = (uint16_t)i;
        src[bitrev(b, c->logn)]   = (uint16_t)i;
        fwd[i]                    = (uint16_t)bitrev(a, c->logn);
        fwd[i + n / 2]            = (uint16_t)bitrev(b, c->logn);
        pos                       = (pos * 3) & (m - 1);
    }
    CU(cudaMalloc(&c->d_src_map, n * sizeof(uint16_t)));
    CU(cudaMemcpy(c->d_src_map, src.data(), n * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c->d_index_map, n * sizeof(uint16_t)));
    CU(cudaMemcpy(c->d_index_map, fwd.data(), n * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" seb_ctx *seb_create(size_t n, size_t nprimes, const uint32_t *primes, const uint32_t *psis,
                               double scale, int asym, int device)
{
    int logn = 0;
    while (((size_t)1 << logn) < n) logn++;
    if (((size_t)1 << logn) != n || logn < 10 || logn > 14 || nprimes < 1 || nprimes > SEB_MAX_PRIMES)
    {
        fail(SE_ERR_INVALD_ARGUMENT, "unsupported parameters n=%zu nprimes=%zu", n, nprimes);
        return nullptr;
    }
    seb_ctx *c = new seb_ctx();
    c->n       = n;
    c->np      = nprimes;
    c->logn    = logn;
    c->asym    = asym != 0;
    if (primes)
    {
        for (size_t i = 0; i < nprimes; i++)
        {
            c->primes[i] = primes[i];
            c->psis[i]   = psis ? psis[i] : default_psi(n, primes[i]);
            if (!c->psis[i]) c->psis[i] = seb_minimal_psi(n, primes[i]);
        }
    }
    else
    {
        if (!default_primes(n, nprimes, c->primes))
        {
            fail(SE_ERR_INVALD_ARGUMENT, "no default parameter set for n=%zu nprimes=%zu", n, nprimes);
            delete c;
            return nullptr;
        }
        for (size_t i = 0; i < nprimes; i++) c->psis[i] = default_psi(n, c->primes[i]);
    }
    for (size_t i = 0; i < nprimes; i++)
    {
        const uint32_t q = c->primes[i], psi = c->psis[i];
        // q < 2^30 keeps lazy values below 2^32; psi^n = -1 makes psi a primitive 2n-th root
        // (a composite q can satisfy psi^n = -1 as well: caller-supplied moduli are tested for primality, so that the
        // NTT is invertible and the ciphertexts decrypt)
        if (q < 3 || q >= (1u << 30) || (q - 1) % (2 * n) != 0 || !is_prime32(q) || psi == 0 || psi >= q ||
            powmod(psi, n, q) != q - 1)
        {
            fail(SE_ERR_INVALD_ARGUMENT, "modulus %u / root %u unusable for n=%zu (need a prime q < 2^30, q = 1 mod 2n, "
                                         "psi a primitive 2n-th root of unity)", q, psi, n);
            delete c;
            return nullptr;
        }
        for (size_t j = 0; j < i; j++)
            if (c->primes[j] == q)
            {
                fail(SE_ERR_INVALD_ARGUMENT, "modulus %u appears twice in the chain", q);
                delete c;
                return nullptr;
            }
        const uint64_t ratio = (uint64_t)(((unsigned __int128)1 << 64) / q);
        c->mods.m[i]         = SebModulus{q, 2 * q, (uint32_t)ratio, (uint32_t)(ratio >> 32)};
    }
    // parameters.c:197-225: the default scale is tied to the degree
    c->scale = scale > 0 ? scale : (n == 1024 ? 1048576.0 : 33554432.0);
    // 1.9 % of the words are rejected under a 30-bit prime: n/8 entries is > 40 standard deviations of
    // head-room; overflowing lists fall back to a scan (SEB_UNIFORM_LIST_CAP lets tests force that)
    c->rej_cap = (uint32_t)(n / 8);
    if (const char *v = getenv("SEB_UNIFORM_LIST_CAP"))
        if (*v && (uint32_t)atoi(v) < c->rej_cap) c->rej_cap = (uint32_t)atoi(v);
    // the remaining A/B switches: read here, never on the call path (seb_set_option changes them at run time)
    auto env_int = [](const char *name, long dflt) -> long {
        const char *v = getenv(name);
        return (v && *v) ? atol(v) : dflt;
    };
    c->knobs.uniform_coop     = (int)env_int("SEB_UNIFORM_COOP", -1);
    c->knobs.uniform_fix_wide = (int)env_int("SEB_UNIFORM_FIX_WIDE", -1);
    c->knobs.uniform_spec     = (int)env_int("SEB_UNIFORM_SPEC", -1);
    c->knobs.uniform_pair     = (int)env_int("SEB_UNIFORM_PAIR", -1);
    c->knobs.host_chunk       = env_int("SEB_HOST_CHUNK", 0);

    memset(&c->spec_plan, 0, sizeof c->spec_plan);
    if (!c->asym && nprimes >= 2)
    {
        // 5 sigma: a true counter falls outside with probability 6e-7 per prime (and is then squeezed on the spot);
        // 6 sigma measured 13 % slower at n = 16384, 4 sigma no faster.  SEB_UNIFORM_SPEC_SIGMAS overrides (tests force
        // the fallback with 0).
        double sigmas = 5.0;
        if (const char *v = getenv("SEB_UNIFORM_SPEC_SIGMAS"))
            if (*v) sigmas = atof(v);
        seb_uniform_spec_plan((int)n, c->mods, (int)nprimes, sigmas, &c->spec_plan);
    }

    auto bail = [&](const char *what, cudaError_t e) -> seb_ctx * {
        fail(SE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        seb_destroy(c);
        return nullptr;
    };
    cudaError_t e;
    if (device >= 0 && (e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaGetDevice(&c->device)) != cudaSuccess) return bail("cudaGetDevice", e);
    // a blocking stream: it orders itself after work the caller queued on the legacy default stream
    // (cudaMemset / cudaMemcpy / a framework's default-stream fills of the buffers handed to us)
    if ((e = cudaStreamCreate(&c->own_stream)) != cudaSuccess)
        return bail("cudaStreamCreate", e);
    c->stream = c->own_stream;
    if ((e = seb_encode_configure(logn)) != cudaSuccess) return bail("encode kernel attributes", e);
    if ((e = seb_encrypt_configure(logn)) != cudaSuccess) return bail("encrypt kernel attributes", e);
    if ((e = seb_verify_configure((int)n)) != cudaSuccess) return bail("verifier kernel attributes", e);
    if ((e = cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, c->device)) != cudaSuccess)
        return bail("cudaDeviceGetAttribute", e);
    c->verify_ctas = c->sms * (n <= 4096 ? 4 : n <= 8192 ? 2 : 1);
    c->knobs.sms   = c->sms;
    // the second stream of the mixed bulk squeeze (seb_launch_uniform); without it the squeeze is one kernel
    if (!asym)
    {
        if (cudaStreamCreateWithFlags(&c->knobs.aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->knobs.aux_ev[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->knobs.aux_ev[1], cudaEventDisableTiming) != cudaSuccess)
        {
            cudaGetLastError();
            if (c->knobs.aux_stream) cudaStreamDestroy(c->knobs.aux_stream);
            c->knobs.aux_stream = nullptr;
        }
    }
    if (build_tables(c) != 0)
    {
        seb_destroy(c);
        return nullptr;
    }
    return c;
}

// ---------------------------------------------------------------------------------------------
// SM partition (CUDA green contexts)
// ---------------------------------------------------------------------------------------------
// A symmetric batch of ~16k items at a large degree spends two thirds of its time in the uniform sampler's bulk squeeze:
// one sequential sponge per (ciphertext, prime), ONE warp per SM sub-partition, and nothing else can share those
// sub-partitions without stretching the chain (profiles/README.md: every co-residency experiment lost).  But 512 such
// warps occupy 128 SMs; the other SMs idle for the whole chain.  The device is therefore split into two green contexts -
// SEB_PART_SIDE_SMS SMs on the side, the rest for the chain - and while the chain runs on the big partition, the encode and
// the CBD sampler of a share of the batch run on the side partition, where they disturb nobody.  The driver entry points
// are fetched with cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda; if anything is
// missing the path simply stays serial.
#define SEB_PART_SIDE_SMS 16
typedef CUresult (*pfn_cuDeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType);
typedef CUresult (*pfn_cuDevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *,
                                                    unsigned int, unsigned int);
typedef CUresult (*pfn_cuDevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
typedef CUresult (*pfn_cuGreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
typedef CUresult (*pfn_cuGreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int);
typedef CUresult (*pfn_cuGreenCtxDestroy)(CUgreenCtx);
typedef CUresult (*pfn_cuDeviceGet)(CUdevice *, int);

template <class F>
static bool seb_driver_fn(const char *name, F *fn)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
    {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

static void seb_partition_destroy(seb_ctx *c)
{
    for (cudaEvent_t &e : c->part_ev)
        if (e) cudaEventDestroy(e), e = nullptr;
    for (cudaStream_t &s : c->part_stream)
        if (s) cudaStreamDestroy(s), s = nullptr;
    pfn_cuGreenCtxDestroy destroy = nullptr;
    if ((c->part_ctx[0] || c->part_ctx[1]) && seb_driver_fn("cuGreenCtxDestroy", &destroy))
        for (CUgreenCtx &g : c->part_ctx)
            if (g) destroy(g), g = nullptr;
    c->part_state = 0;
}

// true when the partition exists (created on the first call); never fails the caller
static bool seb_partition_ready(seb_ctx *c)
{
    if (c->part_state) return c->part_state > 0;
    c->part_state = -1;
    pfn_cuDeviceGet dev_get = nullptr;
    pfn_cuDeviceGetDevResource get_res = nullptr;
    pfn_cuDevSmResourceSplitByCount split = nullptr;
    pfn_cuDevResourceGenerateDesc gen_desc = nullptr;
    pfn_cuGreenCtxCreate ctx_create = nullptr;
    pfn_cuGreenCtxStreamCreate stream_create = nullptr;
    if (!seb_driver_fn("cuDeviceGet", &dev_get) || !seb_driver_fn("cuDeviceGetDevResource", &get_res) ||
        !seb_driver_fn("cuDevSmResourceSplitByCount", &split) || !seb_driver_fn("cuDevResourceGenerateDesc", &gen_desc) ||
        !seb_driver_fn("cuGreenCtxCreate", &ctx_create) || !seb_driver_fn("cuGreenCtxStreamCreate", &stream_create))
        return false;
    CUdevice dev;
    CUdevResource all, side, rest;
    unsigned int groups = 1;
    if (dev_get(&dev, c->device) != CUDA_SUCCESS || get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    if (split(&side, &groups, &all, &rest, 0, SEB_PART_SIDE_SMS) != CUDA_SUCCESS || groups != 1) return false;
    if (side.sm.smCount == 0 || rest.sm.smCount == 0) return false;
    CUdevResource *parts[2] = {&rest, &side};  // [0]: the sampler chain's partition, [1]: the side work's
    for (int i = 0; i < 2; i++)
    {
        CUdevResourceDesc desc;
        CUstream st = nullptr;
        if (gen_desc(&desc, parts[i], 1) != CUDA_SUCCESS ||
            ctx_create(&c->part_ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS ||
            stream_create(&st, c->part_ctx[i], CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS)
        {
            seb_partition_destroy(c);
            c->part_state = -1;
            return false;
        }
        c->part_stream[i] = reinterpret_cast<cudaStream_t>(st);
        c->part_sms[i]    = (int)parts[i]->sm.smCount;
    }
    for (cudaEvent_t &e : c->part_ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
        {
            seb_partition_destroy(c);
            c->part_state = -1;
            return false;
        }
    c->part_state = 1;
    return true;
}

extern "C" void seb_destroy(seb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_scratch(c->dev, c->n);
    for (auto &s : c->hslot) free_scratch(s, c->n);
    if (c->d_roots_sym != c->d_roots1) cudaFree(c->d_roots_sym);
    if (c->d_roots1 != c->d_roots) cudaFree(c->d_roots1);
    cudaFree(c->d_roots);
    cudaFree(c->d_tw);
    cudaFree(c->d_src_map);
    cudaFree(c->d_spec_misses);
    cudaFree(c->d_pk0);
    cudaFree(c->d_pk1);
    cudaFree(c->d_ntt_s);
    cudaFree(c->d_ntt_s_nat);
    cudaFree(c->d_iroots);
    cudaFree(c->d_ninv);
    cudaFree(c->d_index_map);
    cudaFree(c->d_work);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    seb_partition_destroy(c);
    for (cudaEvent_t e : c->knobs.aux_ev)
        if (e) cudaEventDestroy(e);
    if (c->knobs.aux_stream) cudaStreamDestroy(c->knobs.aux_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

// How many (ciphertext, prime) squeezes of the lone-call path found their true counter outside the speculated
// window and were redone on the spot (diagnostic; synchronises the stream).  Negative on error.
extern "C" long seb_uniform_spec_misses(seb_ctx *c)
{
    if (!c) return fail(SE_ERR_INVALD_ARGUMENT, "null context");
    if (!c->d_spec_misses) return 0;
    uint32_t v = 0;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(&v, c->d_spec_misses, sizeof v, cudaMemcpyDeviceToHost));
    return (long)v;
}

extern "C" int seb_set_stream(seb_ctx *c, void *cuda_stream)
{
    if (!c) return fail(SE_ERR_INVALD_ARGUMENT, "null context");
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return 0;
}

// Run-time access to the A/B switches (SebKnobs).  value < 0 restores the automatic choice.
extern "C" int seb_set_option(seb_ctx *c, const char *name, long value)
{
    if (!c || !name) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    const int v = value < 0 ? -1 : (int)value;
    if (!strcmp(name, "uniform_coop")) c->knobs.uniform_coop = v;
    else if (!strcmp(name, "uniform_fix_wide")) c->knobs.uniform_fix_wide = v;
    else if (!strcmp(name, "uniform_spec")) c->knobs.uniform_spec = v;
    else if (!strcmp(name, "uniform_pair")) c->knobs.uniform_pair = v;
    else if (!strcmp(name, "uniform_fix_lanes")) c->knobs.uniform_fix_lanes = v;
    else if (!strcmp(name, "uniform_fix_stream")) c->knobs.uniform_fix_stream = v;
    else if (!strcmp(name, "uniform_mix")) c->knobs.uniform_mix = v;
    else if (!strcmp(name, "sym_partition")) c->knobs.sym_partition = v;
    else if (!strcmp(name, "sym_side_percent")) c->knobs.sym_side_percent = value < 0 ? -1 : value > 90 ? 90 : (int)value;
    else if (!strcmp(name, "host_chunk")) c->knobs.host_chunk = value < 0 ? 0 : value;
    else return fail(SE_ERR_INVALD_ARGUMENT, "unknown option '%s'", name);
    return 0;
}

extern "C" size_t seb_degree(const seb_ctx *c) { return c ? c->n : 0; }
extern "C" size_t seb_nprimes(const seb_ctx *c) { return c ? c->np : 0; }
extern "C" double seb_scale(const seb_ctx *c) { return c ? c->scale : 0; }
extern "C" uint32_t seb_prime(const seb_ctx *c, size_t i) { return (c && i < c->np) ? c->primes[i] : 0; }
extern "C" uint64_t seb_launch_count(const seb_ctx *c) { return c ? c->launches : 0; }

// key: the plan whose epilogue order the table is laid out in (logn for pk0/pk1); < 0: the symmetric kernel's order (ntt(s))
static int upload_shoup(seb_ctx *c, const uint32_t *host, seb_oct **dst, int key)
{
    std::vector<uint2> nat(c->n);
    std::vector<seb_oct> tab(c->np * (c->n / 4));
    for (size_t p = 0; p < c->np; p++)
    {
        for (size_t i = 0; i < c->n; i++)
        {
            const uint32_t w = host[p * c->n + i];
            if (w >= c->primes[p]) return fail(SE_ERR_INVALD_ARGUMENT, "key coefficient %u >= modulus", w);
            nat[i] = make_uint2(w, shoup(w, c->primes[p]));
        }
        if (key < 0)  // the symmetric kernel's order
            seb_host_build_epi_sym(c->logn, nat.data(), tab.data() + p * (c->n / 4));
        else
            seb_host_build_epi(key, nat.data(), tab.data() + p * (c->n / 4));
    }
    if (!*dst) CU(cudaMalloc(dst, tab.size() * sizeof(seb_oct)));
    CU(cudaMemcpy(*dst, tab.data(), tab.size() * sizeof(seb_oct), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int seb_set_public_key(seb_ctx *c, const uint32_t *pk0, const uint32_t *pk1)
{
    if (!c || !pk0 || !pk1) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    CU(cudaSetDevice(c->device));
    int r = upload_shoup(c, pk0, &c->d_pk0, c->logn);
    if (r) return r;
    r = upload_shoup(c, pk1, &c->d_pk1, c->logn);
    if (r) return r;
    c->have_pk = true;
    return 0;
}

// ntt(s) is the same for every ciphertext, so it is computed once here instead of per
// encryption as the reference does (ckks_sym.c:254-268).
extern "C" int seb_set_secret_key(seb_ctx *c, const uint8_t *sk)
{
    if (!c || !sk) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    CU(cudaSetDevice(c->device));
    const size_t n = c->n;
    std::vector<uint32_t> s(c->np * n);
    // the expanded key is secret: the host copy is wiped on every way out, the device copy before it is freed
    struct Wipe
    {
        std::vector<uint32_t> &v;
        ~Wipe() { seb_wipe(v.data(), v.size() * sizeof(uint32_t)); }
    } wipe{s};
    for (size_t p = 0; p < c->np; p++)
        for (size_t i = 0; i < n; i++)
        {
            const uint32_t t = (sk[i / 4] >> (6 - 2 * (i % 4))) & 3u;  // sample.c:89-96
            if (t > 2) return fail(SE_ERR_INVALD_ARGUMENT, "secret key field %zu is not ternary", i);
            s[p * n + i] = t == 0 ? c->primes[p] - 1 : t - 1;          // sample.c:98-116
        }
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, s.size() * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpy(d, s.data(), s.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        e = seb_launch_ntt(c->logn, d, c->d_roots1, c->mods, (int)c->np, c->np, c->stream);
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(s.data(), d, s.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    wipe_free(d, s.size() * sizeof(uint32_t));
    if (e != cudaSuccess) return fail(SE_ERR_CUDA, "ntt(s): %s", cudaGetErrorString(e));
    int r = upload_shoup(c, s.data(), &c->d_ntt_s, -1);
    if (r) return r;
    if (!c->d_ntt_s_nat) CU(cudaMalloc(&c->d_ntt_s_nat, s.size() * sizeof(uint32_t)));
    CU(cudaMemcpy(c->d_ntt_s_nat, s.data(), s.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    c->have_sk = true;
    return 0;
}

// ckks_setup_s with parms->sample_s set (device/lib/ckks_sym.c:162-173): s = sample_small_poly_ternary_prng_96 from
// PRNG(seed) at counter 0 (sample.c:218-242) — the same sampler as the asymmetric path's u, so the same kernel.
// sk_out receives the n/4 packed bytes (sk_<n>.dat format); the key is installed as by seb_set_secret_key.
extern "C" int seb_gen_secret_key(seb_ctx *c, const uint8_t *seed, uint8_t *sk_out)
{
    if (!c || !seed || !sk_out) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    CU(cudaSetDevice(c->device));
    const size_t n = c->n;
    uint8_t *d     = nullptr;  // [64 seed][n/4 key][4 counter]
    const size_t bytes = SEB_SEED_BYTES + n / 4 + sizeof(uint32_t);
    CU(cudaMalloc(&d, bytes));
    cudaError_t e = cudaMemcpyAsync(d, seed, SEB_SEED_BYTES, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess)
    {
        seb_launch_sample_ternary(d, d + SEB_SEED_BYTES, reinterpret_cast<uint32_t *>(d + SEB_SEED_BYTES + n / 4), (int)n, 1,
                                  c->stream);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(sk_out, d + SEB_SEED_BYTES, n / 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    wipe_free(d, bytes);
    if (e != cudaSuccess) return fail(SE_ERR_CUDA, "seb_gen_secret_key: %s", cudaGetErrorString(e));
    return seb_set_secret_key(c, sk_out);
}

// gen_pk (device/lib/ckks_asym.c:159-171) for every prime: a symmetric encryption of zero whose error
// polynomial ep is CBD(PRNG(ep_seed)) (shared by all primes) and whose `a` for prime p comes from a
// fresh PRNG seeded with a_seed_base where byte 0 is replaced by p:
//   pk0_p = -(a_p (.) ntt(s)) + ntt(ep),  pk1_p = a_p.
// Built from the same kernels as the symmetric path (plaintext 0).  The key is returned to the host
// and, on an asymmetric context, installed.
extern "C" int seb_gen_public_key(seb_ctx *c, const uint8_t *sk_packed, const uint8_t *ep_seed,
                                  const uint8_t *a_seed_base, uint32_t *pk0, uint32_t *pk1)
{
    if (!c || !sk_packed || !ep_seed || !a_seed_base || !pk0 || !pk1) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    int r = seb_set_secret_key(c, sk_packed);
    if (r) return r;
    const size_t n = c->n, np = c->np;
    std::vector<uint8_t> seeds((np + 1) * SEB_SEED_BYTES);
    memcpy(seeds.data(), ep_seed, SEB_SEED_BYTES);
    for (size_t p = 0; p < np; p++)
    {
        memcpy(seeds.data() + (p + 1) * SEB_SEED_BYTES, a_seed_base, SEB_SEED_BYTES);
        seeds[(p + 1) * SEB_SEED_BYTES] = (uint8_t)p;
    }
    uint8_t *d_seeds = nullptr, *d_e = nullptr;
    int64_t *d_pt = nullptr;
    uint32_t *d_small = nullptr, *d_out = nullptr;  // d_small: mag, ctr, rej_cnt
    uint16_t *d_rej = nullptr;
    const uint32_t cap = c->rej_cap ? c->rej_cap : 1;
    cudaStream_t st = c->stream;
    auto cleanup = [&]() {
        wipe_free(d_seeds, seeds.size());  // ep_seed and ep are the key pair's secret error
        wipe_free(d_e, n);
        cudaFree(d_pt);
        cudaFree(d_small);
        cudaFree(d_out);
        cudaFree(d_rej);
        seb_wipe(seeds.data(), seeds.size());
    };
#define CUK(call)                                                                          \
    do                                                                                     \
    {                                                                                      \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
        {                                                                                  \
            cleanup();                                                                     \
            return fail(SE_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));      \
        }                                                                                  \
    } while (0)
    CUK(cudaMalloc(&d_seeds, seeds.size()));
    CUK(cudaMalloc(&d_e, n));
    CUK(cudaMalloc(&d_pt, n * sizeof(int64_t)));
    CUK(cudaMalloc(&d_small, 3 * sizeof(uint32_t)));
    CUK(cudaMalloc(&d_out, 2 * np * n * sizeof(uint32_t)));
    CUK(cudaMalloc(&d_rej, cap * sizeof(uint16_t)));
    CUK(cudaMemcpyAsync(d_seeds, seeds.data(), seeds.size(), cudaMemcpyHostToDevice, st));
    CUK(cudaMemsetAsync(d_pt, 0, n * sizeof(int64_t), st));
    CUK(cudaMemsetAsync(d_small, 0, 3 * sizeof(uint32_t), st));
    seb_launch_sample_cbd(d_seeds, nullptr, reinterpret_cast<int8_t *>(d_e), (int)n, 1, 1, st);
    for (size_t p = 0; p < np; p++)
    {
        CUK(cudaMemsetAsync(d_small + 1, 0, sizeof(uint32_t), st));  // a fresh PRNG per prime: counter 0
        seb_launch_uniform(d_seeds + (p + 1) * SEB_SEED_BYTES, d_small + 1, d_out + (2 * p + 1) * n, 2 * np * n, (int)n,
                           c->mods.m[p], 1, d_rej, d_small + 2, cap, c->knobs, st);
    }
    CUK(cudaGetLastError());
    CUK(seb_launch_encrypt_sym(c->logn, d_pt, d_small, reinterpret_cast<int8_t *>(d_e), c->d_roots_sym, c->d_ntt_s, c->mods,
                               (int)np, d_out + n, d_out, 2 * np * n, 2 * n, 0, 1, st));
    c->launches += 2 + 2 * np;
    std::vector<uint32_t> out(2 * np * n);
    CUK(cudaMemcpyAsync(out.data(), d_out, out.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUK(cudaStreamSynchronize(st));
#undef CUK
    cleanup();
    for (size_t p = 0; p < np; p++)
    {
        memcpy(pk0 + p * n, out.data() + (2 * p) * n, n * sizeof(uint32_t));
        memcpy(pk1 + p * n, out.data() + (2 * p + 1) * n, n * sizeof(uint32_t));
    }
    if (c->asym) return seb_set_public_key(c, pk0, pk1);
    return 0;
}

extern "C" int seb_reserve(seb_ctx *c, size_t batch)
{
    if (!c) return fail(SE_ERR_INVALD_ARGUMENT, "null context");
    CU(cudaSetDevice(c->device));
    return ensure_scratch(c, c->dev, batch);
}

// ---------------------------------------------------------------------------------------------
// stage-level entry points
// ---------------------------------------------------------------------------------------------
static int check_vlen(seb_ctx *c, size_t vlen)
{
    if (vlen > c->n / 2) return fail(SE_ERR_INVALD_ARGUMENT, "vlen %zu exceeds n/2 = %zu", vlen, c->n / 2);
    return 0;
}

static int run_encode(seb_ctx *c, const float *d_values, size_t vlen, size_t batch, int64_t *d_pt, int *d_fail,
                      uint32_t *d_mag, cudaStream_t st)
{
    CU(cudaMemsetAsync(d_fail, 0, batch * sizeof(int), st));
    if (d_mag) CU(cudaMemsetAsync(d_mag, 0, batch * sizeof(uint32_t), st));
    CU(seb_launch_encode(c->logn, d_values, vlen, (int)vlen, c->d_src_map, c->d_tw, c->scale / (double)c->n, d_pt,
                         d_fail, d_mag, (int)batch, st));
    c->launches++;
    return 0;
}

extern "C" int seb_encode_device(seb_ctx *c, const float *d_values, size_t vlen, size_t batch, int64_t *d_pt)
{
    if (!c || !d_values || !d_pt) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    int r = check_vlen(c, vlen);
    if (r) return r;
    if ((r = ensure_scratch(c, c->dev, batch))) return r;
    c->last_batch = batch;
    return run_encode(c, d_values, vlen, batch, d_pt, c->dev.fail, nullptr, c->stream);
}

extern "C" int seb_sample_asym_device(seb_ctx *c, const uint8_t *d_seeds, size_t batch, uint8_t *d_u, int8_t *d_e,
                                      uint32_t *d_ctr)
{
    if (!c || !d_seeds || !d_u || !d_e || !d_ctr) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    seb_launch_sample_ternary(d_seeds, d_u, d_ctr, (int)c->n, (int)batch, c->stream);
    seb_launch_sample_cbd(d_seeds, d_ctr, d_e, (int)c->n, 2, (int)batch, c->stream);
    c->launches += 2;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int seb_sample_cbd_device(seb_ctx *c, const uint8_t *d_seeds, const uint32_t *d_ctr, size_t npoly,
                                     size_t batch, int8_t *d_e)
{
    if (!c || !d_seeds || !d_e) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    seb_launch_sample_cbd(d_seeds, d_ctr, d_e, (int)c->n, (int)npoly, (int)batch, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int seb_sample_uniform_device(seb_ctx *c, const uint8_t *d_seeds, uint32_t *d_ctr, size_t prime_idx,
                                         size_t batch, uint32_t *d_out, size_t ct_stride)
{
    if (!c || !d_seeds || !d_ctr || !d_out || prime_idx >= c->np)
        return fail(SE_ERR_INVALD_ARGUMENT, "bad argument");
    if (c->asym) return fail(SE_ERR_INVALD_ARGUMENT, "the uniform sampler belongs to a symmetric context");
    int r = ensure_scratch(c, c->dev, batch);
    if (r) return r;
    seb_launch_uniform(d_seeds, d_ctr, d_out, ct_stride, (int)c->n, c->mods.m[prime_idx], (int)batch,
                       c->dev.rej_idx, c->dev.rej_cnt, c->rej_cap, c->knobs, c->stream);
    c->launches += 2;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int seb_ntt_device(seb_ctx *c, uint32_t *d_polys, size_t batch)
{
    if (!c || !d_polys) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    CU(seb_launch_ntt(c->logn, d_polys, c->d_roots1, c->mods, (int)c->np, batch * c->np, c->stream));
    c->launches++;
    return 0;
}

extern "C" int seb_intt_device(seb_ctx *c, uint32_t *d_polys, size_t batch)
{
    if (!c || !d_polys) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    CU(seb_launch_intt(d_polys, c->d_iroots, c->d_ninv, c->mods, (int)c->n, (int)c->np, batch * c->np, c->sms * 8,
                       c->stream));
    c->launches++;
    return 0;
}

extern "C" int seb_decrypt_decode_device(seb_ctx *c, const uint32_t *d_ct, size_t batch, size_t prime_idx,
                                         size_t vlen, float *d_values_out)
{
    if (!c || !d_ct || !d_values_out || prime_idx >= c->np) return fail(SE_ERR_INVALD_ARGUMENT, "bad argument");
    if (!c->have_sk) return fail(SE_ERR_NO_KEY, "no secret key loaded");
    int r = check_vlen(c, vlen);
    if (r) return r;
    if (!c->d_work) CU(cudaMalloc(&c->d_work, (size_t)c->verify_ctas * c->n * sizeof(double2)));
    CU(seb_launch_decrypt_decode(d_ct, c->d_ntt_s_nat, c->d_iroots, c->d_ninv, c->d_tw, c->d_index_map, c->mods,
                                 (int)c->n, (int)c->np, (int)prime_idx, c->scale, c->d_work, c->verify_ctas, (int)vlen,
                                 d_values_out, batch, c->stream));
    c->launches++;
    return 0;
}

// Per-item 64-bit digest of d_words [items][words_per_item] (seb_verify.cu: k_digest): lets a parity test compare
// every ciphertext of a full-size batch with the compiled reference by moving 8 bytes per item.
extern "C" int seb_digest_device(seb_ctx *c, const uint32_t *d_words, size_t words_per_item, size_t items,
                                 uint64_t *d_digests)
{
    if (!c || !d_words || !d_digests) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    if (words_per_item % 4 != 0) return fail(SE_ERR_INVALD_ARGUMENT, "words_per_item must be a multiple of 4");
    CU(seb_launch_digest(d_words, words_per_item, items, d_digests, c->stream));
    c->launches++;
    return 0;
}

// The integer-issue ceilings of this device, measured now with register-only loops of the path's two inner operations
// (seb_verify.cu): Keccak-f[1600] permutations/s (ALU pipe) and lazy NTT butterflies/s (FMA pipe).  Synchronises.
extern "C" int seb_measure_ceilings(seb_ctx *c, double *keccak_f_per_s, double *butterflies_per_s)
{
    if (!c || !keccak_f_per_s || !butterflies_per_s) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    CU(cudaSetDevice(c->device));
    void *scratch = nullptr;
    CU(cudaMalloc(&scratch, (size_t)c->sms * 16 * 256 * 8));
    cudaError_t e = seb_measure_ceilings(c->sms, c->d_roots, c->primes[0], scratch, keccak_f_per_s, butterflies_per_s, c->stream);
    cudaFree(scratch);
    c->launches += 8;
    if (e != cudaSuccess) return fail(SE_ERR_CUDA, "seb_measure_ceilings: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int seb_prng_blocks_device(seb_ctx *c, const uint8_t *d_seeds, const uint64_t *d_counters, size_t count,
                                      uint64_t *d_out)
{
    if (!c || !d_seeds || !d_counters || !d_out) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    seb_launch_prng_blocks(d_seeds, d_counters, d_out, (int)count, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// full path, device pointers
// ---------------------------------------------------------------------------------------------
// seal_embedded.c:98-215 (asymmetric branch) for a batch: encode, sample u/e0/e1, then one fused
// kernel per (ciphertext, prime).
static int encrypt_asym_on(seb_ctx *c, Scratch &s, const float *d_values, size_t vlen, const uint8_t *d_seeds,
                           size_t batch, uint32_t *d_out, cudaStream_t st)
{
    const int n = (int)c->n;
    prof_mark(c, st, 0);
    int r       = run_encode(c, d_values, vlen, batch, s.pt, s.fail, s.mag, st);
    if (r) return r;
    prof_mark(c, st, 1);
    seb_launch_sample_ternary(d_seeds, s.u, s.ctr, n, (int)batch, st);
    prof_mark(c, st, 2);
    seb_launch_sample_cbd(d_seeds, s.ctr, s.e, n, 2, (int)batch, st);
    CU(cudaGetLastError());
    prof_mark(c, st, 3);
    CU(seb_launch_encrypt_asym(c->logn, s.pt, s.mag, s.e, s.u, c->d_roots, c->d_pk0, c->d_pk1, c->mods, (int)c->np, d_out,
                               (int)batch, st));
    prof_mark(c, st, 4);
    prof_next(c, st);
    c->launches += 3;
    return 0;
}

// seal_embedded.c:98-215 (symmetric branch): encode, e, then per prime sample a (the shareable
// PRNG's counter runs on across primes, ckks_sym.c:219), then the fused c0 kernel.
// Calls this small run every prime's squeeze at once on speculated counters (seb_sample.cu).  All candidate sponges
// of the call run concurrently, one warp each, and the warp-cooperative permutation slows down as the machine
// fills (2.9 us up to ~256 warps, 4.8 us at 2048, 7.9 us at 4096: profiles/r01_ab_uniform_coop.txt), while walking
// the chain serially costs np squeezes at 2.9 us: speculation pays while the candidates number less than about
// 1400 x np warps — 16-20 ciphertexts at n = 4096 x 3 primes (~215 candidates each), 4-5 at n = 16384 x 6 (~1500
// each); measured in profiles/r01_ab_sym_small_batches.txt.
// The "uniform_spec" option (0/1) forces the choice for up to 64 ciphertexts (tests, A/B measurements).
#define SEB_SPEC_MAX_BATCH 64
#define SEB_SPEC_WARPS_PER_PRIME 1400

// sample_poly_uniform for every prime of `batch` ciphertexts (the shareable PRNG's counter runs on across the
// primes, ckks_sym.c:219): a_p0 = row of (item 0, prime 0), prime p is p_stride words further, item b ct_stride.
static int run_uniform_chain(seb_ctx *c, Scratch &s, const uint8_t *d_sseeds, size_t batch, uint32_t *a_p0, size_t ct_stride,
                             size_t p_stride, cudaStream_t st)
{
    const int n = (int)c->n;
    CU(cudaMemsetAsync(s.ctr_a, 0, batch * sizeof(uint32_t), st));
    const bool fits = batch <= SEB_SPEC_MAX_BATCH &&
                      batch * ((size_t)c->spec_plan.total + 1) <= (size_t)SEB_SPEC_WARPS_PER_PRIME * c->np;
    const int force = c->knobs.uniform_spec;
    const bool spec = c->spec_plan.total > 0 && (force >= 0 ? (force != 0 && batch <= SEB_SPEC_MAX_BATCH) : fits);
    if (spec)
    {
        if (s.cand_cap < batch)
        {
            cudaFree(s.cand_rows);
            cudaFree(s.cand_list);
            cudaFree(s.cand_cnt);
            s.cand_rows = nullptr;
            s.cand_list = nullptr;
            s.cand_cnt  = nullptr;
            s.cand_cap  = 0;
            const size_t items = batch;
            const size_t slots = items * c->spec_plan.total;
            CU(cudaMalloc(&s.cand_rows, slots * c->n * sizeof(uint32_t)));
            CU(cudaMalloc(&s.cand_list, slots * (size_t)(c->rej_cap ? c->rej_cap : 1) * sizeof(uint16_t)));
            CU(cudaMalloc(&s.cand_cnt, slots * sizeof(uint32_t)));
            s.cand_cap = items;
        }
        if (!c->d_spec_misses)
        {
            CU(cudaMalloc(&c->d_spec_misses, sizeof(uint32_t)));
            CU(cudaMemsetAsync(c->d_spec_misses, 0, sizeof(uint32_t), st));
        }
        seb_launch_uniform_chain_spec(d_sseeds, s.ctr_a, a_p0, ct_stride, p_stride, n, c->mods, (int)c->np, c->spec_plan,
                                      (int)batch, s.cand_rows, s.cand_list, s.cand_cnt, s.rej_idx, s.rej_cnt, c->rej_cap,
                                      c->d_spec_misses, c->knobs, st);
        c->launches += 2 * c->np;  // one squeeze launch, a select per later prime, a fix-up per prime
    }
    else
    {
        for (size_t p = 0; p < c->np; p++)
            seb_launch_uniform(d_sseeds, s.ctr_a, a_p0 + p * p_stride, ct_stride, n, c->mods.m[p], (int)batch, s.rej_idx,
                               s.rej_cnt, c->rej_cap, c->knobs, st);
        c->launches += 2 * c->np;
    }
    CU(cudaGetLastError());
    return 0;
}

// How many items' encode + CBD go to the side partition: as many as finish there in 90 % of the time the sampler chain takes
// on the big one (a forced share through the "sym_side_percent" option).  Both times from the measured rates of the kernels
// (profiles/README.md): a bulk squeeze is 4n/136 sequential permutations of 5.1 us (3.4 us in the two-lane kernel), a fix-up n x (rejection rate of the
// prime) candidates per item at 4.4 G/s on the whole device; the CBD sampler n/16 permutations per item at 4.41 G/s, the
// encode 25 / 69 / 160 ns per item at n = 4096 / 8192 / 16384.
static size_t seb_side_items(const seb_ctx *c, size_t batch)
{
    double share;
    if (c->knobs.sym_side_percent >= 0)
        share = c->knobs.sym_side_percent / 100.0;
    else
    {
        const double n = (double)c->n, sms = c->knobs.sms, big = c->part_sms[0], side = c->part_sms[1];
        // a permutation of the chain: 5.1 us in the thread-per-sponge kernel, ~3.4 us where the two-lane kernel still has a
        // sub-partition per warp (seb_launch_uniform picks it there)
        const double perm = (double)batch <= 16.0 * 4.0 * big ? 3.4e-6 : 5.1e-6;
        double chain = 0.0;
        for (size_t p = 0; p < c->np; p++)
        {
            const double q   = c->primes[p];
            const double rej = (4294967296.0 - floor(4294967296.0 / q) * q) / 4294967296.0;
            chain += ceil(4.0 * n / 136.0) * perm + (double)batch * n * rej * 1.07 / 4.4e9 * (sms / big);
        }
        const double enc = c->logn <= 12 ? 25e-9 * n / 4096.0 : c->logn == 13 ? 69e-9 : 160e-9;
        const double ec  = (double)batch * (n / 16.0 / 4.41e9 + enc);
        share            = 0.9 * chain * side / (sms * ec);
    }
    if (share > 0.9) share = 0.9;
    if (share < 0.0) share = 0.0;
    return (size_t)((double)batch * share) & ~(size_t)7;
}

// seedct: the seed-compressed form (SE_ENABLE_SYM_SEED_CT, seal_embedded.c:184-194) — `a` goes to
// scratch instead of the c1 slots and d_out receives c0 only, [batch][np][n]; the receiver regenerates a
// from the 64-byte shareable seed (seb_expand_seedct_device).
static int encrypt_sym_on(seb_ctx *c, Scratch &s, const float *d_values, size_t vlen, const uint8_t *d_sseeds,
                          const uint8_t *d_seeds, size_t batch, uint32_t *d_out, int quirk, bool seedct,
                          cudaStream_t st)
{
    const int n = (int)c->n;
    if (seedct && s.a_cap < batch)
    {
        cudaFree(s.a_buf);
        s.a_buf = nullptr;
        s.a_cap = 0;
        CU(cudaMalloc(&s.a_buf, batch * c->np * c->n * sizeof(uint32_t)));
        s.a_cap = batch;
    }
    uint32_t *a_base       = seedct ? s.a_buf : d_out + c->n;
    const size_t ct_stride = (seedct ? 1 : 2) * c->np * c->n;
    const size_t p_stride  = (seedct ? 1 : 2) * c->n;
    // The latency-bound regime of the sampler chain (one thread-per-sponge warp per SM sub-partition of the big partition,
    // at a degree where the chain is long): chain and encode / CBD on disjoint SMs (seb_partition_ready).  Only for calls
    // on the context's own scratch (the host-pointer pipeline already overlaps its chunks).
    int r = 0;
    bool split = false;
    if (&s == &c->dev && c->knobs.sym_partition != 0 && c->logn >= 13)
    {
        // at most one thread-per-sponge warp per SM sub-partition of the big partition, and enough items that the
        // thread / two-lane kernels are the ones that run (below ~2k items the warp-per-sponge kernels fill the machine)
        const size_t warps = (batch + 31) / 32;
        const bool shape   = c->knobs.sms >= 64 && batch >= 2048 && warps <= (size_t)4 * (c->knobs.sms - SEB_PART_SIDE_SMS);
        split = (c->knobs.sym_partition > 0 || shape) && seb_partition_ready(c);
    }
    prof_mark(c, st, 0);
    if (!split)
    {
        r = run_encode(c, d_values, vlen, batch, s.pt, s.fail, s.mag, st);
        if (r) return r;
        prof_mark(c, st, 1);
        seb_launch_sample_cbd(d_seeds, nullptr, s.e, n, 1, (int)batch, st);
        prof_mark(c, st, 2);
        if ((r = run_uniform_chain(c, s, d_sseeds, batch, a_base, ct_stride, p_stride, st))) return r;
    }
    else
    {
        // items [0, b1): encode and CBD on the side partition, under the chain; items [b1, batch): on the whole device first
        const size_t b1 = seb_side_items(c, batch);
        cudaStream_t sa = c->part_stream[0], sb = c->part_stream[1];
        if (b1 < batch)
        {
            r = run_encode(c, d_values + b1 * vlen, vlen, batch - b1, s.pt + b1 * c->n, s.fail + b1, s.mag + b1, st);
            if (r) return r;
        }
        prof_mark(c, st, 1);
        if (b1 < batch) seb_launch_sample_cbd(d_seeds + b1 * SEB_SEED_BYTES, nullptr, s.e + b1 * c->n, n, 1, (int)(batch - b1), st);
        prof_mark(c, st, 2);
        CU(cudaEventRecord(c->part_ev[0], st));
        CU(cudaStreamWaitEvent(sa, c->part_ev[0], 0));
        CU(cudaStreamWaitEvent(sb, c->part_ev[0], 0));
        const int sms_all = c->knobs.sms;
        c->knobs.sms      = c->part_sms[0];  // the chain's kernel choices are made for the partition it runs on
        cudaStream_t aux  = c->knobs.aux_stream;
        c->knobs.aux_stream = nullptr;       // ... and nothing of it leaves that partition
        r                 = run_uniform_chain(c, s, d_sseeds, batch, a_base, ct_stride, p_stride, sa);
        c->knobs.sms      = sms_all;
        c->knobs.aux_stream = aux;
        if (r) return r;
        CU(cudaEventRecord(c->part_ev[1], sa));
        if (b1 > 0)
        {
            r = run_encode(c, d_values, vlen, b1, s.pt, s.fail, s.mag, sb);
            if (r) return r;
            seb_launch_sample_cbd(d_seeds, nullptr, s.e, n, 1, (int)b1, sb);
        }
        CU(cudaEventRecord(c->part_ev[2], sb));
        CU(cudaStreamWaitEvent(st, c->part_ev[1], 0));
        CU(cudaStreamWaitEvent(st, c->part_ev[2], 0));
    }
    prof_mark(c, st, 3);
    CU(seb_launch_encrypt_sym(c->logn, s.pt, s.mag, s.e, c->d_roots_sym, c->d_ntt_s, c->mods, (int)c->np, a_base, d_out,
                              ct_stride, p_stride, seedct ? 0 : quirk, (int)batch, st));
    prof_mark(c, st, 4);
    prof_next(c, st);
    c->launches += 2 + 2 * c->np;
    return 0;
}

extern "C" int seb_encrypt_asym_device(seb_ctx *c, const float *d_values, size_t vlen, const uint8_t *d_seeds,
                                       size_t batch, uint32_t *d_out)
{
    if (!c || !d_values || !d_seeds || !d_out) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    if (!c->have_pk) return fail(SE_ERR_NO_KEY, "no public key loaded");
    int r = check_vlen(c, vlen);
    if (r) return r;
    if ((r = ensure_scratch(c, c->dev, batch))) return r;
    c->last_batch = batch;
    return encrypt_asym_on(c, c->dev, d_values, vlen, d_seeds, batch, d_out, c->stream);
}

static int encrypt_sym_device(seb_ctx *c, const float *d_values, size_t vlen, const uint8_t *d_sseeds,
                              const uint8_t *d_seeds, size_t batch, uint32_t *d_out, int quirk, bool seedct)
{
    if (!c || !d_values || !d_seeds || !d_sseeds || !d_out) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    if (c->asym) return fail(SE_ERR_INVALD_ARGUMENT, "symmetric encryption on an asymmetric context");
    if (!c->have_sk) return fail(SE_ERR_NO_KEY, "no secret key loaded");
    int r = check_vlen(c, vlen);
    if (r) return r;
    if ((r = ensure_scratch(c, c->dev, batch))) return r;
    c->last_batch = batch;
    return encrypt_sym_on(c, c->dev, d_values, vlen, d_sseeds, d_seeds, batch, d_out, quirk, seedct, c->stream);
}

extern "C" int seb_encrypt_sym_device(seb_ctx *c, const float *d_values, size_t vlen, const uint8_t *d_sseeds,
                                      const uint8_t *d_seeds, size_t batch, uint32_t *d_out, int quirk)
{
    return encrypt_sym_device(c, d_values, vlen, d_sseeds, d_seeds, batch, d_out, quirk, false);
}

extern "C" int seb_encrypt_sym_seedct_device(seb_ctx *c, const float *d_values, size_t vlen, const uint8_t *d_sseeds,
                                             const uint8_t *d_seeds, size_t batch, uint32_t *d_c0_out)
{
    return encrypt_sym_device(c, d_values, vlen, d_sseeds, d_seeds, batch, d_c0_out, 0, true);
}

// Receiver side of the seed-compressed form: out[b][p][0] = c0[b][p], out[b][p][1] = a regenerated from the
// shareable seed exactly as the sender sampled it (sample.c:39-57, counter running on across primes).
// d_c0 == NULL leaves the c0 slots of d_out untouched.
extern "C" int seb_expand_seedct_device(seb_ctx *c, const uint8_t *d_sseeds, const uint32_t *d_c0, size_t batch,
                                        uint32_t *d_out)
{
    if (!c || !d_sseeds || !d_out) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    if (c->asym) return fail(SE_ERR_INVALD_ARGUMENT, "seed-compressed ciphertexts are symmetric");
    if (batch == 0) return 0;
    int r;
    if ((r = ensure_scratch(c, c->dev, batch))) return r;
    Scratch &s      = c->dev;
    cudaStream_t st = c->stream;
    const size_t n  = c->n;
    if (d_c0)
        CU(cudaMemcpy2DAsync(d_out, 2 * n * sizeof(uint32_t), d_c0, n * sizeof(uint32_t), n * sizeof(uint32_t),
                             batch * c->np, cudaMemcpyDeviceToDevice, st));
    return run_uniform_chain(c, s, d_sseeds, batch, d_out + n, 2 * c->np * n, 2 * n, st);
}

// Per-kernel timing of the next `max_steps` full-path *_device calls: CUDA events are recorded on
// the context's stream around each kernel (asym: encode, sample_ternary, sample_cbd, encrypt;
// sym: encode, sample_cbd, sample_uniform (all primes), encrypt).
extern "C" int seb_profile_begin(seb_ctx *c, int max_steps)
{
    if (!c || max_steps < 0) return fail(SE_ERR_INVALD_ARGUMENT, "bad argument");
    const size_t need = (size_t)max_steps * (SEB_PROF_SEGMENTS + 1);
    while (c->prof_ev.size() < need)
    {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        c->prof_ev.push_back(e);
    }
    c->prof_max  = max_steps;
    c->prof_step = 0;
    c->prof_on   = max_steps > 0;
    return 0;
}

// Synchronises the stream, stops profiling and writes ms[step][4]; returns the number of steps.
extern "C" int seb_profile_end(seb_ctx *c, float *ms)
{
    if (!c || !ms) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    CU(cudaStreamSynchronize(c->stream));
    c->prof_on = false;
    for (int s = 0; s < c->prof_step; s++)
        for (int k = 0; k < SEB_PROF_SEGMENTS; k++)
            CU(cudaEventElapsedTime(ms + s * SEB_PROF_SEGMENTS + k, c->prof_ev[(size_t)s * (SEB_PROF_SEGMENTS + 1) + k],
                                    c->prof_ev[(size_t)s * (SEB_PROF_SEGMENTS + 1) + k + 1]));
    return c->prof_step;
}

extern "C" int seb_encode_failures(seb_ctx *c)
{
    if (!c) return fail(SE_ERR_INVALD_ARGUMENT, "null context");
    CU(cudaStreamSynchronize(c->stream));
    if (!c->last_batch) return 0;
    std::vector<int> f(c->last_batch);
    CU(cudaMemcpy(f.data(), c->dev.fail, f.size() * sizeof(int), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int v : f) bad += v != 0;
    return bad;
}

// ---------------------------------------------------------------------------------------------
// full path, host pointers: the batch is cut in chunks, two in flight, each on its own stream, so the
// H2D of chunk k+1 and the D2H of chunk k-1 overlap the kernels of chunk k.  With pinned caller buffers
// every copy is asynchronous.  Pageable buffers are handed to cudaMemcpyAsync as they are (the runtime
// stages them and the call blocks); chunk k+1 is launched before chunk k's blocking read-back, so the
// kernels still overlap the copies.
// ---------------------------------------------------------------------------------------------
static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Items per chunk.  Large enough to keep the copy engines busy (>= 64 MiB of ciphertext) AND to fill the
// GPU: the asymmetric path's ternary sampler is a warp per ciphertext, and the symmetric path's uniform
// sampler is ONE sequential sponge per ciphertext whose latency (~2 ms per prime at n = 16384) is paid
// per chunk whatever its size — 85-item chunks made config D's host path 12x slower than its device
// path (profiles/README.md).  Chunks are balanced so the last one is not a sliver.  The "host_chunk"
// option (SEB_HOST_CHUNK at seb_create) overrides the item count.
static size_t host_chunk(const seb_ctx *c, bool sym, size_t batch)
{
    const size_t per_ct = 2 * c->np * c->n * sizeof(uint32_t);
    size_t chunk        = (64u << 20) / per_ct;
    const size_t fill   = sym ? 8192 : 2048;
    if (chunk < fill) chunk = fill;
    if (c->knobs.host_chunk > 0) chunk = (size_t)c->knobs.host_chunk;
    if (chunk >= batch) return batch;
    const size_t nchunks = (batch + chunk - 1) / chunk;
    return (batch + nchunks - 1) / nchunks;
}

static int ensure_io(seb_ctx *c, Scratch &s, size_t chunk, bool sym, size_t per_ct)
{
    (void)sym;
    if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!s.done) CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    if (chunk <= s.io_cap && per_ct <= s.io_out_words) return 0;
    if (chunk < s.io_cap) chunk = s.io_cap;
    if (per_ct < s.io_out_words) per_ct = s.io_out_words;
    // allocate-then-swap (like ensure_scratch): a failed grow keeps the old staging buffers
    Scratch f;
    AllocGroup g;
    g.want((void **)&f.d_values, chunk * (c->n / 2) * sizeof(float));
    g.want((void **)&f.d_seeds, chunk * SEB_SEED_BYTES);
    g.want((void **)&f.d_sseeds, chunk * SEB_SEED_BYTES);
    g.want((void **)&f.d_out, chunk * per_ct * sizeof(uint32_t));
    int *h_fail = nullptr;
    if (g.err == cudaSuccess) g.err = cudaMallocHost(&h_fail, chunk * sizeof(int));
    if (g.failed())
        return fail(SE_ERR_NO_MEMORY, "staging for %zu-item chunks: %s", chunk, cudaGetErrorString(g.err));
    g.commit();
    wipe_free(s.d_values, s.io_cap * (c->n / 2) * sizeof(float));
    wipe_free(s.d_seeds, s.io_cap * SEB_SEED_BYTES);
    cudaFree(s.d_sseeds);
    cudaFree(s.d_out);
    cudaFreeHost(s.h_fail);
    s.d_values = f.d_values, s.d_seeds = f.d_seeds, s.d_sseeds = f.d_sseeds, s.d_out = f.d_out;
    s.h_fail       = h_fail;
    s.io_cap       = chunk;
    s.io_out_words = per_ct;
    return 0;
}

static int encrypt_host(seb_ctx *c, bool sym, const float *values, size_t vlen, const uint8_t *sseeds,
                        const uint8_t *seeds, size_t batch, uint32_t *out, int quirk, bool seedct = false,
                        bool packed = false)
{
    if (!c) return fail(SE_ERR_INVALD_ARGUMENT, "null context");
    if (batch == 0) return 0;
    if (!values || !seeds || !out || (sym && !sseeds)) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    if (sym == c->asym) return fail(SE_ERR_INVALD_ARGUMENT, "encryption type does not match the context");
    if (sym ? !c->have_sk : !c->have_pk) return fail(SE_ERR_NO_KEY, "key material not loaded");
    int r = check_vlen(c, vlen);
    if (r) return r;
    CU(cudaSetDevice(c->device));
    const size_t chunk  = host_chunk(c, sym, batch);
    const size_t per_ct = (seedct ? 1 : 2) * c->np * c->n;  // words of output per ciphertext on the device
    const size_t per_out = packed ? per_ct / 16 * 15 : per_ct;  // ... and in the caller's buffer
    const bool pin_out  = is_pinned(out);
    const size_t nslots = chunk < batch ? 2 : 1;
    for (size_t k = 0; k < nslots; k++)
    {
        if ((r = ensure_scratch(c, c->hslot[k], chunk))) return r;
        if ((r = ensure_io(c, c->hslot[k], chunk, sym, per_ct))) return r;
        Scratch &s = c->hslot[k];
        if (packed && s.pack_words < chunk * per_out)
        {
            cudaFree(s.d_pack);
            s.d_pack     = nullptr;
            s.pack_words = 0;
            CU(cudaMalloc(&s.d_pack, chunk * per_out * sizeof(uint32_t)));
            s.pack_words = chunk * per_out;
        }
    }
    struct Pending
    {
        size_t first = 0, count = 0;
        bool live = false;
    } pend[2];
    int failures = 0;

    // wait for chunk k's results to be in the caller's buffer
    auto drain = [&](int k) -> int {
        if (!pend[k].live) return 0;
        Scratch &s = c->hslot[k];
        if (!pin_out)  // blocking copy, ordered after the chunk's kernels on its stream
            CU(cudaMemcpyAsync(out + pend[k].first * per_out, packed ? s.d_pack : s.d_out,
                               pend[k].count * per_out * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventSynchronize(s.done));
        for (size_t i = 0; i < pend[k].count; i++) failures += s.h_fail[i] != 0;
        pend[k].live = false;
        return 0;
    };

    int k = 0;
    for (size_t first = 0; first < batch; first += chunk, k ^= 1)
    {
        const size_t count = batch - first < chunk ? batch - first : chunk;
        Scratch &s         = c->hslot[k];  // free: its previous chunk was drained one iteration ago
        const float *hv    = values + first * vlen;
        const uint8_t *hs  = seeds + first * SEB_SEED_BYTES;
        const uint8_t *hss = sym ? sseeds + first * SEB_SEED_BYTES : nullptr;
        CU(cudaMemcpyAsync(s.d_values, hv, count * vlen * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        CU(cudaMemcpyAsync(s.d_seeds, hs, count * SEB_SEED_BYTES, cudaMemcpyHostToDevice, s.stream));
        if (sym) CU(cudaMemcpyAsync(s.d_sseeds, hss, count * SEB_SEED_BYTES, cudaMemcpyHostToDevice, s.stream));
        r = sym ? encrypt_sym_on(c, s, s.d_values, vlen, s.d_sseeds, s.d_seeds, count, s.d_out, quirk, seedct, s.stream)
                : encrypt_asym_on(c, s, s.d_values, vlen, s.d_seeds, count, s.d_out, s.stream);
        if (r) return r;
        if (packed)
        {
            CU(seb_launch_pack30(s.d_out, s.d_pack, count * per_ct, s.stream));
            c->launches++;
        }
        if (pin_out)
            CU(cudaMemcpyAsync(out + first * per_out, packed ? s.d_pack : s.d_out, count * per_out * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(s.h_fail, s.fail, count * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventRecord(s.done, s.stream));
        pend[k].first = first;
        pend[k].count = count;
        pend[k].live  = true;
        if ((r = drain(k ^ 1))) return r;  // the previous chunk, while this one computes
    }
    if ((r = drain(0))) return r;
    if ((r = drain(1))) return r;
    if (failures) return fail(SE_ERR_ENCODE_RANGE, "%d of %zu items exceed the int64 range when encoded", failures, batch);
    return 0;
}

extern "C" int seb_encrypt_asym_host(seb_ctx *c, const float *values, size_t vlen, const uint8_t *seeds,
                                     size_t batch, uint32_t *out)
{
    return encrypt_host(c, false, values, vlen, nullptr, seeds, batch, out, 0);
}

// The optional packed wire form (SURVEY 8f-1 "wire formats"): every residue is below 2^30, so
// the [nprimes][2][n] words of a ciphertext travel as 15/16 of their size, 30 bits per residue.  out_packed receives
// seb_packed30_words(ctx) words per ciphertext; seb_unpack30 / seb_unpack30_device restore the full form bit for bit.
extern "C" size_t seb_packed30_words(const seb_ctx *c) { return c ? 2 * c->np * c->n / 16 * 15 : 0; }

extern "C" int seb_encrypt_asym_host_packed30(seb_ctx *c, const float *values, size_t vlen, const uint8_t *seeds,
                                              size_t batch, uint32_t *out_packed)
{
    return encrypt_host(c, false, values, vlen, nullptr, seeds, batch, out_packed, 0, false, true);
}

extern "C" int seb_encrypt_sym_host_packed30(seb_ctx *c, const float *values, size_t vlen, const uint8_t *sseeds,
                                             const uint8_t *seeds, size_t batch, uint32_t *out_packed, int quirk)
{
    return encrypt_host(c, true, values, vlen, sseeds, seeds, batch, out_packed, quirk, false, true);
}

// d_packed [words * 15 / 16] -> d_out [words] (words a multiple of 16), on the context's stream
extern "C" int seb_unpack30_device(seb_ctx *c, const uint32_t *d_packed, size_t words, uint32_t *d_out)
{
    if (!c || !d_packed || !d_out) return fail(SE_ERR_INVALD_ARGUMENT, "null argument");
    if (words % 16) return fail(SE_ERR_INVALD_ARGUMENT, "words must be a multiple of 16");
    CU(seb_launch_unpack30(d_packed, d_out, words, c->stream));
    c->launches++;
    return 0;
}

// the same on the host (no GPU involved): packed [words * 15 / 16] -> out [words]
extern "C" int seb_unpack30(const uint32_t *packed, size_t words, uint32_t *out)
{
    if (!packed || !out || words % 16) return SE_ERR_INVALD_ARGUMENT;
    for (size_t g = 0; g < words / 16; g++)
    {
        const uint32_t *w = packed + 15 * g;
        for (int i = 0; i < 16; i++)
        {
            const int k = (30 * i) / 32, sh = 30 * i - 32 * k;
            const uint64_t two = ((uint64_t)(k < 14 ? w[k + 1] : 0u) << 32) | w[k];
            out[16 * g + i]    = (uint32_t)(two >> sh) & 0x3FFFFFFFu;
        }
    }
    return 0;
}

extern "C" int seb_encrypt_sym_host(seb_ctx *c, const float *values, size_t vlen, const uint8_t *sseeds,
                                    const uint8_t *seeds, size_t batch, uint32_t *out, int quirk)
{
    return encrypt_host(c, true, values, vlen, sseeds, seeds, batch, out, quirk);
}

extern "C" int seb_encrypt_sym_seedct_host(seb_ctx *c, const float *values, size_t vlen, const uint8_t *sseeds,
                                           const uint8_t *seeds, size_t batch, uint32_t *c0_out)
{
    if (c && c->asym) return fail(SE_ERR_INVALD_ARGUMENT, "seed-compressed ciphertexts are symmetric");
    return encrypt_host(c, true, values, vlen, sseeds, seeds, batch, c0_out, 0, true);
}
