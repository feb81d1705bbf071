// emul.cpp — sequential g++ emulation of the per-thread device functions in
// seal-embedded_b200/csrc/*.cuh (NTT passes, IFFT passes, Keccak, sampler bit tricks).
//
// TEST INFRASTRUCTURE: lets the CPU test-suite check the kernels' index math, swizzles and bit
// manipulation against the oracle without a GPU.  "Threads" run one after another and a pass
// boundary (__syncthreads in the kernel) is the end of the loop over threads.  Nothing here is
// part of the product and nothing here computes a product result.
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

#include "seb_encode.cuh"
#include "seb_ntt.cuh"
#include "seb_sample.cuh"

// ---------------------------------------------------------------------------------------------
// NTT
// ---------------------------------------------------------------------------------------------
struct HostLoad
{
    const uint32_t *src;
    int n;
    uint32_t operator()(int p, uint32_t pos) const { return src[(size_t)p * n + pos]; }
};

// LOGN below is a plan KEY (seb_ntt.cuh NttCfg): log2(n) for 16 coefficients per thread, 16 + log2(n) for 32
template <int LOGN, int NPOLY, int P>
static void ntt_passes(std::vector<std::array<uint32_t[NttCfg<LOGN>::E], NPOLY>> &regs, uint32_t *smem, const seb_oct *tw,
                       uint32_t q, HostLoad &ld)
{
    constexpr int T = NttCfg<LOGN>::T;
    constexpr int E = NttCfg<LOGN>::E;
    for (int t = 0; t < T; t++)
    {
        uint32_t(&x)[NPOLY][E] = *reinterpret_cast<uint32_t(*)[NPOLY][E]>(regs[t].data());
        seb_ntt_pass<LOGN, P, NPOLY>(x, smem, t, tw, q, 2 * q, ld);
    }
    if constexpr (P + 1 < NttPlan<LOGN>::NPASS) ntt_passes<LOGN, NPOLY, P + 1>(regs, smem, tw, q, ld);
}

template <int LOGN, int NPOLY>
static void ntt_emul(const uint32_t *in, const uint2 *tw, uint32_t q, uint32_t *out)
{
    constexpr int N = 1 << NttCfg<LOGN>::LOGN;
    constexpr int T = NttCfg<LOGN>::T;
    std::vector<std::array<uint32_t[NttCfg<LOGN>::E], NPOLY>> regs(T);
    std::vector<uint32_t> smem_store((size_t)NPOLY * NttSmem<LOGN>::WORDS + 4, 0xDEADBEEFu);
    uint32_t *smem_al = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(smem_store.data()) + 15) & ~uintptr_t(15));
    HostLoad ld{in, N};
    std::vector<seb_oct> tabs(NttTwSize<LOGN>::OCTS + 1);
    memset(tabs.data(), 0, tabs.size() * sizeof(seb_oct));
    seb_build_tw<LOGN>(tw, tabs.data());
    ntt_passes<LOGN, NPOLY, 0>(regs, smem_al, tabs.data(), q, ld);
    // outputs are handed over through an epilogue-ordered identity table, the way the kernels
    // read their key tables: out[pos] must come back as pos
    using O = NttOut<LOGN>;
    std::vector<uint2> ident(N);
    for (int i = 0; i < N; i++) ident[i] = make_uint2((uint32_t)i, ~(uint32_t)i);
    std::vector<seb_oct> epi(N / 4);
    seb_build_epi<LOGN>(ident.data(), epi.data());
    for (int t = 0; t < T; t++)
        for (int i = 0; i < O::GPL; i++)
            for (int j = 0; j < O::RUN; j++)
            {
                const seb_oct &o   = epi[seb_epi_index<LOGN>(t, i, j / 4)];
                const uint32_t pos = o.v[2 * (j % 4)];
                if (pos != O::pos(t, i) + j || o.v[2 * (j % 4) + 1] != ~pos) throw 1;
                for (int p = 0; p < NPOLY; p++)
                    out[(size_t)p * N + pos] = seb_final_reduce(regs[t][p][i * O::RUN + j], q, 2 * q);
            }
}

extern "C" int emul_ntt(int logn, int npoly, const uint32_t *in, const uint32_t *roots_w, const uint32_t *roots_wq,
                        uint32_t q, uint32_t *out)
{
    const int n = 1 << (logn & 15);
    std::vector<uint2> tw(n);
    for (int i = 0; i < n; i++) tw[i] = make_uint2(roots_w[i], roots_wq[i]);
#define CASE(L)                                                  \
    case L:                                                      \
        if (npoly == 1)                                          \
            ntt_emul<L, 1>(in, tw.data(), q, out);               \
        else if (npoly == 3)                                     \
            ntt_emul<L, 3>(in, tw.data(), q, out);               \
        else                                                     \
            return -1;                                           \
        return 0;
    switch (logn)
    {
        CASE(10)
        CASE(11)
        CASE(12)
        CASE(13)
        CASE(14)
        CASE(29)
        CASE(30)
    }
#undef CASE
    return -1;
}

extern "C" uint32_t emul_pad(int logn, uint32_t a)
{
    switch (logn)
    {
        case 10: return seb_pad<10>(a);
        case 11: return seb_pad<11>(a);
        case 12: return seb_pad<12>(a);
        case 13: return seb_pad<13>(a);
        case 14: return seb_pad<14>(a);
        case 29: return seb_pad<29>(a);
        case 30: return seb_pad<30>(a);
    }
    return a;
}

extern "C" uint32_t emul_smem_words(int logn)
{
    switch (logn)
    {
        case 10: return NttSmem<10>::WORDS;
        case 11: return NttSmem<11>::WORDS;
        case 12: return NttSmem<12>::WORDS;
        case 13: return NttSmem<13>::WORDS;
        case 14: return NttSmem<14>::WORDS;
        case 29: return NttSmem<29>::WORDS;
        case 30: return NttSmem<30>::WORDS;
    }
    return 0;
}

extern "C" int emul_plan(int logn, int *radices)
{
#define CASE(L)                                                                 \
    case L:                                                                     \
        for (int i = 0; i < NttPlan<L>::NPASS; i++) radices[i] = NttPlan<L>::R[i]; \
        return NttPlan<L>::NPASS;
    switch (logn)
    {
        CASE(10)
        CASE(11)
        CASE(12)
        CASE(13)
        CASE(14)
        CASE(29)
        CASE(30)
    }
#undef CASE
    return 0;
}

// coefficient index of element j of group i of thread t in pass P; -1 outside the plan
extern "C" int64_t emul_ntt_elem(int logn, int pass, uint32_t t, uint32_t i, uint32_t j)
{
#define CASEP(L, P)                                                                                   \
    if (logn == L && pass == P)                                                                       \
    {                                                                                                 \
        if (P >= NttPlan<L>::NPASS) return -1;                                                        \
        constexpr int PP = P < NttPlan<L>::NPASS ? P : 0;                                             \
        constexpr int R  = NttPlan<L>::R[PP];                                                         \
        constexpr int LS = NttCfg<L>::LOGN - NttS0<L, PP>::value - R;                                 \
        if (i >= (uint32_t)(NttCfg<L>::E >> R) || j >= (1u << R)) return -1;                          \
        return (int64_t)(seb_ntt_group_base<L, PP>(t, i) | (j << LS));                                \
    }
#define CASEL(L) CASEP(L, 0) CASEP(L, 1) CASEP(L, 2) CASEP(L, 3)
    CASEL(10) CASEL(11) CASEL(12) CASEL(13) CASEL(14) CASEL(29) CASEL(30)
#undef CASEL
#undef CASEP
    return -1;
}

// scope of the barrier after pass `pass`: SEB_SYNC_CTA -> 0, SEB_SYNC_WARP -> 32, SEB_SYNC_GROUP -> threads
// per named barrier
extern "C" int emul_ntt_sync_scope(int logn, int pass)
{
#define CASEP(L, P)                                                                                  \
    if (logn == L && pass == P)                                                                      \
        return NttSync<L, P>::value == SEB_SYNC_CTA ? 0 : NttSync<L, P>::value == SEB_SYNC_WARP ? 32 : NttSync<L, P>::GROUP;
#define CASEL(L) CASEP(L, 0) CASEP(L, 1) CASEP(L, 2) CASEP(L, 3)
    CASEL(10) CASEL(11) CASEL(12) CASEL(13) CASEL(14) CASEL(29) CASEL(30)
#undef CASEL
#undef CASEP
    return 0;
}

extern "C" uint32_t emul_barrett64(uint64_t x, uint32_t q)
{
    const uint64_t ratio = (uint64_t)(((unsigned __int128)1 << 64) / q);
    SebModulus m{q, 2 * q, (uint32_t)ratio, (uint32_t)(ratio >> 32)};
    return seb_barrett64(x, m);
}
extern "C" uint32_t emul_barrett32(uint32_t x, uint32_t q)
{
    const uint64_t ratio = (uint64_t)(((unsigned __int128)1 << 64) / q);
    SebModulus m{q, 2 * q, (uint32_t)ratio, (uint32_t)(ratio >> 32)};
    return seb_barrett32(x, m);
}
extern "C" uint32_t emul_shoup_lazy(uint32_t x, uint32_t w, uint32_t q)
{
    return seb_mul_shoup_lazy(x, w, (uint32_t)(((uint64_t)w << 32) / q), q);
}

// ---------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------
template <int LOGN, int LOGNL, int P>
static void enc_passes(std::vector<std::array<double, 2 * ENC_E>> &regs, double *sre, double *sim, uint32_t pos0,
                       const float *svals, const uint16_t *src_map, const double2 *tw)
{
    constexpr int T = (1 << LOGNL) / ENC_E;
    for (int t = 0; t < T; t++)
    {
        double(&xr)[ENC_E] = *reinterpret_cast<double(*)[ENC_E]>(regs[t].data());
        double(&xi)[ENC_E] = *reinterpret_cast<double(*)[ENC_E]>(regs[t].data() + ENC_E);
        enc_pass<LOGN, LOGNL, P>(xr, xi, sre, sim, t, pos0, svals, src_map, tw);
    }
    if constexpr (P + 1 < enc_npass(LOGNL)) enc_passes<LOGN, LOGNL, P + 1>(regs, sre, sim, pos0, svals, src_map, tw);
}

static uint32_t g_emul_mag = 0;  // max |coefficient| (clipped) of the last emulated encode
extern "C" uint32_t emul_encode_mag(void) { return g_emul_mag; }

template <int LOGN, int CL>
static int enc_emul(const float *vals, int vlen, const uint16_t *src_map, const double2 *tw, double n_inv,
                    int64_t *out)
{
    constexpr int N     = 1 << LOGN;
    constexpr int NL    = N / CL;
    constexpr int LOGNL = (CL == 2) ? LOGN - 1 : LOGN;
    constexpr int T     = NL / ENC_E;
    constexpr int RL    = enc_r(LOGNL, enc_npass(LOGNL) - 1);
    constexpr int LSL   = ENC_LR * (enc_npass(LOGNL) - 1);
    int bad             = 0;
    g_emul_mag          = 0;
    std::vector<double> sre[2];  // 2*NL doubles per CTA, like the kernel's shared memory: re[NL] im[NL] or NL (re, im) pairs
    std::vector<float> svals(EncVals<LOGN>::WORDS, 0.0f);  // the kernel's staged, zero-padded, skewed message
    for (int i = 0; i < vlen && i < N / 2; i++) svals[enc_vskew<LOGN>((uint32_t)i)] = vals[i];
    for (int rank = 0; rank < CL; rank++)
    {
        sre[rank].assign(2 * NL + 2, 1e300);
        double *re0 = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(sre[rank].data()) + 15) & ~uintptr_t(15));
        double *im0 = re0 + NL;
        std::vector<std::array<double, 2 * ENC_E>> regs(T);
        enc_passes<LOGN, LOGNL, 0>(regs, re0, im0, rank * NL, svals.data(), src_map, tw);
        for (int t = 0; t < T; t++)
            for (int i = 0; i < (ENC_E >> RL); i++)
            {
                const uint32_t g    = (uint32_t)t + (uint32_t)i * T;
                const uint32_t off  = g & ((1u << LSL) - 1u);
                const uint32_t base = ((g >> LSL) << (LSL + RL)) | off;
                for (int j = 0; j < (1 << RL); j++)
                {
                    const uint32_t pos = base | ((uint32_t)j << LSL);
                    if (CL == 1)
                        out[pos] = enc_finish(regs[t][i * (1 << RL) + j], n_inv, bad, g_emul_mag);
                    else
                        enc_st(re0, im0, pos, regs[t][i * (1 << RL) + j], regs[t][ENC_E + i * (1 << RL) + j]);
                }
            }
    }
    if (CL == 2)
        for (uint32_t rank = 0; rank < 2; rank++)
            for (uint32_t k = 0; k < (uint32_t)NL; k++)
            {
                auto base = [&](uint32_t r) {
                    return reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(sre[r].data()) + 15) & ~uintptr_t(15));
                };
                double ar, ai, br, bi;
                enc_ld(base(rank), base(rank) + NL, k, ar, ai);
                enc_ld(base(rank ^ 1), base(rank ^ 1) + NL, k, br, bi);
                const double re    = enc_cross_re(rank, ar, ai, br, bi, tw[1]);
                out[rank * NL + k] = enc_finish(re, n_inv, bad, g_emul_mag);
            }
    return bad;
}

// tw: interleaved (re, im) pairs, n of them
extern "C" int emul_encode(int logn, const float *vals, int vlen, const uint16_t *src_map, const double *tw,
                           double n_inv, int64_t *out)
{
    const size_t n = (size_t)1 << logn;
    std::vector<double2> ext(enc_tw_entries(n));  // natural table + the pass-0 copies the kernel reads
    memcpy(ext.data(), tw, n * sizeof(double2));
    enc_build_tw0(n, ext.data());
    const double2 *t2 = ext.data();
    switch (logn)
    {
        case 10: return enc_emul<10, 1>(vals, vlen, src_map, t2, n_inv, out);
        case 11: return enc_emul<11, 1>(vals, vlen, src_map, t2, n_inv, out);
        case 12: return enc_emul<12, 1>(vals, vlen, src_map, t2, n_inv, out);
        case 13: return enc_emul<13, 1>(vals, vlen, src_map, t2, n_inv, out);
        case 14: return enc_emul<14, 2>(vals, vlen, src_map, t2, n_inv, out);
    }
    return -1;
}

// local position of element j of slot i of thread t in encode pass `pass` (-1 outside the plan); lognl = log2 of
// the positions one CTA holds.  Mirrors enc_pass's addressing (g = t + i*T).
extern "C" int64_t emul_enc_pos(int lognl, int pass, uint32_t t, uint32_t i, uint32_t j)
{
    if (pass >= enc_npass(lognl)) return -1;
    const int R = enc_r(lognl, pass), LS = ENC_LR * pass;
    const uint32_t T = (1u << lognl) / ENC_E;
    if (i >= (uint32_t)(ENC_E >> R) || j >= (1u << R) || t >= T) return -1;
    const uint32_t g = t + i * T, off = g & ((1u << LS) - 1u), blk = g >> LS;
    return (int64_t)(((blk << (LS + R)) | off) | (j << LS));
}
extern "C" int emul_enc_sync_width(int lognl, int pass) { return enc_sync_width(lognl, pass); }
extern "C" int emul_enc_e(void) { return ENC_E; }

// physical word of message slot s in the staged (skewed) layout, and the size of that buffer
extern "C" uint32_t emul_enc_vskew(int logn, uint32_t slot)
{
    switch (logn)
    {
        case 10: return enc_vskew<10>(slot);
        case 11: return enc_vskew<11>(slot);
        case 12: return enc_vskew<12>(slot);
        case 13: return enc_vskew<13>(slot);
        case 14: return enc_vskew<14>(slot);
    }
    return 0;
}
extern "C" uint32_t emul_enc_vwords(int logn)
{
    switch (logn)
    {
        case 10: return EncVals<10>::WORDS;
        case 11: return EncVals<11>::WORDS;
        case 12: return EncVals<12>::WORDS;
        case 13: return EncVals<13>::WORDS;
        case 14: return EncVals<14>::WORDS;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Keccak and sampler blocks
// ---------------------------------------------------------------------------------------------
extern "C" void emul_keccak(uint64_t *st)
{
    uint64_t a[25];
    memcpy(a, st, sizeof a);
    seb_keccak_f1600(a);
    memcpy(st, a, sizeof a);
}

// first rate block of SHAKE256(seed || LE64(ctr)) as 17 words
extern "C" void emul_prng_block(const uint8_t *seed, uint64_t ctr, uint64_t *out17)
{
    uint64_t s[8], a[25];
    memcpy(s, seed, 64);
    seb_prng_init(a, s, ctr);
    seb_keccak_f1600(a);
    memcpy(out17, a, 17 * 8);
}

extern "C" void emul_ternary_block(const uint8_t *seed, uint64_t ctr, uint32_t *packed6, uint32_t *mask3)
{
    uint64_t s[8], a[25];
    memcpy(s, seed, 64);
    seb_prng_init(a, s, ctr);
    seb_keccak_f1600<12, true>(a);  // the pruned permutation the ternary sampler uses (round 0 specialised)
    uint32_t p[6];
    seb_ternary_block(a, p, mask3[0], mask3[1], mask3[2]);
    memcpy(packed6, p, sizeof p);
}

// the same on a caller-supplied 96-byte block (exhaustive byte-value checks)
extern "C" void emul_ternary_block_raw(const uint8_t *bytes96, uint32_t *packed6, uint32_t *mask3)
{
    uint64_t a[25] = {0};
    memcpy(a, bytes96, 96);
    uint32_t p[6];
    seb_ternary_block(a, p, mask3[0], mask3[1], mask3[2]);
    memcpy(packed6, p, sizeof p);
}

// the product's form: bit-interleaved state, samples counted in place
extern "C" void emul_cbd_block(const uint8_t *seed, uint64_t ctr, uint32_t *out4)
{
    uint64_t s[8];
    memcpy(s, seed, 64);
    uint32_t se[8], so[8], e[25], o[25];
    for (int i = 0; i < 8; i++) se[i] = seb_half_bits(s[i], 0), so[i] = seb_half_bits(s[i], 1);
    seb_prng_init_il(e, o, se, so, ctr);
    seb_keccak_f1600_il12(e, o);
    uint32_t out[4];
    seb_cbd_block_il(e, o, out);
    memcpy(out4, out, sizeof out);
}

// the plain form (64-bit lanes), kept as a second opinion
extern "C" void emul_cbd_block_plain(const uint8_t *seed, uint64_t ctr, uint32_t *out4)
{
    uint64_t s[8], a[25];
    memcpy(s, seed, 64);
    seb_prng_init(a, s, ctr);
    seb_keccak_f1600<12>(a);  // the pruned permutation the samplers use
    uint32_t o[4];
    seb_cbd_block(a, o);
    memcpy(out4, o, sizeof o);
}

// a 4-byte redraw of the uniform sampler the way the fix-up kernels compute it
extern "C" uint32_t emul_prng_word_il(const uint8_t *seed, uint64_t ctr)
{
    uint64_t s[8];
    memcpy(s, seed, 64);
    uint32_t se[8], so[8];
    for (int i = 0; i < 8; i++) se[i] = seb_half_bits(s[i], 0), so[i] = seb_half_bits(s[i], 1);
    return seb_prng_word_il(se, so, ctr);
}

// one full interleaved permutation against the plain one: state in / state out as 64-bit lanes
extern "C" void emul_keccak_il(uint64_t *a25)
{
    uint32_t e[25], o[25];
    for (int i = 0; i < 25; i++) e[i] = seb_half_bits(a25[i], 0), o[i] = seb_half_bits(a25[i], 1);
    for (int round = 0; round < 24; round++) seb_keccak_round_il<25>(e, o, round);
    for (int i = 0; i < 25; i++)
    {
        uint64_t w = 0;
        for (int k = 0; k < 32; k++)
            w |= ((uint64_t)((e[i] >> k) & 1u) << (2 * k)) | ((uint64_t)((o[i] >> k) & 1u) << (2 * k + 1));
        a25[i] = w;
    }
}

extern "C" uint32_t emul_mod3_bytes(uint32_t x) { return seb_mod3_bytes(x); }
