// seb_sample.cu — SHAKE256-driven samplers, bit-exact with device/lib/sample.c.
//
//   k_prng_blocks      : raw PRNG output blocks (tests / KATs)
//   k_sample_ternary   : sample_small_poly_ternary_prng_96 (sample.c:218-242), one warp per ciphertext
//   k_sample_cbd       : sample_*_cbd_generic_prng_16 (sample.c:263-356), one thread per 96-byte call
//   k_uniform_bulk/fix : sample_poly_uniform (sample.c:39-57), thread per ciphertext + warp per ciphertext
//
// PRNG consumption order is the reference's: every prng_fill_buffer call is
// a fresh SHAKE256(seed || LE64(counter)) and bumps the counter, including the data-dependent
// single-value redraws of the rejection samplers.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "seb_kernels.h"
#include "seb_sample.cuh"

// batches up to this size take the warp-cooperative bulk sampler (25 lanes per sponge): it beats the two-lane kernel
// up to ~1400 items at n = 4096 and 16384 alike (profiles/r02_ab_uniform_pair.txt; against the thread-per-sponge
// kernel alone the crossover was ~3000, profiles/r01_ab_uniform_coop.txt)
#ifndef SEB_UNIFORM_COOP_MAX_BATCH
#define SEB_UNIFORM_COOP_MAX_BATCH 1280
#endif

__device__ __forceinline__ void load_seed(const uint8_t *seeds, size_t b, uint64_t (&s)[8])
{
    const uint64_t *p = reinterpret_cast<const uint64_t *>(seeds + b * SEB_SEED_BYTES);
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = __ldg(p + i);
}

// The seed of ciphertext b split into even and odd bits (seb_keccak.cuh, interleaved state) by a whole warp that serves
// this one ciphertext: lane = 16 * parity + 8 * (high word) + seed word, every lane compresses ONE 32-bit piece into 16
// bits, lane L and lane L + 8 make a half lane, lanes 0..7 / 16..23 end up with the even / odd halves of seed word L & 7,
// and 16 shuffles hand them round.  All 32 lanes must call it.
__device__ __forceinline__ void seb_seed_split_warp(const uint8_t *seeds, size_t b, const int lane, uint32_t (&se)[8],
                                                    uint32_t (&so)[8])
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t piece    = __ldg(reinterpret_cast<const uint32_t *>(seeds + b * SEB_SEED_BYTES) + 2 * (lane & 7) +
                                    ((lane >> 3) & 1));
    const uint32_t v        = seb_half_bits32(piece, lane >> 4);
    const uint32_t comb     = __byte_perm(v, __shfl_down_sync(FULL, v, 8), 0x5410);
#pragma unroll
    for (int i = 0; i < 8; i++) se[i] = __shfl_sync(FULL, comb, i), so[i] = __shfl_sync(FULL, comb, 16 + i);
}

// ---------------------------------------------------------------------------------------------
// raw blocks: out[i][0..136) = first rate block of SHAKE256(seed[b_i] || LE64(ctr_i))
// ---------------------------------------------------------------------------------------------
__global__ void k_prng_blocks(const uint8_t *__restrict__ seeds, const uint64_t *__restrict__ counters,
                              uint64_t *__restrict__ out, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t s[8], a[25];
    load_seed(seeds, (size_t)i, s);
    seb_prng_init(a, s, counters[i]);
    seb_keccak_f1600(a);
#pragma unroll
    for (int k = 0; k < 17; k++) out[(size_t)i * 17 + k] = a[k];
}

// ---------------------------------------------------------------------------------------------
// ternary u (packed 2 bits/coefficient, MSB-first in each byte: sample.c:61-87)
// ---------------------------------------------------------------------------------------------
// One warp per ciphertext.  The sampler's PRNG counter walk is sequential (block j+1's counter
// depends on how many redraws blocks <= j needed), but X(seed,c,96) and X(seed,c,1) are the same
// SHAKE stream, so the warp computes 32 consecutive counters' streams per wave and then decides,
// in the reference's order, which counter was a 96-byte block and which a 1-byte redraw (seb_tern_walk).
// ---------------------------------------------------------------------------------------------
// the same sampler, TWO ciphertexts per warp sharing the last wave (n = 4096)
// ---------------------------------------------------------------------------------------------
// At n = 4096 a ciphertext consumes 43 blocks + 32.25 +- 5.7 redraws = 75 PRNG counters: three waves of 32 compute 96
// permutations for it, and the kernel is ALU-pipe bound (95 %), so a fifth of its time is spent on counters nobody
// reads.  Here a warp owns ciphertexts 2w and 2w+1: two full waves each (counters 0..63), then ONE wave shared by
// both from counter 64, its lanes split by what each still owes; whoever is not done afterwards gets full waves
// of its own.  ~2.55 waves per ciphertext instead of 3.  Same walk, same bytes, same counters as k_sample_ternary.
struct SebTernWalk
{
    int j = 0;       // next block index
    int need = 0;    // redraws still owed to the current block
    int cur = 0;     // current block index
    int served = 0;  // redraws already served to the current block
    uint32_t cm0 = 0, cm1 = 0, cm2 = 0;  // reject masks of the current block (all of them: served ones included)
    uint32_t consumed = 0;
};

// position of the ord-th (0-based) set bit of the 96-bit mask (a0, a1, a2)
__device__ __forceinline__ uint32_t seb_nth_set96(uint32_t a0, uint32_t a1, uint32_t a2, int ord)
{
    const int c0 = __popc(a0), c1 = __popc(a1);
    uint32_t word = a0, base = 0;
    if (ord >= c0 + c1)
        word = a2, base = 64u, ord -= c0 + c1;
    else if (ord >= c0)
        word = a1, base = 32u, ord -= c0;
#pragma unroll 1  // trip counts of 0-2: an unrolled loop only adds prologue work
    for (; ord > 0; ord--) word &= word - 1u;  // a block has a handful of rejected bytes (2 in 256): a short loop
    return base + (uint32_t)__ffs(word) - 1u;
}

// Resolve the counters held by lanes [lo, hi) of this wave in the reference's order (sample.c:223-241: a counter is a
// 96-byte block, or a 1-byte redraw owed to the block before it).  Two steps: (1) ROLES - "lane i is block j" / "lane i is
// the k-th redraw of block j" - from a ballot of the acceptable redraw bytes and each lane's number of rejected bytes, by
// pointer doubling over "next block" links (see inside); (2) DATA, in parallel: block lanes store their 24 packed bytes,
// redraw lanes look up the k-th rejected position in their block's masks and write their two bits there.
// (Round 1 walked the lanes one by one and moved the data in the same loop - a shared-memory byte update by lane 0 per
// redraw, three shuffles and six predicated stores per block: a quarter of the kernel's instructions.)
__device__ __forceinline__ void seb_tern_walk(SebTernWalk &w, const int lo, const int hi, const int lane, const int n,
                                              const int nblocks, uint32_t *usm, const uint32_t (&packed)[6],
                                              const uint32_t m0, const uint32_t m1, const uint32_t m2, const uint32_t b0)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t acc  = __ballot_sync(FULL, b0 < 0xFEu);  // counters whose first byte is acceptable as a redraw
    const int vlast     = n - 96 * (nblocks - 1);           // bytes of the last block that are used: 32, 64 or 96
    const int need_full = __popc(m0) + __popc(m1) + __popc(m2);
    const int need_last = __popc(m0) + (vlast > 32 ? __popc(m1) : 0) + (vlast > 64 ? __popc(m2) : 0);
    const uint32_t e0 = w.cm0, e1 = w.cm1, e2 = w.cm2;  // masks of the block carried over from the previous wave
    const int j_in = w.j, cur_in = w.cur, served_in = w.served;
    // Which lanes are blocks and which are accepted redraws, WITHOUT walking the lanes one by one (round 2 did, in a
    // warp-uniform loop of ~14 instructions per lane: a tenth of the kernel).  Every lane answers "if I were a block,
    // which lane would the next block be?" - the lane after the need-th acceptable redraw above it, found with a
    // clear-lowest-bit loop of need - 1 steps (0.75 on average) - and the blocks of this wave are the lanes reachable
    // from the first one by that map: five rounds of pointer doubling, each one warp OR-reduction and one shuffle.
    auto below_bits = [](int p) { return p >= 32 ? 0xFFFFFFFFu : (1u << p) - 1u; };
    const uint32_t A = acc & below_bits(hi) & ~below_bits(lo);  // acceptable redraws among this wave's counters
    // (1) the block carried over from the previous wave takes the first w.need of them
    int p0 = lo, need_c = 0;  // first lane that can be a block; what the carried block still owes afterwards
    if (w.need > 0)
    {
        const int have = __popc(A);
        if (have < w.need)
            p0 = 32, need_c = w.need - have;
        else
        {
            uint32_t t = A;
#pragma unroll 1
            for (int k = 1; k < w.need; k++) t &= t - 1u;
            p0 = __ffs(t);  // the lane after the last redraw it needed
        }
    }
    // (2) the chain of blocks from p0
    uint32_t is_blk   = 0;
    const int avail   = nblocks - w.j;
    if (p0 < hi && avail > 0)
    {
        const uint32_t above = A & ~below_bits(lane + 1);
        int nxt;
        if (need_full == 0)
            nxt = lane + 1;
        else if (__popc(above) < need_full)
            nxt = 32;
        else
        {
            uint32_t t = above;
#pragma unroll 1
            for (int k = 1; k < need_full; k++) t &= t - 1u;
            nxt = __ffs(t);
        }
        if (nxt >= hi) nxt = 32;  // the next block belongs to the next wave
        is_blk = 1u << p0;
#pragma unroll
        for (int round = 0; round < 5; round++)
        {
            is_blk |= __reduce_or_sync(FULL, (((is_blk >> lane) & 1u) && nxt < 32) ? 1u << nxt : 0u);
            const int far = __shfl_sync(FULL, nxt, nxt & 31);
            nxt           = nxt < 32 ? far : 32;
        }
        if (__popc(is_blk) > avail) is_blk &= (1u << __fns(is_blk, 0, avail + 1)) - 1u;  // the ciphertext ends in this wave
    }
    // (3) the state after this wave
    const int nb = __popc(is_blk);
    int owner    = -1;  // lane of this wave that holds the current block (-1: carried over)
    int end      = hi;  // counters of [lo, end) are consumed
    w.j += nb;
    if (nb > 0)
    {
        owner               = 31 - __clz(is_blk);
        const int n_b       = __shfl_sync(FULL, w.j == nblocks ? need_last : need_full, owner);
        const uint32_t rest = A & ~below_bits(owner + 1);
        const int got       = __popc(rest);
        w.cur               = w.j - 1;
        if (got >= n_b)
        {
            w.need = 0, w.served = n_b;
            if (w.j == nblocks)  // that was the last block, and it is complete: the walk ends behind its last redraw
            {
                uint32_t t = rest;
#pragma unroll 1
                for (int k = 1; k < n_b; k++) t &= t - 1u;
                end = n_b == 0 ? owner + 1 : __ffs(t);
            }
        }
        else
            w.need = n_b - got, w.served = got;
    }
    else
    {
        w.served += __popc(A & below_bits(p0));
        w.need = need_c;
        if (need_c == 0 && w.j == nblocks) end = p0;  // only redraws of the last block were left
    }
    w.consumed += (uint32_t)(end - lo);
    const uint32_t is_red = A & ~is_blk & below_bits(end);
    // every lane derives its role from the masks: block lanes count the blocks below them; redraw lanes belong to the
    // nearest block below them (or to the carried block) and count the redraws in between
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t bb    = is_blk & below;
    const int my_owner   = bb ? 31 - __clz(bb) : -1;
    int kind = 0, blk = 0, ord = 0;  // 1 = block `blk`, 2 = redraw number `ord` of block `blk`
    if ((is_blk >> lane) & 1u)
        kind = 1, blk = j_in + __popc(bb);
    else if ((is_red >> lane) & 1u)
    {
        kind = 2;
        if (my_owner >= 0)
            blk = j_in + __popc(bb) - 1, ord = __popc(is_red & below & ~((2u << my_owner) - 1u));
        else
            blk = cur_in, ord = served_in + __popc(is_red & below);
    }
    const int valid = min(96, n - 96 * blk);  // multiple of 32 for every legal n
    if (kind == 1)
    {
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (k * 16 < valid) usm[blk * 6 + k] = packed[k];
    }
    __syncwarp();
    {
        const int src = my_owner < 0 ? lane : my_owner;
        uint32_t a0 = __shfl_sync(FULL, m0, src), a1 = __shfl_sync(FULL, m1, src), a2 = __shfl_sync(FULL, m2, src);
        if (kind == 2)
        {
            if (my_owner < 0) a0 = e0, a1 = e1, a2 = e2;
            if (valid <= 32) a1 = 0u;
            if (valid <= 64) a2 = 0u;
            const uint32_t pos  = seb_nth_set96(a0, a1, a2, ord);
            const uint32_t byte = (uint32_t)blk * 24u + (pos >> 2);
            const uint32_t sh   = 8u * (byte & 3u) + 6u - 2u * (pos & 3u);
            atomicAnd(usm + (byte >> 2), ~(3u << sh));  // the block left a don't-care value in a rejected field
            atomicOr(usm + (byte >> 2), (b0 % 3u) << sh);
        }
    }
    if (w.need > 0 && owner >= 0)  // warp-uniform: the block that continues into the next wave started in this one
    {
        w.cm0 = __shfl_sync(FULL, m0, owner);
        w.cm1 = __shfl_sync(FULL, m1, owner);
        w.cm2 = __shfl_sync(FULL, m2, owner);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128) k_sample_ternary(const uint8_t *__restrict__ seeds,
                                                        uint8_t *__restrict__ u_out,
                                                        uint32_t *__restrict__ ctr_out, int n, int batch)
{
    extern __shared__ uint32_t usm_all[];
    const int lane      = threadIdx.x & 31;
    const int warp      = threadIdx.x >> 5;
    const int b         = blockIdx.x * (blockDim.x >> 5) + warp;
    const int words_per = n / 16;  // n/4 bytes
    uint32_t *usm       = usm_all + warp * words_per;
    if (b >= batch) return;

    uint64_t seed[8];
    load_seed(seeds, (size_t)b, seed);
    const int nblocks = (n + 95) / 96;
    SebTernWalk w;
    for (uint64_t cbase = 0; w.j < nblocks || w.need > 0; cbase += 32)
    {
        uint64_t a[25];
        seb_prng_init(a, seed, cbase + (uint64_t)lane);
        seb_keccak_f1600<12, true>(a);  // 96 bytes (or 1) are read from each call
        uint32_t m0, m1, m2, packed[6];
        seb_ternary_block(a, packed, m0, m1, m2);
        const uint32_t b0 = (uint32_t)a[0] & 0xFFu;  // value if this counter was a 1-byte redraw
        seb_tern_walk(w, 0, 32, lane, n, nblocks, usm, packed, m0, m1, m2, b0);
    }
    uint32_t *dst = reinterpret_cast<uint32_t *>(u_out + (size_t)b * (n / 4));
    for (int k = lane; k < words_per; k += 32) dst[k] = usm[k];
    if (lane == 0) ctr_out[b] = w.consumed;
}

__global__ void __launch_bounds__(128) k_sample_ternary_pair(const uint8_t *__restrict__ seeds, uint8_t *__restrict__ u_out,
                                                             uint32_t *__restrict__ ctr_out, int n, int batch)
{
    extern __shared__ uint32_t usm_all[];
    const int lane      = threadIdx.x & 31;
    const int warp      = threadIdx.x >> 5;
    const int pair      = blockIdx.x * (blockDim.x >> 5) + warp;
    const int words_per = n / 16;  // n/4 bytes
    if (2 * pair >= batch) return;
    const int nct       = 2 * pair + 1 < batch ? 2 : 1;  // an odd batch leaves the last warp one ciphertext
    const int nblocks   = (n + 95) / 96;
    uint32_t *usm0      = usm_all + (2 * warp) * words_per, *usm1 = usm0 + words_per;
    SebTernWalk w0, w1;  // two named states, no run-time indexed array: they stay in registers

    // The schedule is a loop with ONE instance of the wave body (Keccak + block extraction + walk): four inlined
    // copies fell out of the instruction cache and ran 25 % slower than the unpaired kernel.
    // A wave serves ciphertext c0 on lanes [0, split) from counter base0 and ciphertext c1 on lanes [split, 32) from base1.
    constexpr uint32_t FULL = 64;  // counters every ciphertext gets from full waves of its own
    uint32_t next0 = 0, next1 = 0;  // next counter of each ciphertext
    bool tail_done = nct == 1;
    uint64_t seed[8];
    int seed_of = -1;  // whose seed this lane holds
    for (;;)
    {
        const bool open0 = w0.j < nblocks || w0.need > 0;
        const bool open1 = nct == 2 && (w1.j < nblocks || w1.need > 0);
        int first, split;  // ciphertext on lanes [0, split); lanes [split, 32) serve the other one
        if (open0 && (next0 < FULL || tail_done))
            first = 0, split = 32;
        else if (open1 && (next1 < FULL || tail_done))
            first = 1, split = 32;
        else if (!tail_done && (open0 || open1))
        {
            // The shared wave.  What each ciphertext still owes is known up to the redraws to come: blocks left + redraws
            // owed now + 3/4 redraw per block left (96 bytes x 2/256) + 1; the spare lanes are split evenly.  If the two
            // estimates do not fit 32 lanes the wave is cut in the middle.
            const int rem0 = nblocks - w0.j, rem1 = nblocks - w1.j;
            const int est0 = open0 ? rem0 + w0.need + (3 * rem0 + 3) / 4 + 1 : 0;
            const int est1 = open1 ? rem1 + w1.need + (3 * rem1 + 3) / 4 + 1 : 0;
            split          = est0 + est1 <= 32 ? est0 + (32 - est0 - est1) / 2 : 16;
            first = 0, tail_done = true;
        }
        else
            break;
        const int mine       = lane < split ? first : 1;
        const uint32_t count = (mine ? next1 : next0) + (uint32_t)(lane < split ? lane : lane - split);
        uint64_t a[25];
        if (seed_of != mine)  // only the lanes that change ciphertext reload (L1-resident 64 bytes)
        {
            load_seed(seeds, (size_t)(2 * pair + mine), seed);
            seed_of = mine;
        }
        seb_prng_init(a, seed, (uint64_t)count);
        seb_keccak_f1600<12, true>(a);  // 96 bytes (or 1) are read from each call
        uint32_t m0, m1, m2, packed[6];
        seb_ternary_block(a, packed, m0, m1, m2);
        const uint32_t b0 = (uint32_t)a[0] & 0xFFu;  // value if this counter was a 1-byte redraw
        if (first == 0)
        {
            seb_tern_walk(w0, 0, split, lane, n, nblocks, usm0, packed, m0, m1, m2, b0);
            next0 += (uint32_t)split;
        }
        if (first == 1 || split < 32)
        {
            seb_tern_walk(w1, first == 1 ? 0 : split, 32, lane, n, nblocks, usm1, packed, m0, m1, m2, b0);
            next1 += (uint32_t)(first == 1 ? 32 : 32 - split);
        }
    }

    __syncwarp();
    for (int c = 0; c < nct; c++)
    {
        uint32_t *dst       = reinterpret_cast<uint32_t *>(u_out + (size_t)(2 * pair + c) * (n / 4));
        const uint32_t *src = c ? usm1 : usm0;
        for (int k = lane; k < words_per; k += 32) dst[k] = src[k];
        if (lane == 0) ctr_out[2 * pair + c] = c ? w1.consumed : w0.consumed;
    }
}

// ---------------------------------------------------------------------------------------------
// centered binomial (k=21): one thread per 96-byte PRNG call = 16 samples
// ---------------------------------------------------------------------------------------------
// e_out: [batch][npoly][n] int8; polynomial k of ciphertext b uses counters
// ctr_base[b] + k*n/16 + (0 .. n/16)   (ckks_asym.c:199-200: e0 then e1 from the same PRNG)
//
// The permutation runs on a bit-interleaved state (seb_keccak.cuh: 174 ALU operations per round instead of 180) because
// the samples are popcounts and can be taken from the interleaved block as it is.  The 32 threads of a warp serve the
// same ciphertext, so its seed is split into even and odd bits once per warp - one 32-bit piece per lane, 17 shuffles hand
// the halves round - instead of 32 times.
// Grid: x = ciphertext, y = slice of its calls (blockDim.x calls each; the launcher picks a block size that divides
// n/16 * npoly), so no index is divided and a warp never straddles two ciphertexts.
__global__ void __launch_bounds__(128) k_sample_cbd(const uint8_t *__restrict__ seeds,
                                                    const uint32_t *__restrict__ ctr_base,
                                                    int8_t *__restrict__ e_out, int n, int npoly, int batch)
{
    const uint32_t b = blockIdx.x;
    const uint32_t r = blockIdx.y * blockDim.x + threadIdx.x;
    const int lane   = threadIdx.x & 31;
    (void)batch;

    uint32_t se[8], so[8];
    seb_seed_split_warp(seeds, (size_t)b, lane, se, so);
    uint32_t e[25], o[25];
    seb_prng_init_il(e, o, se, so, (uint64_t)(ctr_base ? ctr_base[b] : 0u) + r);
    seb_keccak_f1600_il12(e, o);  // 96 bytes per call

    uint32_t out[4];
    seb_cbd_block_il(e, o, out);
    uint4 *dst = reinterpret_cast<uint4 *>(e_out + (size_t)b * npoly * n + (size_t)r * 16);
    *dst       = make_uint4(out[0], out[1], out[2], out[3]);
}

// ---------------------------------------------------------------------------------------------
// uniform mod q with rejection (symmetric `a`): bulk 4n-byte squeeze, then ordered redraws
// ---------------------------------------------------------------------------------------------
// Thread per ciphertext: the 4n-byte squeeze is one sequential sponge.  Accepted words are
// stored already reduced mod q (< q <= max_multiple); rejected words are stored raw
// (>= max_multiple) and their indices appended, in ascending order, to the ciphertext's reject list
// (rej_idx[b][0..cap), rej_cnt[b] = how many there were, which may exceed cap).
// Rounds per iteration of the bulk squeeze's permutation loop: 1.  Unrolling by 2 (the loop's seven instructions - counter,
// two round-constant loads, compare, branch - are 3.5 % of a rolled round) measured 21.51 against 21.49 ms for
// configuration D's six primes: they run on the uniform datapath and cost the lone warp nothing.
#ifndef SEB_BULK_UNROLL_N
#define SEB_BULK_UNROLL_N 1
#endif
constexpr int SEB_BULK_UNROLL = SEB_BULK_UNROLL_N;

// A value the compiler must keep in a register: it otherwise re-reads the modulus constants from the parameter bank in
// front of every use (68 LDC per block of 34 words), and at the batch sizes where this kernel runs ONE in-order warp per
// SM sub-partition every instruction is an issue slot of the sequential sponge.
__device__ __forceinline__ uint32_t seb_pin(uint32_t v)
{
    asm volatile("" : "+r"(v));
    return v;
}

// One word of the squeeze in six instructions: v = x >= max_multiple ? x (raw, for the fix-up) : x mod q, and the word's
// bit ORed into the reject mask under the same predicate.  Spelled in PTX because the compiler turns the predicated OR
// into a select and an add, and `x - hi * q` into a negation and a multiply-add (negq = -q avoids it).
template <uint32_t BIT>
__device__ __forceinline__ uint32_t seb_uniform_word(const uint32_t x, const uint32_t negq, const uint32_t q,
                                                     const uint32_t ratio, const uint32_t max_multiple, uint32_t &mask)
{
    const uint32_t red = seb_csub(__umulhi(x, ratio) * negq + x, q);  // seb_barrett32
    uint32_t v;
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u32 p, %2, %3;\n\t@p or.b32 %0, %0, %5;\n\tselp.b32 %1, %2, %4, p;\n\t}"
        : "+r"(mask), "=r"(v)
        : "r"(x), "r"(max_multiple), "r"(red), "n"(BIT));
    return v;
}
template <int K>
struct SebUniformLanes  // rate lanes K..16 of a full block, unrolled by recursion (BIT must be a constant expression)
{
    __device__ __forceinline__ static void run(const uint32_t (&lo)[25], const uint32_t (&hi)[25], uint2 *__restrict__ dst,
                                               const uint32_t negq, const uint32_t q, const uint32_t ratio,
                                               const uint32_t mm, uint32_t &m0, uint32_t &m1)
    {
        uint32_t &m       = K < 16 ? m0 : m1;
        constexpr int sh  = K < 16 ? 2 * K : 0;
        const uint32_t vl = seb_uniform_word<(1u << sh)>(lo[K], negq, q, ratio, mm, m);
        const uint32_t vh = seb_uniform_word<(2u << sh)>(hi[K], negq, q, ratio, mm, m);
        dst[K]            = make_uint2(vl, vh);
        if constexpr (K < 16) SebUniformLanes<K + 1>::run(lo, hi, dst, negq, q, ratio, mm, m0, m1);
    }
};

// the 34 words of one rate block (m0 bit 2k / 2k+1: low / high word of rate lane k rejected; m1: lane 16): a full block
// without per-lane bound checks, or the tail of nk < 17 lanes in plain C.  Every lane writes its own row, 17 64-bit
// stores per block; pairing them into 128-bit stores (the odd lane carried into the next block to stay 16-byte aligned)
// measured SLOWER: 21.83 against 21.49 ms for configuration D's six primes.
template <bool FULL_BLOCK>
__device__ __forceinline__ void seb_uniform_block(const uint32_t (&lo)[25], const uint32_t (&hi)[25], const int nk,
                                                  uint2 *__restrict__ dst, const uint32_t negq, const uint32_t q,
                                                  const uint32_t ratio, const uint32_t max_multiple, uint32_t &m0,
                                                  uint32_t &m1)
{
    // Branch-free over the 17 rate lanes (rejections are 2 % of the words: a branch per word costs more than the
    // work it skips); the rejected words of the block are collected in a bit mask and listed afterwards.
    m0 = 0, m1 = 0;
    if constexpr (FULL_BLOCK)
        SebUniformLanes<0>::run(lo, hi, dst, negq, q, ratio, max_multiple, m0, m1);
    else
    {
#pragma unroll
        for (int k = 0; k < 17; k++)
        {
            const bool in = k < nk;
            const bool rl = in && lo[k] >= max_multiple, rh = in && hi[k] >= max_multiple;
            if (k < 16)
            {
                if (rl) m0 |= 1u << (2 * k);
                if (rh) m0 |= 2u << (2 * k);
            }
            else
            {
                if (rl) m1 |= 1u;
                if (rh) m1 |= 2u;
            }
            const uint32_t tl = __umulhi(lo[k], ratio) * negq + lo[k], th = __umulhi(hi[k], ratio) * negq + hi[k];
            const uint32_t vl = rl ? lo[k] : seb_csub(tl, q), vh = rh ? hi[k] : seb_csub(th, q);
            if (in) dst[k] = make_uint2(vl, vh);
        }
    }
}

__global__ void __launch_bounds__(32) k_uniform_bulk(const uint8_t *__restrict__ seeds,
                                                      const uint32_t *__restrict__ ctr, uint32_t *__restrict__ out,
                                                      size_t ct_stride, int n, SebModulus mod,
                                                      uint32_t max_multiple, int batch,
                                                      uint16_t *__restrict__ rej_idx, uint32_t *__restrict__ rej_cnt,
                                                      uint32_t cap)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    uint32_t lo[25], hi[25];  // the state as 32-bit halves throughout: no repacking between permutation and squeeze
    {
        uint64_t s[8], a[25];
        load_seed(seeds, (size_t)b, s);
        seb_prng_init(a, s, (uint64_t)ctr[b]);
#pragma unroll
        for (int i = 0; i < 25; i++) lo[i] = (uint32_t)a[i], hi[i] = (uint32_t)(a[i] >> 32);
    }
    const uint32_t q = seb_pin(mod.q), negq = seb_pin(0u - mod.q), ratio = seb_pin(mod.ratio_hi), mm = seb_pin(max_multiple);
    uint2 *dst     = reinterpret_cast<uint2 *>(out + (size_t)b * ct_stride);
    uint16_t *list = rej_idx + (size_t)b * cap;
    uint32_t cnt   = 0;
    uint32_t word  = 0;    // index of the next word of the polynomial
    int left       = n / 2;  // 64-bit lanes still to emit
    while (left > 0)
    {
#pragma unroll SEB_BULK_UNROLL
        for (int round = 0; round < 24; round++) seb_keccak_round<false>(lo, hi, round);
        uint32_t m0, m1;
        if (left >= 17)
            seb_uniform_block<true>(lo, hi, 17, dst, negq, q, ratio, mm, m0, m1);
        else
            seb_uniform_block<false>(lo, hi, left, dst, negq, q, ratio, mm, m0, m1);
        while (m0)
        {
            const int bit = __ffs(m0) - 1;
            if (cnt < cap) list[cnt] = (uint16_t)(word + bit);
            cnt++;
            m0 &= m0 - 1;
        }
        while (m1)
        {
            const int bit = __ffs(m1) - 1;
            if (cnt < cap) list[cnt] = (uint16_t)(word + 32 + bit);
            cnt++;
            m1 &= m1 - 1;
        }
        dst += 17;
        word += 34;
        left -= 17;
    }
    rej_cnt[b] = cnt;
}

// ---------------------------------------------------------------------------------------------
// the same bulk squeeze for SMALL batches: one WARP per ciphertext, the Keccak state spread over 25 lanes
// ---------------------------------------------------------------------------------------------
// A 4n-byte squeeze is 4n/136 DEPENDENT permutations (121 at n = 4096, 482 at n = 16384, times the primes: the
// counter is chained).  One thread runs a permutation in ~6.9 us, so a lone symmetric se_encrypt call spent
// 2.4 ms here against 1.2 ms for the whole call on a CPU core (profiles/r01_latency_single_call.txt).  With
// lane 5y+x holding A[x][y] a round is 16 ALU operations and 18 shuffles instead of 180 ALU operations issued in
// order by one thread: 2.4x lower latency per permutation (2.9 us) at several times the issue slots per permutation,
// so it wins whenever the batch leaves the machine under-filled (seb_launch_uniform picks by batch size).
// Output, reject lists and counters are exactly those of k_uniform_bulk (sample.c:39-57).
struct SebCoopLane
{
    int col[4];      // the other four lanes of this lane's column
    int xm1, xp1;    // lanes (x-1, y) and (x+1, y)
    int s0, s1, s2;  // pre-pi source lanes of B[x][y], B[x+1][y], B[x+2][y]
    uint32_t rot;    // rho offset of the word this lane holds
};

__device__ __forceinline__ SebCoopLane seb_coop_setup(const int lane)
{
    // rho offsets indexed by lane = 5y + x (the SRC/ROT columns of seb_keccak.cuh's round macro)
    const uint32_t rho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    SebCoopLane c;
    const int l = lane < 25 ? lane : 0;  // lanes 25..31 shadow lane 0 (their results are never used)
    const int x = l % 5, y = l / 5;
#pragma unroll
    for (int k = 0; k < 4; k++) c.col[k] = (l + 5 * (k + 1)) % 25;
    c.xm1 = 5 * y + (x + 4) % 5;
    c.xp1 = 5 * y + (x + 1) % 5;
    // B[X][Y] = rotl(A[x][y]) with X = y, Y = (2x + 3y) % 5  =>  y = X, x = (3Y + X) % 5: source lane 5X + (3Y + X) % 5
    auto src = [](int X, int Y) { return 5 * X + (3 * Y + X) % 5; };
    c.s0  = src(x, y);
    c.s1  = src((x + 1) % 5, y);
    c.s2  = src((x + 2) % 5, y);
    c.rot = 0;
#pragma unroll
    for (int i = 0; i < 25; i++)
        if (i == l) c.rot = rho[i];
    return c;
}

// One Keccak-f[1600] on the state held as (lo, hi) of lane 5y + x: 18 shuffles in three dependent levels and 16
// ALU operations per round.  (Exchanging through shared memory instead — one 64-bit store, __syncwarp(), 10 + 3
// independent 64-bit loads per round — measured 13 % slower: profiles/README.md.)
//   theta : C = xor of the column (4 shuffles per half), D = C[x-1] ^ rotl(C[x+1], 1) (2 shuffles per half)
//   rho   : every lane rotates its own word by its own offset
//   pi+chi: lane (x,y) fetches B[x][y], B[x+1][y], B[x+2][y] straight from the lanes that hold them before pi
//           (3 shuffles per half)
__device__ __forceinline__ void seb_keccak_coop(uint32_t &lo, uint32_t &hi, const SebCoopLane &c, const int lane)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
#pragma unroll 1
    for (int round = 0; round < 24; round++)
    {
        // theta
        const uint32_t a1l = __shfl_sync(FULL, lo, c.col[0]), a1h = __shfl_sync(FULL, hi, c.col[0]);
        const uint32_t a2l = __shfl_sync(FULL, lo, c.col[1]), a2h = __shfl_sync(FULL, hi, c.col[1]);
        const uint32_t a3l = __shfl_sync(FULL, lo, c.col[2]), a3h = __shfl_sync(FULL, hi, c.col[2]);
        const uint32_t a4l = __shfl_sync(FULL, lo, c.col[3]), a4h = __shfl_sync(FULL, hi, c.col[3]);
        const uint32_t cl  = seb_xor3(seb_xor3(lo, a1l, a2l), a3l, a4l);
        const uint32_t ch  = seb_xor3(seb_xor3(hi, a1h, a2h), a3h, a4h);
        const uint32_t cml = __shfl_sync(FULL, cl, c.xm1), cmh = __shfl_sync(FULL, ch, c.xm1);
        const uint32_t cpl = __shfl_sync(FULL, cl, c.xp1), cph = __shfl_sync(FULL, ch, c.xp1);
        const uint32_t tl  = seb_xor3(lo, cml, __funnelshift_l(cph, cpl, 1));
        const uint32_t th  = seb_xor3(hi, cmh, __funnelshift_l(cpl, cph, 1));
        // rho: rotl64 by this lane's offset (swap the halves for offsets >= 32, then funnel by offset % 32)
        const bool sw      = c.rot >= 32;
        const uint32_t ul  = sw ? th : tl, uh = sw ? tl : th;
        const uint32_t bl  = __funnelshift_l(uh, ul, c.rot);  // shift counts are taken modulo 32
        const uint32_t bh  = __funnelshift_l(ul, uh, c.rot);
        // pi + chi
        const uint32_t b0l = __shfl_sync(FULL, bl, c.s0), b0h = __shfl_sync(FULL, bh, c.s0);
        const uint32_t b1l = __shfl_sync(FULL, bl, c.s1), b1h = __shfl_sync(FULL, bh, c.s1);
        const uint32_t b2l = __shfl_sync(FULL, bl, c.s2), b2h = __shfl_sync(FULL, bh, c.s2);
        lo = seb_chi(b0l, b1l, b2l);
        hi = seb_chi(b0h, b1h, b2h);
        // iota
        if (lane == 0)
        {
            lo ^= c_keccak_rc_lo[round];
            hi ^= c_keccak_rc_hi[round];
        }
    }
}

// the 4n-byte squeeze of SHAKE256(seed || LE64(counter)) by one warp: accepted words reduced into `row`, rejected
// ones raw, their indices (ascending) in `list`; returns how many were rejected
__device__ __forceinline__ uint32_t seb_coop_bulk_row(const uint8_t *__restrict__ seed, uint32_t counter,
                                                      uint32_t *__restrict__ row, uint16_t *__restrict__ list, int n,
                                                      const SebModulus &mod, uint32_t max_multiple, uint32_t cap,
                                                      const SebCoopLane &c, const int lane)
{
    // absorb seed || LE64(counter), pad (seb_prng_init): lanes 0..7 seed, 8 counter, 9 0x1F, 16 the final bit
    uint32_t lo = 0, hi = 0;
    if (lane < 8)
    {
        const uint2 w = __ldg(reinterpret_cast<const uint2 *>(seed) + lane);
        lo = w.x;
        hi = w.y;
    }
    else if (lane == 8)
        lo = counter;
    else if (lane == 9)
        lo = 0x1Fu;
    else if (lane == 16)
        hi = 0x80000000u;
    uint32_t cnt         = 0;
    const uint32_t below = (1u << lane) - 1u;
    for (int word = 0; word < n; word += 34)
    {
        seb_keccak_coop(lo, hi, c, lane);
        // lanes 0..16 hold the 136-byte rate block: words word + 2*lane, + 1
        const int w0     = word + 2 * lane;
        const bool valid = lane < 17 && w0 < n;  // n is even: the pair is valid or not as a whole
        const bool rl    = valid && lo >= max_multiple, rh = valid && hi >= max_multiple;
        const uint32_t ml = __ballot_sync(0xFFFFFFFFu, rl), mh = __ballot_sync(0xFFFFFFFFu, rh);
        if (valid)
        {
            // ascending word order = (lane 0 lo, lane 0 hi, lane 1 lo, ...)
            uint32_t pos = cnt + (uint32_t)__popc(ml & below) + (uint32_t)__popc(mh & below);
            if (rl)
            {
                if (pos < cap) list[pos] = (uint16_t)w0;
                pos++;
            }
            if (rh && pos < cap) list[pos] = (uint16_t)(w0 + 1);
            const uint32_t vl = rl ? lo : seb_barrett32(lo, mod), vh = rh ? hi : seb_barrett32(hi, mod);
            *reinterpret_cast<uint2 *>(row + w0) = make_uint2(vl, vh);
        }
        cnt += (uint32_t)__popc(ml) + (uint32_t)__popc(mh);
    }
    return cnt;
}

__global__ void __launch_bounds__(128) k_uniform_bulk_coop(const uint8_t *__restrict__ seeds,
                                                           const uint32_t *__restrict__ ctr, uint32_t *__restrict__ out,
                                                           size_t ct_stride, int n, SebModulus mod, uint32_t max_multiple,
                                                           int batch, uint16_t *__restrict__ rej_idx,
                                                           uint32_t *__restrict__ rej_cnt, uint32_t cap)
{
    const int lane = threadIdx.x & 31;
    const int b    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const SebCoopLane c = seb_coop_setup(lane);
    const uint32_t cnt  = seb_coop_bulk_row(seeds + (size_t)b * SEB_SEED_BYTES, ctr[b], out + (size_t)b * ct_stride,
                                            rej_idx + (size_t)b * cap, n, mod, max_multiple, cap, c, lane);
    if (lane == 0) rej_cnt[b] = cnt;
}

// ---------------------------------------------------------------------------------------------
// the bulk squeeze for MID-SIZE batches: TWO lanes per ciphertext, the Keccak state bit-interleaved
// ---------------------------------------------------------------------------------------------
// A shard of a few thousand to a few tens of thousands of ciphertexts (configuration D on eight GPUs: 16384 items)
// gives the thread-per-sponge kernel ONE warp per SM sub-partition: every instruction waits for the one before
// it (ncu: ALU pipe 57 %, profiles/README.md), and the 25-lane kernel above costs six times the issue slots.  Here a
// sponge is split over two adjacent lanes in the classic bit-interleaved representation: the EVEN lane holds bits
// 0,2,4,... of each of the 25 words (32 bits each), the ODD lane bits 1,3,5,...  Every logic step is bitwise, so
// each lane works on its own half; a 64-bit rotation by 2k is a 32-bit rotation by k of each half, and a rotation
// by 2k+1 swaps the halves: even' = rotl32(odd, k+1), odd' = rotl32(even, k).  Twelve of the 24 rho offsets and
// theta's rotation by one are odd: 17 exchanges (shfl.bfly 1) per round.  Per lane and round: 90 ALU operations
// (10 column parities, 5 + 24 rotations, 25 + 25 three-input logic, 1 iota) — the SAME 180 per sponge as one thread
// computes, on twice the warps with half the dependency depth.  The squeezed words are de-interleaved with one more
// exchange and a 4-step perfect shuffle each (the even lane rebuilds the low 32 bits of every rate word, the odd
// lane the high 32 bits).  Output, reject lists and counters are exactly those of k_uniform_bulk (sample.c:39-57).

// theta + rho + pi of word SRC on this lane's half; e = 1 on the even lane, 0 on the odd lane
#define SEB_PAIR_RP(SRC, DST, ROT)                                                               \
    {                                                                                            \
        const uint32_t t_ = seb_xor3(s[SRC], c[((SRC) % 5 + 4) % 5], r[((SRC) % 5 + 1) % 5]);    \
        if (((ROT) & 1) == 0)                                                                    \
            b[DST] = ((ROT) / 2) ? __funnelshift_l(t_, t_, (ROT) / 2) : t_;                       \
        else                                                                                     \
        {                                                                                        \
            const uint32_t p_ = __shfl_xor_sync(0xFFFFFFFFu, t_, 1);                             \
            b[DST]            = __funnelshift_l(p_, p_, (ROT) / 2 + e);                          \
        }                                                                                        \
    }

__device__ __forceinline__ void seb_keccak_pair(uint32_t (&s)[25], const uint32_t e)
{
#pragma unroll 1
    for (int round = 0; round < 24; round++)
    {
        uint32_t c[5], r[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = seb_xor3(seb_xor3(s[x], s[x + 5], s[x + 10]), s[x + 15], s[x + 20]);
        // rotl64(C, 1): even half = rotl32(odd half, 1), odd half = even half
#pragma unroll
        for (int x = 0; x < 5; x++)
        {
            const uint32_t pc = __shfl_xor_sync(0xFFFFFFFFu, c[x], 1);
            r[x]              = __funnelshift_l(pc, pc, e);
        }
        SEB_PAIR_RP(0, 0, 0) SEB_PAIR_RP(1, 10, 1) SEB_PAIR_RP(2, 20, 62) SEB_PAIR_RP(3, 5, 28) SEB_PAIR_RP(4, 15, 27)
        SEB_PAIR_RP(5, 16, 36) SEB_PAIR_RP(6, 1, 44) SEB_PAIR_RP(7, 11, 6) SEB_PAIR_RP(8, 21, 55) SEB_PAIR_RP(9, 6, 20)
        SEB_PAIR_RP(10, 7, 3) SEB_PAIR_RP(11, 17, 10) SEB_PAIR_RP(12, 2, 43) SEB_PAIR_RP(13, 12, 25) SEB_PAIR_RP(14, 22, 39)
        SEB_PAIR_RP(15, 23, 41) SEB_PAIR_RP(16, 8, 45) SEB_PAIR_RP(17, 18, 15) SEB_PAIR_RP(18, 3, 21) SEB_PAIR_RP(19, 13, 8)
        SEB_PAIR_RP(20, 14, 18) SEB_PAIR_RP(21, 24, 2) SEB_PAIR_RP(22, 9, 61) SEB_PAIR_RP(23, 19, 56) SEB_PAIR_RP(24, 4, 14)
#pragma unroll
        for (int y = 0; y < 25; y += 5)
#pragma unroll
            for (int x = 0; x < 5; x++) s[y + x] = seb_chi(b[y + x], b[y + (x + 1) % 5], b[y + (x + 2) % 5]);
        s[0] ^= e ? c_keccak_rc_even[round] : c_keccak_rc_odd[round];
    }
}
#undef SEB_PAIR_RP

__global__ void __launch_bounds__(32) k_uniform_bulk_pair(const uint8_t *__restrict__ seeds, const uint32_t *__restrict__ ctr,
                                                           uint32_t *__restrict__ out, size_t ct_stride, int n,
                                                           SebModulus mod, uint32_t max_multiple, int batch,
                                                           uint16_t *__restrict__ rej_idx, uint32_t *__restrict__ rej_cnt,
                                                           uint32_t cap)
{
    const int gt     = blockIdx.x * blockDim.x + threadIdx.x;
    const int odd    = gt & 1;         // which half of the state this lane holds
    const uint32_t e = 1u - (uint32_t)odd;
    const bool live  = (gt >> 1) < batch;
    const int b      = live ? (gt >> 1) : batch - 1;  // idle pairs of the last warp shadow the last ciphertext
    // absorb seed || LE64(counter), pad (seb_prng_init), this lane's half of every word
    uint32_t s[25];
    {
        const uint64_t *sp = reinterpret_cast<const uint64_t *>(seeds + (size_t)b * SEB_SEED_BYTES);
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = seb_half_bits(__ldg(sp + i), odd);
        s[8] = seb_half_bits((uint64_t)ctr[b], odd);
        s[9] = odd ? 0x3u : 0x7u;  // 0x1F: bits 0, 2, 4 | bits 1, 3
#pragma unroll
        for (int i = 10; i < 25; i++) s[i] = 0u;
        s[16] = odd ? 0x80000000u : 0u;  // bit 63
    }
    uint32_t *row   = out + (size_t)b * ct_stride;
    uint16_t *list  = rej_idx + (size_t)b * cap;
    uint32_t cnt    = 0;
    // De-interleave: the even lane rebuilds the LOW 32 bits of a rate word from the low 16 bits of the two halves, the
    // odd lane the HIGH 32 bits from their high 16 bits.  One byte permute of (own, partner) picks the four bytes in
    // the order [even.lo, odd.lo, even.hi, odd.hi] of those 16-bit values; own is the even half on the even lane
    // and the odd half on the odd lane, hence the two selectors.
    const uint32_t sel = odd ? 0x3726u : 0x5140u;
    for (int word = 0; word < n; word += 34)
    {
        seb_keccak_pair(s, e);
        uint32_t rm       = 0;  // bit k: this lane's word of rate lane k was rejected
        // the rate block is 17 (even, odd) word pairs; n is even, so only the LAST block of a row is partial and a pair
        // is inside or outside as a whole: nk = pairs of this block that belong to the row.  Idle lanes store nothing.
        const int nk  = live ? min(17, (n - word) >> 1) : 0;
        uint32_t *dst = row + word + odd;
#pragma unroll
        for (int k = 0; k < 17; k++)
        {
            const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, s[k], 1);
            const uint32_t w     = seb_interleave_tail(__byte_perm(s[k], other, sel));
            const bool in        = k < nk;
            const bool rej       = in && w >= max_multiple;
            rm |= rej ? (1u << k) : 0u;
            const uint32_t red = seb_barrett32(w, mod);
            if (in) dst[2 * k] = rej ? w : red;
        }
        // ascending word order within the block: (even lane k = 0, odd lane k = 0, even lane k = 1, ...)
        const uint32_t pm = __shfl_xor_sync(0xFFFFFFFFu, rm, 1);
        uint32_t m        = rm;
        while (m)
        {
            const int k          = __ffs(m) - 1;
            const uint32_t below = (1u << k) - 1u;
            const uint32_t pos   = cnt + (uint32_t)__popc(rm & below) + (uint32_t)__popc(pm & (odd ? (below << 1) | 1u : below));
            if (live && pos < cap) list[pos] = (uint16_t)(word + 2 * k + odd);
            m &= m - 1;
        }
        cnt += (uint32_t)__popc(rm) + (uint32_t)__popc(pm);
    }
    if (live && !odd) rej_cnt[b] = cnt;
}

// ---------------------------------------------------------------------------------------------
// lone calls (a handful of ciphertexts): the primes' squeezes in PARALLEL, speculating on the counters
// ---------------------------------------------------------------------------------------------
// The squeeze of prime p starts at counter c_p = p + (redraw calls of primes < p), which is only known once those
// primes are done — but it is sharply distributed: the redraw calls of prime i have mean n r_i / (1 - r_i) and
// variance n r_i / (1 - r_i)^2, r_i = the word rejection rate under q_i (1-2 %).  For a few ciphertexts the machine
// is empty, so EVERY counter within +-5 sigma of the mean is squeezed at once (~90 candidates for prime 1 at
// n = 4096, ~1500 over the five later primes at n = 16384), one warp each, beside prime 0; afterwards the chain is
// resolved prime by prime: k_uniform_select copies the candidate the true counter points at into the output (or,
// outside the window, squeezes it on the spot) and the usual fix-up follows.  Same bytes, same counters; the
// dependent work shrinks from all primes' squeezes to one.
// (SebSpecPrime / SebSpecPlan: seb_kernels.h — per prime p >= 1 the first speculated counter `lo`, how many
// counters `width`, and `first`, the index of its first candidate among the `total` candidates of a ciphertext)

// grid: warp (b, k), k = 0: prime 0 at counter 0 into the output row; k >= 1: candidate k - 1
__global__ void __launch_bounds__(128)
    k_uniform_spec_bulk(const uint8_t *__restrict__ seeds, uint32_t *__restrict__ out, size_t ct_stride, size_t p_stride,
                        int n, const __grid_constant__ SebModuli mods, int np, const __grid_constant__ SebSpecPlan plan,
                        int batch, uint32_t *__restrict__ cand_rows, uint16_t *__restrict__ cand_list,
                        uint32_t *__restrict__ cand_cnt, uint16_t *__restrict__ rej_idx, uint32_t *__restrict__ rej_cnt,
                        uint32_t cap)
{
    const int lane       = threadIdx.x & 31;
    const size_t warp    = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t per_ct  = (size_t)plan.total + 1;
    if (warp >= per_ct * (size_t)batch) return;
    const size_t b       = warp / per_ct;
    const uint32_t k     = (uint32_t)(warp % per_ct);
    const SebCoopLane c  = seb_coop_setup(lane);
    const uint8_t *seed  = seeds + b * SEB_SEED_BYTES;
    if (k == 0)
    {
        const SebModulus m  = mods.m[0];
        const uint32_t maxm = 0xFFFFFFFFu - (0xFFFFFFFFu % m.q) - 1u;
        const uint32_t cnt  = seb_coop_bulk_row(seed, 0u, out + b * ct_stride, rej_idx + b * cap, n, m, maxm, cap, c, lane);
        if (lane == 0) rej_cnt[b] = cnt;
        return;
    }
    int p = 1;
    while (p + 1 < np && k - 1 >= plan.p[p + 1].first) p++;
    const uint32_t j    = k - 1 - plan.p[p].first;
    const SebModulus m  = mods.m[p];
    const uint32_t maxm = 0xFFFFFFFFu - (0xFFFFFFFFu % m.q) - 1u;
    const size_t slot   = b * plan.total + (k - 1);
    const uint32_t cnt  = seb_coop_bulk_row(seed, plan.p[p].lo + j, cand_rows + slot * (size_t)n, cand_list + slot * cap, n, m,
                                            maxm, cap, c, lane);
    if (lane == 0) cand_cnt[slot] = cnt;
}

// warp per ciphertext, before the fix-up of prime p >= 1: bring the squeeze that starts at the TRUE counter
// ctr[b] into the output row and the fix-up's reject list
__global__ void __launch_bounds__(128)
    k_uniform_select(const uint8_t *__restrict__ seeds, const uint32_t *__restrict__ ctr, uint32_t *__restrict__ out_p,
                     size_t ct_stride, int n, SebModulus mod, uint32_t max_multiple, SebSpecPrime sp, uint32_t total,
                     int batch, const uint32_t *__restrict__ cand_rows, const uint16_t *__restrict__ cand_list,
                     const uint32_t *__restrict__ cand_cnt, uint16_t *__restrict__ rej_idx, uint32_t *__restrict__ rej_cnt,
                     uint32_t cap, uint32_t *__restrict__ misses)
{
    const int lane = threadIdx.x & 31;
    const int b    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const uint32_t c0 = ctr[b];
    uint32_t *row     = out_p + (size_t)b * ct_stride;
    uint16_t *list    = rej_idx + (size_t)b * cap;
    if (c0 >= sp.lo && c0 - sp.lo < sp.width)
    {
        const size_t slot   = (size_t)b * total + sp.first + (c0 - sp.lo);
        const uint4 *src    = reinterpret_cast<const uint4 *>(cand_rows + slot * (size_t)n);
        uint4 *dst          = reinterpret_cast<uint4 *>(row);
        for (int i = lane; i < n / 4; i += 32) dst[i] = src[i];
        const uint32_t cnt  = cand_cnt[slot];
        const uint32_t keep = cnt < cap ? cnt : cap;
        for (uint32_t i = lane; i < keep; i += 32) list[i] = cand_list[slot * cap + i];
        if (lane == 0) rej_cnt[b] = cnt;
    }
    else
    {
        // outside the speculated window (probability ~6e-7 per prime at 5 sigma): squeeze it now
        const SebCoopLane c = seb_coop_setup(lane);
        const uint32_t cnt  = seb_coop_bulk_row(seeds + (size_t)b * SEB_SEED_BYTES, c0, row, list, n, mod, max_multiple, cap, c, lane);
        if (lane == 0)
        {
            rej_cnt[b] = cnt;
            if (misses) atomicAdd(misses, 1u);
        }
    }
}

// Warp per ciphertext: the k-th rejected index (ascending) receives the k-th accepted
// candidate LE32(X(seed, c0+1+t, 4)), t = 0,1,...; the counter ends one past the last candidate
// consumed (sample.c:49-56).  Candidates are generated 32 counters at a time; with the reject list
// of the bulk kernel every accepted candidate of a wave is placed in parallel (its rank among the
// accepted candidates so far selects the list entry).  A list that overflowed its capacity (never
// with SHAKE output and cap = n/8, but it must stay correct) falls back to scanning the row.
// the fix-up of ciphertext b by one warp
__device__ __forceinline__ void seb_uniform_fix_warp(const int b, const int lane, const uint8_t *__restrict__ seeds,
                                                     uint32_t *__restrict__ ctr, uint32_t *__restrict__ out,
                                                     size_t ct_stride, int n, const SebModulus &mod, uint32_t max_multiple,
                                                     const uint16_t *__restrict__ rej_idx,
                                                     const uint32_t *__restrict__ rej_cnt, uint32_t cap)
{
    uint32_t *row = out + (size_t)b * ct_stride;
    uint32_t se[8], so[8];  // every candidate is one permutation on the bit-interleaved state (seb_prng_word_il)
    seb_seed_split_warp(seeds, (size_t)b, lane, se, so);

    const uint64_t c0  = (uint64_t)ctr[b];
    uint64_t wave_base = c0 + 1;  // counter of lane 0's candidate in the next wave to generate
    uint64_t last_used = c0;      // counter of the last candidate consumed
    const uint32_t cnt = rej_cnt[b];

    if (cnt <= cap)
    {
        const uint16_t *list = rej_idx + (size_t)b * cap;
        uint32_t done        = 0;  // rejected words already replaced
        while (done < cnt)
        {
            const uint32_t cand  = seb_prng_word_il(se, so, wave_base + (uint64_t)lane);  // 4 bytes per call
            const bool ok        = cand < max_multiple;
            const uint32_t avail = __ballot_sync(0xFFFFFFFFu, ok);
            const uint32_t rank  = done + (uint32_t)__popc(avail & ((1u << lane) - 1u));  // this candidate's turn
            const bool used      = ok && rank < cnt;
            if (used) row[list[rank]] = seb_barrett32(cand, mod);
            const uint32_t um = __ballot_sync(0xFFFFFFFFu, used);
            if (um) last_used = wave_base + (uint64_t)(31 - __clz(um));
            done += (uint32_t)__popc(um);
            wave_base += 32;
        }
    }
    else
    {
        uint32_t avail    = 0;  // accepted, unused candidates of the current wave (bit = lane)
        uint32_t cand     = 0;
        uint64_t cur_base = 0;
        for (int base = 0; base < n; base += 32)
        {
            const uint32_t w = row[base + lane];
            uint32_t rm      = __ballot_sync(0xFFFFFFFFu, w >= max_multiple);
            while (rm)
            {
                while (avail == 0)
                {
                    cand     = seb_prng_word_il(se, so, wave_base + (uint64_t)lane);
                    avail    = __ballot_sync(0xFFFFFFFFu, cand < max_multiple);
                    cur_base = wave_base;
                    wave_base += 32;
                }
                const int rp     = __ffs(rm) - 1;
                const int cl     = __ffs(avail) - 1;
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, cand, cl);
                if (lane == rp) row[base + rp] = seb_barrett32(v, mod);
                rm &= rm - 1;
                avail &= avail - 1;
                last_used = cur_base + (uint64_t)cl;
            }
        }
    }
    if (lane == 0) ctr[b] = (uint32_t)(last_used + 1);
}

__global__ void __launch_bounds__(128) k_uniform_fix(const uint8_t *__restrict__ seeds, uint32_t *__restrict__ ctr,
                                                     uint32_t *__restrict__ out, size_t ct_stride, int n,
                                                     SebModulus mod, uint32_t max_multiple, int batch,
                                                     const uint16_t *__restrict__ rej_idx,
                                                     const uint32_t *__restrict__ rej_cnt, uint32_t cap)
{
    const int lane = threadIdx.x & 31;
    const int b    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    seb_uniform_fix_warp(b, lane, seeds, ctr, out, ct_stride, n, mod, max_multiple, rej_idx, rej_cnt, cap);
}

// The same fix-up with L lanes per ciphertext (32/L ciphertexts per warp), for parameter sets that reject a handful of
// words: a wave of 32 candidates per ciphertext computes 32 permutations for the ~1.5 rejections of n = 1024 under a
// 27-bit prime - the fix-up cost as much as the whole squeeze of configuration A.  Candidates of ciphertext b are still LE32(X(seed, c0+1+t, 4)), t = 0, 1, ... in order; a group
// draws L of them per wave and stops when its list is served.  Ciphertexts whose reject list overflowed (cnt > cap)
// are handed, one after the other, to the scanning path of seb_uniform_fix_warp by the whole warp.
template <int L>
__global__ void __launch_bounds__(128) k_uniform_fix_sub(const uint8_t *__restrict__ seeds, uint32_t *__restrict__ ctr,
                                                         uint32_t *__restrict__ out, size_t ct_stride, int n, SebModulus mod,
                                                         uint32_t max_multiple, int batch,
                                                         const uint16_t *__restrict__ rej_idx,
                                                         const uint32_t *__restrict__ rej_cnt, uint32_t cap)
{
    constexpr int G   = 32 / L;  // ciphertexts per warp
    const int lane    = threadIdx.x & 31;
    const int group   = lane / L, sub = lane % L;
    const int warp    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int b0      = warp * G;
    if (b0 >= batch) return;
    const int b       = b0 + group;
    const bool live   = b < batch;
    const uint32_t cnt = live ? rej_cnt[b] : 0u;
    const bool scan    = cnt > cap;  // list overflow: the one-warp scanning path below
    const uint32_t gmask = (L == 32 ? 0xFFFFFFFFu : ((1u << L) - 1u)) << (group * L);
    const uint32_t below = gmask & ((1u << lane) - 1u);

    uint32_t se[8], so[8];
    seb_seed_split(seeds, (size_t)(live ? b : b0), se, so);
    const uint64_t c0    = live ? (uint64_t)ctr[b] : 0ull;
    uint64_t wave_base   = c0 + 1;  // counter of this group's first candidate in the next wave
    uint64_t last_used   = c0;
    const uint32_t want  = scan ? 0u : cnt;  // rejected words this group's waves serve
    const uint16_t *list = rej_idx + (size_t)(live ? b : b0) * cap;
    uint32_t *row        = out + (size_t)(live ? b : b0) * ct_stride;
    uint32_t done        = 0;
    while (__any_sync(0xFFFFFFFFu, done < want))
    {
        const uint32_t cand  = seb_prng_word_il(se, so, wave_base + (uint64_t)sub);  // 4 bytes per call
        const bool ok        = done < want && cand < max_multiple;
        const uint32_t avail = __ballot_sync(0xFFFFFFFFu, ok);
        const uint32_t rank  = done + (uint32_t)__popc(avail & below);  // this candidate's turn within its ciphertext
        const bool used      = ok && rank < want;
        if (used) row[list[rank]] = seb_barrett32(cand, mod);
        const uint32_t um = __ballot_sync(0xFFFFFFFFu, used) & gmask;
        if (um) last_used = wave_base + (uint64_t)(31 - __clz(um) - group * L);
        if (done < want) wave_base += L;
        done += (uint32_t)__popc(um);
    }
    if (live && !scan && sub == 0) ctr[b] = (uint32_t)(last_used + 1);
    // overflowed lists (never with SHAKE output and cap = n/8, but it must stay correct): whole warp, one at a time
    const uint32_t scans = __ballot_sync(0xFFFFFFFFu, scan && sub == 0);
    for (uint32_t m = scans; m; m &= m - 1)
    {
        const int g = (__ffs(m) - 1) / L;
        seb_uniform_fix_warp(b0 + g, lane, seeds, ctr, out, ct_stride, n, mod, max_multiple, rej_idx, rej_cnt, cap);
    }
}

// The fix-up as a STREAM: one warp serves K consecutive ciphertexts, and a wave's 32 candidates go to the current
// ciphertext only as far as it still has rejected words to fill - the lanes behind them already draw the first
// candidates of the next ciphertext.  A warp per ciphertext (k_uniform_fix) computes whole waves, so a ciphertext that
// needs 76 +- 9 candidates (n = 4096, 30-bit primes) pays for 96 and one that needs 250 (n = 16384) for 270-290; here
// the only partly used wave is the last one of the K-th ciphertext.  Candidates of ciphertext b are still
// LE32(X(seed_b, c0 + 1 + t, 4)), t = 0, 1, ... consumed in order (sample.c:39-57); candidates drawn beyond the last one
// consumed are dropped, the counter continues behind the last one consumed.
template <int K>
__global__ void __launch_bounds__(128) k_uniform_fix_stream(const uint8_t *__restrict__ seeds, uint32_t *__restrict__ ctr,
                                                            uint32_t *__restrict__ out, size_t ct_stride, int n,
                                                            SebModulus mod, uint32_t max_multiple, int batch,
                                                            const uint16_t *__restrict__ rej_idx,
                                                            const uint32_t *__restrict__ rej_cnt, uint32_t cap)
{
    constexpr uint32_t FULL = 0xFFFFFFFFu;
    const int lane  = threadIdx.x & 31;
    const int first = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * K;
    if (first >= batch) return;
    const int last = min(first + K, batch);  // ciphertexts [first, last)
    const uint32_t below = (1u << lane) - 1u;

    // A = the ciphertext being filled, B = the one behind it (bB == last: none)
    int bA = first, bB = first;
    uint32_t seA[8], soA[8], seB[8], soB[8];
    uint32_t cntA = 0, doneA = 0, cntB = 0, doneB = 0;
    uint64_t baseA = 0, lastA = 0, baseB = 0, lastB = 0;  // next candidate's counter, counter of the last one consumed
    // fetches the ciphertext behind B into B; overflowed reject lists (never with SHAKE output and cap = n/8, but it must
    // stay correct) are served on the spot by the scanning path and skipped
    auto fetch = [&](int from) {
        int b = from;
        while (b < last)
        {
            const uint32_t c = rej_cnt[b];
            if (c <= cap)
            {
                cntB = c, doneB = 0;
                lastB = (uint64_t)ctr[b], baseB = lastB + 1;
                seb_seed_split_warp(seeds, (size_t)b, lane, seB, soB);
                break;
            }
            seb_uniform_fix_warp(b, lane, seeds, ctr, out, ct_stride, n, mod, max_multiple, rej_idx, rej_cnt, cap);
            b++;
        }
        bB = b;
    };
    auto shift = [&]() {  // B becomes A
        bA = bB, cntA = cntB, doneA = doneB, baseA = baseB, lastA = lastB;
#pragma unroll
        for (int i = 0; i < 8; i++) seA[i] = seB[i], soA[i] = soB[i];
        fetch(bB + 1);
    };
    fetch(first);
    shift();
    while (bA < last)
    {
        if (doneA == cntA)  // warp-uniform: A is served (or had nothing rejected)
        {
            __syncwarp();
            if (lane == 0) ctr[bA] = (uint32_t)(lastA + 1);
            shift();
            continue;
        }
        const uint32_t s   = min(32u, cntA - doneA);                       // lanes [0, s) draw for A
        const uint32_t nB  = bB < last ? min(32u - s, cntB - doneB) : 0u;  // lanes [s, s + nB) for B
        const bool forA    = (uint32_t)lane < s;
        const bool forB    = !forA && (uint32_t)lane < s + nB;
        uint32_t se[8], so[8];
#pragma unroll
        for (int i = 0; i < 8; i++) se[i] = forA ? seA[i] : seB[i], so[i] = forA ? soA[i] : soB[i];
        const uint64_t counter = forA ? baseA + (uint64_t)lane : baseB + (uint64_t)((uint32_t)lane - s);
        const uint32_t cand    = seb_prng_word_il(se, so, counter);  // 4 bytes per call
        const bool ok          = cand < max_multiple;
        const uint32_t okA     = __ballot_sync(FULL, ok && forA), okB = __ballot_sync(FULL, ok && forB);
        // every accepted candidate is consumed: A's s lanes never exceed what A still needs, nor do B's nB
        if (ok && forA)
            (out + (size_t)bA * ct_stride)[(rej_idx + (size_t)bA * cap)[doneA + (uint32_t)__popc(okA & below)]] =
                seb_barrett32(cand, mod);
        if (ok && forB)
            (out + (size_t)bB * ct_stride)[(rej_idx + (size_t)bB * cap)[doneB + (uint32_t)__popc(okB & below)]] =
                seb_barrett32(cand, mod);
        if (okA) lastA = baseA + (uint64_t)(31 - __clz(okA));
        if (okB) lastB = baseB + (uint64_t)(31 - __clz(okB)) - (uint64_t)s;
        doneA += (uint32_t)__popc(okA), baseA += s;
        doneB += (uint32_t)__popc(okB), baseB += nB;
    }
}

// The fix-up for a handful of ciphertexts: one CTA per ciphertext, every thread one candidate, so that the ~n/50
// candidates a polynomial needs come out of ONE round of permutations instead of n/1600 dependent 32-candidate
// waves (10 at n = 16384: 69 us of a lone call's 85 us per prime).  Ranks are counted across the CTA (ballot per
// warp, prefix over the warps); a second round only if the first did not yield enough accepted candidates.
// Ciphertexts whose reject list overflowed take the one-warp scanning path.
__global__ void __launch_bounds__(512) k_uniform_fix_wide(const uint8_t *__restrict__ seeds, uint32_t *__restrict__ ctr,
                                                          uint32_t *__restrict__ out, size_t ct_stride, int n,
                                                          SebModulus mod, uint32_t max_multiple, int batch,
                                                          const uint16_t *__restrict__ rej_idx,
                                                          const uint32_t *__restrict__ rej_cnt, uint32_t cap)
{
    __shared__ uint32_t s_count[16];
    __shared__ uint32_t s_last;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (b >= batch) return;
    const uint32_t cnt = rej_cnt[b];
    if (cnt > cap)  // CTA-uniform
    {
        if (warp == 0) seb_uniform_fix_warp(b, lane, seeds, ctr, out, ct_stride, n, mod, max_multiple, rej_idx, rej_cnt, cap);
        return;
    }
    uint32_t *row        = out + (size_t)b * ct_stride;
    const uint16_t *list = rej_idx + (size_t)b * cap;
    uint32_t se[8], so[8];
    seb_seed_split_warp(seeds, (size_t)b, lane, se, so);
    const uint32_t c0  = ctr[b];
    uint32_t wave_base = c0 + 1;  // counter of thread 0's candidate in the next round
    uint32_t done      = 0;
    if (tid == 0) s_last = c0;
    __syncthreads();  // ctr[b] is read by everybody before thread 0 overwrites it at the end
    while (done < cnt)
    {
        const uint32_t cand  = seb_prng_word_il(se, so, (uint64_t)wave_base + (uint64_t)tid);  // 4 bytes per call
        const bool ok        = cand < max_multiple;
        const uint32_t avail = __ballot_sync(0xFFFFFFFFu, ok);
        if (lane == 0) s_count[warp] = (uint32_t)__popc(avail);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < nwarps; w++)
        {
            const uint32_t cw = s_count[w];
            before += w < warp ? cw : 0u;
            total += cw;
        }
        const uint32_t rank = done + before + (uint32_t)__popc(avail & ((1u << lane) - 1u));  // this candidate's turn
        const bool used     = ok && rank < cnt;
        if (used) row[list[rank]] = seb_barrett32(cand, mod);
        const uint32_t um = __ballot_sync(0xFFFFFFFFu, used);
        if (um && lane == 0) atomicMax(&s_last, wave_base + (uint32_t)(warp * 32 + 31 - __clz(um)));
        done += total < cnt - done ? total : cnt - done;
        wave_base += blockDim.x;
        __syncthreads();  // s_count is rewritten by the next round; s_last is complete
    }
    if (tid == 0) ctr[b] = s_last + 1;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
void seb_launch_prng_blocks(const uint8_t *seeds, const uint64_t *counters, uint64_t *out, int count,
                            cudaStream_t st)
{
    if (count <= 0) return;
    k_prng_blocks<<<(count + 127) / 128, 128, 0, st>>>(seeds, counters, out, count);
}

void seb_launch_sample_ternary(const uint8_t *seeds, uint8_t *u_out, uint32_t *ctr_out, int n, int batch,
                               cudaStream_t st)
{
    if (batch <= 0) return;
    const int warps = 4;
    if (n == 4096 && batch >= 2)
    {
        // two ciphertexts per warp sharing their third wave (the degree where a third of the last wave is unused)
        const int pairs = (batch + 1) / 2;
        k_sample_ternary_pair<<<(pairs + warps - 1) / warps, warps * 32, (size_t)2 * warps * (n / 4), st>>>(seeds, u_out, ctr_out, n,
                                                                                                       batch);
        return;
    }
    const size_t sm = (size_t)warps * (n / 4);
    k_sample_ternary<<<(batch + warps - 1) / warps, warps * 32, sm, st>>>(seeds, u_out, ctr_out, n, batch);
}

void seb_launch_sample_cbd(const uint8_t *seeds, const uint32_t *ctr_base, int8_t *e_out, int n, int npoly,
                           int batch, cudaStream_t st)
{
    if (batch <= 0) return;
    const int per_ct  = npoly * (n / 16);  // a multiple of 32: the degrees are powers of two >= 512
    const int threads = per_ct % 128 == 0 ? 128 : per_ct % 64 == 0 ? 64 : 32;
    k_sample_cbd<<<dim3((unsigned)batch, (unsigned)(per_ct / threads)), threads, 0, st>>>(seeds, ctr_base, e_out, n, npoly,
                                                                                         batch);
}

// fix-up launch: a CTA per ciphertext with one round of ~n/50 + 32 candidates for a handful of ciphertexts (latency),
// a warp per ciphertext with 32-candidate waves otherwise (throughput).  knobs.uniform_fix_wide = 0/1 forces the choice.
#define SEB_FIX_WIDE_MAX_BATCH 64
static void seb_launch_uniform_fix(const uint8_t *seeds, uint32_t *ctr, uint32_t *out, size_t ct_stride, int n,
                                   const SebModulus &mod, uint32_t max_multiple, int batch, uint16_t *rej_idx,
                                   uint32_t *rej_cnt, uint32_t rej_cap, const SebKnobs &knobs, cudaStream_t st)
{
    const bool wide = knobs.uniform_fix_wide >= 0 ? knobs.uniform_fix_wide != 0 : batch <= SEB_FIX_WIDE_MAX_BATCH;
    if (wide)
    {
        // expected rejections: 2 % of n at most (30-bit primes), plus head-room for the candidates' own rejections
        int threads = ((n / 50 + 32 + 31) / 32) * 32;
        if (threads > 512) threads = 512;
        k_uniform_fix_wide<<<batch, threads, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx, rej_cnt,
                                                      rej_cap);
    }
    else
    {
        // Lanes per ciphertext by the expected number of rejected words, n x the prime's rejection rate: 1.5 and 3 at
        // n = 1024 / 2048 under the 27-bit prime (a wave of 32 would compute 32 permutations for them), 76 and more
        // under 30-bit primes at n >= 4096.  A warp of G groups runs until its slowest group is served, which costs about
        // mean + z_G sd + L/2 permutations per ciphertext: at 76 +- 9 every L lands on 91-94, so only the small
        // expectations leave the full-warp form.  knobs.uniform_fix_lanes forces 4 / 8 / 32.
        const double expect = (double)n * (4294967296.0 - (double)max_multiple) / 4294967296.0;
        const int lanes     = knobs.uniform_fix_lanes > 0 ? knobs.uniform_fix_lanes : expect < 2.5 ? 4 : expect < 7.0 ? 8 : 32;
        if (lanes == 4)
            k_uniform_fix_sub<4><<<(batch + 31) / 32, 128, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx,
                                                                    rej_cnt, rej_cap);
        else if (lanes == 8)
            k_uniform_fix_sub<8><<<(batch + 15) / 16, 128, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx,
                                                                    rej_cnt, rej_cap);
        else
        {
            // Ciphertexts per warp of the streamed form: the one partly used wave is shared by K of them, but the warps
            // must still fill the machine - 2048 warps of 8 ciphertexts (configuration D's 16384-item shard) ran 5 %
            // SLOWER than 16384 warps of one (3.5 warps per sub-partition, no balancing left).  profiles/
            // r02_ab_fix_stream.txt: the best K is the largest of 8 / 4 that keeps ~7 warps per SM sub-partition (K = 2
            // never pays), and only where the unused half wave is a sizeable part of a ciphertext's candidates: at ~76
            // of them (n = 4096) the streamed form takes 4 % off the whole sampler at 65536 items, at ~150 (n = 8192)
            // 1.5 %, at ~250 (n = 16384) nothing - its own overhead (two seeds per wave, more registers) cancels it.
            // knobs.uniform_fix_stream forces K = 2 / 4 / 8 (2, 4, any other positive value) or the plain form (0).
            const int sms      = knobs.sms > 0 ? knobs.sms : 148;
            const int min_warps = 6 * 4 * sms;
            int K = 1;
            if (knobs.uniform_fix_stream >= 0)
                K = knobs.uniform_fix_stream == 2 || knobs.uniform_fix_stream == 4 ? knobs.uniform_fix_stream
                    : knobs.uniform_fix_stream                                     ? 8
                                                                                   : 1;
            else if (expect <= 160.0)
                for (int k = 8; k >= 4; k >>= 1)
                    if (batch / k >= min_warps)
                    {
                        K = k;
                        break;
                    }
            const int blocks = ((batch + K - 1) / K + 3) / 4;
            if (K == 8)
                k_uniform_fix_stream<8><<<blocks, 128, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx,
                                                                rej_cnt, rej_cap);
            else if (K == 4)
                k_uniform_fix_stream<4><<<blocks, 128, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx,
                                                                rej_cnt, rej_cap);
            else if (K == 2)
                k_uniform_fix_stream<2><<<blocks, 128, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx,
                                                                rej_cnt, rej_cap);
            else
                k_uniform_fix<<<blocks, 128, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx, rej_cnt,
                                                      rej_cap);
        }
    }
}

// Speculation windows for a parameter set (host, once per context).  sigmas: half-width in standard deviations.
void seb_uniform_spec_plan(int n, const SebModuli &mods, int np, double sigmas, SebSpecPlan *plan)
{
    double mean = 0.0, var = 0.0;
    uint32_t first = 0;
    memset(plan, 0, sizeof *plan);
    for (int p = 1; p < np; p++)
    {
        const uint32_t q    = mods.m[p - 1].q;
        const uint32_t maxm = 0xFFFFFFFFu - (0xFFFFFFFFu % q) - 1u;
        const double r      = (4294967296.0 - (double)maxm) / 4294967296.0;  // word rejection rate under prime p-1
        mean += 1.0 + (double)n * r / (1.0 - r);                               // its squeeze + its redraw calls
        var += (double)n * r / ((1.0 - r) * (1.0 - r));
        const double half = sigmas * sqrt(var) + 2.0;
        const double lo   = mean - half;
        plan->p[p].lo     = lo < (double)p ? (uint32_t)p : (uint32_t)lo;
        plan->p[p].width  = (uint32_t)(mean + half) - plan->p[p].lo + 1u;
        plan->p[p].first  = first;
        first += plan->p[p].width;
    }
    plan->total = first;
}

// The whole chain for a handful of ciphertexts: every prime's squeeze in one launch (prime 0 and the speculated
// counters of the others), then select + fix-up per prime.  out_p0 = row of (item 0, prime 0); prime p's rows are
// p_stride words further.  cand_* = scratch for batch * plan.total candidates.  Leaves ctr[b] = the final counter.
void seb_launch_uniform_chain_spec(const uint8_t *seeds, uint32_t *ctr, uint32_t *out_p0, size_t ct_stride, size_t p_stride,
                                   int n, const SebModuli &mods, int np, const SebSpecPlan &plan, int batch,
                                   uint32_t *cand_rows, uint16_t *cand_list, uint32_t *cand_cnt, uint16_t *rej_idx,
                                   uint32_t *rej_cnt, uint32_t rej_cap, uint32_t *misses, const SebKnobs &knobs,
                                   cudaStream_t st)
{
    if (batch <= 0) return;
    const size_t warps = ((size_t)plan.total + 1) * (size_t)batch;
    k_uniform_spec_bulk<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(seeds, out_p0, ct_stride, p_stride, n, mods, np, plan, batch,
                                                                    cand_rows, cand_list, cand_cnt, rej_idx, rej_cnt, rej_cap);
    for (int p = 0; p < np; p++)
    {
        const SebModulus &mod       = mods.m[p];
        const uint32_t max_multiple = 0xFFFFFFFFu - (0xFFFFFFFFu % mod.q) - 1u;
        uint32_t *out_p             = out_p0 + (size_t)p * p_stride;
        if (p > 0)
            k_uniform_select<<<(batch + 3) / 4, 128, 0, st>>>(seeds, ctr, out_p, ct_stride, n, mod, max_multiple, plan.p[p],
                                                              plan.total, batch, cand_rows, cand_list, cand_cnt, rej_idx,
                                                              rej_cnt, rej_cap, misses);
        seb_launch_uniform_fix(seeds, ctr, out_p, ct_stride, n, mod, max_multiple, batch, rej_idx, rej_cnt, rej_cap, knobs,
                               st);
    }
}

void seb_launch_uniform(const uint8_t *seeds, uint32_t *ctr, uint32_t *out, size_t ct_stride, int n,
                        const SebModulus &mod, int batch, uint16_t *rej_idx, uint32_t *rej_cnt, uint32_t rej_cap,
                        const SebKnobs &knobs, cudaStream_t st)
{
    if (batch <= 0) return;
    // max_multiple = 0xFFFFFFFF - (0xFFFFFFFF mod q) - 1 (sample.c:45-46)
    const uint32_t max_multiple = 0xFFFFFFFFu - (0xFFFFFFFFu % mod.q) - 1u;
    // Small batches: a warp per ciphertext with the sponge spread over its lanes (latency of the dependent
    // permutations / 4); otherwise one sequential sponge per thread, in single-warp CTAs that spread the batch over
    // all SM sub-partitions (131072-item config D leaves 16384 items = 512 warps per GPU for 592 sub-partitions).
    // knobs.uniform_coop = 0/1 forces either (tests, A/B measurements).
    const bool coop = knobs.uniform_coop >= 0 ? knobs.uniform_coop != 0 : batch <= SEB_UNIFORM_COOP_MAX_BATCH;
    // Between the two: two lanes per sponge (bit-interleaved halves).  Both kernels are latency machines at these
    // sizes: a launch takes as long as the most loaded SM sub-partition, kt = ceil(warps / sub-partitions) warps of 32
    // sponges for the thread kernel, kp for the pair kernel's warps of 16.  A linear fit of the measured chain times
    // (profiles/r02_ab_uniform_pair.txt: thread 23 / 41 / 85 ms for kt = 1 / 2 / 4 at n = 16384 x 6, pair 15.8 / 25 /
    // 36.6 / 49 for kp = 1..4) picks the kernel: pair while 12 kp + 3.5 < 21 kt + 2, i.e. up to one pair warp per
    // sub-partition (9472 items on 148 SMs) and again in the windows where the thread kernel has just spilled into
    // a second or third warp.  knobs.uniform_pair = 0/1 forces the choice.
    const int subparts = 4 * (knobs.sms > 0 ? knobs.sms : 148);
    const int kt       = ((batch + 31) / 32 + subparts - 1) / subparts;
    const int kp       = ((batch + 15) / 16 + subparts - 1) / subparts;
    bool pair          = knobs.uniform_pair >= 0 ? knobs.uniform_pair != 0 : (!coop && 12.0 * kp + 3.5 < 21.0 * kt + 2.0);
    // The thread kernel's warps come in layers of one per sub-partition, and the launch lasts as long as the sub-partitions
    // with the most warps: 65536 sponges are 3.46 layers, the machine idles 13 % of the squeeze.  MIXED: the full layers
    // go to the thread kernel and the sponges beyond them to the two-lane kernel on a second stream - one two-lane warp (16
    // sponges, 0.6 of a thread warp's time) on top of k thread warps instead of a (k+1)-th thread warp.  Same sponges,
    // same outputs; knobs.uniform_mix = 0/1 forces it off / on where it applies.
    const int warps_t = (batch + 31) / 32, layers = warps_t / subparts;
    const int full    = layers * subparts * 32;  // sponges of the full layers
    const int rest    = batch - full;
    const int kr      = ((rest + 15) / 16 + subparts - 1) / subparts;  // two-lane warps per sub-partition for the rest
    // by the same linear fit: `layers` thread warps and kr two-lane warps on the fullest sub-partition
    const double cost_mix = 21.0 * layers + 12.0 * kr + 3.5;
    bool mix = !coop && knobs.aux_stream && knobs.uniform_mix != 0 && knobs.uniform_pair < 0 && layers >= 1 && rest >= 512 &&
               (knobs.uniform_mix > 0 || cost_mix < (pair ? 12.0 * kp + 3.5 : 21.0 * kt + 2.0));
    if (mix) pair = false;
    if (mix)
    {
        cudaEventRecord(knobs.aux_ev[0], st);
        cudaStreamWaitEvent(knobs.aux_stream, knobs.aux_ev[0], 0);
        k_uniform_bulk<<<full / 32, 32, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, full, rej_idx, rej_cnt, rej_cap);
        k_uniform_bulk_pair<<<(2 * rest + 31) / 32, 32, 0, knobs.aux_stream>>>(
            seeds + (size_t)full * SEB_SEED_BYTES, ctr + full, out + (size_t)full * ct_stride, ct_stride, n, mod, max_multiple,
            rest, rej_idx + (size_t)full * rej_cap, rej_cnt + full, rej_cap);
        cudaEventRecord(knobs.aux_ev[1], knobs.aux_stream);
        cudaStreamWaitEvent(st, knobs.aux_ev[1], 0);
    }
    else if (pair)
        k_uniform_bulk_pair<<<(2 * batch + 31) / 32, 32, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch,
                                                                  rej_idx, rej_cnt, rej_cap);
    else if (coop)
        k_uniform_bulk_coop<<<(batch + 3) / 4, 128, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch,
                                                             rej_idx, rej_cnt, rej_cap);
    else
        k_uniform_bulk<<<(batch + 31) / 32, 32, 0, st>>>(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch,
                                                         rej_idx, rej_cnt, rej_cap);
    seb_launch_uniform_fix(seeds, ctr, out, ct_stride, n, mod, max_multiple, batch, rej_idx, rej_cnt, rej_cap, knobs, st);
}
