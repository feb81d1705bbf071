// seb_encode.cuh — per-thread pieces of the CKKS encode (FP64 inverse FFT), shared by the kernel
// in seb_encode.cu and by the g++ host emulation in tests/host_emul.
#pragma once

#include "seb_common.cuh"

#ifndef __CUDACC__
#include <cmath>
// host emulation: plain IEEE double ops (compiled with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
#endif

// complex values per thread: 8 (radix-8 passes of three fused radix-2 stages) or 16 (radix-16 passes of four: one
// shared-memory round trip fewer at n >= 2048; a compile-time choice for the whole library, see profiles/README.md)
#ifndef ENC_E
#define ENC_E 16
#endif
#define ENC_LR (ENC_E == 16 ? 4 : 3)  // stages per full pass
static_assert(ENC_E == 8 || ENC_E == 16, "ENC_E: 8 or 16 values per thread");

__host__ __device__ constexpr int enc_npass(int lognl) { return (lognl + ENC_LR - 1) / ENC_LR; }
__host__ __device__ constexpr int enc_r(int lognl, int p) { return (lognl - ENC_LR * p) >= ENC_LR ? ENC_LR : (lognl - ENC_LR * p); }

// The complex vector in shared memory: two arrays of doubles, re[] and im[], with an XOR swizzle (ENC_C2 = 0, the
// product).  ENC_C2 = 1 is the measured alternative: ONE array of (re, im) pairs read and written with 128-bit accesses
// (half the shared-memory instructions; swizzle i ^ ((i >> 3) & 7)) - no faster at n = 4096 (1.782 vs 1.789 ms per 65536
// messages) and slower above (2.51 vs 2.45 ms at n = 8192, 4.15 vs 3.35 ms at n = 16384): profiles/README.md, round 2.
#ifndef ENC_C2
#define ENC_C2 0
#endif
__device__ __forceinline__ uint32_t enc_swz(uint32_t i)
{
#if ENC_C2
    return i ^ ((i >> 3) & 7u);
#elif ENC_E == 16
    // pass 0: lane l holds elements 16 l + j -> low four bits j ^ l, distinct over a half-warp; later passes: lanes on
    // consecutive elements with (i >> 4) & 15 constant over a half-warp
    return i ^ ((i >> 4) & 15u);
#else
    return i ^ ((i >> 4) & 7u) ^ (((i >> 6) & 1u) << 3);
#endif
}
// element `pos` of the vector whose storage starts at sre (ENC_C2: 2*NL doubles as NL pairs; else sre[NL], sim[NL])
__device__ __forceinline__ void enc_ld(const double *sre, const double *sim, uint32_t pos, double &re, double &im)
{
#if ENC_C2
    (void)sim;
    const double2 v = reinterpret_cast<const double2 *>(sre)[enc_swz(pos)];
    re = v.x;
    im = v.y;
#else
    re = sre[enc_swz(pos)];
    im = sim[enc_swz(pos)];
#endif
}
__device__ __forceinline__ void enc_st(double *sre, double *sim, uint32_t pos, double re, double im)
{
#if ENC_C2
    (void)sim;
    reinterpret_cast<double2 *>(sre)[enc_swz(pos)] = make_double2(re, im);
#else
    sre[enc_swz(pos)] = re;
    sim[enc_swz(pos)] = im;
#endif
}

// Layout of the staged message in shared memory.  Pass 0 gathers values[src_map[pos]] for 8 consecutive
// positions per thread; the slots a warp asks for in one instruction are far from random (the index map is
// 3^i mod 2n, bit-reversed): in a linear layout they collide 2- (n = 1024) to 32-way (n = 16384) on the 32
// banks.  Skewing slot s to s + (s >> (log2 n - 9)) makes every gather instruction of every warp conflict
// free at every degree (exhaustive check: tests/test_host_logic.py::test_encode_gather_conflict_free).
template <int LOGN>
__host__ __device__ __forceinline__ constexpr uint32_t enc_vskew(uint32_t s)
{
    // 8 positions per thread: shift log2(n) - 9; 16 per thread: log2(n) - 10 (n = 1024: 1, two-way at worst)
    return s + (s >> (LOGN - 6 - ENC_LR > 0 ? LOGN - 6 - ENC_LR : 1));
}
// floats of shared memory the staged message occupies
template <int LOGN>
struct EncVals
{
    static constexpr uint32_t WORDS = (enc_vskew<LOGN>((1u << (LOGN - 1)) - 1u) + 4u) & ~3u;
};

// Pass-0 twiddles.  Stage r of pass 0 (butterfly distance 2^r) gives the thread that owns positions
// 8g .. 8g+7 the roots tw[(n >> (r+1)) + (g << (2-r)) + m], m < 2^(2-r): seven per thread, all distinct
// across threads, at a lane stride of 64 / 32 / 16 bytes in the natural table — up to 16 cache lines per
// 128-bit load instruction.  They are therefore stored a second time behind the natural table
// (tw[n + slot*(n/8) + g], slot = 0..3 for r = 0, 4..5 for r = 1, 6 for r = 2) so that the lanes of a warp read
// consecutive 16-byte entries.  Later passes share each root between >= 8 consecutive threads and use
// the natural table.
#define ENC_TW0_SLOTS (ENC_E - 1)
// slot of root m of stage r: stage 0's E/2 roots first, then stage 1's E/4, ...
__host__ __device__ __forceinline__ constexpr int enc_tw0_slot(int r, int m) { return ENC_E - (ENC_E >> r) + m; }
// entries of the whole table: n natural + (E - 1) * n/E pass-0 copies
__host__ __device__ __forceinline__ constexpr size_t enc_tw_entries(size_t n) { return n + ENC_TW0_SLOTS * (n / ENC_E); }
template <class D2>
inline void enc_build_tw0(size_t n, D2 *tw)  // tw[0..n) filled; appends the pass-0 copies
{
    const size_t G = n / ENC_E;
    for (size_t g = 0; g < G; g++)
        for (int r = 0; r < ENC_LR; r++)
            for (int m = 0; m < (1 << (ENC_LR - 1 - r)); m++)
                tw[n + (size_t)enc_tw0_slot(r, m) * G + g] = tw[(n >> (r + 1)) + (g << (ENC_LR - 1 - r)) + m];
}

// one pass = R fused Gentleman-Sande stages starting at butterfly distance S = 2^LS
// (fft.c:119-143: vec[k] = u + v; vec[k+tt] = (u - v) * s)
// svals: the message of this ciphertext zero-padded to n/2 floats, value i at svals[enc_vskew(i)] (staged
// in shared memory by the kernel: the scatter of ckks_common.c:139-153 is done as a gather from it)
template <int LOGN, int LOGNL, int P>
__device__ __forceinline__ void enc_pass(double (&xr)[ENC_E], double (&xi)[ENC_E], double *sre, double *sim,
                                         const int t, const uint32_t cta_pos0, const float *svals,
                                         const uint16_t *__restrict__ src_map, const double2 *__restrict__ tw)
{
    constexpr int NL   = 1 << LOGNL;
    constexpr int T    = NL / ENC_E;
    constexpr int R    = enc_r(LOGNL, P);
    constexpr int LS   = ENC_LR * P;
    constexpr int GP   = ENC_E >> R;
    constexpr bool LAST = (P == enc_npass(LOGNL) - 1);

#pragma unroll
    for (int i = 0; i < GP; i++)
    {
        const uint32_t g    = (uint32_t)t + (uint32_t)i * T;
        const uint32_t off  = g & ((1u << LS) - 1u);
        const uint32_t blk  = g >> LS;
        const uint32_t base = (blk << (LS + R)) | off;  // local position of element j = 0
        if (P == 0)
        {
            // pass 0 works on E consecutive positions (LS = 0, R = log2 E): their map entries are one or two
            // 128-bit loads; both conjugate slots of value i receive values[i]
            static_assert(P != 0 || (LS == 0 && R == ENC_LR), "pass 0 is expected to be a full pass on contiguous data");
            uint32_t mw[ENC_E / 2];
#pragma unroll
            for (int k = 0; k < ENC_E / 8; k++)
            {
                const uint4 mp = __ldg(reinterpret_cast<const uint4 *>(src_map + cta_pos0 + base) + k);
                mw[4 * k] = mp.x, mw[4 * k + 1] = mp.y, mw[4 * k + 2] = mp.z, mw[4 * k + 3] = mp.w;
            }
#pragma unroll
            for (int j = 0; j < ENC_E; j++)
            {
                const uint32_t slot = (mw[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                xr[i * ENC_E + j]   = (double)svals[enc_vskew<LOGN>(slot)];
                xi[i * ENC_E + j]   = 0.0;
            }
        }
        else
        {
#pragma unroll
            for (int j = 0; j < (1 << R); j++)
            {
                const uint32_t pos = base | ((uint32_t)j << LS);
                enc_ld(sre, sim, pos, xr[i * (1 << R) + j], xi[i * (1 << R) + j]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
        {
            // stage with tt = 2^(LS+r): group index of global position p is p >> (LS+r+1)
            const uint32_t h     = 1u << (LOGN - LS - r - 1);
            const uint32_t jbase = (cta_pos0 >> (LS + r + 1)) + (blk << (R - r - 1));
#pragma unroll
            for (int m = 0; m < (1 << (R - r - 1)); m++)
            {
                // pass 0: this thread's own roots, from the lane-contiguous copy behind the natural table
                const double2 s = (P == 0) ? __ldg(tw + (1u << LOGN) + (uint32_t)enc_tw0_slot(r, m) * ((1u << LOGN) / ENC_E) +
                                                   (cta_pos0 >> ENC_LR) + g)
                                           : __ldg(tw + h + jbase + m);
#pragma unroll
                for (int k = 0; k < (1 << r); k++)
                {
                    const int ia    = i * (1 << R) + (m << (r + 1)) + k;
                    const int ib    = ia + (1 << r);
                    if (P == 0 && k == 0)
                    {
                        // The message is real (ckks_common.c:139-153 stores values[i] in both conjugate slots, imaginary
                        // parts 0), and after r stages the elements whose low r position bits are 0 still are: u and v
                        // are both real here.  di = 0 - 0 = +0, so the reference's dr*s.x - di*s.y and dr*s.y + di*s.x
                        // equal dr*s.x and dr*s.y up to the SIGN OF A ZERO result, which no later operation turns into a
                        // different non-zero value and the final round-to-int64 maps to 0 either way: 4 FP64
                        // operations instead of 10 for 30 of the 64 butterflies of this pass.
                        const double ur = xr[ia], vr = xr[ib];
                        const double dr = __dsub_rn(ur, vr);
                        xr[ia]          = __dadd_rn(ur, vr);
                        xr[ib]          = __dmul_rn(dr, s.x);
                        xi[ib]          = __dmul_rn(dr, s.y);
                        continue;
                    }
                    const double ur = xr[ia], ui = xi[ia], vr = xr[ib], vi = xi[ib];
                    const double dr = __dsub_rn(ur, vr), di = __dsub_rn(ui, vi);
                    xr[ia]          = __dadd_rn(ur, vr);
                    xi[ia]          = __dadd_rn(ui, vi);
                    xr[ib]          = __dsub_rn(__dmul_rn(dr, s.x), __dmul_rn(di, s.y));
                    xi[ib]          = __dadd_rn(__dmul_rn(dr, s.y), __dmul_rn(di, s.x));
                }
            }
        }
        if (!LAST)
        {
#pragma unroll
            for (int j = 0; j < (1 << R); j++)
            {
                const uint32_t pos = base | ((uint32_t)j << LS);
                enc_st(sre, sim, pos, xr[i * (1 << R) + j], xi[i * (1 << R) + j]);
            }
        }
    }
}

// Barrier scope between pass P and pass P+1 (thread t owns group t in the radix-8 passes).  Pass 1 reads, for
// block t/8, what the eight threads 8*(t/8) .. +7 wrote in pass 0: warp-local.  Pass 2 reads, for block t/64,
// what threads 64*(t/64) .. +63 wrote in pass 1: local to an aligned group of 64 threads, synchronised with a
// named barrier (1 + group; CTAs of 1024 threads use groups of 128 to stay within hardware barriers 1..15).
// Later boundaries span 512 threads or the CTA.  Returns 0 for a CTA-wide barrier, else the unit's width.
// Checked exhaustively by tests/test_host_logic.py::test_encode_barrier_scopes.
// With 16 values per thread: pass 1 reads what the sixteen threads 16*(t/16) .. +15 wrote (warp-local), pass 2 what
// threads 256*(t/256) .. +255 wrote: a named barrier over 256 threads when the CTA has more, else the CTA barrier.
__host__ __device__ __forceinline__ constexpr int enc_sync_width(int lognl, int p)
{
    if (ENC_E == 16) return p == 0 ? 32 : (p == 1 && ((1 << lognl) / ENC_E) > 256) ? 256 : 0;
    return p == 0 ? 32 : p == 1 ? (((1 << lognl) / ENC_E) > 960 ? 128 : 64) : 0;
}
template <int LOGNL, int P>
__device__ __forceinline__ void enc_sync(const int t)
{
    constexpr int W = enc_sync_width(LOGNL, P);
    if constexpr (W == 32)
        __syncwarp();
    else if constexpr (W > 32)
    {
#if defined(__CUDACC__)
        seb_group_barrier<W, (1 << LOGNL) / ENC_E, true>(t);
#endif
    }
    else
        __syncthreads();
}

template <int LOGN, int LOGNL, int P>
struct EncRun
{
    __device__ __forceinline__ static void run(double (&xr)[ENC_E], double (&xi)[ENC_E], double *sre, double *sim,
                                               int t, uint32_t cta_pos0, const float *svals,
                                               const uint16_t *src_map, const double2 *tw)
    {
        enc_pass<LOGN, LOGNL, P>(xr, xi, sre, sim, t, cta_pos0, svals, src_map, tw);
        if (P + 1 < enc_npass(LOGNL))
        {
            enc_sync<LOGNL, P>(t);
            EncRun<LOGN, LOGNL, (P + 1 < enc_npass(LOGNL) ? P + 1 : P)>::run(xr, xi, sre, sim, t, cta_pos0, svals,
                                                                             src_map, tw);
        }
    }
};

// coeff = round(Re * scale/n); |coeff| > 2^63 fails the encode; the int64 conversion follows x86
// (ckks_common.c:183-206)
__device__ __forceinline__ int64_t enc_finish(double re, double n_inv, int &bad, uint32_t &mag)
{
    const double c = round(__dmul_rn(re, n_inv));
    if (fabs(c) > 9223372036854775808.0) bad = 1;
    // magnitude class for the encrypt kernels' 32-bit reduction path: |c| clipped to 2^32 - 1
    // (NaN compares false and clips too)
    const double a = fabs(c);
    const uint32_t m32 = a < 4294967295.0 ? (uint32_t)a : 0xFFFFFFFFu;
    mag                = mag > m32 ? mag : m32;
    if (!(c < 9223372036854775808.0)) return (int64_t)0x8000000000000000ULL;  // NaN / 2^63: "indefinite"
    return (int64_t)c;
}


// Last stage of the n = 16384 transform (tt = n/2, one group, twiddle index 1), real part only:
// rank 0 owns element k (u), rank 1 owns element k + n/2 (v).  fft.c:134-141.
__device__ __forceinline__ double enc_cross_re(uint32_t rank, double own_re, double own_im, double oth_re,
                                               double oth_im, const double2 s)
{
    if (rank == 0) return __dadd_rn(own_re, oth_re);  // Re(u + v)
    const double dr = __dsub_rn(oth_re, own_re), di = __dsub_rn(oth_im, own_im);  // u - v
    return __dsub_rn(__dmul_rn(dr, s.x), __dmul_rn(di, s.y));                      // Re((u - v) * s)
}
