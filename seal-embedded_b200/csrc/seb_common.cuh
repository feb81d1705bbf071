// seb_common.cuh — shared device helpers for the B200 CKKS encode+encrypt path.
//
// Limbs are 32-bit with 64-bit products, as in the reference (device/lib/defines.h:365-379);
// primes are < 2^30 so lazy values in [0,4q) fit a word (device/lib/uintmodarith.h:293-331).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

// Host shims: the per-thread device functions in these headers are also compiled with plain g++
// by tests/host_emul (a sequential emulation used by the CPU test-suite to check index math and
// bit tricks without a GPU).  Under nvcc none of this is active.
#ifndef __CUDACC__
#include <algorithm>
#include <cstring>
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#define SEB_HOST_EMUL 1
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)  // selector nibbles 0..7 only
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r       = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline uint32_t __vcmpgeu4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xFF) >= ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
    return r;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() {}
static inline void __syncwarp() {}
using std::min;
#define SEB_CONSTANT static const
#else
#define SEB_CONSTANT __constant__
#endif

#define SEB_MAX_PRIMES 13
#define SEB_SEED_BYTES 64

// one RNS prime with the constants the kernels need
struct SebModulus
{
    uint32_t q;         // prime value (device/lib/modulus.h:22-30 `value`)
    uint32_t two_q;     // 2q
    uint32_t ratio_lo;  // floor(2^64/q) low word  (const_ratio[0], device/lib/modulus.c:30-47)
    uint32_t ratio_hi;  // floor(2^64/q) high word (const_ratio[1])
};

// ---------------------------------------------------------------------------------------------
// modular arithmetic
// ---------------------------------------------------------------------------------------------
// one conditional subtraction (device/lib/modulo.h:21-32 shift_result), branch-free
__device__ __forceinline__ uint32_t seb_csub(uint32_t x, uint32_t q)
{
    return min(x, x - q);  // x - q wraps above x when x < q
}

// Shoup / "MUMO" lazy product (device/lib/uintmodarith.h:308-331): w < q, wq = floor(w*2^32/q);
// result in [0,2q) for ANY 32-bit x.
__device__ __forceinline__ uint32_t seb_mul_shoup_lazy(uint32_t x, uint32_t w, uint32_t wq, uint32_t q)
{
    return x * w - __umulhi(x, wq) * q;
}

// exact x mod q for a 64-bit x (device/lib/modulo.h:84-116): t = floor(x*floor(2^64/q)/2^64),
// r = lo32(x) - lo32(t)*q, one correction.
__device__ __forceinline__ uint32_t seb_barrett64(uint64_t x, const SebModulus &m)
{
    const uint64_t ratio = ((uint64_t)m.ratio_hi << 32) | m.ratio_lo;
    const uint32_t t     = (uint32_t)__umul64hi(x, ratio);
    return seb_csub((uint32_t)x - t * m.q, m.q);
}

// 32-bit input variant (device/lib/modulo.h:43-75)
__device__ __forceinline__ uint32_t seb_barrett32(uint32_t x, const SebModulus &m)
{
    const uint32_t t = __umulhi(x, m.ratio_hi);
    return seb_csub(x - t * m.q, m.q);
}

// [0,4q) -> [0,q) (device/lib/ntt.c:176-185)
__device__ __forceinline__ uint32_t seb_final_reduce(uint32_t x, uint32_t q, uint32_t two_q)
{
    return seb_csub(seb_csub(x, two_q), q);
}

// 256-bit accesses (sm_100: LDG.E.256 / STG.E.256): one full 32-byte sector per lane
struct __align__(32) seb_oct
{
    uint32_t v[8];
};
#ifdef __CUDACC__
__device__ __forceinline__ seb_oct seb_ldg256(const seb_oct *p)
{
    seb_oct r;  // read-only tables: plain (non-volatile) asm so the compiler may schedule it freely
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void seb_stg256_stream(seb_oct *p, const seb_oct &r)
{
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]),
                 "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
                 : "memory");
}
#else
static inline seb_oct seb_ldg256(const seb_oct *p) { return *p; }
static inline void seb_stg256_stream(seb_oct *p, const seb_oct &r) { *p = r; }
#endif

// ---------------------------------------------------------------------------------------------
// cache-hinted global accesses for streamed (touched-once) data
// ---------------------------------------------------------------------------------------------
#ifndef __CUDACC__
static inline uint4 seb_ldg_stream(const uint4 *p) { return *p; }
static inline uint32_t seb_ldg_stream(const uint32_t *p) { return *p; }
static inline void seb_stg_stream(uint4 *p, const uint4 &v) { *p = v; }
#else
__device__ __forceinline__ uint4 seb_ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t seb_ldg_stream(const uint32_t *p)
{
    uint32_t r;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void seb_stg_stream(uint4 *p, const uint4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------
// named barrier over an aligned group of W threads of a CTA of T threads: hardware barrier 1 + t/W.
// Hardware barriers are an occupancy resource and ptxas reserves all 16 for a kernel whose barrier number is a
// register.  IMM = true selects the number with a switch over immediates, so exactly T/W + 1 are reserved (the
// encode at n = 1024 runs 8 CTAs per SM and lost 13 % with 16 reserved); IMM = false keeps the register form,
// which is 2 % faster where occupancy is not at stake (the NTT at n >= 8192, 2-4 CTAs per SM).
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
template <int W, int T, bool IMM>
__device__ __forceinline__ void seb_group_barrier(const int t)
{
    static_assert(T % W == 0 && T / W >= 1 && T / W <= 15, "named barriers 1..15");
    if constexpr (!IMM)
        asm volatile("bar.sync %0, %1;" ::"r"(1 + t / W), "n"(W) : "memory");
    else
    {
#define SEB_BAR_CASE(G)                                                                          \
    case G:                                                                                      \
        if (G < T / W) asm volatile("bar.sync %0, %1;" ::"n"(G + 1), "n"(W) : "memory");          \
        break;
        switch (t / W)
        {
            SEB_BAR_CASE(0) SEB_BAR_CASE(1) SEB_BAR_CASE(2) SEB_BAR_CASE(3) SEB_BAR_CASE(4) SEB_BAR_CASE(5)
            SEB_BAR_CASE(6) SEB_BAR_CASE(7) SEB_BAR_CASE(8) SEB_BAR_CASE(9) SEB_BAR_CASE(10) SEB_BAR_CASE(11)
            SEB_BAR_CASE(12) SEB_BAR_CASE(13) SEB_BAR_CASE(14)
            default: break;
        }
#undef SEB_BAR_CASE
    }
}
#endif
