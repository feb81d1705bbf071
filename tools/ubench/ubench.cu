// ubench.cu — integer-pipe throughput probes for sm_100a (B200) and Keccak-f[1600] variants.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench ubench.cu
// Not part of the product: it answers "which pipe does the sampler/NTT code saturate".
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int ITER = 4096;
constexpr int CH = 8;  // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(256) k_op(uint32_t *out, uint32_t seed)
{
    uint32_t a[CH], b[CH];
    uint64_t w[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + blockIdx.x + threadIdx.x * 13; w[i] = a[i]; }
    const uint32_t c = (seed | 5u) + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITER; it++)
    {
#pragma unroll
        for (int i = 0; i < CH; i++)
        {
            if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(c));
            if (OP == 4) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 6) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 7)  // 1:1 lop3 + mad.lo
            {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(c));
            }
            if (OP == 8)  // 1:1 lop3 + mad.wide
            {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(c));
            }
            if (OP == 9)  // 1:1 lop3 + shf
            {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(b[i]) : "r"(c));
            }
            if (OP == 10) asm volatile("prmt.b32 %0, %0, %1, 0x3021;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 11)  // 2:1 lop3 + mad.wide
            {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(c));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(c), "r"(b[i]));
            }
            if (OP == 12)  // 1:1 lop3 + mad.hi
            {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(c));
            }
            if (OP == 13) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b[i]));
        }
        if (OP == 13)
        {
#pragma unroll
            for (int i = 0; i < CH; i++) a[i] ^= (uint32_t)(w[i] >> 32);  // keep a dependence (extra ALU op)
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (r == 0x12345678u) out[0] = r;
}

static const char *op_names[] = {"lop3", "shf.wrap", "mad.lo(IMAD)", "mad.wide(IMAD.WIDE)", "mad.hi(IMAD.HI)", "add", "min",
                                 "lop3+mad.lo", "lop3+mad.wide", "lop3+shf", "prmt", "2lop3+mad.wide", "lop3+mad.hi",
                                 "mul.wide+xor"};
static const int op_instrs[] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 3, 2, 2};

template <int OP>
static void run_op(uint32_t *d_out, int sms, double clk_ghz)
{
    const int blocks = sms * 8;
    k_op<OP><<<blocks, 256>>>(d_out, 1);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k_op<OP><<<blocks, 256>>>(d_out, 1);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = 5.0 * blocks * 256.0 * ITER * CH * op_instrs[OP];
    const double tops     = lane_ops / (ms * 1e-3) / 1e12;
    printf("op %-22s %8.3f T lane-instr/s  = %6.1f lanes/clk/SM (at %.3f GHz)\n", op_names[OP], tops,
           tops * 1e12 / (sms * clk_ghz * 1e9), clk_ghz);
}

// ---------------------------------------------------------------------------------------------
// Keccak-f[1600] variants on 32-bit halves
// ---------------------------------------------------------------------------------------------
__constant__ uint32_t c_rc_lo[24] = {0x00000001u, 0x00008082u, 0x0000808au, 0x80008000u, 0x0000808bu, 0x80000001u,
                                     0x80008081u, 0x00008009u, 0x0000008au, 0x00000088u, 0x80008009u, 0x8000000au,
                                     0x8000808bu, 0x0000008bu, 0x00008089u, 0x00008003u, 0x00008002u, 0x00000080u,
                                     0x0000800au, 0x8000000au, 0x80008081u, 0x00008080u, 0x80000001u, 0x80008008u};
__constant__ uint32_t c_rc_hi[24] = {0, 0, 0x80000000u, 0x80000000u, 0, 0, 0x80000000u, 0x80000000u, 0, 0, 0, 0,
                                     0, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0x80000000u, 0, 0x80000000u,
                                     0x80000000u, 0x80000000u, 0, 0x80000000u};

// multipliers 2^r fetched as constant-bank operands so ptxas cannot strength-reduce the multiply
// into LEA/SHF (ALU pipe): the point is to run rotations on the FMA pipe
__constant__ uint32_t c_pow2[32] = {1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,
                                    1u << 8,  1u << 9,  1u << 10, 1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15,
                                    1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21, 1u << 22, 1u << 23,
                                    1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t chi(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;  // a ^ (~b & c)
    asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// MODE 0: funnel shifts (ALU pipe). MODE 1: mul.wide + mad.wide (pair swap / 64-bit add left to ptxas).
// MODE 2: two mul.wide + two mad.lo by a constant-bank 1 (everything on the FMA pipe).
template <int R, int MODE>
__device__ __forceinline__ void rotl64(uint32_t lo, uint32_t hi, uint32_t &olo, uint32_t &ohi)
{
    if constexpr (R == 0) { olo = lo; ohi = hi; }
    else if constexpr (R == 32) { olo = hi; ohi = lo; }
    else if constexpr (R > 32) { rotl64<R - 32, MODE>(hi, lo, olo, ohi); }
    else if constexpr (MODE == 0)
    {
        ohi = __funnelshift_l(lo, hi, R);
        olo = __funnelshift_l(hi, lo, R);
    }
    else if constexpr (MODE == 1)
    {
        uint64_t q, p;
        uint32_t qlo, qhi;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(hi), "r"(c_pow2[R]));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(qlo), "=r"(qhi) : "l"(q));
        uint64_t add;
        asm("mov.b64 %0, {%1,%2};" : "=l"(add) : "r"(qhi), "r"(qlo));
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(lo), "r"(c_pow2[R]), "l"(add));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(olo), "=r"(ohi) : "l"(p));
    }
    else
    {
        // two wide products, halves combined by multiply-adds with a constant-bank 1 (FMA pipe)
        uint64_t q, p;
        uint32_t qlo, qhi, plo, phi;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(q) : "r"(hi), "r"(c_pow2[R]));
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(lo), "r"(c_pow2[R]));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(qlo), "=r"(qhi) : "l"(q));
        asm("mov.b64 {%0,%1}, %2;" : "=r"(plo), "=r"(phi) : "l"(p));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(olo) : "r"(plo), "r"(c_pow2[0]), "r"(qhi));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ohi) : "r"(qlo), "r"(c_pow2[0]), "r"(phi));
    }
}

// MODE_RHO: rotation mode for the 24 rho rotations; MODE_C: for the five rot-by-1 in theta;
// MIX: lanes with index < MIX use funnel shifts regardless (pipe balancing knob).
template <int MODE_RHO, int MODE_C, int MIX, int UNROLL>
__device__ __forceinline__ void keccak_f(uint32_t (&lo)[25], uint32_t (&hi)[25])
{
#pragma unroll UNROLL
    for (int round = 0; round < 24; round++)
    {
        uint32_t cl[5], ch[5], rl[5], rh[5];
#pragma unroll
        for (int x = 0; x < 5; x++)
        {
            cl[x] = xor3(xor3(lo[x], lo[x + 5], lo[x + 10]), lo[x + 15], lo[x + 20]);
            ch[x] = xor3(xor3(hi[x], hi[x + 5], hi[x + 10]), hi[x + 15], hi[x + 20]);
        }
#pragma unroll
        for (int x = 0; x < 5; x++) rotl64<1, MODE_C>(cl[x], ch[x], rl[x], rh[x]);
        uint32_t bl[25], bh[25];
#define RP(src, dst, rot)                                                                              \
    {                                                                                                  \
        const int x_   = (src) % 5;                                                                    \
        const uint32_t tl = xor3(lo[src], cl[(x_ + 4) % 5], rl[(x_ + 1) % 5]);                         \
        const uint32_t th = xor3(hi[src], ch[(x_ + 4) % 5], rh[(x_ + 1) % 5]);                         \
        if ((src) < MIX)                                                                               \
            rotl64<rot, 0>(tl, th, bl[dst], bh[dst]);                                                  \
        else                                                                                           \
            rotl64<rot, MODE_RHO>(tl, th, bl[dst], bh[dst]);                                           \
    }
        RP(0, 0, 0) RP(1, 10, 1) RP(2, 20, 62) RP(3, 5, 28) RP(4, 15, 27)
        RP(5, 16, 36) RP(6, 1, 44) RP(7, 11, 6) RP(8, 21, 55) RP(9, 6, 20)
        RP(10, 7, 3) RP(11, 17, 10) RP(12, 2, 43) RP(13, 12, 25) RP(14, 22, 39)
        RP(15, 23, 41) RP(16, 8, 45) RP(17, 18, 15) RP(18, 3, 21) RP(19, 13, 8)
        RP(20, 14, 18) RP(21, 24, 2) RP(22, 9, 61) RP(23, 19, 56) RP(24, 4, 14)
#undef RP
#pragma unroll
        for (int y = 0; y < 25; y += 5)
#pragma unroll
            for (int x = 0; x < 5; x++)
            {
                lo[y + x] = chi(bl[y + x], bl[y + (x + 1) % 5], bl[y + (x + 2) % 5]);
                hi[y + x] = chi(bh[y + x], bh[y + (x + 1) % 5], bh[y + (x + 2) % 5]);
            }
        lo[0] ^= c_rc_lo[round];
        hi[0] ^= c_rc_hi[round];
    }
}

// reference: straightforward 64-bit version (what the product kernels used at the time of writing)
__constant__ uint64_t c_rc64[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
__device__ __forceinline__ uint64_t rol(uint64_t x, int r) { return r ? (x << r) | (x >> (64 - r)) : x; }
__device__ __forceinline__ void keccak64(uint64_t (&a)[25])
{
    const int rho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    const int pi[25]  = {0, 10, 20, 5, 15, 16, 1, 11, 21, 6, 7, 17, 2, 12, 22, 23, 8, 18, 3, 13, 14, 24, 9, 19, 4};
#pragma unroll 1
    for (int round = 0; round < 24; round++)
    {
        uint64_t c[5], d[5], b[25];
#pragma unroll
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; i++) b[pi[i]] = rol(a[i] ^ d[i % 5], rho[i]);
#pragma unroll
        for (int y = 0; y < 25; y += 5)
#pragma unroll
            for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
        a[0] ^= c_rc64[round];
    }
}

constexpr int KREP = 16;  // permutations per thread

template <int VARIANT>
__global__ void __launch_bounds__(128) k_keccak(uint64_t *out, uint64_t seed)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t acc       = 0;
    if (VARIANT == 0)
    {
        uint64_t a[25];
#pragma unroll
        for (int i = 0; i < 25; i++) a[i] = seed * (i + 1) + tid;
#pragma unroll 1
        for (int r = 0; r < KREP; r++)
        {
            keccak64(a);
            acc ^= a[3];
            a[7] ^= r;
        }
#pragma unroll
        for (int i = 0; i < 25; i++) acc ^= a[i];
    }
    else
    {
        uint32_t lo[25], hi[25];
#pragma unroll
        for (int i = 0; i < 25; i++)
        {
            const uint64_t v = seed * (i + 1) + tid;
            lo[i] = (uint32_t)v;
            hi[i] = (uint32_t)(v >> 32);
        }
#pragma unroll 1
        for (int r = 0; r < KREP; r++)
        {
            if (VARIANT == 1) keccak_f<0, 0, 0, 1>(lo, hi);
            if (VARIANT == 2) keccak_f<1, 1, 0, 1>(lo, hi);
            if (VARIANT == 3) keccak_f<2, 2, 0, 1>(lo, hi);
            if (VARIANT == 4) keccak_f<2, 0, 0, 1>(lo, hi);
            if (VARIANT == 5) keccak_f<2, 2, 4, 1>(lo, hi);
            if (VARIANT == 6) keccak_f<2, 2, 8, 1>(lo, hi);
            if (VARIANT == 7) keccak_f<2, 2, 12, 1>(lo, hi);
            if (VARIANT == 8) keccak_f<0, 0, 0, 2>(lo, hi);
            if (VARIANT == 9) keccak_f<2, 2, 0, 2>(lo, hi);
            if (VARIANT == 10) keccak_f<1, 1, 8, 1>(lo, hi);
            if (VARIANT == 11) keccak_f<2, 2, 16, 1>(lo, hi);
            acc ^= ((uint64_t)hi[3] << 32) | lo[3];
            lo[7] ^= r;
        }
#pragma unroll
        for (int i = 0; i < 25; i++) acc ^= ((uint64_t)hi[i] << 32) | lo[i];
    }
    out[tid] = acc;
}

static const char *kvariants[] = {"64-bit C (product baseline)", "halves: funnel, xor3-theta", "wide x2 + add64 (ptxas picks)",
                                  "wide x2 + mad x2 (all FMA)", "FMA rho, funnel theta", "FMA, 4 lanes funnel", "FMA, 8 lanes funnel",
                                  "FMA, 12 lanes funnel", "funnel, unroll 2", "FMA all, unroll 2", "add64 form, 8 lanes funnel",
                                  "FMA, 16 lanes funnel"};

template <int V>
static uint64_t run_keccak(uint64_t *d_out, uint64_t *h_out, int sms)
{
    const int blocks = sms * 64, threads = 128;
    k_keccak<V><<<blocks, threads>>>(d_out, 0x9E3779B97F4A7C15ULL);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k_keccak<V><<<blocks, threads>>>(d_out, 0x9E3779B97F4A7C15ULL);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(h_out, d_out, (size_t)blocks * threads * 8, cudaMemcpyDeviceToHost));
    uint64_t sum = 0;
    for (size_t i = 0; i < (size_t)blocks * threads; i++) sum = sum * 1099511628211ULL + h_out[i];
    const double perms = 5.0 * blocks * threads * KREP;
    printf("keccak v%-2d %-36s %7.3f G perm/s   checksum %016llx\n", V, kvariants[V], perms / (ms * 1e-3) / 1e9,
           (unsigned long long)sum);
    return sum;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz / 1e6;
    printf("%s, %d SMs, max clock %.3f GHz\n", prop.name, prop.multiProcessorCount, ghz);
    const int sms = prop.multiProcessorCount;
    uint32_t *d_out;
    CK(cudaMalloc(&d_out, 1 << 20));
    run_op<0>(d_out, sms, ghz); run_op<1>(d_out, sms, ghz); run_op<2>(d_out, sms, ghz); run_op<3>(d_out, sms, ghz);
    run_op<4>(d_out, sms, ghz); run_op<5>(d_out, sms, ghz); run_op<6>(d_out, sms, ghz); run_op<7>(d_out, sms, ghz);
    run_op<8>(d_out, sms, ghz); run_op<9>(d_out, sms, ghz); run_op<10>(d_out, sms, ghz); run_op<11>(d_out, sms, ghz);
    run_op<12>(d_out, sms, ghz); run_op<13>(d_out, sms, ghz);

    const size_t nthreads = (size_t)sms * 64 * 128;
    uint64_t *d_k, *h_k = (uint64_t *)malloc(nthreads * 8);
    CK(cudaMalloc(&d_k, nthreads * 8));
    uint64_t ref = run_keccak<0>(d_k, h_k, sms);
    uint64_t s[12];
    s[1] = run_keccak<1>(d_k, h_k, sms); s[2] = run_keccak<2>(d_k, h_k, sms); s[3] = run_keccak<3>(d_k, h_k, sms);
    s[4] = run_keccak<4>(d_k, h_k, sms); s[5] = run_keccak<5>(d_k, h_k, sms); s[6] = run_keccak<6>(d_k, h_k, sms);
    s[7] = run_keccak<7>(d_k, h_k, sms); s[8] = run_keccak<8>(d_k, h_k, sms); s[9] = run_keccak<9>(d_k, h_k, sms);
    s[10] = run_keccak<10>(d_k, h_k, sms); s[11] = run_keccak<11>(d_k, h_k, sms);
    int bad = 0;
    for (int i = 1; i < 12; i++) bad += s[i] != ref;
    printf("variants disagreeing with v0: %d\n", bad);
    return bad != 0;
}
