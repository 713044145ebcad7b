// ubench.cu -- integer-pipe microbenchmark for B200 (sm_100a).
//
// The Skein/SHA/fused kernels are bounded by the integer issue rate, not by HBM, so the roofline
// denominator has to be measured: how many thread-instructions per clock per SM do IADD3 / LOP3 / SHF
// (ALU pipe) and IMAD / IMAD.WIDE (FMA pipe) sustain alone and mixed.  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

enum { K_IADD3, K_LOP3, K_SHF, K_IMAD, K_IMADWIDE, K_LOP3_IMADWIDE, K_MIX_DEFAULT, K_MIX_FUNNEL, K_MIX_MULWIDE, K_NKINDS };

// ---- three ways to write rotl64(b, r) ^ a -------------------------------------------------------
__device__ __forceinline__ uint64_t rotxor_default(uint64_t b, uint64_t a, int r, const uint32_t *)
{
    return ((b << r) | (b >> (64 - r))) ^ a;
}
__device__ __forceinline__ uint64_t rotxor_funnel(uint64_t b, uint64_t a, int r, const uint32_t *)
{
    uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
    if (r >= 32) {
        uint32_t t = lo;
        lo = hi;
        hi = t;
        r -= 32;
    }
    uint32_t nlo = __funnelshift_l(hi, lo, r), nhi = __funnelshift_l(lo, hi, r);
    return (((uint64_t)nhi << 32) | nlo) ^ a;
}
// rotate on the FMA pipe: x * 2^r as a 32x32->64 multiply yields (x << r, x >> (32-r)) in one IMAD.WIDE;
// the two OR-merges fold into the XOR's LOP3.  pw[r] = 2^r comes from memory so ptxas cannot turn the
// multiply back into shifts.
__device__ __forceinline__ uint64_t rotxor_mulwide(uint64_t b, uint64_t a, int r, const uint32_t *pw)
{
    uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
    if (r >= 32) {
        uint32_t t = lo;
        lo = hi;
        hi = t;
        r -= 32;
    }
    const uint32_t m = pw[r];
    uint64_t P, Q;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(P) : "r"(lo), "r"(m));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(Q) : "r"(hi), "r"(m));
    uint32_t nlo = ((uint32_t)P | (uint32_t)(Q >> 32)) ^ (uint32_t)a;
    uint32_t nhi = ((uint32_t)Q | (uint32_t)(P >> 32)) ^ (uint32_t)(a >> 32);
    return ((uint64_t)nhi << 32) | nlo;
}

#define UB_MIX(a, b, r) a += b; b = ROTXOR(b, a, r, pw);
#define UB_8ROUNDS(x)                                                                       \
    UB_MIX(x[0], x[1], 38) UB_MIX(x[2], x[3], 30) UB_MIX(x[4], x[5], 50) UB_MIX(x[6], x[7], 53) \
    UB_MIX(x[2], x[1], 48) UB_MIX(x[4], x[7], 20) UB_MIX(x[6], x[5], 43) UB_MIX(x[0], x[3], 31) \
    UB_MIX(x[4], x[1], 34) UB_MIX(x[6], x[3], 14) UB_MIX(x[0], x[5], 15) UB_MIX(x[2], x[7], 27) \
    UB_MIX(x[6], x[1], 26) UB_MIX(x[0], x[7], 12) UB_MIX(x[2], x[5], 58) UB_MIX(x[4], x[3], 7)  \
    UB_MIX(x[0], x[1], 33) UB_MIX(x[2], x[3], 49) UB_MIX(x[4], x[5], 8) UB_MIX(x[6], x[7], 42)  \
    UB_MIX(x[2], x[1], 39) UB_MIX(x[4], x[7], 27) UB_MIX(x[6], x[5], 41) UB_MIX(x[0], x[3], 14) \
    UB_MIX(x[4], x[1], 29) UB_MIX(x[6], x[3], 26) UB_MIX(x[0], x[5], 11) UB_MIX(x[2], x[7], 9)  \
    UB_MIX(x[6], x[1], 33) UB_MIX(x[0], x[7], 51) UB_MIX(x[2], x[5], 39) UB_MIX(x[4], x[3], 35)

#define MIX_ITERS 512  // x 32 MIX each

template <int KIND>
__global__ void __launch_bounds__(256) ubench_kernel(uint32_t *out, uint32_t seed, uint32_t mul, const uint32_t *__restrict__ pw_g)
{
    __shared__ uint32_t pw_s[32];
    if (threadIdx.x < 32) pw_s[threadIdx.x] = pw_g[threadIdx.x];
    __syncthreads();
    if (KIND >= K_MIX_DEFAULT) {
        uint32_t pw[32];
#pragma unroll
        for (int i = 0; i < 32; i++) pw[i] = pw_s[i];
        uint64_t x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = ((uint64_t)(seed + i) << 32) | (threadIdx.x * 977u + i + blockIdx.x);
        for (int it = 0; it < MIX_ITERS; it++) {
            if (KIND == K_MIX_DEFAULT) {
#define ROTXOR rotxor_default
                UB_8ROUNDS(x)
#undef ROTXOR
            } else if (KIND == K_MIX_FUNNEL) {
#define ROTXOR rotxor_funnel
                UB_8ROUNDS(x)
#undef ROTXOR
            } else {
#define ROTXOR rotxor_mulwide
                UB_8ROUNDS(x)
#undef ROTXOR
            }
        }
        uint64_t r = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) r ^= x[i];
        if (r == 0x12345678u) out[0] = (uint32_t)r;
        return;
    }
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        a[i] = seed + threadIdx.x * 7 + i;
        b[i] = seed * 3 + i * 11 + blockIdx.x + threadIdx.x * 13;
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (KIND == K_IADD3) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
            } else if (KIND == K_LOP3) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(mul));
            } else if (KIND == K_SHF) {
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(b[i]) : "r"(a[i]));
            } else if (KIND == K_IMAD) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(mul), "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(mul), "r"(a[i]));
            } else if (KIND == K_IMADWIDE) {
                uint64_t t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(mul));
                a[i] = (uint32_t)(t >> 32);
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(b[i]), "r"(mul));
                b[i] = (uint32_t)(t >> 32);
            } else if (KIND == K_LOP3_IMADWIDE) {
                uint64_t t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(mul));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"((uint32_t)t), "r"((uint32_t)(t >> 32)), "r"(b[i]));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(b[i]), "r"(mul));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(b[i]) : "r"((uint32_t)t), "r"((uint32_t)(t >> 32)), "r"(a[i]));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) r ^= a[i] ^ b[i];
    if (r == 0x12345678u) out[0] = r;
}

__global__ void clock_kernel(unsigned long long *out)
{
    unsigned long long t0, c0, t1, c1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    c0 = clock64();
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < 2000000ull);
    c1 = clock64();
    out[0] = c1 - c0;
    out[1] = t1 - t0;
}

template <int KIND>
static double run(int blocks, int threads, uint32_t *d_out, const uint32_t *d_pw)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    ubench_kernel<KIND><<<blocks, threads>>>(d_out, 12345u, 64u, d_pw);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        ubench_kernel<KIND><<<blocks, threads>>>(d_out, 12345u + rep, 64u, d_pw);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    uint32_t *d_out;
    cudaMalloc(&d_out, 64);
    unsigned long long *d_clk, h_clk[2];
    cudaMalloc(&d_clk, 16);
    const int sms = prop.multiProcessorCount;
    const int threads = 256, blocks = sms * 8;  // 64 warps per SM
    // instructions per (iteration, chain) of the hand-written part
    uint32_t h_pw[32], *d_pw;
    for (int i = 0; i < 32; i++) h_pw[i] = 1u << i;
    cudaMalloc(&d_pw, sizeof h_pw);
    cudaMemcpy(d_pw, h_pw, sizeof h_pw, cudaMemcpyHostToDevice);
    const char *names[K_NKINDS] = {"add_u32(ptxas: iadd3+imad.iadd)", "lop3", "shf", "imad", "imad_wide", "lop3+imad_wide",
                                   "threefish_mix_default", "threefish_mix_funnel", "threefish_mix_mulwide"};
    // thread-instructions (simple kinds) or MIX operations (threefish kinds) per thread
    const double per_thread[K_NKINDS] = {2.0 * ITERS * CHAINS, 2.0 * ITERS * CHAINS, 2.0 * ITERS * CHAINS, 2.0 * ITERS * CHAINS,
                                         2.0 * ITERS * CHAINS, 4.0 * ITERS * CHAINS, 32.0 * MIX_ITERS, 32.0 * MIX_ITERS, 32.0 * MIX_ITERS};
    double ms[K_NKINDS];
    ms[K_IADD3] = run<K_IADD3>(blocks, threads, d_out, d_pw);
    ms[K_LOP3] = run<K_LOP3>(blocks, threads, d_out, d_pw);
    ms[K_SHF] = run<K_SHF>(blocks, threads, d_out, d_pw);
    ms[K_IMAD] = run<K_IMAD>(blocks, threads, d_out, d_pw);
    ms[K_IMADWIDE] = run<K_IMADWIDE>(blocks, threads, d_out, d_pw);
    ms[K_LOP3_IMADWIDE] = run<K_LOP3_IMADWIDE>(blocks, threads, d_out, d_pw);
    ms[K_MIX_DEFAULT] = run<K_MIX_DEFAULT>(blocks, threads, d_out, d_pw);
    ms[K_MIX_FUNNEL] = run<K_MIX_FUNNEL>(blocks, threads, d_out, d_pw);
    ms[K_MIX_MULWIDE] = run<K_MIX_MULWIDE>(blocks, threads, d_out, d_pw);
    // SM clock right after load
    ubench_kernel<K_IADD3><<<blocks, threads>>>(d_out, 1u, 64u, d_pw);
    clock_kernel<<<1, 1>>>(d_clk);
    cudaMemcpy(h_clk, d_clk, 16, cudaMemcpyDeviceToHost);
    double mhz = (double)h_clk[0] / (double)h_clk[1] * 1000.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_mhz_measured\": %.1f, \"clock_rate_khz\": %d", prop.name, sms, mhz,
           prop.clockRate);
    for (int k = 0; k < K_NKINDS; k++) {
        double units = (double)blocks * threads * per_thread[k];
        double per_clk_sm = units / (ms[k] * 1e-3) / (mhz * 1e6) / sms;
        printf(", \"%s\": {\"ms\": %.4f, \"%s_per_clk_per_sm\": %.2f}", names[k], ms[k], k >= K_MIX_DEFAULT ? "mix" : "thread_instr", per_clk_sm);
    }
    printf("}\n");
    return 0;
}
