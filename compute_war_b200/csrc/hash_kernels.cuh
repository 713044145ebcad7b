// hash_kernels.cuh -- per-chunk digest kernels: one chunk per thread.
//
// Device-side replacement for doHashing (HashAndCompress.cpp:119-134) / HashBlockSkein256 /
// HashBlockSHA256 (src/hashing_perf/hash.cpp:5-46) applied to n_chunks chunks at once.
//
// Layout: thread g hashes chunk g (src + g*chunk), digest g goes to digests + g*D.  A thread walks its chunk
// 64 (or 32) bytes at a time with LDG.128; those are two (one) full 32-byte sectors per request, so HBM
// efficiency is not hurt by the 8 KB stride between neighbouring lanes, and at ~30 integer instructions
// per byte the loads are far off the critical path.  Blocks are small (64 threads) so that the block
// scheduler keeps the tail of the last wave short.
#pragma once
#include "sha256.cuh"
#include "skein.cuh"

struct CwSkeinIV {
    uint64_t iv512[8];  // Skein-512, 512-bit output
    uint64_t iv256[4];  // Skein-256, 128-bit output
};

enum { CW_K_SKEIN512 = 0, CW_K_SKEIN256_128 = 1, CW_K_SHA256 = 2, CW_K_SKEIN256_FULL = 3 /* whole 32-byte output block */ };

#ifndef CW_SKEIN_UNROLL72
#define CW_SKEIN_UNROLL72 true
#endif
#ifndef CW_SKEIN_PREFETCH
#define CW_SKEIN_PREFETCH true  // load the next block before processing this one (skein.cuh; A/B in profiles/r01_ab_skein_mix.txt)
#endif

// digest of chunk `p` (len bytes) written to `out` (D bytes); OUT_ALIGNED: out is 16-byte aligned
// PREFETCH: Skein loads the next block before it processes the current one (16 more registers)
// COMPACT: one call site of the Skein block function (skein.cuh; no prefetch) -- the fused kernels, for their code size
template <int ALG, bool ALIGNED, bool OUT_ALIGNED, bool PREFETCH = CW_SKEIN_PREFETCH, bool COMPACT = false>
__device__ __forceinline__ void cw_hash_one(const CwSkeinIV &iv, const uint8_t *__restrict__ p, uint32_t len,
                                            uint8_t *__restrict__ out)
{
    if constexpr (ALG == CW_K_SKEIN512) {
        uint64_t X[8];
        if constexpr (COMPACT) cw_skein_chunk_compact<8, ALIGNED, CW_SKEIN_UNROLL72>(iv.iv512, p, len, X);
        else cw_skein_chunk<8, 64, ALIGNED, CW_SKEIN_UNROLL72, PREFETCH>(iv.iv512, p, len, X);
        if (OUT_ALIGNED) {
            uint4 *o = reinterpret_cast<uint4 *>(out);
#pragma unroll
            for (int i = 0; i < 4; i++)
                o[i] = make_uint4((uint32_t)X[2 * i], (uint32_t)(X[2 * i] >> 32), (uint32_t)X[2 * i + 1],
                                  (uint32_t)(X[2 * i + 1] >> 32));
        } else {
#pragma unroll
            for (int i = 0; i < 64; i++) out[i] = (uint8_t)(X[i >> 3] >> (8 * (i & 7)));
        }
    } else if constexpr (ALG == CW_K_SKEIN256_128) {
        uint64_t X[4];
        if constexpr (COMPACT) cw_skein_chunk_compact<4, ALIGNED, true>(iv.iv256, p, len, X);
        else cw_skein_chunk<4, 16, ALIGNED, true, PREFETCH>(iv.iv256, p, len, X);
        if (OUT_ALIGNED) {
            *reinterpret_cast<uint4 *>(out) =
                make_uint4((uint32_t)X[0], (uint32_t)(X[0] >> 32), (uint32_t)X[1], (uint32_t)(X[1] >> 32));
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) out[i] = (uint8_t)(X[i >> 3] >> (8 * (i & 7)));
        }
    } else if constexpr (ALG == CW_K_SKEIN256_FULL) {
        uint64_t X[4];
        cw_skein_chunk<4, 32, ALIGNED, true, PREFETCH>(iv.iv256, p, len, X);
        if (OUT_ALIGNED) {
            uint4 *o = reinterpret_cast<uint4 *>(out);
            o[0] = make_uint4((uint32_t)X[0], (uint32_t)(X[0] >> 32), (uint32_t)X[1], (uint32_t)(X[1] >> 32));
            o[1] = make_uint4((uint32_t)X[2], (uint32_t)(X[2] >> 32), (uint32_t)X[3], (uint32_t)(X[3] >> 32));
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) out[i] = (uint8_t)(X[i >> 3] >> (8 * (i & 7)));
        }
    } else {
        uint32_t H[8];
        cw_sha256_chunk<ALIGNED>(p, len, H);
        if (OUT_ALIGNED) {
            uint4 *o = reinterpret_cast<uint4 *>(out);
            o[0] = make_uint4(cw_bswap32(H[0]), cw_bswap32(H[1]), cw_bswap32(H[2]), cw_bswap32(H[3]));
            o[1] = make_uint4(cw_bswap32(H[4]), cw_bswap32(H[5]), cw_bswap32(H[6]), cw_bswap32(H[7]));
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) out[i] = (uint8_t)(H[i >> 2] >> (24 - 8 * (i & 3)));
        }
    }
}

template <int ALG>
__host__ __device__ constexpr int cw_digest_len()
{
    return ALG == CW_K_SKEIN512 ? 64 : (ALG == CW_K_SKEIN256_128 ? 16 : 32);  // SHA-256 and the full Skein-256 block: 32
}

#define CW_HASH_THREADS 64

template <int ALG, bool ALIGNED>
__global__ void __launch_bounds__(CW_HASH_THREADS)
    cw_hash_kernel(const __grid_constant__ CwSkeinIV iv, const uint8_t *__restrict__ src, uint32_t chunk,
                   uint64_t n_chunks, uint8_t *__restrict__ digests)
{
    const uint64_t g = (uint64_t)blockIdx.x * CW_HASH_THREADS + threadIdx.x;
    if (g >= n_chunks) return;
    constexpr int D = cw_digest_len<ALG>();
    cw_hash_one<ALG, ALIGNED, ALIGNED>(iv, src + g * chunk, chunk, digests + g * D);
}
