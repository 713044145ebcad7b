// skein.cuh -- Skein v1.1 (NIST round-1 constants) as a register-resident Threefish UBI chain.
//
// Replaces, for one chunk per thread, what the reference computes with
//   Skein_512_Init/Update/Final   reference_code/skein/Optimized_64bit/skein.c:226-266, 329-373, 377-408
//   Skein_512_Process_Block       .../skein_block.c:227-418   (loop form: Reference_Implementation/skein_block.c:127-219)
//   Skein_256_Init/Update/Final + Skein_256_Process_Block (skein.c:31-210, skein_block.c:42-210) with the
//   128-bit digest the tools ask for (src/hashing_perf/hash.h:16-17, HashAndCompress.cpp:123-129).
// Rotation constants / key-schedule parity are the round-1 values of skein.h:196,272-292 -- NOT the final
// Skein 1.3 ones; the KATs in reference_code/skein/KAT_MCT pin exactly these.
//
// B200 notes: the chain is serial inside a chunk, so one thread owns one chunk and the whole Threefish
// state (8 x u64 X, 9 x u64 key schedule, 3 tweak words, 8 x u64 message block) lives in registers.
// 64-bit add = IADD3 + IADD3.X, rotl64 = 2 x SHF.L.W, xor = 2 x LOP3: everything is on the integer
// pipes; there is no shared memory and no tensor-core work in this path.  The 72 rounds are emitted
// either fully unrolled or as 3 x 24 rounds with the key schedule rotated in registers (6 injections per
// trip, 6 = 0 mod 3 keeps the tweak schedule fixed); the latter is ~11 KB of SASS and stays inside the
// SM's 32 KB instruction cache.
#pragma once
#include "cw_device.cuh"

#define CW_SKEIN_T1_FIRST (1ull << 62)
#define CW_SKEIN_T1_FINAL (1ull << 63)
#define CW_SKEIN_T1_MSG (48ull << 56)
#define CW_SKEIN_T1_OUT (63ull << 56)
#define CW_SKEIN_T1_CFG (4ull << 56)
#define CW_SKEIN_PARITY 0x5555555555555555ull
#define CW_SKEIN_SCHEMA 0x0000000133414853ull

CW_HOSTDEV uint64_t cw_rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// How a MIX (a += b; b = rotl64(b, r) ^ a) is issued on the device (CW_SKEIN_MIXPOLICY):
//   0  plain C, ptxas decides: 64-bit shifts + ORs folded into the XOR's LOP3 = 5.2 ALU + 1.9 FMA instructions
//   1  funnel shifts: IADD3, IMAD.X, 2 SHF.L.W, 2 LOP3 = 5 ALU + 1 FMA (19 % fewer instructions; measured
//      676 vs 650 GB/s for Skein-512 on B200, profiles/r01_ab_skein_mix.txt)
// A third form that rotates two MIXes out of three on the FMA pipe (x * 2^r as a 32x32->64 IMAD.WIDE) balances
// the two integer pipes on paper (tools/ubench.cu) but measured slower than (1) in the real kernel (661 GB/s,
// IMAD.WIDE is half rate) and is not kept.
#ifndef CW_SKEIN_MIXPOLICY
#define CW_SKEIN_MIXPOLICY 1
#endif

CW_HOSTDEV uint64_t cw_rotxor_funnel(uint64_t b, uint64_t a, int r)
{
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
    if (r >= 32) {
        const uint32_t t = lo;
        lo = hi;
        hi = t;
        r -= 32;
    }
    const uint32_t nlo = __funnelshift_l(hi, lo, r), nhi = __funnelshift_l(lo, hi, r);
    return (((uint64_t)nhi << 32) | nlo) ^ a;
#else
    return cw_rotl64(b, r) ^ a;
#endif
}

#if CW_SKEIN_MIXPOLICY == 0
#define CW_MIXF(a, b, r) a += b; b = cw_rotl64(b, r) ^ a;
#else
#define CW_MIXF(a, b, r) a += b; b = cw_rotxor_funnel(b, a, r);
#endif
#define CW_MIXM(a, b, r) CW_MIXF(a, b, r)
#define CW_MIX(a, b, r) CW_MIXF(a, b, r)

// rounds 8d+0..3 and 8d+4..7 of Threefish-512 (word pairings and rotations, skein_block.c:175-216);
#define CW_TF512_R0123(x)                                                                           \
    CW_MIXF(x[0], x[1], 38) CW_MIXM(x[2], x[3], 30) CW_MIXM(x[4], x[5], 50) CW_MIXF(x[6], x[7], 53) \
    CW_MIXM(x[2], x[1], 48) CW_MIXM(x[4], x[7], 20) CW_MIXF(x[6], x[5], 43) CW_MIXM(x[0], x[3], 31) \
    CW_MIXM(x[4], x[1], 34) CW_MIXF(x[6], x[3], 14) CW_MIXM(x[0], x[5], 15) CW_MIXM(x[2], x[7], 27) \
    CW_MIXF(x[6], x[1], 26) CW_MIXM(x[0], x[7], 12) CW_MIXM(x[2], x[5], 58) CW_MIXF(x[4], x[3], 7)
#define CW_TF512_R4567(x)                                                                           \
    CW_MIXM(x[0], x[1], 33) CW_MIXM(x[2], x[3], 49) CW_MIXF(x[4], x[5], 8) CW_MIXM(x[6], x[7], 42)  \
    CW_MIXM(x[2], x[1], 39) CW_MIXF(x[4], x[7], 27) CW_MIXM(x[6], x[5], 41) CW_MIXM(x[0], x[3], 14) \
    CW_MIXF(x[4], x[1], 29) CW_MIXM(x[6], x[3], 26) CW_MIXM(x[0], x[5], 11) CW_MIXF(x[2], x[7], 9)  \
    CW_MIXM(x[6], x[1], 33) CW_MIXM(x[0], x[7], 51) CW_MIXF(x[2], x[5], 39) CW_MIXM(x[4], x[3], 35)

// key injection number s (1..18) given ks index base u (= s mod 9 pattern) -- InjectKey, skein_block.c:29-35
#define CW_TF512_INJECT(x, ks, ts, u, sval)         \
    x[0] += ks[((u) + 0) % 9];                      \
    x[1] += ks[((u) + 1) % 9];                      \
    x[2] += ks[((u) + 2) % 9];                      \
    x[3] += ks[((u) + 3) % 9];                      \
    x[4] += ks[((u) + 4) % 9];                      \
    x[5] += ks[((u) + 5) % 9] + ts[(u) % 3];        \
    x[6] += ks[((u) + 6) % 9] + ts[((u) + 1) % 3];  \
    x[7] += ks[((u) + 7) % 9] + (uint64_t)(sval);

// One UBI block: X <- Threefish512(key = X, tweak = (T0, T1), plaintext = w) ^ w.
// UNROLL72 = true: 72 rounds straight-line; false: 3 trips of 24 rounds.
template <bool UNROLL72>
CW_HOSTDEV void cw_tf512_block(uint64_t (&X)[8], const uint64_t (&w)[8], uint64_t T0, uint64_t T1)
{
    uint64_t ks[9], ts[3], x[8];
    ks[8] = CW_SKEIN_PARITY;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ks[i] = X[i];
        ks[8] ^= X[i];
    }
    ts[0] = T0;
    ts[1] = T1;
    ts[2] = T0 ^ T1;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = w[i] + ks[i];
    x[5] += ts[0];
    x[6] += ts[1];
    if (UNROLL72) {
#pragma unroll
        for (int d = 0; d < 9; d++) {
            CW_TF512_R0123(x)
            CW_TF512_INJECT(x, ks, ts, 2 * d + 1, 2 * d + 1)
            CW_TF512_R4567(x)
            CW_TF512_INJECT(x, ks, ts, 2 * d + 2, 2 * d + 2)
        }
    } else {
#pragma unroll 1
        for (int t = 0; t < 18; t += 6) {
            CW_TF512_R0123(x)
            CW_TF512_INJECT(x, ks, ts, 1, t + 1)
            CW_TF512_R4567(x)
            CW_TF512_INJECT(x, ks, ts, 2, t + 2)
            CW_TF512_R0123(x)
            CW_TF512_INJECT(x, ks, ts, 3, t + 3)
            CW_TF512_R4567(x)
            CW_TF512_INJECT(x, ks, ts, 4, t + 4)
            CW_TF512_R0123(x)
            CW_TF512_INJECT(x, ks, ts, 5, t + 5)
            CW_TF512_R4567(x)
            CW_TF512_INJECT(x, ks, ts, 6, t + 6)
            // rotate the key schedule by 6 so the next trip can reuse the same indices
            uint64_t k0 = ks[0], k1 = ks[1], k2 = ks[2], k3 = ks[3], k4 = ks[4], k5 = ks[5];
            ks[0] = ks[6];
            ks[1] = ks[7];
            ks[2] = ks[8];
            ks[3] = k0;
            ks[4] = k1;
            ks[5] = k2;
            ks[6] = k3;
            ks[7] = k4;
            ks[8] = k5;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) X[i] = x[i] ^ w[i];
}

// Threefish-256 rounds (skein_block.c:76-104)
#define CW_TF256_R0123(x)                                                                            \
    CW_MIXF(x[0], x[1], 5) CW_MIXM(x[2], x[3], 56) CW_MIXM(x[0], x[3], 36) CW_MIXF(x[2], x[1], 28)   \
    CW_MIXM(x[0], x[1], 13) CW_MIXM(x[2], x[3], 46) CW_MIXF(x[0], x[3], 58) CW_MIXM(x[2], x[1], 44)
#define CW_TF256_R4567(x)                                                                            \
    CW_MIXM(x[0], x[1], 26) CW_MIXF(x[2], x[3], 20) CW_MIXM(x[0], x[3], 53) CW_MIXM(x[2], x[1], 35)  \
    CW_MIXF(x[0], x[1], 11) CW_MIXM(x[2], x[3], 42) CW_MIXM(x[0], x[3], 59) CW_MIXF(x[2], x[1], 50)
#define CW_TF256_INJECT(x, ks, ts, s)              \
    x[0] += ks[((s) + 0) % 5];                     \
    x[1] += ks[((s) + 1) % 5] + ts[(s) % 3];       \
    x[2] += ks[((s) + 2) % 5] + ts[((s) + 1) % 3]; \
    x[3] += ks[((s) + 3) % 5] + (uint64_t)(s);

CW_HOSTDEV void cw_tf256_block(uint64_t (&X)[4], const uint64_t (&w)[4], uint64_t T0, uint64_t T1)
{
    uint64_t ks[5], ts[3], x[4];
    ks[4] = CW_SKEIN_PARITY;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ks[i] = X[i];
        ks[4] ^= X[i];
    }
    ts[0] = T0;
    ts[1] = T1;
    ts[2] = T0 ^ T1;
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = w[i] + ks[i];
    x[1] += ts[0];
    x[2] += ts[1];
#pragma unroll
    for (int d = 0; d < 9; d++) {
        CW_TF256_R0123(x)
        CW_TF256_INJECT(x, ks, ts, 2 * d + 1)
        CW_TF256_R4567(x)
        CW_TF256_INJECT(x, ks, ts, 2 * d + 2)
    }
#pragma unroll
    for (int i = 0; i < 4; i++) X[i] = x[i] ^ w[i];
}

// Initial chaining value for (state, hash_bits): UBI of the configuration block from the zero state
// (Skein_512_Init default branch, skein.c:246-257; the precomputed tables of skein_iv.h are this value).
template <int WORDS>
CW_HOSTDEV void cw_skein_iv(uint64_t (&X)[WORDS], unsigned hash_bits)
{
    uint64_t w[WORDS];
#pragma unroll
    for (int i = 0; i < WORDS; i++) {
        X[i] = 0;
        w[i] = 0;
    }
    w[0] = CW_SKEIN_SCHEMA;
    w[1] = hash_bits;
    const uint64_t T1 = CW_SKEIN_T1_FIRST | CW_SKEIN_T1_FINAL | CW_SKEIN_T1_CFG;
    if constexpr (WORDS == 8) cw_tf512_block<true>(X, w, 32, T1);
    else cw_tf256_block(X, w, 32, T1);
}

// message word loaders: p is 8-byte aligned on the fast path
template <int WORDS, bool ALIGNED>
CW_HOSTDEV void cw_skein_load(uint64_t (&w)[WORDS], const uint8_t *p, unsigned avail)
{
    if (ALIGNED && avail >= (unsigned)WORDS * 8u) {
#if defined(__CUDA_ARCH__)
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
        for (int i = 0; i < WORDS / 2; i++) {
            uint4 v = __ldg(q + i);
            w[2 * i] = (uint64_t)v.x | ((uint64_t)v.y << 32);
            w[2 * i + 1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
        }
#else
        for (int i = 0; i < WORDS; i++) {
            uint64_t v = 0;
            for (int k = 7; k >= 0; k--) v = (v << 8) | p[8 * i + k];
            w[i] = v;
        }
#endif
        return;
    }
    if (avail >= (unsigned)WORDS * 8u) {  // a whole block behind an unaligned pointer: plain byte loads
#pragma unroll
        for (int i = 0; i < WORDS; i++) {
            uint64_t v = 0;
#pragma unroll
            for (int k = 7; k >= 0; k--) v = (v << 8) | p[8 * i + k];
            w[i] = v;
        }
        return;
    }
    // the last, partial block of a chunk (once per chunk): one short counted loop per word.  (The fully unrolled form
    // "byte o of the block if o < avail" needs 64 predicates at once; nvcc 12.9 / ptxas produced wrong words from it for
    // a few tail lengths in some kernels -- found by the ragged-size tests -- so the bound is a loop count here.)
#pragma unroll
    for (int i = 0; i < WORDS; i++) {
        const unsigned nb = avail > 8u * i ? (avail - 8u * i < 8u ? avail - 8u * i : 8u) : 0u;
        uint64_t v = 0;
#pragma unroll 1
        for (unsigned k = 0; k < nb; k++) v |= (uint64_t)p[8u * i + k] << (8u * k);
        w[i] = v;
    }
}

// The same digest with ONE call site of the block function: message blocks, the final block and the output block go
// through one loop.  The fused kernels use this form -- three inlined copies of the 72 rounds are 33 KB of SASS, and the
// instruction cache of an SM whose other warps run the compressor holds 32 KB.
template <int WORDS, bool ALIGNED, bool UNROLL72>
CW_HOSTDEV void cw_skein_chunk_compact(const uint64_t (&IV)[WORDS], const uint8_t *p, uint32_t len, uint64_t (&X)[WORDS])
{
    constexpr uint32_t B = WORDS * 8;
    uint64_t w[WORDS];
#pragma unroll
    for (int i = 0; i < WORDS; i++) X[i] = IV[i];
    uint64_t T1 = CW_SKEIN_T1_FIRST | CW_SKEIN_T1_MSG;
    uint32_t off = 0;
    bool out_block = false;
#pragma unroll 1
    for (;;) {
        uint64_t T0;
        if (!out_block) {  // the next 1..B message bytes (an empty message is one zero block); the last block is FINAL
            const uint32_t rest = len - off, nb = rest > B ? B : rest;
            cw_skein_load<WORDS, ALIGNED>(w, p + off, nb);
            off += nb;
            T0 = off;
            if (off == len) T1 |= CW_SKEIN_T1_FINAL;
        } else {  // output block: counter 0 under the final chaining value (skein.c:393-399)
#pragma unroll
            for (int i = 0; i < WORDS; i++) w[i] = 0;
            T0 = 8;
            T1 = CW_SKEIN_T1_FIRST | CW_SKEIN_T1_FINAL | CW_SKEIN_T1_OUT;
        }
        if constexpr (WORDS == 8) cw_tf512_block<UNROLL72>(X, w, T0, T1);
        else cw_tf256_block(X, w, T0, T1);
        if (out_block) break;
        T1 = CW_SKEIN_T1_MSG;
        out_block = off == len;
    }
}

// Hash one chunk of `len` bytes.  IV = cw_skein_iv(...) computed once on the host.  out receives
// OUT_BYTES (<= 8*WORDS) digest bytes, little-endian words (Skein_Put64_LSB_First, skein_port.h:57).
// PREFETCH: the next block's words are loaded before the current block is processed.  The lone hash warp of the fused
// kernel needs this -- its loads come from L2 and nothing else on its scheduler hides them (+18 % when hashing is the
// limit); with thousands of threads per SM it is still worth 2.5 % (stand-alone kernel: 675 -> 692 GB/s).
template <int WORDS, int OUT_BYTES, bool ALIGNED, bool UNROLL72, bool PREFETCH = false>
CW_HOSTDEV void cw_skein_chunk(const uint64_t (&IV)[WORDS], const uint8_t *p, uint32_t len,
                               uint64_t (&X)[WORDS])
{
    constexpr uint32_t B = WORDS * 8;
    uint64_t w[WORDS];
#pragma unroll
    for (int i = 0; i < WORDS; i++) X[i] = IV[i];
    uint64_t T1 = CW_SKEIN_T1_FIRST | CW_SKEIN_T1_MSG;
    uint32_t off = 0;
    // all blocks but the last are full (Skein_512_Update keeps the last 1..B bytes for Final, skein.c:354-370)
    if constexpr (PREFETCH) {
        uint64_t wn[WORDS];
        cw_skein_load<WORDS, ALIGNED>(w, p, len > B ? B : len);
        while (len - off > B) {
            const uint32_t rest = len - off - B;  // bytes behind this block: 1..B of them form the next one
            cw_skein_load<WORDS, ALIGNED>(wn, p + off + B, rest > B ? B : rest);
            off += B;
            if constexpr (WORDS == 8) cw_tf512_block<UNROLL72>(X, w, off, T1);
            else cw_tf256_block(X, w, off, T1);
            T1 = CW_SKEIN_T1_MSG;
#pragma unroll
            for (int i = 0; i < WORDS; i++) w[i] = wn[i];
        }
    } else {
        while (len - off > B) {
            cw_skein_load<WORDS, ALIGNED>(w, p + off, B);
            off += B;
            if constexpr (WORDS == 8) cw_tf512_block<UNROLL72>(X, w, off, T1);
            else cw_tf256_block(X, w, off, T1);
            T1 = CW_SKEIN_T1_MSG;
        }
        // final (zero padded) block
        cw_skein_load<WORDS, ALIGNED>(w, p + off, len - off);
    }
    // final block, T0 = total length (skein.c:383-387)
    if constexpr (WORDS == 8) cw_tf512_block<UNROLL72>(X, w, len, T1 | CW_SKEIN_T1_FINAL);
    else cw_tf256_block(X, w, len, T1 | CW_SKEIN_T1_FINAL);
    // output block: counter 0 under the final chaining value (skein.c:393-399)
#pragma unroll
    for (int i = 0; i < WORDS; i++) w[i] = 0;
    if constexpr (WORDS == 8) cw_tf512_block<UNROLL72>(X, w, 8, CW_SKEIN_T1_FIRST | CW_SKEIN_T1_FINAL | CW_SKEIN_T1_OUT);
    else cw_tf256_block(X, w, 8, CW_SKEIN_T1_FIRST | CW_SKEIN_T1_FINAL | CW_SKEIN_T1_OUT);
}
