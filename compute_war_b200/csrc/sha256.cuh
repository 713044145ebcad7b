// sha256.cuh -- SHA-256 (FIPS 180-4) of one chunk per thread, multi-block, state and the rolling 16-word
// message schedule in registers.
//
// Replaces HashBlockSHA256 -> OpenSSL SHA256() (src/hashing_perf/hash.cpp:28-46, call :35) and the
// *intended* behaviour of the isa-l multi-buffer path doSHA256MBHashing / HashBlockSHA256MB
// (HashAndCompress.cpp:136-158, hash.cpp:48-77), whose reference code never reads a digest back
// (SURVEY.md section 3.1); digests are the plain SHA-256 of each chunk, big-endian words.
//
// B200 notes: rotr = SHF.R.W, Ch/Maj = one LOP3 each, the five-input round sums go to IADD3.  All on the
// integer pipes; per 64-byte block ~1450 SASS instructions.
#pragma once
#include "cw_device.cuh"

// device-only under nvcc (the round constants live in __constant__ memory); plain inline for the host tests
#define CW_SHAFN CW_DEV

CW_SHAFN uint32_t cw_rotr32(uint32_t x, int n)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}
CW_SHAFN uint32_t cw_bswap32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}

#if defined(__CUDACC__)
__constant__
#endif
    static const uint32_t CW_K256[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

// one compression; w[16] holds the big-endian message words and is consumed (rolling schedule)
CW_SHAFN void cw_sha256_compress(uint32_t (&H)[8], uint32_t (&w)[16])
{
    uint32_t a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5], g = H[6], h = H[7];
#pragma unroll
    for (int t = 0; t < 64; t++) {
        if (t >= 16) {
            const uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
            const uint32_t s0 = cw_rotr32(w15, 7) ^ cw_rotr32(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = cw_rotr32(w2, 17) ^ cw_rotr32(w2, 19) ^ (w2 >> 10);
            w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
        }
        const uint32_t S1 = cw_rotr32(e, 6) ^ cw_rotr32(e, 11) ^ cw_rotr32(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t t1 = h + S1 + ch + CW_K256[t] + w[t & 15];
        const uint32_t S0 = cw_rotr32(a, 2) ^ cw_rotr32(a, 13) ^ cw_rotr32(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        h = g;
        g = f;
        f = e;
        e = d + t1;
        d = c;
        c = b;
        b = a;
        a = t1 + S0 + mj;
    }
    H[0] += a;
    H[1] += b;
    H[2] += c;
    H[3] += d;
    H[4] += e;
    H[5] += f;
    H[6] += g;
    H[7] += h;
}

template <bool ALIGNED>
CW_SHAFN void cw_sha256_load(uint32_t (&w)[16], const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    if constexpr (ALIGNED) {
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint4 v = __ldg(q + i);
            w[4 * i] = cw_bswap32(v.x);
            w[4 * i + 1] = cw_bswap32(v.y);
            w[4 * i + 2] = cw_bswap32(v.z);
            w[4 * i + 3] = cw_bswap32(v.w);
        }
    } else
#endif
    {
#pragma unroll
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    }
}

// SHA-256 of one chunk; H receives the eight state words (digest = their big-endian bytes)
template <bool ALIGNED>
CW_SHAFN void cw_sha256_chunk(const uint8_t *p, uint32_t len, uint32_t (&H)[8])
{
    uint32_t w[16];
    H[0] = 0x6a09e667u; H[1] = 0xbb67ae85u; H[2] = 0x3c6ef372u; H[3] = 0xa54ff53au;
    H[4] = 0x510e527fu; H[5] = 0x9b05688cu; H[6] = 0x1f83d9abu; H[7] = 0x5be0cd19u;
    uint32_t off = 0;
    // full blocks, the next one loaded before the current one is compressed (as in cw_skein_chunk)
    if (len >= 64u) {
        uint32_t wn[16];
        cw_sha256_load<ALIGNED>(w, p);
        for (; off + 128 <= len; off += 64) {
            cw_sha256_load<ALIGNED>(wn, p + off + 64);
            cw_sha256_compress(H, w);
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = wn[i];
        }
        cw_sha256_compress(H, w);
        off += 64;
    }
    // padding: 0x80, zeros, 64-bit big-endian bit length
    const uint32_t rem = len - off;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t o = 4u * i + k;
            const uint32_t byte = o < rem ? p[off + o] : (o == rem ? 0x80u : 0u);
            v = (v << 8) | byte;
        }
        w[i] = v;
    }
    if (rem >= 56) {
        cw_sha256_compress(H, w);
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = 0;
    }
    w[14] = len >> 29;
    w[15] = len << 3;
    cw_sha256_compress(H, w);
}
