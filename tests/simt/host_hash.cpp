// host_hash.cpp -- compiles the product's scalar per-chunk hash code (csrc/skein.cuh, csrc/sha256.cuh) as
// plain C++ so the build container (no GPU) can check the exact device arithmetic against the oracle.
// Test infrastructure only; never linked into libcomputewar.so.
#include <cstring>
#include "../../compute_war_b200/csrc/skein.cuh"
#include "../../compute_war_b200/csrc/sha256.cuh"

extern "C" void host_skein512(const uint8_t *p, uint32_t len, uint8_t *out, int unroll72)
{
    uint64_t IV[8], X[8];
    cw_skein_iv<8>(IV, 512);
    // unroll72: 0 = 3 x 24-round body, 1 = 72 rounds unrolled, 2 / 3 = the same two with the next-block prefetch loop,
    // 4 / 5 = the same two in the single-call-site form of the fused kernels
    if (unroll72 == 4) cw_skein_chunk_compact<8, false, false>(IV, p, len, X);
    else if (unroll72 == 5) cw_skein_chunk_compact<8, false, true>(IV, p, len, X);
    else if (unroll72 == 1) cw_skein_chunk<8, 64, false, true>(IV, p, len, X);
    else if (unroll72 == 0) cw_skein_chunk<8, 64, false, false>(IV, p, len, X);
    else if (unroll72 == 3) cw_skein_chunk<8, 64, false, true, true>(IV, p, len, X);
    else cw_skein_chunk<8, 64, false, false, true>(IV, p, len, X);
    memcpy(out, X, 64);
}
extern "C" void host_skein256_128(const uint8_t *p, uint32_t len, uint8_t *out, int prefetch)
{
    uint64_t IV[4], X[4];
    cw_skein_iv<4>(IV, 128);
    if (prefetch == 2) cw_skein_chunk_compact<4, false, true>(IV, p, len, X);
    else if (prefetch) cw_skein_chunk<4, 16, false, true, true>(IV, p, len, X);
    else cw_skein_chunk<4, 16, false, true>(IV, p, len, X);
    memcpy(out, X, 16);
}
extern "C" void host_sha256(const uint8_t *p, uint32_t len, uint8_t *out)
{
    uint32_t H[8];
    cw_sha256_chunk<false>(p, len, H);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(H[i] >> 24);
        out[4 * i + 1] = (uint8_t)(H[i] >> 16);
        out[4 * i + 2] = (uint8_t)(H[i] >> 8);
        out[4 * i + 3] = (uint8_t)H[i];
    }
}
