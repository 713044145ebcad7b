// hash_kernels.cu -- instantiation and launch of the per-chunk digest kernels (hash_kernels.cuh).
#include "hash_kernels.cuh"
#include "launch.h"

void cw_make_skein_iv(CwSkeinIV *iv)
{
    // the same UBI code the device runs, evaluated once on the host (Skein_512_Init / Skein_256_Init,
    // skein.c:226-266 and :31-71)
    cw_skein_iv<8>(iv->iv512, 512);
    cw_skein_iv<4>(iv->iv256, 128);
}

template <int ALG>
static cudaError_t launch_alg(const CwSkeinIV &iv, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_digests,
                              cudaStream_t st)
{
    const bool aligned = (chunk % 16u == 0) && ((reinterpret_cast<uintptr_t>(d_src) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(d_digests) & 15u) == 0);
    const unsigned blocks = (unsigned)((n + CW_HASH_THREADS - 1) / CW_HASH_THREADS);
    if (aligned) cw_hash_kernel<ALG, true><<<blocks, CW_HASH_THREADS, 0, st>>>(iv, d_src, chunk, n, d_digests);
    else cw_hash_kernel<ALG, false><<<blocks, CW_HASH_THREADS, 0, st>>>(iv, d_src, chunk, n, d_digests);
    return cudaGetLastError();
}

cudaError_t cw_launch_hash(int alg, const CwSkeinIV &iv, const uint8_t *d_src, uint32_t chunk, uint64_t n,
                           uint8_t *d_digests, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    switch (alg) {
    case CW_K_SKEIN512: return launch_alg<CW_K_SKEIN512>(iv, d_src, chunk, n, d_digests, st);
    case CW_K_SKEIN256_128: return launch_alg<CW_K_SKEIN256_128>(iv, d_src, chunk, n, d_digests, st);
    case CW_K_SHA256: return launch_alg<CW_K_SHA256>(iv, d_src, chunk, n, d_digests, st);
    case CW_K_SKEIN256_FULL: return launch_alg<CW_K_SKEIN256_FULL>(iv, d_src, chunk, n, d_digests, st);
    }
    return cudaErrorInvalidValue;
}

// ---- Skein with any digest length up to the state size: Skein_512_Init / Skein_256_Init(ctx, hashBitLen)
// (skein.c:226-266, :31-71) followed by Update / Final.  A digest shorter than the state is the head of the one output
// block computed under the initial chaining value of THAT length, so the chunk kernels above do the work: they are
// given the other chaining value and write whole output blocks (64 / 32 bytes), which the caller truncates. ----
void cw_make_skein_iv_bits(CwSkeinIV *iv, unsigned state_bits, unsigned hash_bits)
{
    cw_make_skein_iv(iv);
    if (state_bits == 512) cw_skein_iv<8>(iv->iv512, hash_bits);
    else cw_skein_iv<4>(iv->iv256, hash_bits);
}
