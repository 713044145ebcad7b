// decode_kernels.cu -- LZ4 / LZF decompression of independent blocks: one warp per block, persistent CTAs.
//
// Device-side counterpart of LZ4_decompress_safe / lzf_decompress as the reference calls them after each
// compression (experiment.cpp:256, :118).  Slot per warp = [compressed block][decoded chunk] in shared
// memory; per-block algorithm in decode_warp.cuh.  out_lens[i] = decoded size, 0 for a malformed block.
// An LZF block whose length is 0 ("did not fit", lzf.h:59-64) was stored raw by the caller: it is copied.
#include "decode_warp.cuh"
#include "launch.h"

#define CW_DEC_MAX_WARPS 16

__global__ void __launch_bounds__(CW_DEC_MAX_WARPS * 32, 1)
    cw_decode_kernel(int alg, const uint8_t *__restrict__ blocks, uint64_t stride, const uint32_t *__restrict__ lens,
                     uint32_t chunk, uint64_t n_chunks, uint8_t *__restrict__ out, uint32_t *__restrict__ out_lens,
                     unsigned long long *work, uint32_t in_cap, uint32_t slot_bytes)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint8_t *sin = smem + (size_t)warp * slot_bytes;
    uint8_t *sout = sin + in_cap;
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(work, 1ull);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_chunks) break;
        const uint8_t *b = blocks + g * stride;
        uint8_t *o = out + g * chunk;
        uint32_t len = lens[g];
        uint32_t got;
        __syncwarp();
        if (len == 0u && alg == 1) {  // LZF "stored raw"
            for (uint32_t i = lane; i < chunk; i += 32u) o[i] = b[i];
            got = chunk;
        } else if (len > in_cap || (uint64_t)len > stride) {  // longer than the codec's bound or than its own slot: malformed
            got = 0u;
        } else {
            if (((reinterpret_cast<uintptr_t>(b) & 3u) == 0)) {
                const uint32_t *b4 = reinterpret_cast<const uint32_t *>(b);
                uint32_t *s4 = reinterpret_cast<uint32_t *>(sin);
                for (uint32_t i = lane; i < len / 4u; i += 32u) s4[i] = __ldg(b4 + i);
                if (lane < (len & 3u)) sin[(len & ~3u) + lane] = b[(len & ~3u) + lane];  // never read past the block
            } else {
                for (uint32_t i = lane; i < len; i += 32u) sin[i] = b[i];
            }
            __syncwarp();
            got = alg == 0 ? cw_lz4_decode_warp(sin, len, sout, chunk) : cw_lzf_decode_warp(sin, len, sout, chunk);
            if (got == chunk) {
                if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0 && (chunk & 15u) == 0) {
                    const uint4 *s4 = reinterpret_cast<const uint4 *>(sout);
                    uint4 *o4 = reinterpret_cast<uint4 *>(o);
                    for (uint32_t i = lane; i < chunk / 16u; i += 32u) o4[i] = s4[i];
                } else {
                    for (uint32_t i = lane; i < chunk; i += 32u) o[i] = sout[i];
                }
            } else
                got = 0u;  // short or malformed: the safe decoders report failure
        }
        if (lane == 0) out_lens[g] = got;
    }
}

static int g_dec_max_smem = 0;

cudaError_t cw_decode_configure(const CwDeviceInfo &dev)
{
    g_dec_max_smem = (int)dev.smem_optin;
    return cudaFuncSetAttribute(cw_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_dec_max_smem);
}

cudaError_t cw_launch_decode(const CwDeviceInfo &dev, int alg, const uint8_t *d_blocks, uint64_t stride, const uint32_t *d_lens,
                             uint32_t chunk, uint64_t n, uint8_t *d_out, uint32_t *d_out_lens, unsigned long long *work,
                             cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    // a valid block is never longer than the codec's bound; the word-wise staging may read up to 3 bytes past len
    const uint32_t bound = alg == 0 ? chunk + chunk / 255u + 16u : chunk;
    const uint32_t in_cap = (bound + 15u) & ~15u;
    const uint32_t slot = in_cap + ((chunk + 15u) & ~15u);
    int warps = g_dec_max_smem / (int)slot;
    if (warps < 1) return cudaErrorInvalidValue;
    if (warps > CW_DEC_MAX_WARPS) warps = CW_DEC_MAX_WARPS;
    uint64_t ctas = (n + (uint64_t)warps - 1) / (uint64_t)warps;
    if (ctas > (uint64_t)dev.sm_count) ctas = (uint64_t)dev.sm_count;
    cw_decode_kernel<<<(unsigned)ctas, warps * 32, (size_t)warps * slot, st>>>(alg, d_blocks, stride, d_lens, chunk, n, d_out,
                                                                             d_out_lens, work, in_cap, slot);
    return cudaGetLastError();
}
