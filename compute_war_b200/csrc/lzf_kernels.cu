// lzf_kernels.cu -- LZF compression of independent chunks: one warp per chunk, persistent CTAs.
//
// Device-side replacement for the lzf doCompression lambda lzf_compress(s, l, d, l-1)
// (HashAndCompress.cpp:342-348) applied to n_chunks chunks; per-chunk algorithm in lzf_warp.cuh.
// lens[i] = bytes written, or 0 when the chunk does not shrink by at least one byte (the caller then stores
// it raw, lzf.h:59-64).
//
// Same structure as lz4_kernels.cu: one CTA per SM, each warp owns a shared-memory slot
// [chunk + 16 B pad][head u16[8192]][next u16[8192]] = 40 KB at 8 KB chunks (5 warps per SM) and pulls
// chunk indices from a global counter.
#include "launch.h"
#include "lzf_warp.cuh"
#include "stage.cuh"

#define CW_LZF_MAX_WARPS 16

static inline uint32_t lzf_slot_bytes(uint32_t chunk) { return ((chunk + 16u + 15u) & ~15u) + (uint32_t)sizeof(CwLzfScratch); }

__global__ void __launch_bounds__(CW_LZF_MAX_WARPS * 32, 1)
    cw_lzf_kernel(const uint8_t *__restrict__ src, uint32_t chunk, uint64_t n_chunks, uint8_t *__restrict__ dst,
                  uint64_t stride, uint32_t *__restrict__ lens, unsigned long long *work, uint32_t slot_bytes)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[32];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint8_t *s = smem + (size_t)warp * slot_bytes;
    const uint32_t chunk_pad = (chunk + 16u + 15u) & ~15u;
    CwLzfScratch *scr = reinterpret_cast<CwLzfScratch *>(s + chunk_pad);
    uint64_t *bar = &bars[warp];
    if (lane == 0) cw_mbar_init(bar, 1);
    for (uint32_t i = chunk + lane; i < chunk_pad; i += 32u) s[i] = 0;
    __syncwarp();
    const bool bulk_ok = ((chunk & 15u) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    uint32_t phase = 0;
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(work, 1ull);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_chunks) break;
        cw_stage_chunk(s, src + g * chunk, chunk, bulk_ok, bar, phase);
        const uint32_t len = cw_lzf_compress_warp(s, chunk, scr, dst + g * stride, chunk - 1u);
        if (lane == 0) lens[g] = len;
    }
}

static int g_lzf_max_smem = 0;

cudaError_t cw_lzf_configure(const CwDeviceInfo &dev)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, cw_lzf_kernel);
    if (e != cudaSuccess) return e;
    g_lzf_max_smem = (int)dev.smem_optin - (int)fa.sharedSizeBytes;
    return cudaFuncSetAttribute(cw_lzf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_lzf_max_smem);
}

cudaError_t cw_launch_lzf(const CwDeviceInfo &dev, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                          uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    const uint32_t slot = lzf_slot_bytes(chunk);
    int warps = g_lzf_max_smem / (int)slot;
    if (warps < 1) return cudaErrorInvalidValue;
    if (warps > CW_LZF_MAX_WARPS) warps = CW_LZF_MAX_WARPS;
    uint64_t ctas = (n + (uint64_t)warps - 1) / (uint64_t)warps;
    if (ctas > (uint64_t)dev.sm_count) ctas = (uint64_t)dev.sm_count;
    cw_lzf_kernel<<<(unsigned)ctas, warps * 32, (size_t)warps * slot, st>>>(d_src, chunk, n, d_dst, stride, d_lens, work,
                                                                          slot);
    return cudaGetLastError();
}
