// lz4_kernels.cu -- LZ4 compression of independent chunks: one warp per chunk, persistent CTAs.
//
// Device-side replacement for the lz4 doCompression lambda (HashAndCompress.cpp:349-355) applied to
// n_chunks chunks; per-chunk algorithm in lz4_warp.cuh.
//
// One CTA per SM; every warp owns a shared-memory slot = [chunk bytes + 16 B zero pad][u16 table 16 KB]
// and pulls chunk indices from a global counter until the batch is exhausted (chunk cost varies by an
// order of magnitude between incompressible and text-like data, so static partitioning would leave SMs
// idle).  The number of warps per CTA is whatever fits in the SM's 227 KB: at 8 KB chunks 9 with the reference's
// 8192-entry table (compat) and 14 with the 4096-entry table of fast mode.
#include "launch.h"
#include "lz4_warp.cuh"
#include "stage.cuh"

#define CW_LZ4_MAX_WARPS 16

static inline uint32_t lz4_chunk_pad(uint32_t chunk) { return (chunk + CW_LZ4_PAD + 15u) & ~15u; }
static inline uint32_t lz4_slot_bytes(uint32_t chunk, int hashlog) { return lz4_chunk_pad(chunk) + (2u << hashlog); }

template <int HASHLOG>
__global__ void __launch_bounds__(CW_LZ4_MAX_WARPS * 32, 1)
    cw_lz4_kernel(const uint8_t *__restrict__ src, uint32_t chunk, uint64_t n_chunks, uint8_t *__restrict__ dst,
                  uint64_t stride, uint32_t *__restrict__ lens, unsigned long long *work, uint32_t slot_bytes)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[32];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint8_t *s = smem + (size_t)warp * slot_bytes;
    const uint32_t chunk_pad = (chunk + CW_LZ4_PAD + 15u) & ~15u;
    uint16_t *table = reinterpret_cast<uint16_t *>(s + chunk_pad);
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
        // zero the table while the chunk is in flight
        __syncwarp();
        if (bulk_ok && lane == 0) cw_bulk_load(s, src + g * chunk, chunk, bar);
        {
            uint4 *t4 = reinterpret_cast<uint4 *>(table);
#pragma unroll 8
            for (uint32_t i = lane; i < (2u << HASHLOG) / 16u; i += 32u) t4[i] = make_uint4(0, 0, 0, 0);
        }
        if (bulk_ok) {
            cw_mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            const uint8_t *gp = src + g * chunk;
            for (uint32_t i = lane; i < chunk; i += 32u) s[i] = gp[i];
        }
        __syncwarp();
        const uint32_t len = cw_lz4_compress_warp<HASHLOG>(s, chunk, table, dst + g * stride);
        if (lane == 0) lens[g] = len;
    }
}

static int g_lz4_max_smem = 0;

cudaError_t cw_lz4_configure(const CwDeviceInfo &dev)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, cw_lz4_kernel<13>);
    if (e != cudaSuccess) return e;
    g_lz4_max_smem = (int)dev.smem_optin - (int)fa.sharedSizeBytes;  // opt-in limit covers static + dynamic
    e = cudaFuncSetAttribute(cw_lz4_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_lz4_max_smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(cw_lz4_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_lz4_max_smem);
}

cudaError_t cw_launch_lz4(const CwDeviceInfo &dev, bool fast, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                          uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    const uint32_t slot = lz4_slot_bytes(chunk, fast ? 12 : 13);
    int warps = g_lz4_max_smem / (int)slot;
    if (warps < 1) return cudaErrorInvalidValue;
    if (warps > CW_LZ4_MAX_WARPS) warps = CW_LZ4_MAX_WARPS;
    uint64_t ctas = (n + (uint64_t)warps - 1) / (uint64_t)warps;
    if (ctas > (uint64_t)dev.sm_count) ctas = (uint64_t)dev.sm_count;
    if (fast)
        cw_lz4_kernel<12><<<(unsigned)ctas, warps * 32, (size_t)warps * slot, st>>>(d_src, chunk, n, d_dst, stride, d_lens, work, slot);
    else
        cw_lz4_kernel<13><<<(unsigned)ctas, warps * 32, (size_t)warps * slot, st>>>(d_src, chunk, n, d_dst, stride, d_lens, work, slot);
    return cudaGetLastError();
}
