// lzf_kernels.cu -- LZF compression of independent chunks: one warp per chunk, persistent CTAs.
//
// Device-side replacement for the lzf doCompression lambda lzf_compress(s, l, d, l-1)
// (HashAndCompress.cpp:342-348) applied to n_chunks chunks; per-chunk algorithm in lzf_warp.cuh.
// lens[i] = bytes written, or 0 when the chunk does not shrink by at least one byte (the caller then stores
// it raw, lzf.h:59-64).
//
// Same structure as lz4_kernels.cu: one CTA per SM; parser warps each own a shared-memory slot
// [chunk + 16 B pad][head u16[NB]][next u16[NB]] (NB = 8192: 40 KB at 8 KB chunks, 5 per SM; NB = 4096 for chunks of
// up to 4 KB: 20 KB, 11 per SM) plus a record queue, and pull
// chunk indices from a global counter; emitter warps (no slot) write the streams (lzf_warp.cuh).
#include "comp_roles.cuh"
#include "launch.h"

#include <cstdlib>

#define CW_LZF_MAX_WARPS 16
#define CW_LZF_EMITTERS_DEFAULT 2

static int lzf_emitters()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CW_LZF_EMITTERS");  // tuning knob for the profiling scripts
        v = e ? atoi(e) : 0;
        if (v > 3) v = 3;
    }
    return v > 0 ? v : CW_LZF_EMITTERS_DEFAULT;
}

// bytes of one parser's slot, without its queue
static inline uint32_t lzf_slot_bytes(uint32_t chunk)
{
    return ((chunk + 16u + 15u) & ~15u) + cw_lzf_table_bytes(cw_lzf_table_format(chunk));
}

template <bool ALIGNED>
__global__ void __launch_bounds__((CW_LZF_MAX_WARPS + 3) * 32, 1)
    cw_lzf_kernel(const uint8_t *__restrict__ src, uint32_t chunk, uint64_t n_chunks, uint8_t *__restrict__ dst,
                  uint64_t stride, uint32_t *__restrict__ lens, unsigned long long *work, uint32_t slot_bytes, uint32_t parsers,
                  uint32_t emitters, uint32_t qlen)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[32];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    CwLz4Queue *queues = reinterpret_cast<CwLz4Queue *>(smem + (size_t)parsers * slot_bytes);
    if (warp < parsers && lane < 8u) reinterpret_cast<uint32_t *>(cw_lz4_queue_at(queues, warp, qlen))[lane] = 0u;
    __syncthreads();
    if (warp < parsers)
        cw_lzf_parser_role<ALIGNED>(smem + (size_t)warp * slot_bytes, cw_lz4_queue_at(queues, warp, qlen), &bars[warp], src, chunk,
                                    n_chunks, work, qlen);
    else
        cw_lzf_emitter_role(warp - parsers, emitters, parsers, smem, slot_bytes, queues, chunk, dst, stride, lens, qlen);
}

static int g_lzf_max_smem = 0;

cudaError_t cw_lzf_configure(const CwDeviceInfo &dev)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, cw_lzf_kernel<true>);
    if (e != cudaSuccess) return e;
    g_lzf_max_smem = (int)dev.smem_optin - (int)fa.sharedSizeBytes;
    e = cudaFuncSetAttribute(cw_lzf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_lzf_max_smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(cw_lzf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_lzf_max_smem);
}

cudaError_t cw_launch_lzf(const CwDeviceInfo &dev, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                          uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    // rings of 32 records: LZF is the kernel shortest of shared memory (8 KiB chunks: 7 parsers per SM instead of 6)
    uint32_t qlen = 32u;
    if (const char *e = getenv("CW_LZ4_QLEN")) qlen = (uint32_t)atoi(e) == 32u ? 32u : CW_LZ4_QLEN;  // tuning knob
    const uint32_t slot = lzf_slot_bytes(chunk), per = slot + cw_lz4_queue_bytes(qlen);
    int parsers = g_lzf_max_smem / (int)per;
    if (parsers < 1) return cudaErrorInvalidValue;
    if (parsers > CW_LZF_MAX_WARPS) parsers = CW_LZF_MAX_WARPS;
    int emitters = lzf_emitters();
    if (emitters > parsers) emitters = parsers;
    uint64_t ctas = (n + (uint64_t)parsers - 1) / (uint64_t)parsers;
    if (ctas > (uint64_t)dev.sm_count) ctas = (uint64_t)dev.sm_count;
    const bool aligned = ((chunk & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_src) & 15u) == 0);
    const unsigned threads = (unsigned)(parsers + emitters) * 32u;
    const size_t dyn = (size_t)parsers * per;
    if (aligned)
        cw_lzf_kernel<true><<<(unsigned)ctas, threads, dyn, st>>>(d_src, chunk, n, d_dst, stride, d_lens, work, slot,
                                                                  (uint32_t)parsers, (uint32_t)emitters, qlen);
    else
        cw_lzf_kernel<false><<<(unsigned)ctas, threads, dyn, st>>>(d_src, chunk, n, d_dst, stride, d_lens, work, slot,
                                                                   (uint32_t)parsers, (uint32_t)emitters, qlen);
    return cudaGetLastError();
}
