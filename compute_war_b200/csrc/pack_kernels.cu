// pack_kernels.cu -- packed output: the strided block array + length table compacted into one contiguous stream.
//
// SURVEY.md 8(f) row 2.  The reference throws its compressed blocks away (HashAndCompress.cpp:259-260); a storage
// front-end keeps them back to back.  offsets[i] = sum over j < i of stored(j), where stored(j) = lens[j], or
// chunk bytes when lens[j] == 0 -- lzf_compress's "did not fit" (lzf.h:59-64), for which the chunk itself is stored raw.
//
// Three small kernels: per-tile exclusive scan of the stored sizes (1024 per CTA, warp shuffles), scan of the tile
// totals by one CTA, then a copy kernel (one warp per chunk, 16-byte loads where the source allows).
#include "launch.h"

#define CW_PACK_TILE 1024

__device__ __forceinline__ uint32_t cw_min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

__device__ __forceinline__ uint32_t cw_stored(const uint32_t *lens, uint64_t i, uint32_t chunk)
{
    const uint32_t l = lens[i];
    return l ? l : chunk;
}

// offsets[i] = exclusive prefix within the tile; tile_sum[tile] = total of the tile
__global__ void __launch_bounds__(CW_PACK_TILE) cw_pack_scan_tiles(const uint32_t *__restrict__ lens, uint32_t chunk, uint64_t n,
                                                                  unsigned long long *__restrict__ offsets,
                                                                  unsigned long long *__restrict__ tile_sum)
{
    __shared__ unsigned long long warp_tot[32];
    const uint64_t i = (uint64_t)blockIdx.x * CW_PACK_TILE + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned long long v = i < n ? cw_stored(lens, i, chunk) : 0ull, inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_tot[lane], winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (unsigned)d) winc += t;
        }
        warp_tot[lane] = winc - w;  // exclusive
        if (lane == 31) tile_sum[blockIdx.x] = winc;
    }
    __syncthreads();
    if (i < n) offsets[i] = warp_tot[warp] + inc - v;
}

// exclusive scan of the tile totals in place (one CTA; tiles <= 2^21 chunks / 1024 = 2048, looped for more)
__global__ void __launch_bounds__(1024) cw_pack_scan_totals(unsigned long long *tile_sum, uint64_t tiles, unsigned long long *total)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < tiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        unsigned long long v = i < tiles ? tile_sum[i] : 0ull, inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned)d) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_tot[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= (unsigned)d) winc += t;
            }
            warp_tot[lane] = winc - w;
        }
        __syncthreads();
        const unsigned long long c = carry;
        if (i < tiles) tile_sum[i] = c + warp_tot[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warp_tot[31] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// one warp per chunk: add the tile base, write offsets[n] = total, copy the stored bytes
__global__ void __launch_bounds__(256) cw_pack_copy(const uint8_t *__restrict__ blocks, uint64_t stride,
                                                    const uint32_t *__restrict__ lens, const uint8_t *__restrict__ src,
                                                    uint32_t chunk, uint64_t n, unsigned long long *__restrict__ offsets,
                                                    const unsigned long long *__restrict__ tile_base,
                                                    const unsigned long long *__restrict__ total, uint8_t *__restrict__ packed,
                                                    uint64_t packed_cap)
{
    const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (g >= n) return;
    const unsigned long long off = offsets[g] + tile_base[g / CW_PACK_TILE];
    __syncwarp();
    if (lane == 0) {
        offsets[g] = off;
        if (g == n - 1) offsets[n] = *total;
    }
    const uint32_t l = lens[g];
    const uint32_t cnt = l ? l : chunk;
    const uint8_t *from = l ? blocks + g * stride : src + g * (uint64_t)chunk;
    if (off + cnt > packed_cap) return;  // caller's buffer too small: offsets still tell how much is needed
    uint8_t *to = packed + off;
    // destination alignment is arbitrary: bytes up to a 4-byte boundary, then words when the source agrees
    const uint32_t head = cw_min_u32((4u - (uint32_t)(reinterpret_cast<uintptr_t>(to) & 3u)) & 3u, cnt);
    if (lane < head) to[lane] = from[lane];
    if (((reinterpret_cast<uintptr_t>(from + head)) & 3u) == 0) {
        const uint32_t words = (cnt - head) >> 2;
        const uint32_t *f4 = reinterpret_cast<const uint32_t *>(from + head);
        uint32_t *t4 = reinterpret_cast<uint32_t *>(to + head);
        for (uint32_t w = lane; w < words; w += 32u) t4[w] = f4[w];
        const uint32_t done = head + 4u * words;
        if (lane < cnt - done) to[done + lane] = from[done + lane];
    } else {
        for (uint32_t x = head + lane; x < cnt; x += 32u) to[x] = from[x];
    }
}

cudaError_t cw_launch_pack(const uint8_t *d_blocks, uint64_t stride, const uint32_t *d_lens, const uint8_t *d_src, uint32_t chunk,
                           uint64_t n, uint8_t *d_packed, uint64_t packed_cap, unsigned long long *d_offsets,
                           unsigned long long *d_scratch, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const uint64_t tiles = (n + CW_PACK_TILE - 1) / CW_PACK_TILE;
    unsigned long long *tile_sum = d_scratch, *total = d_scratch + tiles;
    cw_pack_scan_tiles<<<(unsigned)tiles, CW_PACK_TILE, 0, st>>>(d_lens, chunk, n, d_offsets, tile_sum);
    cw_pack_scan_totals<<<1, 1024, 0, st>>>(tile_sum, tiles, total);
    const uint64_t threads = n * 32;
    cw_pack_copy<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_blocks, stride, d_lens, d_src, chunk, n, d_offsets, tile_sum,
                                                                   total, d_packed, packed_cap);
    return cudaGetLastError();
}
