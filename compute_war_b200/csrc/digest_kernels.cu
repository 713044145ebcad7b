// digest_kernels.cu -- what the digests are for: the byte-value distribution table of the reference's randomness
// study and the duplicate map a deduplicating store needs.  SURVEY.md 8(f) row 3.
//
//  * cw_digest_hist_kernel: hist[b*256 + v] = number of digests whose byte b equals v.  This is the table
//    Hasher::LogResultsInternal builds in a std::map (src/hashing_randomness/hash.cpp:30-58); the statistic it prints
//    from it (hash.cpp:74-101) is a 256-term formula per byte and is evaluated on the host (api.cu).
//    Layout: a CTA owns one 16-byte column group of the digests and a range of rows; 16 x 256 u32 counters live in
//    shared memory (16 KB), one 16-byte load per thread and row, flushed once per CTA with 64-bit global atomics.
//    HBM bytes per digest byte: 1.
//
//  * cw_dedup_insert_kernel / cw_dedup_resolve_kernel: first[i] = smallest j with digest_j == digest_i.
//    Open-addressing table of chunk indices (u32, >= 2 slots per digest).  A slot, once taken, belongs to ONE digest
//    value for good: equal digests lower the stored index with atomicMin, different digests move on to the next slot,
//    so equal digests always end in the same slot and the result does not depend on the order threads ran in.
#include "launch.h"

#define CW_DIG_THREADS 256
#define CW_DEDUP_EMPTY 0xFFFFFFFFu

__global__ void __launch_bounds__(CW_DIG_THREADS)
    cw_digest_hist_kernel(const uint8_t *__restrict__ digests, uint32_t D, uint64_t n, uint64_t rows_per_cta, int vec,
                          unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t cnt[16 * 256];
    for (int i = threadIdx.x; i < 16 * 256; i += CW_DIG_THREADS) cnt[i] = 0;
    __syncthreads();
    const uint32_t col0 = blockIdx.y * 16u;
    const uint32_t cols = D - col0 < 16u ? D - col0 : 16u;
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_cta;
    const uint64_t r1 = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
    for (uint64_t r = r0 + threadIdx.x; r < r1; r += CW_DIG_THREADS) {
        const uint8_t *p = digests + r * D + col0;
        if (vec) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 16; k++) atomicAdd(&cnt[k * 256 + ((w[k >> 2] >> (8 * (k & 3))) & 255u)], 1u);
        } else {
            for (uint32_t k = 0; k < cols; k++) atomicAdd(&cnt[k * 256 + __ldg(p + k)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cols * 256u; i += CW_DIG_THREADS) {
        const uint32_t c = cnt[i];
        if (c) atomicAdd(&hist[(uint64_t)col0 * 256u + i], (unsigned long long)c);
    }
}

__device__ __forceinline__ uint32_t cw_dedup_slot(const uint32_t *w, uint32_t words, uint32_t mask)
{
    // digests of a cryptographic hash are uniform already; the multiply-xorshift only protects the table against
    // callers that hand in structured "digests" (tests do)
    uint64_t h = (uint64_t)w[0] | ((uint64_t)w[words > 1 ? 1 : 0] << 32);
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 29;
    return (uint32_t)h & mask;
}

__device__ __forceinline__ bool cw_digest_equal(const uint32_t *a, const uint32_t *b, uint32_t words)
{
    for (uint32_t k = 0; k < words; k++)
        if (a[k] != b[k]) return false;
    return true;
}

// digests are read as u32 words: D % 4 == 0 and a 4-byte aligned base are required by the API
__global__ void __launch_bounds__(CW_DIG_THREADS)
    cw_dedup_insert_kernel(const uint32_t *__restrict__ digests, uint32_t words, uint64_t n, uint32_t *table, uint32_t mask)
{
    const uint64_t i = (uint64_t)blockIdx.x * CW_DIG_THREADS + threadIdx.x;
    if (i >= n) return;
    const uint32_t *mine = digests + i * words;
    uint32_t h = cw_dedup_slot(mine, words, mask);
    for (;;) {
        uint32_t cur = *reinterpret_cast<volatile uint32_t *>(table + h);
        if (cur == CW_DEDUP_EMPTY) {
            cur = atomicCAS(table + h, CW_DEDUP_EMPTY, (uint32_t)i);
            if (cur == CW_DEDUP_EMPTY) return;
        }
        if (cw_digest_equal(digests + (uint64_t)cur * words, mine, words)) {
            atomicMin(table + h, (uint32_t)i);
            return;
        }
        h = (h + 1u) & mask;
    }
}

__global__ void __launch_bounds__(CW_DIG_THREADS)
    cw_dedup_resolve_kernel(const uint32_t *__restrict__ digests, uint32_t words, uint64_t n, const uint32_t *__restrict__ table,
                            uint32_t mask, uint32_t *__restrict__ first, unsigned long long *__restrict__ n_unique)
{
    const uint64_t i = (uint64_t)blockIdx.x * CW_DIG_THREADS + threadIdx.x;
    bool keep = false;
    if (i < n) {
        const uint32_t *mine = digests + i * words;
        uint32_t h = cw_dedup_slot(mine, words, mask);
        uint32_t cur;
        for (;;) {
            cur = table[h];
            if (cw_digest_equal(digests + (uint64_t)cur * words, mine, words)) break;
            h = (h + 1u) & mask;
        }
        first[i] = cur;
        keep = cur == (uint32_t)i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(n_unique, (unsigned long long)__popc(m));
}

cudaError_t cw_launch_digest_hist(const CwDeviceInfo &dev, const uint8_t *d_digests, uint32_t D, uint64_t n,
                                  unsigned long long *d_hist, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_hist, 0, (size_t)D * 256 * sizeof(unsigned long long), st);
    if (e != cudaSuccess || n == 0) return e;
    const uint32_t groups = (D + 15) / 16;
    // enough CTAs to fill the machine a few times over, few enough that the flush (4096 atomics) stays negligible
    uint64_t ctas = (uint64_t)dev.sm_count * 8 / groups + 1;
    uint64_t rows = (n + ctas - 1) / ctas;
    if (rows < CW_DIG_THREADS) rows = CW_DIG_THREADS;
    ctas = (n + rows - 1) / rows;
    const int vec = (D % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_digests) & 15u) == 0);
    cw_digest_hist_kernel<<<dim3((unsigned)ctas, groups), CW_DIG_THREADS, 0, st>>>(d_digests, D, n, rows, vec, d_hist);
    return cudaGetLastError();
}

cudaError_t cw_launch_dedup(const uint8_t *d_digests, uint32_t D, uint64_t n, uint32_t *d_table, uint32_t table_slots,
                            uint32_t *d_first, unsigned long long *d_n_unique, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_n_unique, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess || n == 0) return e;
    e = cudaMemsetAsync(d_table, 0xFF, (size_t)table_slots * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((n + CW_DIG_THREADS - 1) / CW_DIG_THREADS);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(d_digests);
    cw_dedup_insert_kernel<<<blocks, CW_DIG_THREADS, 0, st>>>(w, D / 4, n, d_table, table_slots - 1);
    cw_dedup_resolve_kernel<<<blocks, CW_DIG_THREADS, 0, st>>>(w, D / 4, n, d_table, table_slots - 1, d_first, d_n_unique);
    return cudaGetLastError();
}
