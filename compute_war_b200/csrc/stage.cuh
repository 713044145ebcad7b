// stage.cuh -- moving one chunk from HBM into a warp's shared-memory slot (CUDA only).
//
// Fast path: one elected lane issues a 1-D bulk async copy (cp.async.bulk, SASS UBLKCP -- the TMA engine
// without a tensor map) that lands the whole chunk in shared memory and signals an mbarrier; the warp
// sleeps on the barrier instead of spending issue slots on LDG/STS pairs.  Needs 16-byte aligned source,
// destination and size.  Slow path (ragged chunk sizes / unaligned buffers): coalesced byte loads.
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t cw_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cw_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cw_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// lane-0 only: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void cw_bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    // order earlier generic-proxy accesses to this slot (the previous chunk's reads) before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cw_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     cw_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(cw_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void cw_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = cw_smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CW_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CW_DONE_%=;\n"
        "bra CW_WAIT_%=;\n"
        "CW_DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}

// Stage chunk `g` (chunk bytes) into smem slot s; s has >= chunk+16 bytes, the 16 bytes after the chunk are
// kept zero by the caller.  `phase` is the warp's barrier phase bit (flipped here).  All 32 lanes call.
__device__ __forceinline__ void cw_stage_chunk(uint8_t *s, const uint8_t *__restrict__ g, uint32_t chunk, bool bulk_ok,
                                               uint64_t *bar, uint32_t &phase)
{
    const unsigned lane = threadIdx.x & 31u;
    __syncwarp();
    if (bulk_ok) {
        if (lane == 0) cw_bulk_load(s, g, chunk, bar);
        cw_mbar_wait(bar, phase);
        phase ^= 1u;
    } else {
        for (uint32_t i = lane; i < chunk; i += 32u) s[i] = g[i];
        __syncwarp();
    }
}
