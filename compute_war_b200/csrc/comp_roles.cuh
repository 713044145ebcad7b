// comp_roles.cuh -- the per-warp loops of the compressor kernels (CUDA only): a PARSER warp owns a
// shared-memory slot and a chunk at a time; an EMITTER warp serves several parsers' record queues
// (lz4_warp.cuh).  Used by lz4_kernels.cu and fused_kernels.cu.
#pragma once
#include "lz4_warp.cuh"
#include "lz4_fast.cuh"
#include "lzf_warp.cuh"
#include "stage.cuh"

__host__ __device__ constexpr uint32_t cw_lz4_chunk_pad(uint32_t chunk) { return (chunk + CW_LZ4_PAD + 15u) & ~15u; }

// LZ4 parser role.  s: this warp's slot [chunk | pad | table] (DIRECT: [table] only, the chunk is read where it
// lies in global memory); q: this warp's queue (zeroed by the caller before the CTA-wide barrier).
template <int HASHLOG, bool ALIGNED, bool DIRECT>
__device__ __forceinline__ void cw_lz4_parser_role(uint8_t *s, CwLz4Queue *q, uint64_t *bar, const uint8_t *__restrict__ src,
                                                   uint32_t chunk, uint64_t n_chunks, unsigned long long *work,
                                                   uint32_t qlen = CW_LZ4_QLEN)
{
    const unsigned lane = threadIdx.x & 31u;
    if (DIRECT) {
        uint16_t *table = reinterpret_cast<uint16_t *>(s);
        uint32_t qh = 0;
        for (;;) {
            unsigned long long g = 0;
            if (lane == 0) g = atomicAdd(work, 1ull);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= n_chunks) break;
            cw_lz4_q_drain(q, qh, lane);  // one chunk index per queue at a time
            uint4 *t4 = reinterpret_cast<uint4 *>(table);
#pragma unroll 8
            for (uint32_t i = lane; i < (2u << HASHLOG) / 16u; i += 32u) t4[i] = make_uint4(0, 0, 0, 0);
            if (lane == 0) {
                q->chunk_lo = (uint32_t)g;
                q->chunk_hi = (uint32_t)(g >> 32);
            }
            __syncwarp();
            cw_lz4_parse_warp<HASHLOG, true>(src + g * chunk, chunk, table, q, qh, qlen);
        }
        cw_lz4_q_drain(q, qh, lane);
        if (lane == 0) {
            cw_fence_cta();
            cw_st_vol(&q->done, 1u);
        }
        return;
    }
    const uint32_t chunk_pad = cw_lz4_chunk_pad(chunk);
    uint16_t *table = reinterpret_cast<uint16_t *>(s + chunk_pad);
    if (lane == 0) cw_mbar_init(bar, 1);
    for (uint32_t i = chunk + lane; i < chunk_pad; i += 32u) s[i] = 0;
    __syncwarp();
    uint32_t phase = 0, qh = 0;
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(work, 1ull);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_chunks) break;
        cw_lz4_q_drain(q, qh, lane);  // the emitter still reads literals of the previous chunk from this slot
        if (ALIGNED && lane == 0) cw_bulk_load(s, src + g * chunk, chunk, bar);
        {
            uint4 *t4 = reinterpret_cast<uint4 *>(table);
#pragma unroll 8
            for (uint32_t i = lane; i < (2u << HASHLOG) / 16u; i += 32u) t4[i] = make_uint4(0, 0, 0, 0);
        }
        if (lane == 0) {
            q->chunk_lo = (uint32_t)g;
            q->chunk_hi = (uint32_t)(g >> 32);
        }
        if (ALIGNED) {
            cw_mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            const uint8_t *gp = src + g * chunk;
            for (uint32_t i = lane; i < chunk; i += 32u) s[i] = gp[i];
        }
        __syncwarp();
        cw_lz4_parse_warp<HASHLOG, false>(s, chunk, table, q, qh, qlen);
    }
    cw_lz4_q_drain(q, qh, lane);
    if (lane == 0) {
        cw_fence_cta();
        cw_st_vol(&q->done, 1u);
    }
}

// LZ4 emitter role: serves the queues e, e + emitters, e + 2*emitters, ... until all their parsers are done.
// DIRECT: literals come from the input buffer `src` itself instead of the parsers' shared-memory slots.
template <bool DIRECT>
__device__ __forceinline__ void cw_lz4_emitter_role(unsigned e, unsigned emitters, unsigned parsers, uint8_t *smem,
                                                    uint32_t slot_bytes, CwLz4Queue *queues, const uint8_t *__restrict__ src,
                                                    uint32_t chunk, uint8_t *__restrict__ dst, uint64_t stride,
                                                    uint32_t *__restrict__ lens, uint32_t qlen = CW_LZ4_QLEN)
{
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        bool all_done = true;
        uint32_t did = 0;
        for (unsigned p = e; p < parsers; p += emitters) {
            CwLz4Queue *q = cw_lz4_queue_at(queues, p, qlen);
            uint32_t d = 0;
            if (lane == 0) d = cw_ld_vol(&q->done);  // read BEFORE the head: done is set after the last publish
            d = __shfl_sync(0xffffffffu, d, 0);
            const uint32_t cnt = cw_lz4_emit_batch<DIRECT>(q, DIRECT ? src : smem + (size_t)p * slot_bytes, chunk, dst, stride, lens, qlen);
            did += cnt;
            if (!(d && cnt == 0u)) all_done = false;
        }
        if (all_done) break;
        if (!did) cw_spin();
    }
}

// LZ4 fast mode (lz4_fast.cuh): a TEAM of nteam warps stages, parses and emits a chunk; no queues, no emitter warps.
// slot of a team: [chunk | pad | table | CwLz4fTeam | nteam x CwLz4fScratch] (DIRECT: no chunk, it is read where it lies).
__host__ __device__ constexpr uint32_t cw_lz4f_slot_bytes(uint32_t chunk, int hashlog, bool direct, uint32_t nteam)
{
    return ((direct ? 0u : cw_lz4_chunk_pad(chunk)) + (2u << hashlog) + (uint32_t)sizeof(CwLz4fTeam) +
            nteam * (uint32_t)sizeof(CwLz4fScratch) + 15u) & ~15u;  // slots stay 16-byte aligned for the bulk copy
}
// Which team a compress warp belongs to when `warps` warps share `teams` slots: the first warps % teams teams get one
// warp more than the others (13 slots of 8 KiB chunks and 32 warps: six teams of 3, seven of 2), so that the register
// file, not the number of slots, decides how many warps an SM runs.
struct CwLz4fPlace {
    unsigned team, rank, nteam;
};
__host__ __device__ __forceinline__ CwLz4fPlace cw_lz4f_place(unsigned warp, unsigned teams, unsigned warps)
{
    const unsigned base = warps / teams, extra = warps - base * teams, cut = extra * (base + 1u);
    if (warp < cut) return {warp / (base + 1u), warp % (base + 1u), base + 1u};
    const unsigned w = warp - cut;
    return {extra + w / base, w % base, base};
}
// barrier of the warps of one team (hardware barrier `id`, 1..15); a team of one warp needs none
__device__ __forceinline__ void cw_lz4f_team_sync(unsigned id, unsigned nteam)
{
    if (nteam == 1u) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nteam * 32u) : "memory");
}
// rank: this warp's index inside its team; bar_id: the team's hardware barrier; bar: the team's mbarrier
// CODEC: CW_F_LZ4 / CW_F_LZF (an LZF block may take chunk - 1 bytes, like the reference's lzf_compress(s, l, d, l - 1) call)
template <int HASHLOG, bool ALIGNED, bool DIRECT, int CODEC = CW_F_LZ4>
__device__ __forceinline__ void cw_lz4f_team_role(uint8_t *s, uint64_t *bar, unsigned rank, unsigned nteam, unsigned bar_id,
                                                  const uint8_t *__restrict__ src, uint32_t chunk, uint64_t n_chunks,
                                                  unsigned long long *work, uint8_t *__restrict__ dst, uint64_t stride,
                                                  uint32_t *__restrict__ lens)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t chunk_pad = DIRECT ? 0u : cw_lz4_chunk_pad(chunk);
    uint16_t *table = reinterpret_cast<uint16_t *>(s + chunk_pad);
    CwLz4fTeam *team = reinterpret_cast<CwLz4fTeam *>(s + chunk_pad + (2u << HASHLOG));
    CwLz4fScratch *scr = reinterpret_cast<CwLz4fScratch *>(team + 1) + rank;
    if (rank == 0u) {
        if (lane == 0) {
            for (unsigned r = 0; r < CW_LZ4F_MAX_TEAM; r++) {
                cw_lz4f_tok_init(&team->tok_table[r]);
                cw_lz4f_tok_init(&team->tok_walk[r]);
                cw_lz4f_tok_init(&team->tok_emit[r]);
            }
            cw_mbar_init(bar, 1);  // also fences the initialisations above
        }
        if (!DIRECT)
            for (uint32_t i = chunk + lane; i < chunk_pad; i += 32u) s[i] = 0;
    }
    cw_lz4f_team_sync(bar_id, nteam);
    uint32_t phase = 0, seen = 0;  // seen: arrivals consumed on this warp's tokens
    for (;;) {
        // the slot is free (barrier at the end of the loop): next chunk index, tokens, bulk copy, zeroed table
        if (rank == 0u && lane == 0) {
            const unsigned long long g = atomicAdd(work, 1ull);
            team->chunk_lo = (uint32_t)g;
            team->chunk_hi = (uint32_t)(g >> 32);
            team->ip = 1u;
            team->op = 0u;
            if (!DIRECT && ALIGNED && g < n_chunks) cw_bulk_load(s, src + g * chunk, chunk, bar);
        }
        {
            uint4 *t4 = reinterpret_cast<uint4 *>(table);
            for (uint32_t i = rank * 32u + lane; i < (2u << HASHLOG) / 16u; i += nteam * 32u) t4[i] = make_uint4(0, 0, 0, 0);
        }
        cw_lz4f_team_sync(bar_id, nteam);
        const unsigned long long g = (unsigned long long)team->chunk_lo | ((unsigned long long)team->chunk_hi << 32);
        if (g >= n_chunks) break;
        if (!DIRECT) {
            if (ALIGNED) {
                cw_mbar_wait(bar, phase);
                phase ^= 1u;
            } else {
                const uint8_t *gp = src + g * chunk;
                for (uint32_t i = rank * 32u + lane; i < chunk; i += nteam * 32u) s[i] = gp[i];
                cw_lz4f_team_sync(bar_id, nteam);
            }
        }
        // the block's address, made opaque: left to itself the compiler recomputes dst + g * stride (a 64-bit multiply-add and
        // four constant loads) in front of every one of the kernel's byte stores instead of holding it in two registers
        uint8_t *outp = dst + g * stride;
        asm volatile("" : "+l"(outp));
        __builtin_assume(__isGlobal(outp));  // (still a global address: STG, not generic stores)
        const uint32_t len =
            cw_lz4f_team_chunk<CODEC, HASHLOG, DIRECT>(DIRECT ? src + g * chunk : s, chunk, table, team, scr, outp, rank, nteam,
                                                       seen, CODEC == CW_F_LZF ? chunk - 1u : 0xffffffffu);
        if (len != CW_LZ4_NONE && lane == 0) lens[g] = len;
        cw_lz4f_team_sync(bar_id, nteam);  // every warp is done with the slot before it is refilled
    }
}

// LZF parser role; slot = [chunk | 16 B pad | table of cw_lzf_table_bytes(format)], one queue per parser as for LZ4
template <bool ALIGNED>
__device__ __forceinline__ void cw_lzf_parser_role(uint8_t *s, CwLz4Queue *q, uint64_t *bar, const uint8_t *__restrict__ src,
                                                   uint32_t chunk, uint64_t n_chunks, unsigned long long *work,
                                                   uint32_t qlen = CW_LZ4_QLEN)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t chunk_pad = (chunk + 16u + 15u) & ~15u;
    uint16_t *tab = reinterpret_cast<uint16_t *>(s + chunk_pad);
    const int bb = cw_lzf_table_format(chunk);
    if (lane == 0) cw_mbar_init(bar, 1);
    for (uint32_t i = chunk + lane; i < chunk_pad; i += 32u) s[i] = 0;
    __syncwarp();
    uint32_t phase = 0, qh = 0;
    for (;;) {
        unsigned long long g = 0;
        if (lane == 0) g = atomicAdd(work, 1ull);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_chunks) break;
        cw_lz4_q_drain(q, qh, lane);  // the emitter still reads literals of the previous chunk from this slot
        if (lane == 0) {
            q->chunk_lo = (uint32_t)g;
            q->chunk_hi = (uint32_t)(g >> 32);
        }
        if (ALIGNED && lane == 0) cw_bulk_load(s, src + g * chunk, chunk, bar);
        if (bb == 12 || bb == 25) cw_lzf_table_init<12>(tab);
        else if (bb == 24) cw_lzf_table_init<11>(tab);
        else if (bb == 13) cw_lzf_table_init<13>(tab);
        else cw_lzf_table_init<0>(tab);
        if (ALIGNED) {
            cw_mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            const uint8_t *gp = src + g * chunk;
            for (uint32_t i = lane; i < chunk; i += 32u) s[i] = gp[i];
        }
        __syncwarp();
        if (bb == 12) cw_lzf_parse_warp<12, 12>(s, chunk, tab, q, qh, qlen);
        else if (bb == 25) cw_lzf_parse_warp<12, 13>(s, chunk, tab, q, qh, qlen);
        else if (bb == 24) cw_lzf_parse_warp<11, 13>(s, chunk, tab, q, qh, qlen);
        else if (bb == 13) cw_lzf_parse_warp<13, 13>(s, chunk, tab, q, qh, qlen);
        else cw_lzf_parse_warp<0, 13>(s, chunk, tab, q, qh, qlen);
    }
    cw_lz4_q_drain(q, qh, lane);
    if (lane == 0) {
        cw_fence_cta();
        cw_st_vol(&q->done, 1u);
    }
}

// LZF emitter role: same protocol as cw_lz4_emitter_role
__device__ __forceinline__ void cw_lzf_emitter_role(unsigned e, unsigned emitters, unsigned parsers, uint8_t *smem,
                                                    uint32_t slot_bytes, CwLz4Queue *queues, uint32_t chunk,
                                                    uint8_t *__restrict__ dst, uint64_t stride, uint32_t *__restrict__ lens,
                                                    uint32_t qlen = CW_LZ4_QLEN)
{
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        bool all_done = true;
        uint32_t did = 0;
        for (unsigned p = e; p < parsers; p += emitters) {
            CwLz4Queue *q = cw_lz4_queue_at(queues, p, qlen);
            uint32_t d = 0;
            if (lane == 0) d = cw_ld_vol(&q->done);  // read BEFORE the head: done is set after the last publish
            d = __shfl_sync(0xffffffffu, d, 0);
            const uint32_t cnt = cw_lzf_emit_batch(q, smem + (size_t)p * slot_bytes, chunk, chunk - 1u, dst, stride, lens, qlen);
            did += cnt;
            if (!(d && cnt == 0u)) all_done = false;
        }
        if (all_done) break;
        if (!did) cw_spin();
    }
}
