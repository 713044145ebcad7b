// lz4_kernels.cu -- LZ4 compression of independent chunks: persistent CTAs of parser and emitter warps.
//
// Device-side replacement for the lz4 doCompression lambda (HashAndCompress.cpp:349-355) applied to
// n_chunks chunks; per-chunk algorithm in lz4_warp.cuh.
//
// One CTA per SM.  Every PARSER warp owns a shared-memory slot = [chunk | 144 B pad | u16 table] and pulls
// chunk indices from a global counter until the batch is exhausted (chunk cost varies by an order of
// magnitude between incompressible and text-like data, so static partitioning would leave SMs idle).  The
// number of parsers is whatever fits in the SM's 227 KB: at 8 KB chunks 9 with the reference's 8192-entry
// table (compat) and 13 with the 4096-entry table of fast mode.  A few EMITTER warps (no slot of their own)
// turn the parsers' match records into output bytes.
// compat mode runs the exact sequential parse (lz4_warp.cuh) with parser and emitter warps; fast mode runs
// cw_lz4f_kernel below: the window-parallel parse of lz4_fast.cuh by self-contained teams of warps.
#include <cstdlib>

#include "comp_roles.cuh"
#include "launch.h"
#include "role_map.h"

#define CW_LZ4_MAX_PARSERS 16
#define CW_LZ4_MAX_EMITTERS 4
#define CW_LZ4_EMITTERS_DEFAULT 3
#define CW_LZ4_DIRECT_ABOVE 16384u  // chunks larger than this are read from global memory instead of being staged
#define CW_LZ4_COMPAT_DIRECT_ABOVE 8192u  // ... compat mode: already above 8 KiB (cw_launch_lz4)

static int lz4_emitters()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CW_LZ4_EMITTERS");  // tuning knob for the profiling scripts
        v = e ? atoi(e) : CW_LZ4_EMITTERS_DEFAULT;
        if (v < 1) v = 1;
        if (v > CW_LZ4_MAX_EMITTERS) v = CW_LZ4_MAX_EMITTERS;
    }
    return v;
}

// DIRECT (chunks > CW_LZ4_DIRECT_ABOVE bytes, e.g. the 64 KB blocks of BASELINE configs[2]): a slot holds only the
// table; parsers and emitters read the chunk where it lies (L2), which keeps 13 chunks per SM in flight instead of 2.
template <int HASHLOG, bool ALIGNED, bool DIRECT>
__global__ void __launch_bounds__((CW_LZ4_MAX_PARSERS + CW_LZ4_MAX_EMITTERS) * 32, 1)
    cw_lz4_kernel(const uint8_t *__restrict__ src, uint32_t chunk, uint64_t n_chunks, uint8_t *__restrict__ dst,
                  uint64_t stride, uint32_t *__restrict__ lens, unsigned long long *work, uint32_t slot_bytes,
                  uint32_t parsers, uint32_t emitters, const __grid_constant__ CwRoleMap roles, uint32_t qlen)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[32];
    const unsigned warp = roles.logical[threadIdx.x >> 5], lane = threadIdx.x & 31u;
    CwLz4Queue *queues = reinterpret_cast<CwLz4Queue *>(smem + (size_t)parsers * slot_bytes);
    if (warp < parsers && lane < 8u) reinterpret_cast<uint32_t *>(cw_lz4_queue_at(queues, warp, qlen))[lane] = 0u;  // header
    __syncthreads();
    if (warp < parsers)
        cw_lz4_parser_role<HASHLOG, ALIGNED, DIRECT>(smem + (size_t)warp * slot_bytes, cw_lz4_queue_at(queues, warp, qlen),
                                                     &bars[warp], src, chunk, n_chunks, work, qlen);
    else
        cw_lz4_emitter_role<DIRECT>(warp - parsers, emitters, parsers, smem, slot_bytes, queues, src, chunk, dst, stride, lens,
                                    qlen);
}

// fast mode (lz4_fast.cuh): teams of `nteam` warps, one chunk per team at a time; every team stages, parses and emits
// on its own.  As many teams as slots fit into the SM's shared memory (and hardware barriers: 15).
#ifndef CW_LZ4F_MAX_WARPS
#define CW_LZ4F_MAX_WARPS 32  // 1024 threads: ptxas may use 64 registers (the kernel needs 62, no spills)
#endif
#define CW_LZ4F_MAX_TEAMS 15
template <int HASHLOG, bool ALIGNED, bool DIRECT, int CODEC = CW_F_LZ4>
__global__ void __launch_bounds__(CW_LZ4F_MAX_WARPS * 32, 1)
    cw_lz4f_kernel(const uint8_t *__restrict__ src, uint32_t chunk, uint64_t n_chunks, uint8_t *__restrict__ dst, uint64_t stride,
                   uint32_t *__restrict__ lens, unsigned long long *work, uint32_t slot_bytes, uint32_t teams)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[CW_LZ4F_MAX_TEAMS];
    const CwLz4fPlace pl = cw_lz4f_place(threadIdx.x >> 5, teams, blockDim.x >> 5);
    cw_lz4f_team_role<HASHLOG, ALIGNED, DIRECT, CODEC>(smem + (size_t)pl.team * slot_bytes, &bars[pl.team], pl.rank, pl.nteam,
                                                       1u + pl.team, src, chunk, n_chunks, work, dst, stride, lens);
}

// Teams and warps of a fast-mode CTA.  Shared memory decides how many chunks an SM holds (`slots`: 13 at 8 KiB, 6 tables
// for chunks read in place), the register file how many warps it runs (`max_warps`); the warps are spread over the
// slots as evenly as they go (cw_lz4f_place).  CW_LZ4F_TEAM = n forces n warps per chunk (profiling scripts).
void cw_lz4f_shape(int slots, int max_warps, int *teams, int *warps)
{
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("CW_LZ4F_TEAM");
        forced = e ? atoi(e) : 0;
        if (forced > CW_LZ4F_MAX_TEAM) forced = CW_LZ4F_MAX_TEAM;
    }
    int t = slots < CW_LZ4F_MAX_TEAMS ? slots : CW_LZ4F_MAX_TEAMS;  // one hardware barrier per team
    if (t > max_warps) t = max_warps;
    int w;
    if (forced > 0) {
        if (t * forced > max_warps) t = max_warps / forced;
        w = t * forced;
    } else {
        w = t * CW_LZ4F_MAX_TEAM < max_warps ? t * CW_LZ4F_MAX_TEAM : max_warps;
    }
    *teams = t;
    *warps = w;
}

static int g_lz4_max_smem = 0, g_lz4f_max_smem = 0;

cudaError_t cw_lz4_configure(const CwDeviceInfo &dev)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, cw_lz4_kernel<13, true, false>);
    if (e != cudaSuccess) return e;
    g_lz4_max_smem = (int)dev.smem_optin - (int)fa.sharedSizeBytes;  // opt-in limit covers static + dynamic
    const int m = g_lz4_max_smem;
#define CW_CFG(HL, AL, DI) \
    if ((e = cudaFuncSetAttribute(cw_lz4_kernel<HL, AL, DI>, cudaFuncAttributeMaxDynamicSharedMemorySize, m)) != cudaSuccess) return e;
    CW_CFG(13, true, false) CW_CFG(13, false, false) CW_CFG(13, true, true)
#undef CW_CFG
    if ((e = cudaFuncGetAttributes(&fa, cw_lz4f_kernel<12, true, false>)) != cudaSuccess) return e;
    g_lz4f_max_smem = (int)dev.smem_optin - (int)fa.sharedSizeBytes;
    const int mf = g_lz4f_max_smem;
#define CW_CFG(HL, AL, DI) \
    if ((e = cudaFuncSetAttribute(cw_lz4f_kernel<HL, AL, DI>, cudaFuncAttributeMaxDynamicSharedMemorySize, mf)) != cudaSuccess) return e;
    CW_CFG(12, true, false) CW_CFG(12, false, false) CW_CFG(13, true, false) CW_CFG(13, false, false)
    CW_CFG(14, true, true) CW_CFG(14, false, false)
#undef CW_CFG
#define CW_CFG(HL, AL, DI)                                                                                                         \
    if ((e = cudaFuncSetAttribute(cw_lz4f_kernel<HL, AL, DI, CW_F_LZF>, cudaFuncAttributeMaxDynamicSharedMemorySize, mf)) != cudaSuccess) \
        return e;
    CW_CFG(13, true, false) CW_CFG(13, false, false) CW_CFG(14, true, false) CW_CFG(14, false, false) CW_CFG(14, true, true)
#undef CW_CFG
    return cudaSuccess;
}

cudaError_t cw_launch_lz4(const CwDeviceInfo &dev, bool fast, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                          uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    const bool aligned = ((chunk & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_src) & 15u) == 0);
    const bool direct = aligned && chunk > CW_LZ4_DIRECT_ABOVE;
    if (fast) {
        // 12 / 13 / 14 table bits by chunk size (lz4_fast.cuh); teams of warps, no emitters
        const int hl = cw_lz4f_hashlog(chunk);
        // a slot is sized for the largest team it may get; the slot count follows from that, then the final shape
        int teams = 0, warps = 0;
        uint32_t fslot = cw_lz4f_slot_bytes(chunk, hl, direct, CW_LZ4F_MAX_TEAM);
        for (int it = 0; it < 2; it++) {
            const int slots = g_lz4f_max_smem / (int)fslot;
            if (slots < 1) return cudaErrorInvalidValue;
            cw_lz4f_shape(slots, CW_LZ4F_MAX_WARPS, &teams, &warps);
            fslot = cw_lz4f_slot_bytes(chunk, hl, direct, (uint32_t)((warps + teams - 1) / teams));
        }
        uint64_t fctas = (n + (uint64_t)teams - 1) / (uint64_t)teams;
        if (fctas > (uint64_t)dev.sm_count) fctas = (uint64_t)dev.sm_count;
#define CW_GOF(HL, AL, DI)                                                                                               \
    cw_lz4f_kernel<HL, AL, DI><<<(unsigned)fctas, warps * 32, (size_t)teams * fslot, st>>>(d_src, chunk, n, d_dst, stride, \
                                                                                          d_lens, work, fslot, (uint32_t)teams)
        if (hl == 12) {
            if (aligned) CW_GOF(12, true, false);
            else CW_GOF(12, false, false);
        } else if (hl == 13) {
            if (aligned) CW_GOF(13, true, false);
            else CW_GOF(13, false, false);
        } else {
            if (direct) CW_GOF(14, true, true);
            else CW_GOF(14, false, false);
        }
#undef CW_GOF
        return cudaGetLastError();
    }
    const int hashlog = 13;  // compat: the reference's 13-bit table
    // compat reads chunks above 8 KiB in place (a slot is then only the table: 13 parsers per SM instead of 6 at 16 KiB,
    // bytes from L1 / L2): 16 KiB chunks +22 % on the mix, +40 % on text; 12 KiB +7 / +23 %; at 8 KiB and below staging wins by
    // 3-16 % (profiles/r02_ab_compat_read_in_place.txt).  CW_LZ4_COMPAT_DIRECT_ABOVE: the threshold, for A/B runs
    static const uint32_t direct_above = [] {
        const char *e = getenv("CW_LZ4_COMPAT_DIRECT_ABOVE");
        return e ? (uint32_t)atoi(e) : CW_LZ4_COMPAT_DIRECT_ABOVE;
    }();
    const bool directc = aligned && chunk > direct_above;
    const uint32_t slot = (directc ? 0u : cw_lz4_chunk_pad(chunk)) + (2u << hashlog);
    // a ring of 32 records instead of 64 where that buys another parser (4 KiB chunks: 11 instead of 10)
    uint32_t qlen = CW_LZ4_QLEN;
    if (g_lz4_max_smem / (int)(slot + cw_lz4_queue_bytes(32u)) > g_lz4_max_smem / (int)(slot + cw_lz4_queue_bytes(CW_LZ4_QLEN))) qlen = 32u;
    // (a ring of 16 would buy chunks read in place a 14th parser: measured, -8 % on the mix, +-0 on text -- the short ring
    // throttles the parsers; CW_LZ4_QLEN=16 runs it, profiles/r02_ab_compat_read_in_place.txt)
    if (const char *e = getenv("CW_LZ4_QLEN")) {  // tuning knob
        const uint32_t v = (uint32_t)atoi(e);
        qlen = v == 16u ? 16u : v == 32u ? 32u : CW_LZ4_QLEN;
    }
    int parsers = g_lz4_max_smem / (int)(slot + cw_lz4_queue_bytes(qlen));
    if (parsers < 1) return cudaErrorInvalidValue;
    if (parsers > CW_LZ4_MAX_PARSERS) parsers = CW_LZ4_MAX_PARSERS;
    if (const char *e = getenv("CW_LZ4_PARSERS")) {  // tuning knob for the profiling scripts
        const int v = atoi(e);
        if (v > 0 && v < parsers) parsers = v;
    }
    int emitters = lz4_emitters();
    if (emitters > parsers) emitters = parsers;
    uint64_t ctas = (n + (uint64_t)parsers - 1) / (uint64_t)parsers;
    if (ctas > (uint64_t)dev.sm_count) ctas = (uint64_t)dev.sm_count;
    const dim3 grid((unsigned)ctas), block((parsers + emitters) * 32);
    const size_t dyn = (size_t)parsers * (slot + cw_lz4_queue_bytes(qlen));
    // emitters share scheduler 3 (warp w issues from scheduler w % 4), parsers fill the others: parser-only schedulers
    // keep their instruction caches to the parser loop (CW_LZ4_ROLE_MAP=0: roles in warp order)
    static const bool grouped = [] {
        const char *e = getenv("CW_LZ4_ROLE_MAP");
        return !(e && atoi(e) == 0);
    }();
    static const int none[4] = {3, 3, 3, 3}, es[4] = {3, 3, 3, 3};
    const CwRoleMap roles = make_role_map(parsers, emitters, 0, grouped ? none : nullptr, grouped ? es : nullptr);
#define CW_GO(HL, AL, DI) \
    cw_lz4_kernel<HL, AL, DI><<<grid, block, dyn, st>>>(d_src, chunk, n, d_dst, stride, d_lens, work, slot, (uint32_t)parsers, (uint32_t)emitters, roles, qlen)
    if (directc) CW_GO(13, true, true);
    else if (aligned) CW_GO(13, true, false);
    else CW_GO(13, false, false);
#undef CW_GO
    return cudaGetLastError();
}

// LZF fast mode: the same team kernel with the LZF flavour of the parse (lz4_fast.cuh); 13 table bits up to 4 KiB chunks,
// 14 above (cw_lzff_hashlog: chosen like the LZ4 ones, tools/lzf_fast_model.c)
cudaError_t cw_launch_lzf_fast(const CwDeviceInfo &dev, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                               uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    const bool aligned = ((chunk & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_src) & 15u) == 0);
    const bool direct = aligned && chunk > CW_LZ4_DIRECT_ABOVE;
    const int hl = cw_lzff_hashlog(chunk);
    int teams = 0, warps = 0;
    uint32_t fslot = cw_lz4f_slot_bytes(chunk, hl, direct, CW_LZ4F_MAX_TEAM);
    for (int it = 0; it < 2; it++) {
        const int slots = g_lz4f_max_smem / (int)fslot;
        if (slots < 1) return cudaErrorInvalidValue;
        cw_lz4f_shape(slots, CW_LZ4F_MAX_WARPS, &teams, &warps);
        fslot = cw_lz4f_slot_bytes(chunk, hl, direct, (uint32_t)((warps + teams - 1) / teams));
    }
    uint64_t fctas = (n + (uint64_t)teams - 1) / (uint64_t)teams;
    if (fctas > (uint64_t)dev.sm_count) fctas = (uint64_t)dev.sm_count;
#define CW_GOF(HL, AL, DI)                                                                                                   \
    cw_lz4f_kernel<HL, AL, DI, CW_F_LZF><<<(unsigned)fctas, warps * 32, (size_t)teams * fslot, st>>>(d_src, chunk, n, d_dst, stride, \
                                                                                                    d_lens, work, fslot, (uint32_t)teams)
    if (hl == 13) {
        if (aligned) CW_GOF(13, true, false);
        else CW_GOF(13, false, false);
    } else if (direct) CW_GOF(14, true, true);
    else if (aligned) CW_GOF(14, true, false);
    else CW_GOF(14, false, false);
#undef CW_GOF
    return cudaGetLastError();
}
