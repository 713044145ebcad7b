// fused_kernels.cu -- ProcessBlock (HashAndCompress.cpp:231-261): compressed block + digest of every chunk
// in ONE pass over the input.
//
// One persistent CTA per SM, warp-specialised:
//   * parser warps (as many as the SM's shared memory holds: 9 for LZ4 / 5 for LZF at 8 KB chunks) each
//     own a shared-memory slot, pull chunk indices from a global counter, stage the chunk with a bulk
//     async copy and run the parser of lz4_warp.cuh / lzf_warp.cuh; 2-3 emitter warps write the streams;
//   * hash warps (1 per CTA by default, no shared memory) take groups of 32 consecutive chunks, one chunk per lane,
//     and run the register-resident Skein / SHA-256 chain of skein.cuh / sha256.cuh straight from global
//     memory.
// The two roles are coupled through the counters: a hash warp starts group g only while it is within 32 MiB of
// the compress counter, so whichever role reads a chunk second finds it in L2 (126 MB), not in HBM;
// the compressors are latency bound and leave most issue slots of the SM idle, which is exactly what the
// integer-bound hash warps consume.
// the 3 x 24-round Skein body (11 KB of SASS instead of 36 KB) keeps hash + compressor code inside the instruction cache
#define CW_SKEIN_UNROLL72 false
#include "comp_roles.cuh"
#include "hash_kernels.cuh"
#include "launch.h"
#include "role_map.h"

#include <cstdlib>

// Hash warps per CTA.  Every hash lane walks its own 8 KB chunk for the whole time it takes to hash it, so the
// bytes that must stay in L2 between the compress warps' bulk copy and the last hash read are
// hash_warps * 32 lanes * chunk * 148 SMs; 1 warp keeps that window at 39 MB for 8 KB chunks (L2: 126 MB)
// while still hashing faster than the compressors consume (measured, profiles/).
#ifndef CW_FUSED_COMPACT_ALWAYS
#define CW_FUSED_COMPACT_ALWAYS false  // A/B knob: the single-call-site Skein loop in the compat kernels too (measured: 72.4 -> 73.0 GB/s
                                       // at 2 GiB, hash-bound data -1 %: left off)
#endif
#define CW_FUSED_HASH_WARPS_DEFAULT 1
#ifndef CW_FUSED_TAIL_HELP
#define CW_FUSED_TAIL_HELP 1  // warps without compress work left join the hash role
#endif
#ifndef CW_FUSED_HASH_WARPS_MAX
#define CW_FUSED_HASH_WARPS_MAX 2  // also sizes __launch_bounds__: 19 warps leave ptxas 96 registers and no spills
#endif
#define CW_FUSED_MAX_COMP_WARPS 14
#define CW_FUSED_MAX_THREADS ((CW_FUSED_HASH_WARPS_MAX + CW_FUSED_MAX_COMP_WARPS + 3) * 32)  // + emitters
// fast mode: teams of warps per chunk (lz4_fast.cuh), no emitters: 26 compress warps + 2 hash warps = 896 threads, which
// leaves ptxas 72 registers -- the hash role then runs the single-call-site Skein loop without the next-block prefetch
#ifndef CW_FUSED_FAST_COMP_WARPS
#define CW_FUSED_FAST_COMP_WARPS 26  // 30 (1024 threads, 64 registers) measured 10 % slower at 8 KiB: the hash rounds want the registers
#endif
#define CW_FUSED_FAST_THREADS ((CW_FUSED_HASH_WARPS_MAX + CW_FUSED_FAST_COMP_WARPS) * 32)
#ifndef CW_FUSED_FAST_COMP_WARPS_FEW_SLOTS
#define CW_FUSED_FAST_COMP_WARPS_FEW_SLOTS 26  // variants with 5-6 slots per SM (32 KB tables); 30 measured +5-9 % on text, -6-8 % on the mix
                                              // (its random chunks are hash bound, and the hash rounds want the registers): profiles/r02_ab_fused_few_slots_30_vs_26_warps.txt
#endif

// measured on the synthetic mix (profiles/r01_tune_emitters_hashwarps.txt): the 9-parser compat kernel is best
// with 1 hash warp + 3 emitters, the 13-parser fast kernel with 2 + 2
static int fused_hash_warps(bool fast)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CW_FUSED_HASH_WARPS");  // tuning knob for the profiling scripts
        v = e ? atoi(e) : 0;
        if (v > CW_FUSED_HASH_WARPS_MAX) v = CW_FUSED_HASH_WARPS_MAX;
    }
    return v > 0 ? v : (fast ? 2 : CW_FUSED_HASH_WARPS_DEFAULT);
}

// compressor variants of the fused kernel: LZ4 compat (the reference's parse, 13-bit table, chunk staged) / LZF /
// LZ4 fast (window-parallel parse of lz4_fast.cuh) with a 12-bit and a 13-bit table / the two LZ4 modes for chunks
// > 16 KiB, which are read where they lie (DIRECT; fast mode then uses a 14-bit table)
// ... and the LZF flavour of the fast parse with 13 / 14 table bits, staged or read in place
enum {
    CW_C_LZ4 = 0, CW_C_LZF = 1, CW_C_LZ4_FAST = 2, CW_C_LZ4_FAST13 = 3, CW_C_LZ4_DIRECT = 4, CW_C_LZ4_FAST_DIRECT = 5,
    CW_C_LZF_FAST13 = 6, CW_C_LZF_FAST14 = 7, CW_C_LZF_FAST_DIRECT = 8
};
__host__ __device__ constexpr bool cw_c_direct(int c) { return c == CW_C_LZ4_DIRECT || c == CW_C_LZ4_FAST_DIRECT || c == CW_C_LZF_FAST_DIRECT; }
__host__ __device__ constexpr bool cw_c_fast(int c) { return c == CW_C_LZ4_FAST || c == CW_C_LZ4_FAST13 || c == CW_C_LZ4_FAST_DIRECT || c >= CW_C_LZF_FAST13; }
__host__ __device__ constexpr int cw_c_codec(int c) { return c >= CW_C_LZF_FAST13 ? CW_F_LZF : CW_F_LZ4; }
__host__ __device__ constexpr int cw_c_hashlog(int c)
{
    return c == CW_C_LZ4_FAST ? 12 : (c == CW_C_LZ4_FAST_DIRECT || c == CW_C_LZF_FAST14 || c == CW_C_LZF_FAST_DIRECT) ? 14 : 13;
}

// compress warps of a fast-mode fused CTA: the kernels whose teams wait on each other take the 1024-thread shape
__host__ __device__ constexpr int cw_c_fast_warps(int c)
{
    return (c == CW_C_LZ4_FAST_DIRECT || c == CW_C_LZF_FAST14 || c == CW_C_LZF_FAST_DIRECT) ? CW_FUSED_FAST_COMP_WARPS_FEW_SLOTS
                                                                                            : CW_FUSED_FAST_COMP_WARPS;
}

#define CW_FUSED_EMITTERS_DEFAULT 3  // emitter warps per CTA
static int fused_emitters(bool fast)
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CW_LZ4_EMITTERS");  // tuning knob for the profiling scripts
        v = e ? atoi(e) : 0;
        if (v > 3) v = 3;
    }
    return v > 0 ? v : 2;
}

// records per parser -> emitter ring: LZF is short of shared memory (32: 7 parsers per SM at 8 KiB instead of 6)
template <int CALG>
__host__ __device__ constexpr uint32_t cw_fused_qlen()
{
    return CALG == CW_C_LZF ? 32u : CW_LZ4_QLEN;
}

template <int CALG>
__host__ __device__ constexpr uint32_t cw_slot_bytes(uint32_t chunk, uint32_t nteam = 1)
{
    if (cw_c_fast(CALG)) return cw_lz4f_slot_bytes(chunk, cw_c_hashlog(CALG), cw_c_direct(CALG), nteam);  // no queue
    return (CALG == CW_C_LZF ? ((chunk + 16u + 15u) & ~15u) + cw_lzf_table_bytes(cw_lzf_table_format(chunk))
                             : (cw_c_direct(CALG) ? 0u : cw_lz4_chunk_pad(chunk)) + (2u << cw_c_hashlog(CALG))) +
           cw_lz4_queue_bytes(cw_fused_qlen<CALG>());
}

// "3;3,3" -> hash warp on scheduler 3, two emitters on scheduler 3 (tuning knob CW_ROLE_MAP; "0" = warp order)
static bool parse_role_env(int *hs, int *es)
{
    const char *e = getenv("CW_ROLE_MAP");
    if (!e) return false;
    int *dst = hs, k = 0;
    for (int i = 0; i < 4; i++) hs[i] = es[i] = 3;
    for (; *e; e++) {
        if (*e == ';') {
            dst = es;
            k = 0;
        } else if (*e >= '0' && *e <= '3' && k < 4)
            dst[k++] = *e - '0';
    }
    return true;
}

// logical warps [0, parsers): parsers; [parsers, parsers + emitters): emitters; the rest: hash
template <int HALG, int CALG, bool ALIGNED>
__global__ void __launch_bounds__(cw_c_fast(CALG) ? (CW_FUSED_HASH_WARPS_MAX + cw_c_fast_warps(CALG)) * 32 : CW_FUSED_MAX_THREADS, 1)
    cw_fused_kernel(const __grid_constant__ CwSkeinIV iv, const uint8_t *__restrict__ src, uint32_t chunk, uint64_t n_chunks,
                    uint8_t *__restrict__ digests, uint8_t *__restrict__ dst, uint64_t stride, uint32_t *__restrict__ lens,
                    unsigned long long *work, uint32_t slot_bytes, uint32_t parsers, uint32_t emitters,
                    const __grid_constant__ CwRoleMap roles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[32];
    const unsigned warp = roles.logical[threadIdx.x >> 5], lane = threadIdx.x & 31u;
    if constexpr (cw_c_fast(CALG)) {
        // fast mode: `emitters` carries the number of teams; a team of warps stages, parses and emits a chunk on its own
        // (lz4_fast.cuh), no queues, no emitter warps
        if (warp < parsers) {
            const CwLz4fPlace pl = cw_lz4f_place(warp, emitters, parsers);
            cw_lz4f_team_role<cw_c_hashlog(CALG), ALIGNED, cw_c_direct(CALG), cw_c_codec(CALG)>(
                smem + (size_t)pl.team * slot_bytes, &bars[pl.team], pl.rank, pl.nteam, 1u + pl.team, src, chunk, n_chunks, &work[0], dst,
                stride, lens);
            if (!CW_FUSED_TAIL_HELP) return;
        }
    } else {
        // slot_bytes excludes the queue: queues sit behind the slots
        CwLz4Queue *queues = reinterpret_cast<CwLz4Queue *>(smem + (size_t)parsers * slot_bytes);
        constexpr uint32_t QL = cw_fused_qlen<CALG>();
        CwLz4Queue *myq = cw_lz4_queue_at(queues, warp < parsers ? warp : 0u, QL);
        if (warp < parsers && lane < 8u) reinterpret_cast<uint32_t *>(myq)[lane] = 0u;
        __syncthreads();
        if (warp < parsers) {
            if (CALG == CW_C_LZF)
                cw_lzf_parser_role<ALIGNED>(smem + (size_t)warp * slot_bytes, myq, &bars[warp], src, chunk, n_chunks, &work[0], QL);
            else
                cw_lz4_parser_role<cw_c_hashlog(CALG), ALIGNED, cw_c_direct(CALG)>(
                    smem + (size_t)warp * slot_bytes, myq, &bars[warp], src, chunk, n_chunks, &work[0], QL);
            if (!CW_FUSED_TAIL_HELP) return;
        } else if (warp < parsers + emitters) {
            if (CALG == CW_C_LZF)
                cw_lzf_emitter_role(warp - parsers, emitters, parsers, smem, slot_bytes, queues, chunk, dst, stride, lens, QL);
            else
                cw_lz4_emitter_role<cw_c_direct(CALG)>(warp - parsers, emitters, parsers, smem, slot_bytes, queues, src, chunk, dst,
                                                       stride, lens, QL);
            if (!CW_FUSED_TAIL_HELP) return;
        }
    }
    // ------------------------------ hash role ------------------------------
    // Parser and emitter warps that have run out of chunks end up here too: when the last chunk has been claimed the
    // hash warp is still up to a group or more behind (one group = 32 chunks = 0.3 ms of hashing for one warp), and on
    // a slice-sized batch that tail is a fifth of the kernel.
    constexpr int D = cw_digest_len<HALG>();
    const unsigned long long lead = (32ull << 20) / chunk;
    for (;;) {
        unsigned long long grp = 0;
        if (lane == 0) grp = atomicAdd(&work[1], 1ull);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        const unsigned long long first = grp * 32ull;
        if (first >= n_chunks) break;
        const unsigned long long need = first + 32ull < n_chunks ? first + 32ull : n_chunks;
        if (lane == 0) {
            // stay close to the compress warps so that whichever role reads a chunk second finds it in L2: a group may
            // start once its last chunk is at most `lead` chunks (32 MiB) ahead of the compress counter.  Waiting for the
            // chunks to be claimed first (lead = 0) leaves the last groups of a batch -- 0.3-0.6 ms of hashing each --
            // to run after the parsers are done.
            while (*reinterpret_cast<volatile unsigned long long *>(&work[0]) + lead < need) __nanosleep(256);
        }
        __syncwarp();
        const unsigned long long idx = first + lane;
        if (idx < n_chunks) cw_hash_one<HALG, ALIGNED, ALIGNED, !cw_c_fast(CALG) && !CW_FUSED_COMPACT_ALWAYS, cw_c_fast(CALG) || CW_FUSED_COMPACT_ALWAYS>(iv, src + idx * chunk, chunk, digests + idx * D);
    }
}

static CwRoleMap fused_role_map(int calg, int parsers, int emitters, int hw)
{
    static int ehs[4], ees[4];
    static const bool have_env = parse_role_env(ehs, ees);
    if (have_env) {
        const char *e = getenv("CW_ROLE_MAP");
        if (e[0] == '0' && e[1] == 0) return make_role_map(parsers, emitters, hw, nullptr, nullptr);
        return make_role_map(parsers, emitters, hw, ehs, ees);
    }
    // measured (profiles/r01_tune_emitters_hashwarps.txt): with 9 parsers the hash warp gets scheduler 3 to itself, the
    // emitters for company, and every other scheduler runs 3 parsers; the 13-parser fast kernel keeps the warp order
    static const int hs3[4] = {3, 3, 3, 3}, es3[4] = {3, 3, 3, 3};
    if (calg == CW_C_LZ4 && hw == 1 && (parsers + emitters + hw) % 4 == 0) return make_role_map(parsers, emitters, hw, hs3, es3);
    return make_role_map(parsers, emitters, hw, nullptr, nullptr);
}

template <int HALG, int CALG, bool ALIGNED>
static cudaError_t launch_one(const CwDeviceInfo &dev, int max_dyn_smem, const CwSkeinIV &iv, const uint8_t *d_src,
                              uint32_t chunk, uint64_t n, uint8_t *d_digests, uint8_t *d_dst, uint64_t stride,
                              uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    auto kern = cw_fused_kernel<HALG, CALG, ALIGNED>;
    const int hw = fused_hash_warps(cw_c_fast(CALG));
    if (cw_c_fast(CALG)) {
        // one slot per team; the kernel's `emitters` argument carries the number of teams (cw_lz4f_place spreads the warps)
        int teams = 0, warps = 0;
        uint32_t fslot = cw_slot_bytes<CALG>(chunk, CW_LZ4F_MAX_TEAM);
        for (int it = 0; it < 2; it++) {
            const int slots = max_dyn_smem / (int)fslot;
            if (slots < 1) return cudaErrorInvalidValue;
            cw_lz4f_shape(slots, cw_c_fast_warps(CALG), &teams, &warps);
            fslot = cw_slot_bytes<CALG>(chunk, (uint32_t)((warps + teams - 1) / teams));
        }
        uint64_t fctas = (n + (uint64_t)teams - 1) / (uint64_t)teams;
        if (fctas > (uint64_t)dev.sm_count) fctas = (uint64_t)dev.sm_count;
        kern<<<(unsigned)fctas, (warps + hw) * 32, (size_t)teams * fslot, st>>>(
            iv, d_src, chunk, n, d_digests, d_dst, stride, d_lens, work, fslot, (uint32_t)warps, (uint32_t)teams,
            make_role_map(warps, 0, hw, nullptr, nullptr));
        return cudaGetLastError();
    }
    const uint32_t per = cw_slot_bytes<CALG>(chunk);  // slot + queue
    const uint32_t slot = per - cw_lz4_queue_bytes(cw_fused_qlen<CALG>());
    int parsers = max_dyn_smem / (int)per;
    if (parsers < 1) return cudaErrorInvalidValue;
    if (parsers > CW_FUSED_MAX_COMP_WARPS) parsers = CW_FUSED_MAX_COMP_WARPS;
    int emitters = fused_emitters(false);
    if (emitters > parsers) emitters = parsers;
    uint64_t ctas = (n + (uint64_t)parsers - 1) / (uint64_t)parsers;
    if (ctas > (uint64_t)dev.sm_count) ctas = (uint64_t)dev.sm_count;
    kern<<<(unsigned)ctas, (parsers + emitters + hw) * 32, (size_t)parsers * per, st>>>(
        iv, d_src, chunk, n, d_digests, d_dst, stride, d_lens, work, slot, (uint32_t)parsers, (uint32_t)emitters,
        fused_role_map(CALG, parsers, emitters, hw));
    return cudaGetLastError();
}

static int g_fused_max_smem = 0;

template <int HALG, int CALG, bool ALIGNED>
static cudaError_t configure_one(int max_dyn)
{
    return cudaFuncSetAttribute(cw_fused_kernel<HALG, CALG, ALIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
}

// opt-in shared memory is a per-device function attribute: set it for every instantiation on the current device
cudaError_t cw_fused_configure(const CwDeviceInfo &dev)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, cw_fused_kernel<CW_K_SKEIN512, CW_C_LZ4, true>);
    if (e != cudaSuccess) return e;
    const int m = (int)dev.smem_optin - (int)fa.sharedSizeBytes;
    g_fused_max_smem = m;
#define CW_CFG(H, C)                                                  \
    if (e == cudaSuccess) e = configure_one<H, C, true>(m);           \
    if (e == cudaSuccess) e = configure_one<H, C, false>(m);
#define CW_CFG_D(H, C) \
    if (e == cudaSuccess) e = configure_one<H, C, true>(m);
#define CW_CFG_H(H)                                                                                                    \
    CW_CFG(H, CW_C_LZ4) CW_CFG(H, CW_C_LZF) CW_CFG(H, CW_C_LZ4_FAST) CW_CFG(H, CW_C_LZ4_FAST13) CW_CFG_D(H, CW_C_LZ4_DIRECT) \
        CW_CFG_D(H, CW_C_LZ4_FAST_DIRECT) CW_CFG(H, CW_C_LZF_FAST13) CW_CFG(H, CW_C_LZF_FAST14) CW_CFG_D(H, CW_C_LZF_FAST_DIRECT)
    CW_CFG_H(CW_K_SKEIN512)
    CW_CFG_H(CW_K_SKEIN256_128)
    CW_CFG_H(CW_K_SHA256)
#undef CW_CFG_H
#undef CW_CFG_D
#undef CW_CFG
    return e;
}

template <int HALG>
static cudaError_t launch_h(int calg, bool aligned, const CwDeviceInfo &dev, const CwSkeinIV &iv, const uint8_t *d_src,
                            uint32_t chunk, uint64_t n, uint8_t *d_digests, uint8_t *d_dst, uint64_t stride, uint32_t *d_lens,
                            unsigned long long *work, cudaStream_t st)
{
    const int m = g_fused_max_smem;
#define CW_ARGS dev, m, iv, d_src, chunk, n, d_digests, d_dst, stride, d_lens, work, st
#define CW_BOTH(C) return aligned ? launch_one<HALG, C, true>(CW_ARGS) : launch_one<HALG, C, false>(CW_ARGS)
    switch (calg) {
    case CW_C_LZ4_FAST: CW_BOTH(CW_C_LZ4_FAST);
    case CW_C_LZ4_FAST13: CW_BOTH(CW_C_LZ4_FAST13);
    case CW_C_LZ4: CW_BOTH(CW_C_LZ4);
    case CW_C_LZF: CW_BOTH(CW_C_LZF);
    case CW_C_LZ4_DIRECT: return launch_one<HALG, CW_C_LZ4_DIRECT, true>(CW_ARGS);
    case CW_C_LZ4_FAST_DIRECT: return launch_one<HALG, CW_C_LZ4_FAST_DIRECT, true>(CW_ARGS);
    case CW_C_LZF_FAST13: CW_BOTH(CW_C_LZF_FAST13);
    case CW_C_LZF_FAST14: CW_BOTH(CW_C_LZF_FAST14);
    case CW_C_LZF_FAST_DIRECT: return launch_one<HALG, CW_C_LZF_FAST_DIRECT, true>(CW_ARGS);
    }
#undef CW_BOTH
#undef CW_ARGS
    return cudaErrorInvalidValue;
}

cudaError_t cw_launch_fused(const CwDeviceInfo &dev, int halg, int calg, bool fast, const CwSkeinIV &iv, const uint8_t *d_src,
                            uint32_t chunk, uint64_t n, uint8_t *d_digests, uint8_t *d_dst, uint64_t stride,
                            uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const bool aligned = (chunk % 16u == 0) && ((reinterpret_cast<uintptr_t>(d_src) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(d_digests) & 15u) == 0);
    // compat reads chunks above 8 KiB in place, like the stand-alone kernel (cw_launch_lz4: 16 KiB chunks +21 % on the mix,
    // +40 % on text); CW_LZ4_COMPAT_DIRECT_ABOVE: the threshold, for A/B runs.  Unaligned buffers / ragged sizes of up to
    // 16 KiB keep the staged kernel
    static const uint32_t compat_direct_above = [] {
        const char *e = getenv("CW_LZ4_COMPAT_DIRECT_ABOVE");
        return e ? (uint32_t)atoi(e) : 8192u;
    }();
    if (calg == CW_C_LZ4 && !fast && aligned && chunk > compat_direct_above && chunk <= 16384u)
        calg = CW_C_LZ4_DIRECT;
    else if (calg == CW_C_LZ4 && chunk > 16384u) {
        // big chunks (the 64 KiB blocks of BASELINE configs[2]) are not staged: parsers, emitters and hash warps all read
        // the chunk where it lies, which keeps 13 (fast: 6) chunks per SM in flight instead of 2, in ONE launch
        if (aligned) calg = fast ? CW_C_LZ4_FAST_DIRECT : CW_C_LZ4_DIRECT;
        else {
            // ragged sizes / unaligned buffers cannot use the in-place parser: stand-alone kernels back to back
            cudaError_t e = cw_launch_lz4(dev, fast, d_src, chunk, n, d_dst, stride, d_lens, &work[0], st);
            if (e != cudaSuccess) return e;
            return cw_launch_hash(halg, iv, d_src, chunk, n, d_digests, st);
        }
    } else if (calg == CW_C_LZ4 && fast)
        calg = cw_lz4f_hashlog(chunk) == 12 ? CW_C_LZ4_FAST : CW_C_LZ4_FAST13;
    else if (calg == CW_C_LZF && fast) {
        if (chunk > 16384u) {
            if (aligned) calg = CW_C_LZF_FAST_DIRECT;
            else {
                cudaError_t e = cw_launch_lzf_fast(dev, d_src, chunk, n, d_dst, stride, d_lens, &work[0], st);
                if (e != cudaSuccess) return e;
                return cw_launch_hash(halg, iv, d_src, chunk, n, d_digests, st);
            }
        } else
            calg = cw_lzff_hashlog(chunk) == 13 ? CW_C_LZF_FAST13 : CW_C_LZF_FAST14;
    }
    switch (halg) {
    case CW_K_SKEIN512: return launch_h<CW_K_SKEIN512>(calg, aligned, dev, iv, d_src, chunk, n, d_digests, d_dst, stride, d_lens, work, st);
    case CW_K_SKEIN256_128: return launch_h<CW_K_SKEIN256_128>(calg, aligned, dev, iv, d_src, chunk, n, d_digests, d_dst, stride, d_lens, work, st);
    case CW_K_SHA256: return launch_h<CW_K_SHA256>(calg, aligned, dev, iv, d_src, chunk, n, d_digests, d_dst, stride, d_lens, work, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t cw_configure_kernels(const CwDeviceInfo &dev)
{
    cudaError_t e = cw_lz4_configure(dev);
    if (e == cudaSuccess) e = cw_lzf_configure(dev);
    if (e == cudaSuccess) e = cw_fused_configure(dev);
    if (e == cudaSuccess) e = cw_decode_configure(dev);
    return e;
}
