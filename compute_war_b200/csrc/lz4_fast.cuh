// lz4_fast.cuh -- LZ4 "fast mode": window-parallel parse AND emission of ONE chunk by a small TEAM of warps.
//
// The reference side of fast mode is only the decoder: every block must decode bit-exactly with
// LZ4_decompress_safe (src/compression_perf/include/lz4/lz4.h:141-149; call site experiment.cpp:256) and the
// compressed size must stay within 1 % of what LZ4_compress_default (HashAndCompress.cpp:349-355) produces.  No
// parse order is imposed, so nothing of the sequential parser is replayed here (that is compat mode,
// lz4_warp.cuh).  The parse is defined position-wise, so that a warp can evaluate it 128 positions at a time, and
// it is a deterministic function of the chunk (the test suite holds a scalar restatement of the rules below; the
// tests compare the GPU output with it byte for byte):
//
//   * the chunk is cut into groups of four windows of 32 consecutive positions, one lane per position.  ALL
//     positions of a window are hashed, looked up and inserted together (the reference inserts only the positions
//     its skipping search probes), whether or not they end up inside a match: candidates therefore do not depend on
//     the parse, which is what lets the groups of a chunk be worked on by several warps at once (below);
//   * candidate of a position: the table entry read before the window's inserts; if a LOWER position of the same
//     window has the same hash, that position instead (store, sync, read back; the lowest position must own the
//     slot, so lanes that lost a same-address store race to a higher position store again); if the 4 bytes two
//     positions back are equal, p - 2 instead (runs and period-2 patterns);
//   * a match is the candidate's full forward extension, and the greedy choice is "first verified position at or
//     after the end of the previous match".
//
// What the warp does NOT do is measure every verified position's match: neighbouring positions of one match have
// neighbouring candidates, so the masks
//     OK   (position verified),  SAME (position x and x + 1 verified with candidates c and c + 1)
// and a per-position count CNT (how many of the bytes x + 4 .. x + 7 equal c + 4 .. c + 7) already hold the exact end of
// almost every match: it is the first position r at or after the start whose SAME bit is clear, plus 4 + CNT[r].
// Every position works out, from those masks alone, where its own match ends (E) and which verified position the greedy
// parse would choose after it (N), and leaves both in a 16-bit entry of a per-warp table in shared memory.  The greedy
// walk over a group is then ONE shared-memory load per chosen match.  Only when CNT[r] is 4 (the bytes go on although
// the neighbour chose another candidate: an "open" match), when a chain of neighbours is longer than 32 positions, or
// when no verified position lies within 32 positions of a match end does the walk leave that fast path: the warp
// extends the match itself (32 bytes, then 128 per step) and looks the next position up in the OK masks.
//
// Emission happens in the same warp as the parse, once per group: the chosen matches are handed to one lane each (backward
// extension, sizes, prefix sum, token / lengths / offset), and the literal bytes are then written position-wise --
// a lane looks up the sequence its position belongs to and stores its own byte -- so a group costs the same number
// of instructions whether it holds two sequences or twenty.  Literals that precede the group (a sequence is only
// written once its match is known) are copied by the whole warp.
//
// A chunk is a dependent chain of several 10^4 instructions, a warp issues one of them every 6-7 cycles, and shared
// memory holds only 13 chunks of 8 KiB per SM, so one warp per chunk leaves most issue slots idle.  A team of NW warps
// therefore shares a chunk: warp r takes groups r, r + NW, ... and only three short sections of a group are ordered
// by tokens in shared memory (CwLz4fTeam): the table look-ups / inserts (group g needs the inserts of g - 1), the
// greedy walk (needs the position the walk of g - 1 ended at) and the output offset (one addition).  Everything
// else of group g + 1 -- fetch, hash, verification, masks, E / N entries, sizes, stores -- overlaps with group g in the
// other warp.  Nothing is copied between the warps: a group's values stay in the registers and the scratch of the warp
// that owns it.
//
// End rules of the block format: a match starts at or before n - 12 and ends at or before
// n - 5; chunks shorter than 13 bytes are one literal run.
//
// Table: zeroed u16[1 << HASHLOG] in shared memory.  An empty slot reads as candidate 0, which is a legal
// source position and is verified like any other.
#pragma once
#include "lz4_warp.cuh"

#define CW_LZ4F_SW 4  // windows per group: 128 positions, which is what the 8-bit fields of the E / N entries hold

// hash table size for a chunk size: chosen on the reference's corpora with tools/lz4_fast_model.c so that the worst
// file stays below 1.01 x the reference's size (4 KiB 1.006, 8 KiB 1.002 with 12 bits; 16 KiB 0.998 with 13;
// 64 KiB 0.990 with 14 -- 13 bits would give 1.012 on world192.txt)
CW_HOSTDEV int cw_lz4f_hashlog(uint32_t chunk) { return chunk <= 8192u ? 12 : chunk <= 16384u ? 13 : 14; }
// LZF flavour (3-byte keys collide more): 4 KiB 1.003 / worst file 1.007 with 13 bits; 8 KiB 1.002 / 1.006, 16 KiB 0.999 / 1.007,
// 64 KiB 0.996 / 1.008 with 14 (tools/lzf_fast_model.c against lzf_compress)
CW_HOSTDEV int cw_lzff_hashlog(uint32_t chunk) { return chunk <= 4096u ? 13 : 14; }

// per-warp scratch in shared memory (344 bytes)
struct CwLz4fScratch {
    union {
        uint16_t pos[128];    // per position of the group.  First the continuation count CNT; then N | (E - x) << 8 [| ONCHAIN]:
                              // the verified position the parse chooses after this one's match (relative to the group, or a
                              // CW_LZ4F_* code below) and the length of that match
        uint32_t map[32][2];  // afterwards: {output offset of literal x, minus x; first literal | end of literals << 16}
    };
    uint32_t okw[6];  // OK masks of the four windows, two zero words behind them
    uint8_t rec[32];   // the matches the speculative walk chose (their positions), in order
    uint8_t rec2[32];  // the matches chosen from the real entry until the two chains met
};
static_assert(sizeof(CwLz4fScratch) % 8 == 0, "the literal map is read with 64-bit loads");
#define CW_LZ4F_LEAVE 128u  // N >= 128: nothing more to choose in this group
#define CW_LZ4F_OPEN 254u   // N: the match goes on beyond E, the warp extends it
#define CW_LZ4F_FAR 255u    // N: no verified position within 32 positions of E, the walk looks it up in the OK masks

// aligned 32-bit word i of the chunk (DIRECT: words at or beyond nw read as 0)
template <bool DIRECT>
CW_DEV uint32_t cw_lz4f_word(const uint32_t *w, uint32_t i, uint32_t nw)
{
    if (!DIRECT) return w[i];
    return i < nw ? cw_ldg32(w + i) : 0u;
}
CW_DEV uint32_t cw_lz4f_ctz(uint32_t v) { return (uint32_t)cw_ffs(v) - 1u; }

// whole-warp forward extension of a match: positions e.. against e - d..; returns the end.  Most continuations are
// a few bytes long, so the first step looks at 32 single bytes; after that 32 lanes x 4 bytes per step.
template <bool DIRECT>
CW_DEV uint32_t cw_lz4f_extend(const uint8_t *s, uint32_t e, uint32_t d, uint32_t matchlimit, uint32_t nw, unsigned lane)
{
    {
        const uint32_t pa = e + lane;
        const bool ne = pa >= matchlimit || cw_lz4_src8<DIRECT>(s, pa) != cw_lz4_src8<DIRECT>(s, pa - d);
        const uint32_t stop = cw_ballot(ne);
        if (stop) return e + cw_lz4f_ctz(stop);
        e += 32u;
    }
    for (;;) {
        const uint32_t pa = e + 4u * lane;
        const uint32_t x = cw_lz4_src32<DIRECT>(s, pa, nw) ^ cw_lz4_src32<DIRECT>(s, pa - d, nw);
        uint32_t eq = x ? cw_lz4f_ctz(x) >> 3 : 4u;
        eq = cw_umin(eq, pa < matchlimit ? matchlimit - pa : 0u);
        const uint32_t stop = cw_ballot(eq < 4u);
        if (stop) {
            const unsigned L = cw_lz4f_ctz(stop);
            return e + 4u * L + cw_shfl(eq, L);
        }
        e += 128u;
    }
}

// shared state of the warps that work on one chunk.  A token says "the group before yours has passed this point": warp
// r waits on its own token, which only the owner of the previous group (warp r - 1) posts, once per group, so a waiting
// warp is never more than one arrival ahead.  A token is an mbarrier on the GPU (arrival count 1: a waiting warp is
// suspended by the hardware instead of polling shared memory -- a polling team mate takes the issue slots the working
// warp needs) and a plain counter on the warp emulator.  The kernels initialise them once; every warp keeps the number
// of arrivals it has consumed (`seen`, the same for its three tokens).
#define CW_LZ4F_MAX_TEAM 8
struct CwLz4fTeam {
    uint64_t tok_table[CW_LZ4F_MAX_TEAM];  // the table inserts of the group before are done
    uint64_t tok_walk[CW_LZ4F_MAX_TEAM];   // its greedy walk is done; ip = where the parse stands
    uint64_t tok_emit[CW_LZ4F_MAX_TEAM];   // its output space is allotted; op = output bytes allotted so far
    uint32_t ip, op;
    uint32_t chunk_lo, chunk_hi;  // kernels: global index of the chunk the team works on
};
static_assert(sizeof(CwLz4fTeam) % 16 == 0, "team state keeps the slots 16-byte aligned");

#if defined(__CUDACC__)
CW_DEV void cw_lz4f_tok_init(uint64_t *tok)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(tok)) : "memory");
}
// wait until k (>= 1) arrivals have happened (k-th phase complete); the caller has consumed k - 1; all lanes call
CW_DEV void cw_lz4f_tok_wait(uint64_t *tok, uint32_t k, unsigned lane)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(tok), parity = (k - 1u) & 1u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CW_TW_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra CW_TD_%=;\n"
        "bra CW_TW_%=;\n"
        "CW_TD_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity), "r"(1000000u)  // suspend-time hint (ns): an upper bound, the arrival wakes the warp; without it the
                                    // waiting warps of a team spend 5 % of the kernel's issue slots re-polling
        : "memory");
}
// everything this warp wrote before becomes visible to the warp that waits for this arrival; all lanes call
CW_DEV void cw_lz4f_tok_post(uint64_t *tok, unsigned lane)
{
    __syncwarp();
    if (lane == 0u)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(tok)) : "memory");
}
#else
CW_DEV void cw_lz4f_tok_init(uint64_t *tok) { *tok = 0; }
CW_DEV void cw_lz4f_tok_wait(uint64_t *tok, uint32_t k, unsigned lane)
{
    if (lane == 0u)
        while (cw_ld_vol(reinterpret_cast<const uint32_t *>(tok)) < k) cw_spin();
    cw_syncwarp();
    cw_fence_cta();
}
CW_DEV void cw_lz4f_tok_post(uint64_t *tok, unsigned lane)
{
    cw_syncwarp();
    if (lane == 0u) {
        cw_fence_cta();
        cw_st_vol(reinterpret_cast<uint32_t *>(tok), cw_ld_vol(reinterpret_cast<const uint32_t *>(tok)) + 1u);
    }
}
#endif

// whole-warp copy of len literal bytes s[src..] -> out[dst..] (long runs: literals in front of a group, the end of the chunk).
// Same stores as cw_lz4_copy_warp, kept as ONE out-of-line loop nest: the inlined, unrolled copies were a quarter of the
// kernel's code, and the fused kernel's instruction cache also has to hold the hash rounds.
#if defined(__CUDACC__)
#define CW_NOINLINE __noinline__
#else
#define CW_NOINLINE
#endif
template <bool DIRECT>
CW_NOINLINE CW_DEV_NI void cw_lz4f_copy(uint8_t *out, uint32_t dst, const uint8_t *s, uint32_t src, uint32_t len, uint32_t nw)
{
    const unsigned lane = cw_lane();
    uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(out + dst) & 3u)) & 3u;
    if (len < 64u) head = len;  // short runs: bytes only
    const uint32_t words = (len - head) >> 2;
    if (lane < head || len < 64u) {
#pragma unroll 1
        for (uint32_t x = lane; x < head; x += 32u) out[dst + x] = cw_lz4_src8<DIRECT>(s, src + x);
    }
    uint32_t *o32 = reinterpret_cast<uint32_t *>(out + dst + head);
#pragma unroll 1
    for (uint32_t w = lane; w < words; w += 32u) o32[w] = cw_lz4_src32<DIRECT>(s, src + head + 4u * w, nw);
    const uint32_t done = head + 4u * words;
    if (lane < len - done) out[dst + done + lane] = cw_lz4_src8<DIRECT>(s, src + done + lane);
}

#define CW_F_LZ4 0
#define CW_F_LZF 1
// LZF literal runs: the first `cnt` literals of a run of `run` bytes s[src..] whose area (a header byte in front of every
// 32 literals, lzf.h) starts at out[w].  Whole warp, out of line like cw_lz4f_copy.
template <bool DIRECT>
CW_NOINLINE CW_DEV_NI void cw_lzff_put_literals(uint8_t *out, uint32_t w, const uint8_t *s, uint32_t src, uint32_t run, uint32_t cnt)
{
#pragma unroll 1
    for (uint32_t k = cw_lane(); k < cnt; k += 32u) {
        const uint32_t o = w + k + (k >> 5);
        out[o + 1u] = cw_lz4_src8<DIRECT>(s, src + k);
        if ((k & 31u) == 0u) out[o] = (uint8_t)(cw_umin(32u, run - k) - 1u);
    }
}

// first verified position of the group at or after rel (< 128), from the OK masks in shared memory; 128: none
CW_DEV uint32_t cw_lz4f_scan(const uint32_t *okw, uint32_t rel)
{
    uint32_t j = rel >> 5;
    uint32_t m = okw[j] & (0xffffffffu << (rel & 31u));
    while (!m) {
        if (++j >= (uint32_t)CW_LZ4F_SW) return 128u;
        m = okw[j];
    }
    return 32u * j + cw_lz4f_ctz(m);
}

// The greedy walk over a group's entries, from verified position x (< 128): records the chosen positions in rec[cnt..] and
// follows N until the chain leaves the group (returns 128) or -- STOP -- reaches a position that carries CW_LZ4F_ONCHAIN
// (returns it, not recorded).  MARK: sets that flag on every position it chooses.  ip: end of the last chosen match.
// The fast path is the inner loop: one load per chosen match; every lane reads the same entry and stores the same record
// byte, so nothing in it is predicated on the lane.  An open match is extended by the whole warp; its end goes to the
// lane whose index is the match's place in rec[] (extE).
#define CW_LZ4F_ONCHAIN 0x8000u
template <bool DIRECT, bool MARK, bool STOP, int SWN>
CW_DEV uint32_t cw_lz4f_walk(CwLz4fScratch *scr, uint8_t *rec, uint32_t x, uint32_t &cnt, uint32_t &ip, uint32_t &extE,
                             const uint32_t (&c)[CW_LZ4F_SW], const uint8_t *s, uint32_t sbase, uint32_t matchlimit, uint32_t nw,
                             unsigned lane)
{
    for (;;) {
        uint32_t t, xs;
#if defined(__CUDACC__)
        if constexpr (!MARK && !STOP) {
            // the loop below as the compiler will not write it: 7 instructions per chosen match instead of 10
            uint32_t rps = (uint32_t)__cvta_generic_to_shared(rec) + cnt;
            const uint32_t pb = (uint32_t)__cvta_generic_to_shared(scr->pos), rp0 = rps;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                ".reg .u32 a;\n"
                ".reg .u16 h;\n"
                "CW_WK_%=:\n"
                "mad.lo.u32 a, %1, 2, %4;\n"
                "ld.shared.u16 h, [a];\n"
                "st.shared.u8 [%2], %1;\n"
                "add.u32 %2, %2, 1;\n"
                "mov.u32 %3, %1;\n"
                "cvt.u32.u16 %0, h;\n"
                "and.b32 %1, %0, 255;\n"
                "setp.lt.u32 p, %1, 128;\n"
                "@p bra CW_WK_%=;\n"
                "}\n"
                : "=r"(t), "+r"(x), "+r"(rps), "=r"(xs)
                : "r"(pb)
                : "memory");
            cnt += rps - rp0;
            ip = sbase + xs + ((t >> 8) & 63u);
        } else
#endif
        {
        volatile uint8_t *rp = rec + cnt;
        do {
            t = scr->pos[x];
            if (STOP && (t & CW_LZ4F_ONCHAIN)) {
                cnt = (uint32_t)(rp - rec);
                return x;
            }
            if (MARK && lane == 0u) scr->pos[x] = (uint16_t)(t | CW_LZ4F_ONCHAIN);
            *rp++ = (uint8_t)x;
            x = t & 255u;
        } while (x < 128u);
        cnt = (uint32_t)(rp - rec);
        xs = rp[-1];  // (read back instead of carried through the loop: fewer instructions per match)
        ip = sbase + xs + ((t >> 8) & 63u);
        }
        if (x < 254u) return 128u;  // nothing more to choose in this group
        if (x == CW_LZ4F_OPEN) {    // the warp extends the match
            uint32_t cnd = 0;
#pragma unroll
            for (int j = 0; j < SWN; j++) {
                const uint32_t q = cw_shfl(c[j], xs & 31u);
                if ((xs >> 5) == (uint32_t)j) cnd = q;
            }
            ip = cw_lz4f_extend<DIRECT>(s, ip, sbase + xs - cnd, matchlimit, nw, lane);
            if (lane == cnt - 1u) extE = ip;
        }
        const uint32_t r2 = ip - sbase;
        if (r2 >= 128u) return 128u;
        x = cw_lz4f_scan(scr->okw, r2);
        if (x >= 128u) return 128u;
    }
}

// Compress chunk s[0..n) (n <= 65536) into out, as warp `rank` of a team of `nteam` warps that all call this function
// with the same arguments.  s: 4-byte aligned; not DIRECT: CW_LZ4_PAD readable bytes behind the chunk.  table: zeroed
// u16[1 << HASHLOG]; team: tokens initialised, ip = 1, op = 0, visible to every warp of the team before the first one
// starts; seen: arrivals this warp has consumed on its tokens so far (kept by the caller from chunk to chunk); scr: this
// warp's own scratch.  Returns the block length in the warp that finishes the chunk (same value in every lane) and CW_LZ4_NONE in
// the others.
CW_HOSTDEV uint32_t cw_lz4f_groups(uint32_t n) { return n >= 13u ? (n - 12u) / (32u * CW_LZ4F_SW) + 1u : 0u; }
CW_HOSTDEV uint32_t cw_lzff_groups(uint32_t n) { return n >= 8u ? (n - 7u) / 96u + 1u : 0u; }
// CODEC: CW_F_LZ4 or CW_F_LZF (the LZF flavour of the same parse: 3-byte keys, lzf.h block format, see the end of this file).
// LZF: `cap` = the most the block may take (the reference's call passes chunk - 1); 0 is returned when it does not fit.
// SPEC: walk ahead of the turn (phase 3b); cw_lz4f_spec() says where the kernels do
CW_HOSTDEV constexpr bool cw_lz4f_spec(int codec, int hashlog, bool direct) { return direct || (codec == CW_F_LZF && hashlog >= 14); }
template <int CODEC, int HASHLOG, bool DIRECT, bool SPEC = cw_lz4f_spec(CODEC, HASHLOG, DIRECT)>
CW_DEV uint32_t cw_lz4f_team_chunk(const uint8_t *s, uint32_t n, uint16_t *table, CwLz4fTeam *team, CwLz4fScratch *scr,
                                   uint8_t *out, uint32_t rank, uint32_t nteam, uint32_t &seen, uint32_t cap = 0xffffffffu)
{
    constexpr bool LZF = CODEC == CW_F_LZF;
    constexpr uint32_t KB = LZF ? 3u : 4u;  // bytes of a key = the shortest match
    // windows per group: a group hands every chosen match to a lane of its own, and 32 matches of KB bytes fill 4 (3) windows.
    // (Positions and codes in the E / N entries stay as for 128 positions; an LZF group simply has no fourth window.)
    constexpr int SWN = LZF ? 3 : CW_LZ4F_SW;
    constexpr uint32_t G = 32u * SWN;
    const uint32_t next = rank + 1u == nteam ? 0u : rank + 1u;  // owner of the group after mine
    const uint32_t nw = n >> 2;
    const unsigned lane = cw_lane();
    const uint32_t lt = (1u << lane) - 1u;  // window positions below mine
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s);
    const unsigned sh = (lane & 3u) * 8u;           // my position's byte offset inside its aligned word, in bits
    const unsigned sh2 = ((lane + 2u) & 3u) * 8u;   // the same for the position two bytes back
    const bool hi2 = (lane & 3u) >= 2u;             // ... which starts in the same aligned word as mine
    const uint32_t NONE = 0xfff00000u;              // "no verified candidate": never one below a real candidate
    uint32_t anchor = 0, op = 0;

    if (n >= (LZF ? 8u : 13u)) {
        // LZ4: a match starts at or before n - 12 and never covers the last 5 bytes.  LZF has no such rules; a match may
        // start where the 7 bytes the parse looks at (key + continuation count) lie inside the chunk, and run to its end
        const uint32_t last = LZF ? n - 7u : n - 12u;
        const uint32_t matchlimit = LZF ? n : n - 5u;
        const uint32_t groups = LZF ? cw_lzff_groups(n) : cw_lz4f_groups(n);
        uint32_t ip = 1u;  // next position the greedy parse looks at (position 0 has no source)
        for (uint32_t g = rank; g < groups; g += nteam) {
            const uint32_t sbase = g * G;  // first position of the group
            uint32_t v[CW_LZ4F_SW], c[CW_LZ4F_SW], h[CW_LZ4F_SW], own4[CW_LZ4F_SW];
            uint32_t okm[CW_LZ4F_SW], nsame[CW_LZ4F_SW + 1];
            bool valid[CW_LZ4F_SW], per2[CW_LZ4F_SW];
            // ---- phase 1: my own bytes in every window of the group (independent loads).  Positions beyond the last
            //      match start are fetched like the others (the chunk is padded / words beyond it read as 0) and are
            //      simply never valid ----
            const uint32_t wi0 = (sbase >> 2) + (lane >> 2);
#pragma unroll
            for (int j = 0; j < SWN; j++) {
                const uint32_t p = sbase + 32u * j + lane, wi = wi0 + 8u * j;
                const uint32_t wm = cw_lz4f_word<DIRECT>(sw, (j == 0 && wi == 0u) ? 0u : wi - 1u, nw);
                const uint32_t w0 = cw_lz4f_word<DIRECT>(sw, wi, nw), w1 = cw_lz4f_word<DIRECT>(sw, wi + 1u, nw);
                const uint32_t w2 = cw_lz4f_word<DIRECT>(sw, wi + 2u, nw);
                v[j] = cw_funnel_r(w0, w1, sh);
                own4[j] = cw_funnel_r(w1, w2, sh);  // bytes p + 4 .. p + 7
                valid[j] = p - 1u < last;           // 1 <= p <= last
                // the 4 bytes two positions back (run / period-2 candidate)
                const uint32_t d2 = cw_funnel_r(hi2 ? w0 : wm, hi2 ? w1 : w0, sh2) ^ v[j];
                per2[j] = (LZF ? (d2 & 0xffffffu) == 0u : d2 == 0u) && (j != 0 || p >= 2u);
                h[j] = cw_lz4_hash<HASHLOG>(LZF ? v[j] & 0xffffffu : v[j]);
            }
            // ---- phase 2a (ordered: after the inserts of group g - 1): look-up and insert, window by window (a window
            //      sees the inserts of the windows before it) ----
            if (g) cw_lz4f_tok_wait(&team->tok_table[rank], seen + 1u, lane);
#pragma unroll
            for (int j = 0; j < SWN; j++) {
                const uint32_t p = sbase + 32u * j + lane;
                uint32_t cj = table[h[j]];
                cw_syncwarp();  // every lane has read the old entry before any lane stores
                if (valid[j]) table[h[j]] = (uint16_t)p;
                cw_syncwarp();
                // the LOWEST position of the window that maps to a slot must own it: lanes that lost to a higher position
                // store again (on B200 the lowest lane's store survives and this loop never iterates)
                uint32_t w = table[h[j]];
                while (cw_ballot(valid[j] && p < w)) {
                    if (valid[j] && p < w) table[h[j]] = (uint16_t)p;
                    cw_syncwarp();
                    w = table[h[j]];
                }
                if (valid[j] && w < p) cj = w;  // an earlier position of this window: more recent than the old entry
                if (per2[j]) cj = p - 2u;
                c[j] = cj;
            }
            if (g + 1u < groups) cw_lz4f_tok_post(&team->tok_table[next], lane);
            // ---- phase 2b: candidate verification, continuation counts and the OK / SAME masks ----
            uint32_t rotprev = NONE;  // lane 0: lane 31's verified candidate in the window before
            uint32_t psm[CW_LZ4F_SW + 1];
#pragma unroll
            for (int j = 0; j < SWN; j++) {
                const uint32_t cj = c[j], a = cj >> 2;
                const unsigned shc = (cj & 3u) * 8u;
                const uint32_t a0 = cw_lz4f_word<DIRECT>(sw, a, nw), a1 = cw_lz4f_word<DIRECT>(sw, a + 1u, nw);
                const uint32_t a2 = cw_lz4f_word<DIRECT>(sw, a + 2u, nw);
                // the 8 bytes at the candidate against my 8: the key decides, the rest is the continuation count
                // (how many of the bytes behind the key go on matching; 4: at least that many)
                const uint32_t x0 = cw_funnel_r(a0, a1, shc) ^ v[j], x = cw_funnel_r(a1, a2, shc) ^ own4[j];
                bool ok;
                uint32_t cntv = x ? cw_lz4f_ctz(x) >> 3 : 4u;
                if (LZF) {
                    const uint32_t m0 = x0 ? cw_lz4f_ctz(x0) >> 3 : 4u;
                    ok = valid[j] && m0 >= 3u;
                    if (n > 8192u) ok = ok && sbase + 32u * j + lane - cj <= 8192u;  // lzf.h: offsets end at 8192
                    cntv = m0 == 4u ? cw_umin(1u + cntv, 4u) : 0u;
                } else
                    ok = valid[j] && x0 == 0u;
                scr->pos[32 * j + lane] = (uint16_t)cntv;
                const uint32_t cc = ok ? cj : NONE;
                const uint32_t rot = cw_shfl(cc, (lane + 31u) & 31u);
                const uint32_t prevc = lane == 0u ? rotprev : rot;
                rotprev = rot;
                okm[j] = cw_ballot(ok);
                psm[j] = cw_ballot(ok && prevc + 1u == cj);  // my match continues the one of the position before me
            }
            psm[SWN] = 0u;
            // nsame[j] bit x: positions x and x + 1 of window j do NOT belong to one match
#pragma unroll
            for (int j = 0; j < SWN; j++) nsame[j] = ~((psm[j] >> 1) | ((psm[j + 1] & 1u) << 31));
            nsame[SWN] = 0xffffffffu;
            const uint32_t okany = okm[0] | okm[1] | okm[2] | (SWN == 4 ? okm[3] : 0u);
            if (lane < 6u) scr->okw[lane] = lane == 0u ? okm[0] : lane == 1u ? okm[1] : lane == 2u ? okm[2] : (SWN == 4 && lane == 3u) ? okm[3] : 0u;
            cw_syncwarp();  // counts and masks are in the scratch
            // ---- phase 3a: every position's entry: E = where its match ends, N = what the parse chooses after it ----
            if (okany) {
                uint32_t ent[CW_LZ4F_SW];
#pragma unroll
                for (int j = 0; j < SWN; j++) {
                    const uint32_t ns = cw_funnel_r(nsame[j], nsame[j + 1], lane);  // bit k: the chain breaks after lane + k
                    const uint32_t dr = ns ? cw_lz4f_ctz(ns) : 32u;
                    const uint32_t r = 32u * j + lane + dr;  // last position of my chain (dr = 32: it goes on)
                    const uint32_t cr = scr->pos[r];
                    const bool opn = dr == 32u || cr == 4u;
                    const uint32_t E = r + KB + (cr & 3u);  // end of my match (a lower bound if open)
                    const uint32_t lo = scr->okw[E >> 5], hi = scr->okw[(E >> 5) + 1u];
                    const uint32_t t = cw_funnel_r(lo, hi, E & 31u);  // verified positions from E on
                    uint32_t N = t ? E + cw_lz4f_ctz(t) : CW_LZ4F_FAR;
                    if (E >= G) N = CW_LZ4F_LEAVE;
                    if (opn) N = CW_LZ4F_OPEN;
                    ent[j] = N | ((E - (32u * j + lane)) << 8);  // N <= 255; the match ends at most 39 positions behind its start
                }
                cw_syncwarp();  // every lane has read the counts
#pragma unroll
                for (int j = 0; j < SWN; j++) scr->pos[32 * j + lane] = (uint16_t)ent[j];
                cw_syncwarp();  // the entries are in place
            }
            // ---- phase 3b: the greedy walk.  Where it enters the group depends on the group before (ip), but chains that
            //      enter a little apart meet after a match or two.  So the warp walks AHEAD of its turn, from the group's
            //      first verified position, and marks what it chooses; when its turn comes it walks from the real entry only
            //      until it steps on a marked position, and takes the rest of the speculative chain from there.  The
            //      ordered section of a group is then a few instructions instead of one step per match ----
            //      (SPEC; worth it where few chunks fit an SM and the chain of walks is what its teams wait for -- chunks read
            //      in place, LZF's 32 KB tables: +30 % -- and not where the kernel is issue bound: -8 % at 8 KiB LZ4)
            uint32_t ptot = 0, sp_end = 0, extE_sp = 0, specpos = 0;
            if (SPEC && okany) {
                const uint32_t x0 = cw_lz4f_scan(scr->okw, 0u);
                cw_lz4f_walk<DIRECT, true, false, SWN>(scr, scr->rec, x0, ptot, sp_end, extE_sp, c, s, sbase, matchlimit, nw, lane);
                cw_syncwarp();  // marks and records are in the scratch
                specpos = lane < ptot ? scr->rec[lane] : 0xffffu;
            }
            // (ordered: after the walk of group g - 1)
            if (g) cw_lz4f_tok_wait(&team->tok_walk[rank], seen + 1u, lane);
            ip = cw_ld_vol(&team->ip);
            cw_syncwarp();  // every lane has read it before lane 0 writes the new value
            anchor = ip == 1u ? 0u : ip;  // the parse stands at the end of the last match (nothing chosen yet: 1 / 0)
            uint32_t tot = 0, k = 0, jm = 0, extE_new = 0;  // k: matches chosen before the chains met, jm: where they met
            if (okany) {
                const uint32_t rel = ip > sbase ? ip - sbase : 0u;
                uint32_t x = rel < 128u ? cw_lz4f_scan(scr->okw, rel) : 128u;
                if (x < 128u)
                    x = cw_lz4f_walk<DIRECT, false, SPEC, SWN>(scr, scr->rec2, x, k, ip, extE_new, c, s, sbase, matchlimit, nw, lane);
                tot = k;
                if (x < 128u) {  // met the speculative chain at x: its place there, and the rest is known
                    jm = cw_lz4f_ctz(cw_ballot(specpos == x));
                    tot = k + ptot - jm;
                    ip = sp_end;
                }
            }
            if (lane == 0u) cw_st_vol(&team->ip, ip);
            if (g + 1u < groups) cw_lz4f_tok_post(&team->tok_walk[next], lane);
            // ---- phase 4: the sequences this group closes ----
            if (!tot) {  // nothing to write: hand the output offset on
                if (g) {
                    cw_lz4f_tok_wait(&team->tok_emit[rank], seen + 1u, lane);
                    seen++;
                }
                op = cw_ld_vol(&team->op);
                if (g + 1u < groups) cw_lz4f_tok_post(&team->tok_emit[next], lane);
                cw_syncwarp();  // the scratch is read before the next group's counts overwrite it
                continue;
            }
            cw_syncwarp();  // the walk's records are in the scratch
            // one lane per chosen match, in position order
            const bool mine = lane < tot;
            uint32_t P = 0, E = 0, C = 0, xr = 0;
            {
                // lane i takes match i: one of the k chosen before the chains met, else number jm + i - k of the speculative chain
                const uint32_t src = (jm + lane - k) & 31u;
                const uint32_t xsp = SPEC ? cw_shfl(specpos, src) : 0u, esp = SPEC ? cw_shfl(extE_sp, src) : 0u;
                if (mine) {
                    xr = lane < k ? (uint32_t)scr->rec2[lane] : xsp;
                    const uint32_t t = scr->pos[xr];
                    P = sbase + xr;
                    E = (t & 255u) == CW_LZ4F_OPEN ? (lane < k ? extE_new : esp) : P + ((t >> 8) & 63u);
                }
            }
            uint32_t sel[CW_LZ4F_SW];  // the chosen positions of every window
#pragma unroll
            for (int j = 0; j < SWN; j++) {
                const uint32_t q = cw_shfl(c[j], xr & 31u);  // my match's candidate sits in the lane that owns the position
                const bool here = mine && (xr >> 5) == (uint32_t)j;
                if (here) C = q;
                sel[j] = cw_reduce_or(here ? 1u << (xr & 31u) : 0u);
            }
            const uint32_t prevE = cw_shfl(E, (lane + 31u) & 31u);
            const uint32_t A = lane == 0u ? anchor : prevE;  // first literal of my sequence
            // backward extension (it changes literal and match lengths, not where the parse continues)
            uint32_t back = 0;
            if (mine) {
                const uint32_t maxb = cw_umin(P - A, C);
                while (back < maxb && cw_lz4_src8<DIRECT>(s, P - 1u - back) == cw_lz4_src8<DIRECT>(s, C - 1u - back)) back++;
            }
            if constexpr (LZF) {
                // ---- lzf.h block format: literal runs of at most 32 bytes behind a header byte (length - 1); a match is
                //      (length - 2) << 5 | offset >> 8, [length - 9], offset & 255 with offset = distance - 1, at most 264
                //      bytes long: longer matches are written as several tokens of the same distance ----
                const uint32_t Pb = P - back, lit = Pb - A, mlen = E - Pb;
                const uint32_t lsz = lit + ((lit + 31u) >> 5);
                uint32_t msz = 0;
                if (mine) {
                    uint32_t l = mlen;
#pragma unroll 1
                    while (l > 264u) {
                        l -= l - 264u < 3u ? 262u : 264u;  // never leave a rest shorter than a match
                        msz += 3u;
                    }
                    msz += l - 2u < 7u ? 2u : 3u;
                }
                const uint32_t sz = mine ? lsz + msz : 0u;
                uint32_t inc = sz;  // inclusive prefix sum over the lanes
#pragma unroll
                for (uint32_t d = 1; d < 32u; d <<= 1) {
                    const uint32_t t = cw_shfl(inc, lane >= d ? lane - d : lane);
                    if (lane >= d) inc += t;
                }
                const uint32_t total = cw_shfl(inc, 31u);
                // (ordered: after the allotment of group g - 1) this group's place in the block
                if (g) {
                    cw_lz4f_tok_wait(&team->tok_emit[rank], seen + 1u, lane);
                    seen++;
                }
                op = cw_ld_vol(&team->op);
                cw_syncwarp();  // every lane has read it before lane 0 writes the new value
                if (lane == 0u) cw_st_vol(&team->op, op + total);
                if (g + 1u < groups) cw_lz4f_tok_post(&team->tok_emit[next], lane);
                // a block that outgrows `cap` is dropped (the caller stores the chunk raw): from the first group that does
                // not fit nothing is written any more -- the offsets only grow -- and the chunk's length comes out as 0
                const bool fits = op + total <= cap;
                const uint32_t start = op + inc - sz;
                cw_syncwarp();  // the entries are read: the scratch now takes the literal map
                if (mine) {
                    scr->map[lane][0] = start;  // literal k of my run goes to out[start + 1 + k + k / 32]
                    scr->map[lane][1] = A | (Pb << 16);
                    if (fits) {
                        uint32_t w = start + lsz, l = mlen;
                        const uint32_t off = P - C - 1u;
#pragma unroll 1
                        while (l > 264u) {
                            const uint32_t take = l - 264u < 3u ? 262u : 264u;
                            out[w] = (uint8_t)((7u << 5) | (off >> 8));
                            out[w + 1u] = (uint8_t)(take - 9u);
                            out[w + 2u] = (uint8_t)off;
                            w += 3u;
                            l -= take;
                        }
                        if (l - 2u < 7u) {
                            out[w] = (uint8_t)(((l - 2u) << 5) | (off >> 8));
                            out[w + 1u] = (uint8_t)off;
                        } else {
                            out[w] = (uint8_t)((7u << 5) | (off >> 8));
                            out[w + 1u] = (uint8_t)(l - 9u);
                            out[w + 2u] = (uint8_t)off;
                        }
                    }
                }
                op += total;
                cw_syncwarp();
                if (fits) {
                    // literals in front of the group (they belong to its first sequence): copied by the whole warp
                    if (anchor < sbase) {
                        const uint32_t m0 = scr->map[0][0], fend = scr->map[0][1] >> 16;
                        const uint32_t stop = cw_umin(sbase, fend);
                        if (stop > anchor) cw_lzff_put_literals<DIRECT>(out, m0, s, anchor, fend - anchor, stop - anchor);
                    }
                    // literals inside the group: every position looks up its run and stores its own byte (and the run header
                    // in front of every 32nd)
                    uint32_t cum = 0;
#pragma unroll
                    for (int j = 0; j < SWN; j++) {
                        const uint32_t x = sbase + 32u * j + lane;
                        const uint32_t idx = cum + (uint32_t)cw_popc(sel[j] & lt);
                        if (idx < tot) {
                            const uint32_t m0 = scr->map[idx][0], m1 = scr->map[idx][1];
                            const uint32_t k = x - (m1 & 0xffffu), run = (m1 >> 16) - (m1 & 0xffffu);
                            if (k < run) {  // (k wraps for positions below the run)
                                const uint32_t w = m0 + k + (k >> 5);
                                out[w + 1u] = (uint8_t)v[j];
                                if ((k & 31u) == 0u) out[w] = (uint8_t)(cw_umin(32u, run - k) - 1u);
                            }
                        }
                        cum += (uint32_t)cw_popc(sel[j]);
                    }
                }
                cw_syncwarp();  // the map is read before the next group's counts overwrite the scratch
            } else {
                const uint32_t lit = P - back - A, mcode = E - P - 4u + back;
                const uint32_t le = cw_lz4_ext_bytes(lit);
                const uint32_t sz = mine ? 3u + le + lit + cw_lz4_ext_bytes(mcode) : 0u;
                uint32_t inc = sz;  // inclusive prefix sum over the lanes
    #pragma unroll
                for (uint32_t d = 1; d < 32u; d <<= 1) {
                    const uint32_t t = cw_shfl(inc, lane >= d ? lane - d : lane);
                    if (lane >= d) inc += t;
                }
                const uint32_t total = cw_shfl(inc, 31u);
                // (ordered: after the allotment of group g - 1) this group's place in the block
                if (g) {
                    cw_lz4f_tok_wait(&team->tok_emit[rank], seen + 1u, lane);
                    seen++;
                }
                op = cw_ld_vol(&team->op);
                cw_syncwarp();  // every lane has read it before lane 0 writes the new value
                if (lane == 0u) cw_st_vol(&team->op, op + total);
                if (g + 1u < groups) cw_lz4f_tok_post(&team->tok_emit[next], lane);
                const uint32_t start = op + inc - sz;
                cw_syncwarp();  // the entries are read: the scratch now takes the literal map
                if (mine) {
                    uint32_t w = start;
                    out[w++] = (uint8_t)((cw_umin(lit, 15u) << 4) | cw_umin(mcode, 15u));
                    if (lit >= 15u) w = cw_lz4_put_ext1(out, w, lit - 15u);
                    scr->map[lane][0] = w - A;  // literal x of my sequence goes to out[w - A + x]
                    scr->map[lane][1] = A | ((P - back) << 16);
                    w += lit;
                    const uint32_t off = P - C;
                    out[w] = (uint8_t)off;
                    out[w + 1u] = (uint8_t)(off >> 8);
                    if (mcode >= 15u) cw_lz4_put_ext1(out, w + 2u, mcode - 15u);
                }
                op += total;
                cw_syncwarp();
                // literals in front of the group (they belong to its first sequence): copied by the whole warp
                if (anchor < sbase) {
                    const uint32_t m0 = scr->map[0][0], fend = scr->map[0][1] >> 16;
                    const uint32_t stop = cw_umin(sbase, fend);
                    if (stop > anchor) {
                        if (stop - anchor <= 32u) {  // the usual case: a few bytes between the last match and the group
                            if (lane < stop - anchor) out[m0 + anchor + lane] = cw_lz4_src8<DIRECT>(s, anchor + lane);
                        } else
                            cw_lz4f_copy<DIRECT>(out, m0 + anchor, s, anchor, stop - anchor, nw);
                    }
                }
                // literals inside the group: every position looks up its sequence and stores its own byte
                uint32_t cum = 0;
    #pragma unroll
                for (int j = 0; j < SWN; j++) {
                    const uint32_t x = sbase + 32u * j + lane;
                    const uint32_t idx = cum + (uint32_t)cw_popc(sel[j] & lt);
                    // (one predicate, one predicated store: entries at or beyond tot are stale but lie inside the map)
                    const uint32_t m0 = scr->map[idx & 31u][0], m1 = scr->map[idx & 31u][1];
                    const uint32_t a0 = m1 & 0xffffu;
                    if (idx < tot && x - a0 < (m1 >> 16) - a0) out[m0 + x] = (uint8_t)v[j];
                    cum += (uint32_t)cw_popc(sel[j]);
                }
                cw_syncwarp();  // the map is read before the next group's counts overwrite the scratch
            }
        }
        if ((groups - 1u) % nteam != rank) return CW_LZ4_NONE;  // another warp owns the last group and finishes the chunk
        anchor = ip == 1u ? 0u : ip;
    } else if (rank != 0u)
        return CW_LZ4_NONE;
    // ---- end of chunk: the final literal run [anchor, n) ----
    const uint32_t lit = n - anchor;
    if constexpr (LZF) {
        const uint32_t total = op + lit + ((lit + 31u) >> 5);
        if (total > cap) return 0u;  // did not shrink: the caller keeps the chunk as it is (lzf.h:59-64)
        cw_lzff_put_literals<DIRECT>(out, op, s, anchor, lit, lit);
        return total;
    } else {
        const uint32_t qe = op + 1u + cw_lz4_ext_bytes(lit);
        if (lane == 0u) {
            out[op] = (uint8_t)(cw_umin(lit, 15u) << 4);
            if (lit >= 15u) cw_lz4_put_ext1(out, op + 1u, lit - 15u);
        }
        cw_lz4f_copy<DIRECT>(out, qe, s, anchor, lit, nw);
        return qe + lit;
    }
}
