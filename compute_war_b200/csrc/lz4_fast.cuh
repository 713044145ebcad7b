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
// and a per-lane count CNT (how many of the bytes x + 4 .. x + 7 equal c + 4 .. c + 7) already hold the exact end of
// almost every match: it is the first position r at or after the start whose SAME bit is clear, plus 4 + CNT[r].
// Only when CNT[r] is 4 (the bytes go on although the neighbour chose another candidate), and for matches that
// outgrow a window, does the warp extend the match itself (32 bytes, then 128 per step).
//
// Emission happens in the same warp as the parse, once per group: the chosen matches are compacted into one lane each (backward
// extension, sizes, prefix sum, token / lengths / offset), and the literal bytes are then written position-wise --
// a lane looks up the sequence its position belongs to and stores its own byte -- so a group costs the same number
// of instructions whether it holds two sequences or twenty.  Literals that precede the group (a sequence is only
// written once its match is known) are copied by the whole warp.
//
// A chunk is a dependent chain of ~10^5 instructions, a warp issues one of them every 6-7 cycles, and shared memory
// holds only 13 chunks of 8 KiB per SM, so one warp per chunk leaves most issue slots idle.  A team of NW warps
// therefore shares a chunk: warp r takes groups r, r + NW, ... and only three short sections of a group are ordered
// by tokens in shared memory (CwLz4fTeam): the table look-ups / inserts (group g needs the inserts of g - 1), the
// greedy walk (needs the position the walk of g - 1 ended at) and the output offset (one addition).  Everything
// else of group g + 1 -- fetch, hash, verification, masks, sizes, stores -- overlaps with group g in the other warp.
// Nothing is copied between the warps: a group's values stay in the registers of the warp that owns it.
//
// End rules of the block format: a match starts at or before n - 12 and ends at or before
// n - 5; chunks shorter than 13 bytes are one literal run.
//
// Table: zeroed u16[1 << HASHLOG] in shared memory.  An empty slot reads as candidate 0, which is a legal
// source position and is verified like any other.
#pragma once
#include "lz4_warp.cuh"

#ifndef CW_LZ4F_SW
#define CW_LZ4F_SW 4  // windows per group (the output does not depend on it)
#endif
#ifndef CW_LZ4F_DOUBLINGS
#define CW_LZ4F_DOUBLINGS 2  // rounds of pointer doubling before the walk: chains of 4 choices per step of the walk
#endif

// hash table size for a chunk size: chosen on the reference's corpora with tools/lz4_fast_model.c so that the worst
// file stays below 1.01 x the reference's size (4 KiB 1.006, 8 KiB 1.002 with 12 bits; 16 KiB 0.998 with 13;
// 64 KiB 0.990 with 14 -- 13 bits would give 1.012 on world192.txt)
CW_HOSTDEV int cw_lz4f_hashlog(uint32_t chunk) { return chunk <= 8192u ? 12 : chunk <= 16384u ? 13 : 14; }

// per-warp scratch in shared memory: the chosen matches of a group, then the literal map of its sequences
union CwLz4fScratch {
    uint32_t rec[32][2];  // {start | end << 16, candidate}
    uint32_t map[32][2];  // {output offset of literal x, minus x; first literal | end of literals << 16}
};

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
#define CW_LZ4F_MAX_TEAM 4
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CW_TD_%=;\n"
        "bra CW_TW_%=;\n"
        "CW_TD_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
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

// Compress chunk s[0..n) (n <= 65536) into out, as warp `rank` of a team of `nteam` warps that all call this function
// with the same arguments.  s: 4-byte aligned; not DIRECT: CW_LZ4_PAD readable bytes behind the chunk.  table: zeroed
// u16[1 << HASHLOG]; team: tokens initialised, ip = 1, op = 0, visible to every warp of the team before the first one
// starts; seen: arrivals this warp has consumed on its tokens so far (kept by the caller from chunk to chunk); scr: this
// warp's own scratch.  Returns the block length in the warp that finishes the chunk (same value in every lane) and CW_LZ4_NONE in
// the others.
CW_HOSTDEV uint32_t cw_lz4f_groups(uint32_t n) { return n >= 13u ? (n - 12u) / (32u * CW_LZ4F_SW) + 1u : 0u; }
template <int HASHLOG, bool DIRECT>
CW_DEV uint32_t cw_lz4f_team_chunk(const uint8_t *s, uint32_t n, uint16_t *table, CwLz4fTeam *team, CwLz4fScratch *scr,
                                   uint8_t *out, uint32_t rank, uint32_t nteam, uint32_t &seen)
{
    const uint32_t next = rank + 1u == nteam ? 0u : rank + 1u;  // owner of the group after mine
    const uint32_t nw = n >> 2;
    const unsigned lane = cw_lane();
    const uint32_t lt = (1u << lane) - 1u;  // window positions below mine
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s);
    const unsigned sh = (lane & 3u) * 8u;  // my position's byte offset inside its aligned word, in bits
    const uint32_t NONE = 0xfff00000u;     // "no verified candidate": never one below a real candidate
    uint32_t anchor = 0, op = 0;

    if (n >= 13u) {
        const uint32_t last = n - 12u;       // last position a match may start at
        const uint32_t matchlimit = n - 5u;  // matches never cover the last 5 bytes
        const uint32_t groups = cw_lz4f_groups(n);
        uint32_t ip = 1u;  // next position the greedy parse looks at (position 0 has no source)
        for (uint32_t g = rank; g < groups; g += nteam) {
            const uint32_t sbase = g * (32u * CW_LZ4F_SW);  // first position of the group
            uint32_t v[CW_LZ4F_SW], c[CW_LZ4F_SW], h[CW_LZ4F_SW], own4[CW_LZ4F_SW], cnt[CW_LZ4F_SW], myend[CW_LZ4F_SW];
            uint32_t okm[CW_LZ4F_SW + 1], psm[CW_LZ4F_SW + 1], same[CW_LZ4F_SW + 1], sel[CW_LZ4F_SW];
            bool live[CW_LZ4F_SW], valid[CW_LZ4F_SW], per2[CW_LZ4F_SW];
            // ---- phase 1: my own bytes in every window of the group (independent loads) ----
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                const uint32_t wbase = sbase + 32u * j, p = wbase + lane;
                live[j] = wbase <= last;  // windows beyond the last match start hold literals only
                valid[j] = live[j] && p <= last && p != 0u;
                v[j] = h[j] = own4[j] = c[j] = cnt[j] = 0u;
                per2[j] = false;
                myend[j] = 0u;
                sel[j] = 0u;
                if (live[j]) {
                    const uint32_t wi = (wbase >> 2) + (lane >> 2);
                    const uint32_t w0 = cw_lz4f_word<DIRECT>(sw, wi, nw), w1 = cw_lz4f_word<DIRECT>(sw, wi + 1u, nw);
                    v[j] = cw_funnel_r(w0, w1, sh);
                    own4[j] = cw_funnel_r(w1, cw_lz4f_word<DIRECT>(sw, wi + 2u, nw), sh);  // bytes p + 4 .. p + 7
                    // the 4 bytes two positions back (run / period-2 candidate)
                    per2[j] = p >= 2u && cw_lz4_src32<DIRECT>(s, p >= 2u ? p - 2u : 0u, nw) == v[j];
                    h[j] = cw_lz4_hash<HASHLOG>(v[j]);
                }
            }
            // ---- phase 2a (ordered: after the inserts of group g - 1): look-up and insert, window by window (a window
            //      sees the inserts of the windows before it) ----
            if (g) cw_lz4f_tok_wait(&team->tok_table[rank], seen + 1u, lane);
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                if (!live[j]) continue;
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
                c[j] = cj;
            }
            if (g + 1u < groups) cw_lz4f_tok_post(&team->tok_table[next], lane);
            // ---- phase 2b: candidate verification and the OK / CONT / chain masks ----
            uint32_t c_last = NONE;  // lane 31's verified candidate in the window before
            okm[CW_LZ4F_SW] = psm[CW_LZ4F_SW] = 0u;
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                okm[j] = psm[j] = 0u;
                if (!live[j]) {
                    c_last = NONE;
                    continue;
                }
                const uint32_t p = sbase + 32u * j + lane;
                uint32_t cj = c[j];
                if (per2[j]) cj = p - 2u;
                const bool ok = valid[j] && cw_lz4_src32<DIRECT>(s, valid[j] ? cj : 0u, nw) == v[j];
                if (ok) {  // how many of the next 4 bytes go on matching (4: at least that many)
                    const uint32_t x = cw_lz4_src32<DIRECT>(s, cj + 4u, nw) ^ own4[j];
                    cnt[j] = x ? cw_lz4f_ctz(x) >> 3 : 4u;
                }
                c[j] = cj;
                const uint32_t cc = ok ? cj : NONE;
                uint32_t prevc = cw_shfl(cc, (lane + 31u) & 31u);
                if (lane == 0u) prevc = c_last;
                c_last = cw_shfl(cc, 31u);
                okm[j] = cw_ballot(ok);
                psm[j] = cw_ballot(ok && prevc + 1u == cj);  // my match continues the one of the position before me
            }
            // SAME[j] bit x: positions x and x + 1 of window j belong to one match
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) same[j] = (psm[j] >> 1) | ((psm[j + 1] & 1u) << 31);
            same[CW_LZ4F_SW] = 0u;
            // ---- phase 3 (ordered: after the walk of group g - 1): greedy walk over the group ----
            // Every verified lane knows from the masks where its own match ends (exactly, unless it is "open") and which
            // lane the greedy parse would choose after it; rounds of pointer doubling give every lane the chain of the next
            // 2^rounds choices that starts at it (a window holds at most 8 matches).  A step of the walk itself is then one
            // pair of shuffles, plus one whole-warp extension if the chain stops at an open match.
            uint32_t pk[CW_LZ4F_SW], R[CW_LZ4F_SW];
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                pk[j] = R[j] = 0u;
                if (!okm[j]) continue;
                const uint32_t ns = cw_funnel_r(~same[j], ~same[j + 1], lane);  // bit k: the chain breaks after lane + k
                const uint32_t dr = ns ? cw_lz4f_ctz(ns) : 32u;
                // the bytes behind the last position of my chain: 0-3 more of them match, or the match is "open"
                const uint32_t r = lane + dr;
                const uint32_t ca = cw_shfl(cnt[j], r & 31u), cb = cw_shfl(j + 1 < CW_LZ4F_SW ? cnt[j + 1] : 0u, r & 31u);
                const uint32_t cr = r < 32u ? ca : cb;
                const bool opn = !ns || cr == 4u;
                const uint32_t endr = r + 4u + (opn ? 0u : cr);  // end of my match, relative to the window (a lower bound if open)
                myend[j] = sbase + 32u * j + endr;
                uint32_t J = 32u;  // next chosen lane after my match; 32: the chain leaves the window or stops at an open match
                if (!opn && endr < 32u) {
                    const uint32_t m = okm[j] & (0xffffffffu << endr);
                    if (m) J = cw_lz4f_ctz(m);
                }
                uint32_t Rj = 1u << lane;                                // lanes chosen on the chain that starts at me
                uint32_t pkj = J | (endr << 8) | (opn ? 1u << 16 : 0u);  // chain: next lane | end | ends in an open match
#pragma unroll
                for (int k = 0; k < CW_LZ4F_DOUBLINGS; k++) {
                    const uint32_t nx = pkj & 0xffu;
                    const uint32_t pk2 = cw_shfl(pkj, nx & 31u), R2 = cw_shfl(Rj, nx & 31u);
                    if (nx < 32u) {
                        Rj |= R2;
                        pkj = pk2;
                    }
                }
                pk[j] = pkj;
                R[j] = Rj;
            }
            if (g) cw_lz4f_tok_wait(&team->tok_walk[rank], seen + 1u, lane);
            ip = cw_ld_vol(&team->ip);
            cw_syncwarp();  // every lane has read it before lane 0 writes the new value
            anchor = ip == 1u ? 0u : ip;  // the parse stands at the end of the last match (nothing chosen yet: 1 / 0)
            uint32_t tot = 0;
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                const uint32_t base = sbase + 32u * j;
                if (!okm[j] || ip >= base + 32u) continue;
                uint32_t rel = ip > base ? ip - base : 0u;
                while (rel < 32u) {
                    const uint32_t m = okm[j] & (0xffffffffu << rel);
                    if (!m) break;
                    const unsigned f = cw_lz4f_ctz(m);
                    const uint32_t pkf = cw_shfl(pk[j], f), Rf = cw_shfl(R[j], f);
                    sel[j] |= Rf;
                    tot += (uint32_t)cw_popc(Rf);
                    uint32_t e = base + ((pkf >> 8) & 0xffu);
                    if ((pkf >> 16) & 1u) {  // the chain stopped at an open match (its highest lane): the warp extends it
                        const unsigned fo = 31u - (unsigned)cw_clz(Rf);
                        const uint32_t cnd = cw_shfl(c[j], fo);
                        e = cw_lz4f_extend<DIRECT>(s, e, base + fo - cnd, matchlimit, nw, lane);
                        if (lane == fo) myend[j] = e;
                    }
                    ip = e;
                    rel = e - base;
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
                continue;
            }
            // one lane per chosen match, in position order
            uint32_t cum = 0;
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                if (!sel[j]) continue;
                if ((sel[j] >> lane) & 1u) {
                    const uint32_t k = cum + (uint32_t)cw_popc(sel[j] & lt);
                    scr->rec[k][0] = (sbase + 32u * j + lane) | (myend[j] << 16);
                    scr->rec[k][1] = c[j];
                }
                cum += (uint32_t)cw_popc(sel[j]);
            }
            cw_syncwarp();
            const bool mine = lane < tot;
            uint32_t P = 0, E = 0, C = 0;
            if (mine) {
                const uint32_t r0 = scr->rec[lane][0];
                C = scr->rec[lane][1];
                P = r0 & 0xffffu;
                E = r0 >> 16;
            }
            const uint32_t prevE = cw_shfl(E, (lane + 31u) & 31u);
            const uint32_t A = lane == 0u ? anchor : prevE;  // first literal of my sequence
            // backward extension (it changes literal and match lengths, not where the parse continues)
            uint32_t back = 0;
            if (mine) {
                const uint32_t maxb = cw_umin(P - A, C);
                while (back < maxb && cw_lz4_src8<DIRECT>(s, P - 1u - back) == cw_lz4_src8<DIRECT>(s, C - 1u - back)) back++;
            }
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
            cw_syncwarp();  // the records are read: the scratch now takes the literal map
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
                if (stop > anchor) cw_lz4_copy_warp<DIRECT>(out, m0 + anchor, s, anchor, stop - anchor, nw, lane);
            }
            // literals inside the group: every position looks up its sequence and stores its own byte
            cum = 0;
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                if (!live[j]) continue;
                const uint32_t x = sbase + 32u * j + lane;
                const uint32_t idx = cum + (uint32_t)cw_popc(sel[j] & lt);
                if (idx < tot) {
                    const uint32_t m0 = scr->map[idx][0], m1 = scr->map[idx][1];
                    if (x >= (m1 & 0xffffu) && x < (m1 >> 16)) out[m0 + x] = (uint8_t)v[j];
                }
                cum += (uint32_t)cw_popc(sel[j]);
            }
            cw_syncwarp();  // the map is read before the next group's matches overwrite the scratch
        }
        if ((groups - 1u) % nteam != rank) return CW_LZ4_NONE;  // another warp owns the last group and finishes the chunk
        anchor = ip == 1u ? 0u : ip;
    } else if (rank != 0u)
        return CW_LZ4_NONE;
    // ---- end of chunk: the final literal run [anchor, n) ----
    const uint32_t lit = n - anchor;
    const uint32_t qe = op + 1u + cw_lz4_ext_bytes(lit);
    if (lane == 0u) {
        out[op] = (uint8_t)(cw_umin(lit, 15u) << 4);
        if (lit >= 15u) cw_lz4_put_ext1(out, op + 1u, lit - 15u);
    }
    cw_lz4_copy_warp<DIRECT>(out, qe, s, anchor, lit, nw, lane);
    return qe + lit;
}
