// lz4_fast.cuh -- LZ4 "fast mode": a window-parallel parse of ONE chunk by ONE warp.
//
// The reference side of fast mode is only the decoder: every block must decode bit-exactly with
// LZ4_decompress_safe (src/compression_perf/include/lz4/lz4.h:141-149; call site experiment.cpp:256) and the
// compressed size must stay within 1 % of what LZ4_compress_default (HashAndCompress.cpp:349-355) produces.  No
// parse order is imposed, so nothing of the sequential parser is replayed here (that is compat mode,
// lz4_warp.cuh); the parse below is laid out for a warp:
//
//   * the chunk is cut into windows of 32 consecutive positions, one lane per position.  ALL 32 positions of a
//     window are hashed, looked up and inserted together (the reference inserts only the positions its skipping
//     search probes).  Four windows (a group of 128 positions) are in flight per loop iteration, so that a warp has
//     four independent dependency chains to overlap: with one window at a time the parse was bound by the latency
//     of its own chain of shared-memory round trips (profiles/r02_lz4_fast_first_run.txt);
//   * in-window candidates: after the look-up every lane stores its position into its slot, the warp syncs, and
//     every lane reads the slot back.  One store per slot survives; lanes with a higher position than the
//     survivor take it as their candidate (it is more recent than the old entry).  The ratio is best when the
//     LOWEST position of a window owns the slot (tools/lz4_fast_model.c: every file of the reference's corpora
//     <= 1.006 x the reference's size that way, <= 1.023 x with the opposite winner).  Which of several stores of
//     a warp to one address survives is undefined, so lanes that lost to a higher position store again until the
//     lowest one owns the slot; on B200 the lowest lane's store survives at once (corpus ratio 0.5835 with
//     positions ascending over the lanes, 0.5849 descending, the reference's parse 0.5844:
//     profiles/r02_lz4_fast_first_run.txt) and the loop never iterates.  That makes the parse a deterministic
//     function of the chunk: the tests hold a scalar restatement of it and compare the GPU output byte for
//     byte.  Correctness of a block never depends on it: every candidate is below its position and verified;
//   * a position whose 4 bytes equal the 4 bytes two positions back takes p - 2 as its candidate (what the
//     reference achieves by inserting ip-2 after a match and re-testing at once: runs and period-2 patterns are
//     caught by one overlapping match instead of being chopped at every old candidate's end);
//   * every lane with a verified candidate measures its own match for up to CW_LZ4F_SPEC more words; the greedy
//     choice "first match at or after the end of the previous one" is then a walk over registers: one shuffle
//     per chosen match.  Only matches that are still open after the speculation are extended by the whole warp
//     (128 bytes per step).  Windows that lie entirely inside a match when their group starts are skipped without
//     being inserted;
//   * the chosen matches of a window are written as records by their own lanes into the parser -> emitter ring of
//     lz4_warp.cuh; the emitter warps turn them into tokens, lengths, offsets and literal copies exactly as for
//     compat mode (including the backward extension, which only improves the ratio here).
//
// End rules of the block format: a match starts at or before n - 12 and ends at or before
// n - 5; chunks shorter than 13 bytes are one literal run.
//
// Table: zeroed u16[1 << HASHLOG] in shared memory.  An empty slot reads as candidate 0, which is a legal
// source position and is verified like any other.
#pragma once
#include "lz4_warp.cuh"

#ifndef CW_LZ4F_SPEC
#define CW_LZ4F_SPEC 3  // per-lane speculative extension: words of 4 bytes after the first 4
#endif
#define CW_LZ4F_SW 4  // windows per group

// hash table size for a chunk size: chosen on the reference's corpora with tools/lz4_fast_model.c so that the worst
// file stays below 1.01 x the reference's size (4 KiB 1.006, 8 KiB 1.002 with 12 bits; 16 KiB 0.998 with 13;
// 64 KiB 0.990 with 14 -- 13 bits would give 1.012 on world192.txt)
CW_HOSTDEV int cw_lz4f_hashlog(uint32_t chunk) { return chunk <= 8192u ? 12 : chunk <= 16384u ? 13 : 14; }

// aligned 32-bit word i of the chunk (DIRECT: words at or beyond nw read as 0)
template <bool DIRECT>
CW_DEV uint32_t cw_lz4f_word(const uint32_t *w, uint32_t i, uint32_t nw)
{
    if (!DIRECT) return w[i];
    return i < nw ? cw_ldg32(w + i) : 0u;
}

template <int HASHLOG, bool DIRECT>
CW_DEV void cw_lz4f_parse_warp(const uint8_t *s, uint32_t n, uint16_t *table, CwLz4Queue *q, uint32_t &qh,
                               uint32_t qlen = CW_LZ4_QLEN)
{
    const uint32_t nw = n >> 2;
    const unsigned lane = cw_lane();
    const uint32_t below_me = (1u << lane) - 1u;  // window positions below mine
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(s);
    const unsigned sh = (lane & 3u) * 8u;  // my position's byte offset inside its aligned word, in bits
    uint32_t anchor = 0, tailc = qh;       // the caller drained the queue: tail == qh
    uint32_t pub = qh;                     // records published so far

    if (n >= 13u) {
        const uint32_t last = n - 12u;       // last position a match may start at
        const uint32_t matchlimit = n - 5u;  // matches never cover the last 5 bytes
        uint32_t ip = 1u;                    // next position the greedy parse looks at (position 0 has no source)
        uint32_t sbase = 0u;                 // first position of the group, a multiple of 128
        while (sbase <= last) {
            uint32_t p[CW_LZ4F_SW], v[CW_LZ4F_SW], own[CW_LZ4F_SW][CW_LZ4F_SPEC], h[CW_LZ4F_SW], c[CW_LZ4F_SW], fw[CW_LZ4F_SW];
            uint32_t okm[CW_LZ4F_SW], openm[CW_LZ4F_SW];
            bool valid[CW_LZ4F_SW], per2[CW_LZ4F_SW];
            // ---- phase 1: my own bytes in every window of the group (independent loads) ----
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                const uint32_t wbase = sbase + 32u * j;
                p[j] = wbase + lane;
                // a window is live unless it lies beyond the last match start or inside a match when the group starts
                const bool live = wbase <= last && ip < wbase + 32u;
                valid[j] = live && p[j] <= last && p[j] != 0u;
                const uint32_t wi = (wbase >> 2) + (lane >> 2);
                uint32_t w0 = cw_lz4f_word<DIRECT>(sw, wi, nw), w1 = cw_lz4f_word<DIRECT>(sw, wi + 1u, nw);
                v[j] = cw_funnel_r(w0, w1, sh);
#pragma unroll
                for (int k = 0; k < CW_LZ4F_SPEC; k++) {
                    w0 = w1;
                    w1 = cw_lz4f_word<DIRECT>(sw, wi + 2u + k, nw);
                    own[j][k] = cw_funnel_r(w0, w1, sh);
                }
                // the 4 bytes two positions back (run / period-2 candidate)
                per2[j] = p[j] >= 2u && cw_lz4_src32<DIRECT>(s, p[j] >= 2u ? p[j] - 2u : 0u, nw) == v[j];
                h[j] = cw_lz4_hash<HASHLOG>(v[j]);
            }
            // ---- phase 2: look-up and insert, window by window (a window sees the inserts of the windows before it) ----
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                c[j] = table[h[j]];
                cw_syncwarp();  // every lane has read the old entry before any lane stores
                if (valid[j]) table[h[j]] = (uint16_t)p[j];
                cw_syncwarp();
                // the LOWEST position of the window that maps to a slot must own it: lanes that lost to a higher position
                // store again (on B200 the lowest lane's store survives and this loop never iterates)
                uint32_t w = table[h[j]];
                while (cw_ballot(valid[j] && p[j] < w)) {
                    if (valid[j] && p[j] < w) table[h[j]] = (uint16_t)p[j];
                    cw_syncwarp();
                    w = table[h[j]];
                }
                if (valid[j] && w < p[j]) c[j] = w;  // an earlier position of this window: more recent than the old entry
                if (per2[j]) c[j] = p[j] - 2u;
            }
            // ---- phase 3: verify the candidate and measure the match, every lane for itself ----
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                const uint32_t ci = c[j] >> 2;
                const unsigned csh = (c[j] & 3u) * 8u;
                uint32_t w0 = cw_lz4f_word<DIRECT>(sw, ci, nw), w1 = cw_lz4f_word<DIRECT>(sw, ci + 1u, nw);
                const bool ok = valid[j] && cw_funnel_r(w0, w1, csh) == v[j];
                uint32_t f = 0;
                bool open = ok;
#pragma unroll
                for (int k = 0; k < CW_LZ4F_SPEC; k++) {
                    w0 = w1;
                    w1 = cw_lz4f_word<DIRECT>(sw, ci + 2u + k, nw);
                    const uint32_t d = cw_funnel_r(w0, w1, csh) ^ own[j][k];
                    const uint32_t eq = d ? (uint32_t)(cw_ffs(d) - 1) >> 3 : 4u;
                    if (open) f += eq;
                    open = open && eq == 4u;
                }
                const uint32_t fwmax = matchlimit - 4u - p[j];  // garbage for invalid lanes, which are never chosen
                if (f >= fwmax) {
                    f = fwmax;
                    open = false;
                }
                fw[j] = f;
                okm[j] = cw_ballot(ok);
                openm[j] = cw_ballot(ok && open);
            }
            // ---- phase 4: greedy walk over the group (uniform) and the records of the chosen matches ----
#pragma unroll
            for (int j = 0; j < CW_LZ4F_SW; j++) {
                const uint32_t base = sbase + 32u * j;
                if (!okm[j] || ip >= base + 32u) continue;
                uint32_t ipl = ip > base ? ip - base : 0u;
                uint32_t my_end = lane + 4u + fw[j];  // window index of the end of my match
                uint32_t sel = 0;
                while (ipl < 32u) {
                    const uint32_t m = okm[j] & (0xffffffffu << ipl);
                    if (!m) break;
                    const unsigned f = (unsigned)cw_ffs(m) - 1u;
                    if ((openm[j] >> f) & 1u) {
                        // still equal after the speculation: the whole warp extends it, 32 lanes x 4 bytes per step
                        const uint32_t cur = base + f, cnd = cw_shfl(c[j], f);
                        uint32_t fwd = 4u * CW_LZ4F_SPEC;
                        for (;;) {
                            const uint32_t pa = cur + 4u + fwd + 4u * lane, pb = cnd + 4u + fwd + 4u * lane;
                            const uint32_t d = cw_lz4_src32<DIRECT>(s, pa, nw) ^ cw_lz4_src32<DIRECT>(s, pb, nw);
                            uint32_t eq = d ? (uint32_t)(cw_ffs(d) - 1) >> 3 : 4u;
                            eq = cw_umin(eq, pa < matchlimit ? matchlimit - pa : 0u);
                            const uint32_t stop = cw_ballot(eq < 4u);
                            if (stop) {
                                const unsigned L = (unsigned)cw_ffs(stop) - 1u;
                                fwd += 4u * L + cw_shfl(eq, L);
                                break;
                            }
                            fwd += 128u;
                        }
                        if (lane == f) {
                            fw[j] = fwd;
                            my_end = lane + 4u + fwd;
                        }
                    }
                    sel |= 1u << f;
                    ipl = cw_shfl(my_end, f);
                }
                if (sel) {
                    // the chosen lanes write their records themselves; a record's anchor is the end of the match chosen
                    // before it (the carried-in anchor for the first of the window)
                    const uint32_t cnt = (uint32_t)cw_popc(sel), below = sel & below_me;
                    cw_lz4_q_reserve(q, qh, tailc, cnt, lane, qlen);
                    const uint32_t prev_end = cw_shfl(my_end, below ? (unsigned)(31 - cw_clz(below)) : lane);
                    if ((sel >> lane) & 1u) {
                        const uint32_t k = (qh + (uint32_t)cw_popc(below)) & (qlen - 1u);
                        q->rec[k][0] = (below ? base + prev_end : anchor) | (p[j] << 16);
                        q->rec[k][1] = c[j] | (fw[j] << 16);
                    }
                    qh += cnt;
                    anchor = ip = base + ipl;
                }
            }
            if (qh - pub >= CW_LZ4_PUB_BATCH) {  // publishing costs a fence: not after every group
                cw_lz4_q_publish(q, qh, lane);
                pub = qh;
            }
            sbase += 32u * CW_LZ4F_SW;
            if (ip >= sbase + 32u * CW_LZ4F_SW) sbase = ip & ~(32u * CW_LZ4F_SW - 1u);  // groups inside the match
        }
    }
    // ---- end of chunk: the final literal run [anchor, n) ----
    cw_lz4_q_push(q, qh, tailc, anchor | (CW_LZ4_REC_END << 16), 0u, lane, 0u, qlen);
    cw_lz4_q_publish(q, qh, lane);
}
