// lzf_warp.cuh -- LZF compression of one chunk by a PARSER warp and an EMITTER warp, byte-identical to the reference.
//
// Replaces the lzf lambda  lzf_compress(s, l, d, l - 1)  (HashAndCompress.cpp:342-348; experiment.cpp:110).
// The reference links a prebuilt liblzf.a (API 0x105) whose source is not in its tree; its build options are
// in src/compression_perf/include/lzf/lzfP.h:54-164 (HLOG 16, VERY_FAST 1, ULTRA_FAST 0, INIT_HTAB 0).
// The algorithm reproduced here is liblzf 3.x lzf_compress with those options and a zeroed table
// (SURVEY.md Appendix A.3), accepted only through byte-equality with the vendored archive.
//
// Sequential semantics: every position p < n-2 that is not inside a match is probed:
//   slot = f(3 bytes at p) (16 bit); ref = htab[slot]; htab[slot] = p;
//   match iff ref > 0 and p-ref-1 < 8192 and the 3 bytes are equal; no backward extension; match length
//   up to 264 with liblzf's unrolled-compare quirk; after a match ending at e the positions e-2 and e-1 are
//   inserted and probing resumes at e.  Literal runs are cut into pieces of <= 32 bytes, each with a
//   one-byte header.  The call returns 0 as soon as the output would not fit in n-1 bytes.
//
// B200 mapping: the reference's table has 65536 entries (128 KB as u16) -- too big for one warp's share of
// shared memory.  The same "most recent inserted position with this 16-bit slot" query is answered exactly
// by a chained table: head[slot & 8191] -> newest position in the bucket, next[pos & 8191] -> previous one;
// a lane walks the chain until it meets a position whose own slot (recomputed from its bytes) equals the
// query slot.  32 KB per chunk instead of 128 KB.  Positions older than 8192 bytes can never match
// (MAX_OFF), so the next[] ring never needs more than 8192 entries whatever the chunk size.
// For chunks of up to 8192 bytes (BB = 13) a table entry is  (slot >> 13) << 13 | position : the three slot bits the
// bucket index does not cover ride along with the 13-bit position (chunks of up to 4096 bytes, BB = 12: 4096 buckets,
// four tag bits, a 4096-entry ring -- 16 KB of table instead of 32, so almost twice the chunks per SM), so a chain walk is one shared-memory load per step
// and no trigram has to be re-read and re-hashed.  Position 0 terminates a chain: it is the oldest position and can
// never be a match source (the reference requires ref > in_data), exactly like an empty slot of its zeroed table.
// As in lz4_warp.cuh the parser warp evaluates 32 consecutive probe positions at once and then replays the
// parser's control flow over the window with an "inserted lanes" mask; it decides where matches start and how
// long they are (the next probe position depends on both) and hands {anchor, position, source, length} records to
// an emitter warp through the shared-memory ring of lz4_warp.cuh.  The emitter owns everything the parse does not
// depend on: liblzf's three output-space checks, the literal runs with their headers, the match tokens.
#pragma once
#include "cw_device.cuh"
#include "lz4_warp.cuh"  // cw_lz4_read32, cw_umin

#define CW_LZF_BUCKETS 8192u
#define CW_LZF_NONE 0xffffu
#define CW_LZF_MAX_OFF 8192u
#define CW_LZF_MAX_REF 264u

// Table formats (layout: u16 head[buckets], then u16 next[ring entries]):
//   0   plain positions, 8192 buckets + 8192-entry ring (any chunk size)
//   13  tagged entries, 8192 buckets + 8192-entry ring, chunks of up to 8192 bytes: tag = slot bits 13..15
//   12  tagged entries, 4096 buckets + 4096-entry ring, chunks of up to 4096 bytes: tag = slot bits 12..15
//   25  tagged entries, 4096 buckets + 8192-entry ring, chunks of up to 8192 bytes: the tag (slot bits 13..15) leaves
//       slot bit 12 open, so a tag match is confirmed by re-hashing the candidate; 24 KB instead of 32
//   24  the same with 2048 buckets (slot bits 11 and 12 open): 20 KB
#ifndef CW_LZF_FORMAT_8K
#define CW_LZF_FORMAT_8K 25
#endif
CW_HOSTDEV int cw_lzf_table_format(uint32_t chunk) { return chunk <= 4096u ? 12 : (chunk <= CW_LZF_BUCKETS ? CW_LZF_FORMAT_8K : 0); }
CW_HOSTDEV uint32_t cw_lzf_table_bytes(int fmt)
{
    return fmt == 12 ? 16384u : (fmt == 25 ? 24576u : (fmt == 24 ? 20480u : 32768u));
}

// 24-bit big-endian trigram at p (liblzf's hval) and its 16-bit table slot (IDX with VERY_FAST, HLOG 16)
CW_DEV uint32_t cw_lzf_hv(const uint8_t *s, uint32_t p)
{
    const uint32_t v = cw_lz4_read32(s, p);
    return ((v & 0xffu) << 16) | (v & 0xff00u) | ((v >> 16) & 0xffu);
}
CW_DEV uint32_t cw_lzf_slot(uint32_t hv) { return ((hv >> 8) - hv * 5u) & 0xffffu; }

// literal stretch: lit bytes s[src..] to out[op..] with a run header every 32 bytes (the header of the first
// run is the already reserved byte out[op-1])
CW_DEV void cw_lzf_put_literals(uint8_t *out, uint32_t op, const uint8_t *s, uint32_t src, uint32_t lit, unsigned lane)
{
    for (uint32_t k = lane; k < lit; k += 32u) out[op + k + (k >> 5)] = s[src + k];
    for (uint32_t j = lane; j * 32u < lit; j += 32u) out[op - 1u + 33u * j] = (uint8_t)(cw_umin(32u, lit - 32u * j) - 1u);
}

// ---------------------------------------------------------------------------------------------------------------
// PARSER.  s[0..n): the chunk (shared memory, >= 8 readable bytes behind it); tab: table prepared by
// cw_lzf_table_init<BB>; q / qh as in cw_lz4_parse_warp.  Records: {anchor | cur << 16, ref | len << 16}; the closing record is
// {CW_LZ4_REC_END << 16, anchor | "the last match ran into the end of the input" << 31}.
// ---------------------------------------------------------------------------------------------------------------
// empty table (the caller runs this while the chunk's bulk copy is in flight, then syncs the warp)
template <int BB>
CW_DEV void cw_lzf_table_init(uint16_t *tab)
{
    uint32_t *h = reinterpret_cast<uint32_t *>(tab);
    const uint32_t f = BB ? 0u : 0xffffffffu;
    for (uint32_t i = cw_lane(); i < (1u << (BB ? BB : 13)) / 2u; i += 32u) h[i] = f;
}

// BB: bucket bits (0 = the untagged format with 8192 buckets), RB: ring / position bits
template <int BB, int RB>
CW_DEV void cw_lzf_parse_warp(const uint8_t *s, uint32_t n, uint16_t *tab, CwLz4Queue *q, uint32_t &qh,
                              uint32_t qlen = CW_LZ4_QLEN)
{
    constexpr bool TAGGED = BB != 0;
    constexpr bool VERIFY = TAGGED && BB < RB;            // the tag does not cover every slot bit above the bucket index
    constexpr uint32_t NB = 1u << (TAGGED ? BB : 13);     // buckets
    constexpr uint32_t BM = NB - 1u;                      // bucket mask
    constexpr uint32_t PM = (1u << (TAGGED ? RB : 13)) - 1u;  // position mask of an entry / ring index mask
    constexpr uint32_t TM = 0xffffu & ~PM;                // tag bits of an entry = the slot bits from RB up
    uint16_t *const head = tab, *const next = tab + NB;
    const unsigned lane = cw_lane();
    const uint32_t lower = (1u << lane) - 1u;
    uint32_t tailc = qh;  // the caller drained the queue
    uint32_t pub = qh;    // records published so far
    uint32_t anchor = 0;                          // first literal of the current stretch
    const uint32_t limit = n > 2u ? n - 2u : 0u;  // positions < limit are probed
    uint32_t base = 0;
    bool ended_by_match = false;
    while (base < limit) {
        // ---------------- phase A ----------------
        const uint32_t pos = base + lane;
        const bool valid = pos < limit;
        const uint32_t hv = cw_lzf_hv(s, valid ? pos : 0u);
        const uint32_t slot = cw_lzf_slot(hv);
        const uint32_t b = slot & BM;
        const uint32_t head_old = head[b];
        uint32_t cq;  // most recent inserted position with this 16-bit slot, CW_LZF_NONE if there is none
        bool tmatch;
        if (TAGGED) {
            uint32_t e = valid ? head_old : 0u;
            for (;;) {
                cq = e & PM;
                if (cq == 0u) break;
                if ((e & TM) == (slot & TM) && (!VERIFY || cw_lzf_slot(cw_lzf_hv(s, cq)) == slot)) break;
                e = next[cq];
            }
            tmatch = cq != 0u && cw_lzf_hv(s, cq) == hv;
            if (cq == 0u) cq = CW_LZF_NONE;
        } else {
            uint32_t hq = 0;
            cq = valid ? head_old : CW_LZF_NONE;
            while (cq != CW_LZF_NONE) {
                if (pos - cq > CW_LZF_MAX_OFF) {  // too far to match, and its ring entry may be stale
                    cq = CW_LZF_NONE;
                    break;
                }
                hq = cw_lzf_hv(s, cq);
                if (cw_lzf_slot(hq) == slot) break;
                cq = next[cq & PM];
            }
            tmatch = valid && cq != CW_LZF_NONE && cq > 0u && (pos - cq - 1u) < CW_LZF_MAX_OFF && hq == hv;
        }
        const uint32_t own = 1u << lane;
        // lanes that share a bucket inside the window ("twins") are detected by tagging the bucket head and reading
        // it back; the chain walk above is done by then, and phase C rewrites every tagged head
        cw_syncwarp();
        if (valid) head[b] = (uint16_t)(0xffe0u | lane);
        cw_syncwarp();
        const bool owner = valid && head[b] == (0xffe0u | lane);
        uint32_t M13 = own, M16 = own, E = own;
        const uint32_t twins = cw_ballot(valid && !owner);
        if (cw_popc(twins) > 8) {  // many groups (e.g. tables of small integers): MATCH.ANY's fixed ~365 cycles each win
            M13 = cw_match_any(valid ? b : 0x10000u + lane);
            M16 = cw_match_any(valid ? slot : 0x10000u + lane);
            E = cw_match_any(valid ? hv : 0x1000000u + lane);
        } else {
            // every bucket shared by several lanes has at least one lane that lost the tag race: walk those, one group
            // of equal trigrams per step (three ballots)
            uint32_t rem = twins;
            while (rem) {
                const uint32_t hvL = cw_shfl(hv, (unsigned)cw_ffs(rem) - 1u);
                const uint32_t slotL = cw_lzf_slot(hvL);
                const bool s13 = valid && b == (slotL & BM), s16 = valid && slot == slotL, se = valid && hv == hvL;
                const uint32_t m13 = cw_ballot(s13), m16 = cw_ballot(s16), me = cw_ballot(se);
                if (s13) M13 = m13;
                if (s16) M16 = m16;
                if (se) E = me;
                rem &= ~me;
            }
        }
        // ---------------- phase B ----------------
        uint32_t ins = 0, i0 = 0, pend0 = CW_LZF_NONE, pend1 = CW_LZF_NONE, e = 0;
        bool restart = false;
        for (;;) {
            const uint32_t range = ~((1u << i0) - 1u);
            const uint32_t cm = M16 & (ins | range) & lower;
            bool hit = tmatch;
            uint32_t cand = cq;
            if (cm) {
                const int jl = 31 - cw_clz(cm);
                cand = base + (uint32_t)jl;
                hit = valid && ((E >> jl) & 1u) && cand > 0u;
            }
            const uint32_t hitm = cw_ballot(hit && lane >= i0);
            const int f = cw_ffs(hitm) - 1;
            if (f < 0) {
                ins |= range;  // invalid lanes are masked at commit
                break;
            }
            ins |= range & ((2u << f) - 1u);
            const uint32_t cur = base + (uint32_t)f;
            const uint32_t ref = cw_shfl(cand, (unsigned)f);
            // ---- match length with liblzf's unrolled-compare rule ----
            const uint32_t maxlen = cw_umin(CW_LZF_MAX_REF, n - cur - 2u);
            uint32_t fm;  // first mismatching index >= 3
            {
                uint32_t idx = 3u + 4u * lane;
                for (;;) {
                    const uint32_t d = cw_lz4_read32(s, cw_umin(cur + idx, n)) ^ cw_lz4_read32(s, cw_umin(ref + idx, n));
                    const uint32_t eq = d ? (uint32_t)(cw_ffs(d) - 1) >> 3 : 4u;
                    const uint32_t stop = cw_ballot(eq < 4u);
                    if (stop) {
                        const unsigned L0 = (unsigned)cw_ffs(stop) - 1u;
                        fm = cw_shfl(idx + eq, L0);
                        break;
                    }
                    idx += 128u;
                    if (idx - 4u * lane > CW_LZF_MAX_REF + 8u) {
                        fm = 0xffffu;
                        break;
                    }
                }
            }
            uint32_t len;
            if (maxlen > 16u) len = fm <= 18u ? fm : (fm < maxlen ? fm : (maxlen > 19u ? maxlen : 19u));
            else len = fm < maxlen ? fm : (maxlen > 3u ? maxlen : 3u);
            cw_lz4_q_push(q, qh, tailc, anchor | (cur << 16), ref | (len << 16), lane, 0u, qlen);
            e = cur + len;
            anchor = e;
            if (e >= limit) {  // "if (ip >= in_end - 2) break"
                ended_by_match = true;
                break;
            }
            // insert e-2 and e-1, resume probing at e
            if (e - 2u <= base + 31u) ins |= 1u << (e - 2u - base);
            else pend0 = e - 2u;
            if (e - 1u <= base + 31u) ins |= 1u << (e - 1u - base);
            else pend1 = e - 1u;
            if (e <= base + 31u) {
                i0 = e - base;
                continue;
            }
            restart = true;
            break;
        }
        if (ended_by_match) break;
        // ---------------- phase C: link the inserted positions into their bucket chains ----------------
        // every bucket touched by a valid lane carries a tag: its owner first puts the old head back, then the
        // inserted lanes link themselves in (highest inserted lane of a bucket becomes the new head)
        if (owner) head[b] = (uint16_t)head_old;
        const uint32_t grp = M13 & ins;
        const uint32_t lo = grp & lower;
        const int jprev = lo ? 31 - cw_clz(lo) : (int)lane;  // nearest inserted lane below this one in the same bucket
        const uint32_t tag_prev = TAGGED && twins ? cw_shfl(slot & TM, (unsigned)jprev) : 0u;
        cw_syncwarp();
        if (valid && ((ins >> lane) & 1u)) {
            const uint32_t prev = lo ? ((base + (uint32_t)jprev) | tag_prev) : head_old;
            next[pos & PM] = (uint16_t)prev;
            if (!(grp & ~((2u << lane) - 1u))) head[b] = (uint16_t)(TAGGED ? (pos | (slot & TM)) : pos);
        }
        cw_syncwarp();
        if (pend0 != CW_LZF_NONE || pend1 != CW_LZF_NONE) {
            if (lane == 0) {
                for (int t = 0; t < 2; t++) {
                    const uint32_t p = t ? pend1 : pend0;
                    if (p == CW_LZF_NONE) continue;
                    const uint32_t ps = cw_lzf_slot(cw_lzf_hv(s, p));
                    const uint32_t pb = ps & BM;
                    next[p & PM] = head[pb];
                    head[pb] = (uint16_t)(TAGGED ? (p | (ps & TM)) : p);
                }
            }
            cw_syncwarp();
        }
        if (qh - pub >= CW_LZ4_PUB_BATCH) {
            cw_lz4_q_publish(q, qh, lane);
            pub = qh;
        }
        base = restart ? e : base + 32u;
    }
    // liblzf's length rule can carry the last match to the very end of the input (anchor == n, 17 bits for a 64 KB
    // chunk), so the closing record keeps its anchor in the second word
    cw_lz4_q_push(q, qh, tailc, CW_LZ4_REC_END << 16, anchor | (ended_by_match ? 0x80000000u : 0u), lane, 0u, qlen);
    cw_lz4_q_publish(q, qh, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// EMITTER: pop up to 32 records of queue q, one lane each, and write their literal runs and match tokens.
// Output position: every chunk starts at op = 1 (the first run header is reserved).  A match record with `lit`
// literals before it moves op by  lit + lit/32 - [lit % 32 == 0] + (2 or 3 token bytes) + 1 (next reserved header).
// The sequential encoder gives up (returns 0) at three points; all of them are of the form "op' >= bound" with op'
// growing from record to record, so once a record fails every later one fails too: each lane checks its own record,
// writes only if it passes, and a sticky flag in the queue turns the chunk's length into 0.
// s / n: the parser's chunk slot and the chunk size; out_len: capacity the reference passes (n - 1).
// ---------------------------------------------------------------------------------------------------------------
#define CW_LZF_LIT_SELF 32u  // literal runs up to this long (one run header) are copied by the recording lane itself

CW_DEV uint32_t cw_lzf_emit_batch(CwLz4Queue *q, const uint8_t *s, uint32_t n, uint32_t out_len, uint8_t *dst_base,
                                  uint64_t stride, uint32_t *lens, uint32_t qlen = CW_LZ4_QLEN)
{
    const unsigned lane = cw_lane();
    uint32_t hd = 0, tl = 0;
    if (lane == 0) {
        hd = cw_ld_vol(&q->head);
        tl = cw_ld_vol(&q->tail);
    }
    hd = cw_shfl(hd, 0);
    tl = cw_shfl(tl, 0);
    const uint32_t avail = hd - tl;
    if (avail == 0u) return 0u;
    cw_fence_cta();
    const uint32_t m = cw_umin(avail, 32u);
    const uint64_t g = (uint64_t)q->chunk_lo | ((uint64_t)q->chunk_hi << 32);
    uint8_t *out = dst_base + g * stride;
    const uint32_t op0 = q->op ? q->op : 1u;  // 0 = first batch of a chunk
    const uint32_t failed0 = q->pad[0];
    const bool mine = lane < m;
    uint32_t anchor = 0, cur = 0, ref = 0, len = 0;
    bool by_match = false;  // closing record: the last match ran into the end of the input
    if (mine) {
        const uint32_t r0 = q->rec[(tl + lane) & (qlen - 1u)][0], r1 = q->rec[(tl + lane) & (qlen - 1u)][1];
        anchor = r0 & 0xffffu;
        cur = r0 >> 16;
        ref = r1 & 0xffffu;
        len = r1 >> 16;
        if (cur == CW_LZ4_REC_END) {
            anchor = r1 & 0x7fffffffu;
            by_match = (r1 >> 31) != 0u;
        }
    }
    const bool is_end = mine && cur == CW_LZ4_REC_END;
    const bool is_match = mine && !is_end;
    const uint32_t lit = is_match ? cur - anchor : (is_end ? n - anchor : 0u);
    const uint32_t code = len - 2u;
    const uint32_t run = lit + (lit >> 5) - ((lit & 31u) == 0u ? 1u : 0u);  // literal bytes + headers beyond the reserved one
    const uint32_t sz = is_match ? run + (code < 7u ? 2u : 3u) + 1u : 0u;
    uint32_t inc = sz;
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const uint32_t t = cw_shfl(inc, lane >= d ? lane - d : lane);
        if (lane >= d) inc += t;
    }
    const uint32_t op = op0 + inc - sz;  // output position when this record starts
    bool ok = true;
    if (is_match) {
        if (lit) {
            const uint32_t k = lit - 1u;
            if (op + k + (k >> 5) >= out_len) ok = false;  // "if (op >= out_end) return 0" at the last literal
        }
        if (op + run + 4u >= out_len) ok = false;  // room for the longest token before a match is coded
    } else if (is_end) {
        // literals the main loop still emits (positions below n - 2), then the 3-byte slack check before the tail
        const uint32_t limit = n > 2u ? n - 2u : 0u;
        const uint32_t lm = (!by_match && limit > anchor) ? limit - anchor : 0u;
        if (lm) {
            const uint32_t k = lm - 1u;
            if (op + k + (k >> 5) >= out_len) ok = false;
        }
        if (op + lm + (lm >> 5) + 3u > out_len) ok = false;
    }
    const uint32_t bad = cw_ballot(mine && !ok);
    const bool write = mine && ok && !failed0 && !(bad & ((1u << lane) - 1u));
    if (write) {
        if (lit <= CW_LZF_LIT_SELF) {
            for (uint32_t k = 0; k < lit; k++) out[op + k + (k >> 5)] = s[anchor + k];
            if (lit) out[op - 1u] = (uint8_t)(lit - 1u);
        }
        if (is_match) {
            const uint32_t p = op + run, off = cur - ref - 1u;
            if (code < 7u) {
                out[p] = (uint8_t)((off >> 8) + (code << 5));
                out[p + 1u] = (uint8_t)off;
            } else {
                out[p] = (uint8_t)((off >> 8) + (7u << 5));
                out[p + 1u] = (uint8_t)(code - 7u);
                out[p + 2u] = (uint8_t)off;
            }
        }
    }
    uint32_t big = cw_ballot(write && lit > CW_LZF_LIT_SELF);
    while (big) {
        const unsigned k = (unsigned)cw_ffs(big) - 1u;
        big &= big - 1u;
        cw_lzf_put_literals(out, cw_shfl(op, k), s, cw_shfl(anchor, k), cw_shfl(lit, k), lane);
    }
    const uint32_t endm = cw_ballot(is_end);
    const uint32_t total = cw_shfl(inc, 31u);
    if (endm) {  // the closing record is the last of its chunk: length = position after the tail literals
        const unsigned k = (unsigned)cw_ffs(endm) - 1u;
        const uint32_t fin = cw_shfl(op + run, k);
        if (lane == 0) {
            lens[g] = (failed0 || bad) ? 0u : fin;
            q->op = 0;
            q->pad[0] = 0;
        }
    } else if (lane == 0) {
        q->op = op0 + total;
        if (bad) q->pad[0] = 1u;
    }
    cw_syncwarp();
    if (lane == 0) {
        cw_fence_cta();  // literal reads of this batch are complete before the parser may reuse the slot
        cw_st_vol(&q->tail, tl + m);
    }
    return m;
}
