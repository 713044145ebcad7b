// lz4_warp.cuh -- LZ4 block compression of ONE chunk by ONE warp.
//
// Replaces the lz4 lambda  LZ4_compress_default(s, d, l, 2*l)  (HashAndCompress.cpp:349-355; also
// experiment.cpp:249).  The reference links a prebuilt liblz4.a 1.8.2 whose source is not in its tree
// (header: src/compression_perf/include/lz4/lz4.h:94-96,126-156); the parse reproduced here is that
// library's greedy single-pass parse for inputs < 65547 bytes (u16 position table, acceleration 1) --
// SURVEY.md Appendix A.2.
//   HASHLOG = 13 ("compat"): the reference's own table size; output byte-identical to the vendored archive
//             (accepted only through that byte-equality: tests/golden, oracle/).
//   HASHLOG = 12 ("fast")  : the same parse over a 4096-entry table.  Still a valid LZ4 block that
//             LZ4_decompress_safe restores bit-exactly; 0.1-0.7 % larger on the reference's corpora at 8 KB
//             chunks, and the 8 KB it saves per chunk lets 14 instead of 9 chunks be resident
//             per SM.
//
// The sequential parser probes one position at a time: h = hash(read32(p)); cand = table[h];
// table[h] = p; match iff read32(cand) == read32(p).  Here a warp evaluates the next 32 probe positions
// of the parser at once ("window"):
//   phase A (all lanes, independent): read32, hash, table look-up, candidate compare; lanes that share a
//            table slot inside the window ("twins") are detected by writing a lane tag into the slot and
//            reading it back; the groups are then resolved with two ballots per group of equal values
//            (MATCH.ANY costs several hundred cycles on sm_100 and is kept for windows with many groups).
//   phase B (uniform, registers only): replay the parser's control flow over the window.  `ins` is the
//            set of window lanes whose position the sequential parser has inserted so far; a lane's true
//            candidate is the highest inserted twin below it, else the table entry read in phase A.  The
//            first hitting lane is the parser's next match; the warp extends it (28 lanes x 4 bytes
//            forward), the match is pushed to the emitter queue, the parser's post-match insert of position e-2 is
//            added to `ins`, and the replay resumes at lane e (positions inside a match are never probed,
//            exactly as in the reference).
//   phase C: every lane that owns its slot writes the slot's final value: its position if the parser
//            inserted it, else the old entry.
// The parser does only what the next decision depends on.  Each match is pushed as an 8-byte record
// (anchor, cur, cand, forward length) into a shared-memory ring; an EMITTER warp of the same CTA pops
// records 32 at a time, one lane per record, and does everything the parse does not depend on: the
// backward extension (it changes literal and match lengths but not where the parser continues), the
// sequence sizes and their prefix sum, token / length bytes / offset, and the literal copies (long runs
// by the whole warp with 32-bit stores).  The compressors are bound by the length of the parser's
// dependent instruction chain, not by issue slots, so moving ~60 % of the per-match instructions to
// another warp shortens the chain by that much.
// After 64 failed probes the reference widens its stride (step = searchNb++ >> 6); lane positions follow
// the closed form of that schedule, so incompressible data is skipped as quickly as the CPU does.
//
// Memory: the chunk (followed by >= CW_LZ4_PAD readable bytes) and the table live in shared memory, staged
// by the caller; output bytes go straight to global memory.
#pragma once
#include "cw_device.cuh"

#define CW_LZ4_NONE 0xffffffffu
#define CW_LZ4_PAD 144u       // readable bytes after the chunk: the forward compare reads up to 131 bytes past n-12
#ifndef CW_LZ4_SPEC_STEPS
#define CW_LZ4_SPEC_STEPS 3     // per-lane speculative forward extension in phase A: 4 bytes per step
#endif
#define CW_LZ4_LIT_SELF 16u   // literal runs up to this long are copied by the recording lane itself
#define CW_LZ4_TAG 0xffe0u    // slot tag | lane, used for twin detection.  Tags never outlive a window (phase C rewrites
                              // every tagged slot) and are only compared with tags, so they may overlap position values

CW_DEV uint32_t cw_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// unaligned little-endian 32-bit read from the staged chunk (base 4-byte aligned, CW_LZ4_PAD readable bytes behind it)
CW_DEV uint32_t cw_lz4_read32(const uint8_t *s, uint32_t pos)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
    const uint32_t i = pos >> 2;
    return cw_funnel_r(w[i], w[i + 1], (pos & 3u) * 8u);
}
// Source access of the compressor.  DIRECT = false: the chunk sits in shared memory with padding behind it.
// DIRECT = true: the chunk is read where it lies in global memory (chunks too big to keep several of them in
// shared memory); words at or beyond nw = n/4 are not touched and read as 0 -- every use of such bytes is masked
// by the match limit anyway.
template <bool DIRECT>
CW_DEV uint32_t cw_lz4_src32(const uint8_t *s, uint32_t pos, uint32_t nw)
{
    if (!DIRECT) return cw_lz4_read32(s, pos);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
    const uint32_t i = pos >> 2;
    const uint32_t lo = i < nw ? cw_ldg32(w + i) : 0u, hi = i + 1u < nw ? cw_ldg32(w + i + 1u) : 0u;
    return cw_funnel_r(lo, hi, (pos & 3u) * 8u);
}
template <bool DIRECT>
CW_DEV uint32_t cw_lz4_src8(const uint8_t *s, uint32_t pos)
{
    return DIRECT ? cw_ldg8(s + pos) : s[pos];
}
template <int HASHLOG>
CW_DEV uint32_t cw_lz4_hash(uint32_t v)
{
    return (v * 2654435761u) >> (32 - HASHLOG);
}

// probe schedule of one search: probe j sits S(j) bytes after the search start and
// the next probe is step(j) further.  j = -1 is the immediate re-test right after a match.
CW_DEV uint32_t cw_lz4_F(uint32_t m)
{
    const uint32_t q = m >> 6, r = m & 63u;
    return 32u * q * (q - 1u) + q * r;  // sum_{x<m} (x >> 6)
}
CW_DEV int32_t cw_lz4_S(int32_t j) { return j <= 0 ? j : 1 + (int32_t)cw_lz4_F(63u + (uint32_t)j); }
CW_DEV uint32_t cw_lz4_step(int32_t j) { return j <= 0 ? 1u : (63u + (uint32_t)j) >> 6; }

// 255-run length extension of a literal / match length, written by ONE lane: bytes for value r (>= 0)
CW_DEV uint32_t cw_lz4_put_ext1(uint8_t *out, uint32_t op, uint32_t r)
{
#pragma unroll 1
    while (r >= 255u) {
        out[op++] = 255u;
        r -= 255u;
    }
    out[op++] = (uint8_t)r;
    return op;
}
CW_DEV uint32_t cw_lz4_ext_bytes(uint32_t len) { return len >= 15u ? (len - 15u) / 255u + 1u : 0u; }

// whole-warp copy of len literal bytes s[src..] -> out[dst..]; 32-bit stores on the aligned body
template <bool DIRECT>
CW_DEV void cw_lz4_copy_warp(uint8_t *out, uint32_t dst, const uint8_t *s, uint32_t src, uint32_t len, uint32_t nw,
                             unsigned lane)
{
    if (len < 64u) {
        for (uint32_t x = lane; x < len; x += 32u) out[dst + x] = cw_lz4_src8<DIRECT>(s, src + x);
        return;
    }
    const uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(out + dst) & 3u)) & 3u;
    if (lane < head) out[dst + lane] = cw_lz4_src8<DIRECT>(s, src + lane);
    const uint32_t words = (len - head) >> 2;
    uint32_t *o32 = reinterpret_cast<uint32_t *>(out + dst + head);
    for (uint32_t w = lane; w < words; w += 32u) o32[w] = cw_lz4_src32<DIRECT>(s, src + head + 4u * w, nw);
    const uint32_t done = head + 4u * words;
    if (lane < len - done) out[dst + done + lane] = cw_lz4_src8<DIRECT>(s, src + done + lane);
}

#define CW_LZ4_QLEN 64u  // records per parser ring (power of two); kernels short of shared memory run rings of 32:
                         // the functions below take the ring length, and a queue then occupies cw_lz4_queue_bytes(qlen)
#ifndef CW_LZ4_PUB_BATCH
#define CW_LZ4_PUB_BATCH 16u // the parser makes its records visible to the emitter in batches of this many
#endif
#ifndef CW_LZ4_CHAIN
#define CW_LZ4_CHAIN 0  // batched replay of twin-free unit windows (cw_lz4_parse_warp, phase B): +4-9 % on real text, but the
                        // bigger loop costs 5 % on the synthetic mix (profiles/r01_ab_lz4_speculation.txt); off by default
#endif
#define CW_LZ4_REC_END 0xffffu  // cur field of the record that closes a chunk (final literals from anchor)

// parser -> emitter ring, one per parser warp, in shared memory
struct CwLz4Queue {
    uint32_t head;               // records published by the parser (monotonic)
    uint32_t tail;               // records consumed by the emitter (monotonic)
    uint32_t chunk_lo, chunk_hi; // global index of the chunk the pending records belong to
    uint32_t op;                 // emitter: bytes of that chunk written so far
    uint32_t done;               // parser has exited
    uint32_t pad[2];
    uint32_t rec[CW_LZ4_QLEN][2];  // {anchor | cur << 16, cand | fwd << 16}; only the first qlen entries exist in a short ring
};
CW_HOSTDEV uint32_t cw_lz4_queue_bytes(uint32_t qlen) { return 32u + 8u * qlen; }
// queue p of an array of rings of qlen records each
CW_DEV CwLz4Queue *cw_lz4_queue_at(CwLz4Queue *base, uint32_t p, uint32_t qlen)
{
    return reinterpret_cast<CwLz4Queue *>(reinterpret_cast<uint8_t *>(base) + (size_t)p * cw_lz4_queue_bytes(qlen));
}

// parser side: push one record; lane `writer` holds the values (qh = records pushed so far, tailc = cached tail)
CW_DEV void cw_lz4_q_push(CwLz4Queue *q, uint32_t &qh, uint32_t &tailc, uint32_t a, uint32_t b, unsigned lane,
                          unsigned writer, uint32_t qlen = CW_LZ4_QLEN)
{
    if (qh - tailc >= qlen) {  // ring looks full: publish what we have and wait for the emitter
        cw_syncwarp();
        if (lane == 0) {
            cw_fence_cta();
            cw_st_vol(&q->head, qh);
        }
        for (;;) {
            uint32_t t = 0;
            if (lane == 0) t = cw_ld_vol(&q->tail);
            tailc = cw_shfl(t, 0);
            if (qh - tailc < qlen) break;
            cw_spin();
        }
    }
    if (lane == writer) {
        q->rec[qh & (qlen - 1u)][0] = a;
        q->rec[qh & (qlen - 1u)][1] = b;
    }
    qh++;
}
// parser side: make room for cnt (<= 32) records that their lanes are about to write themselves
CW_DEV void cw_lz4_q_reserve(CwLz4Queue *q, uint32_t qh, uint32_t &tailc, uint32_t cnt, unsigned lane,
                             uint32_t qlen = CW_LZ4_QLEN)
{
    if (qh + cnt - tailc <= qlen) return;
    cw_syncwarp();
    if (lane == 0) {
        cw_fence_cta();
        cw_st_vol(&q->head, qh);
    }
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = cw_ld_vol(&q->tail);
        tailc = cw_shfl(t, 0);
        if (qh + cnt - tailc <= qlen) break;
        cw_spin();
    }
}
CW_DEV void cw_lz4_q_publish(CwLz4Queue *q, uint32_t qh, unsigned lane)
{
    cw_syncwarp();
    if (lane == 0) {
        cw_fence_cta();
        cw_st_vol(&q->head, qh);
    }
}
// parser side: wait until the emitter has consumed everything (before the chunk slot is reused)
CW_DEV void cw_lz4_q_drain(CwLz4Queue *q, uint32_t qh, unsigned lane)
{
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = cw_ld_vol(&q->tail);
        if (cw_shfl(t, 0) == qh) break;
        cw_spin();
    }
    cw_fence_cta();
}

// EMITTER: pop up to 32 records of queue q (one lane each) and write their sequences.  s / n: the parser's
// chunk slot and chunk size; block of chunk g goes to dst_base + g * stride, its length to lens[g].
// Returns the number of records consumed (same value in every lane).
template <bool DIRECT>
CW_DEV uint32_t cw_lz4_emit_batch(CwLz4Queue *q, const uint8_t *s, uint32_t n, uint8_t *dst_base, uint64_t stride,
                                  uint32_t *lens, uint32_t qlen = CW_LZ4_QLEN)
{
    const uint32_t nw = n >> 2;
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
    cw_fence_cta();  // records were written before head was published
    const uint32_t m = cw_umin(avail, 32u);
    const uint64_t g = (uint64_t)q->chunk_lo | ((uint64_t)q->chunk_hi << 32);
    uint8_t *out = dst_base + g * stride;
    if (DIRECT) s += g * (uint64_t)n;  // DIRECT: s is the base of the whole input buffer
    const uint32_t op0 = q->op;
    const bool mine = lane < m;
    uint32_t anchor = 0, cur = 0, cnd = 0, fwd = 0;
    if (mine) {
        const uint32_t r0 = q->rec[(tl + lane) & (qlen - 1u)][0], r1 = q->rec[(tl + lane) & (qlen - 1u)][1];
        anchor = r0 & 0xffffu;
        cur = r0 >> 16;
        cnd = r1 & 0xffffu;
        fwd = r1 >> 16;
    }
    const bool is_end = mine && cur == CW_LZ4_REC_END;
    // backward extension ("catch up"): per lane, typically 0-2 bytes
    uint32_t back = 0;
    if (mine && !is_end) {
        const uint32_t maxb = cw_umin(cur - anchor, cnd);
        while (back < maxb && cw_lz4_src8<DIRECT>(s, cur - 1u - back) == cw_lz4_src8<DIRECT>(s, cnd - 1u - back)) back++;
    }
    const uint32_t lit = mine ? (is_end ? n - anchor : cur - back - anchor) : 0u;
    const uint32_t mcode = back + fwd;  // match length - 4
    const uint32_t sz = mine ? 1u + cw_lz4_ext_bytes(lit) + lit + (is_end ? 0u : 2u + cw_lz4_ext_bytes(mcode)) : 0u;
    uint32_t inc = sz;  // inclusive prefix sum over the lanes
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const uint32_t t = cw_shfl(inc, lane >= d ? lane - d : lane);
        if (lane >= d) inc += t;
    }
    const uint32_t total = cw_shfl(inc, 31u);
    const uint32_t p0 = op0 + inc - sz;
    if (mine) {
        uint32_t p = p0;
        out[p++] = (uint8_t)((cw_umin(lit, 15u) << 4) | (is_end ? 0u : cw_umin(mcode, 15u)));
        if (lit >= 15u) p = cw_lz4_put_ext1(out, p, lit - 15u);
        if (lit <= CW_LZ4_LIT_SELF)
            for (uint32_t x = 0; x < lit; x++) out[p + x] = cw_lz4_src8<DIRECT>(s, anchor + x);
        p += lit;
        if (!is_end) {
            const uint32_t off = cur - cnd;
            out[p] = (uint8_t)off;
            out[p + 1u] = (uint8_t)(off >> 8);
            if (mcode >= 15u) cw_lz4_put_ext1(out, p + 2u, mcode - 15u);
        }
    }
    uint32_t big = cw_ballot(mine && lit > CW_LZ4_LIT_SELF);
    while (big) {
        const unsigned k = (unsigned)cw_ffs(big) - 1u;
        big &= big - 1u;
        const uint32_t l = cw_shfl(lit, k);
        cw_lz4_copy_warp<DIRECT>(out, cw_shfl(p0, k) + 1u + cw_lz4_ext_bytes(l), s, cw_shfl(anchor, k), l, nw, lane);
    }
    const uint32_t endm = cw_ballot(is_end);
    if (lane == 0) {
        if (endm) {
            lens[g] = op0 + total;
            q->op = 0;
        } else
            q->op = op0 + total;
    }
    cw_syncwarp();
    if (lane == 0) {
        cw_fence_cta();  // literal reads of this batch are complete before the parser may reuse the slot
        cw_st_vol(&q->tail, tl + m);
    }
    return m;
}

// PARSER: parse chunk s[0..n) (n <= 65536) and push its matches, then the end-of-chunk record, to q.
// table: zeroed u16[1 << HASHLOG] in shared memory.  qh: records pushed so far on this queue (kept by the
// caller across chunks).  The caller sets q->chunk_lo/hi before and drains the queue before reusing s.
template <int HASHLOG, bool DIRECT>
CW_DEV void cw_lz4_parse_warp(const uint8_t *s, uint32_t n, uint16_t *table, CwLz4Queue *q, uint32_t &qh,
                              uint32_t qlen = CW_LZ4_QLEN)
{
    const uint32_t nw = n >> 2;
    const unsigned lane = cw_lane();
    const uint32_t own = 1u << lane;
    const uint32_t lower = own - 1u;
    uint32_t anchor = 0, tailc = qh;  // the caller drained the queue: tail == qh
    uint32_t pub = qh;                // records published so far

    if (n >= 13u) {
        const uint32_t mfl1 = n - 11u;       // mflimit + 1: a probe happens only if the NEXT probe position is <= this
        const uint32_t matchlimit = n - 5u;  // matches never cover the last 5 bytes
        uint32_t sbase = 1;                  // first position of the current search
        int32_t j0 = 0;                      // search index of lane 0 in the next window
        bool done = false;
        bool spec_on = true;  // speculate only while matches are short (text-like data); long matches go to the warp-wide path
        while (!done) {
            // ---------------- phase A ----------------
            const bool unit = (j0 + 31 <= 64);  // all 32 probes one byte apart
            uint32_t base, pos;
            bool valid;
            if (unit) {
                base = sbase + (uint32_t)j0;  // S(j) = j while the steps are 1
                pos = base + lane;
                valid = pos < mfl1;
            } else {
                const int32_t j = j0 + (int32_t)lane;
                base = sbase + (uint32_t)cw_lz4_S(j0);
                pos = sbase + (uint32_t)cw_lz4_S(j);
                valid = pos + cw_lz4_step(j) <= mfl1;
            }
            const uint32_t v = cw_lz4_src32<DIRECT>(s, valid ? pos : 0u, nw);
            const uint32_t h = cw_lz4_hash<HASHLOG>(v);
            const uint32_t t = table[h];
            // twin detection: tag the slot, read it back; the lane whose tag survives owns the slot
            uint32_t M = own;
            bool owner = valid;
            uint32_t twins;
            cw_syncwarp();  // every lane has read its old entry before any lane tags a slot
            if (valid) table[h] = (uint16_t)(CW_LZ4_TAG | lane);
            cw_syncwarp();
            owner = valid && table[h] == (CW_LZ4_TAG | lane);
            twins = cw_ballot(valid && !owner);
            const bool tmatch = valid && cw_lz4_src32<DIRECT>(s, t, nw) == v;
            // speculative forward extension against the table candidate, every matching lane for itself:
            // spec = equal bytes after the 4-byte match; bit 15 = "still equal after CW_LZ4_SPEC_STEPS steps"
            const uint32_t tm_all = cw_ballot(tmatch);
            uint32_t spec = 0x80000000u;
            if (CW_LZ4_SPEC_STEPS > 0 && spec_on && tm_all) {
                uint32_t fw = 0;
                bool open = tmatch;
                const uint32_t first = tm_all & (0u - tm_all);  // the lane the replay will look at first
#pragma unroll
                for (int k = 0; k < CW_LZ4_SPEC_STEPS; k++) {
                    if (open) {
                        const uint32_t pa = pos + 4u + fw;
                        const uint32_t d = cw_lz4_src32<DIRECT>(s, pa, nw) ^ cw_lz4_src32<DIRECT>(s, t + 4u + fw, nw);
                        uint32_t eq = d ? (uint32_t)(cw_ffs(d) - 1) >> 3 : 4u;
                        eq = cw_umin(eq, pa < matchlimit ? matchlimit - pa : 0u);
                        fw += eq;
                        open = eq == 4u;
                    }
                    const uint32_t openm = cw_ballot(open);
                    // stop when nobody is left, or when the first match is a long one: the whole warp will
                    // extend it anyway and most of the window lies inside it
                    if (!openm || (openm & first)) break;
                }
                if (tmatch) spec = fw | (open ? 0x80000000u : 0u);
            }
            const uint32_t inv_all = cw_ballot(!valid);
            // valid lanes form a prefix of the window (positions increase with the lane)
            const uint32_t vmask = inv_all ? (1u << (cw_ffs(inv_all) - 1)) - 1u : 0xffffffffu;
            uint32_t E = own, dyn_all = 0;
            if (twins) {
                // every slot shared by several lanes has at least one lane that lost the tag race: walk those, one
                // group of equal 4-byte values per step (two ballots; MATCH.ANY costs ~365 cycles on sm_100)
                uint32_t rem = twins;
                if (cw_popc(twins) > 6) {  // many groups: the fixed cost of two MATCH.ANY is lower
                    M = cw_match_any(valid ? h : 0x10000u + lane);
                    E = cw_match_any(v);
                    rem = 0;
                }
                while (rem) {
                    const unsigned L = (unsigned)cw_ffs(rem) - 1u;
                    const uint32_t hL = cw_shfl(h, L), vL = cw_shfl(v, L);
                    const bool sh = valid && h == hL, sv = valid && v == vL;
                    const uint32_t mh = cw_ballot(sh), mv = cw_ballot(sv);
                    if (sh) M = mh;
                    if (sv) E = mv;
                    rem &= ~mv;
                }
                dyn_all = cw_ballot(valid && (M & lower) != 0u);  // lanes that may have an in-window candidate
            }
            // ---------------- phase B ----------------
            uint32_t ins = 0, i0 = 0, pend = CW_LZ4_NONE;
            bool restart = false;
            if (CW_LZ4_CHAIN && spec_on && unit && !twins && tm_all) {
                // Twin-free unit window: every candidate is the table entry read in phase A, and a lane whose
                // speculation closed already knows where its match ends.  The parser's walk from match to match is then
                // one shuffle per match, and the records are written by their own lanes, all at once.  The walk stops at
                // a match that is still open (the general loop below extends it with the whole warp) and wherever the
                // general loop would leave the window.
                const uint32_t closed_all = cw_ballot(tmatch && !(spec & 0x80000000u));
                const uint32_t my_de = lane + 4u + (spec & 0x7fffffffu);  // lane index of the end of this lane's match
                const uint32_t anchor0 = anchor;
                uint32_t chain = 0;
                for (;;) {
                    const uint32_t range = ~((1u << i0) - 1u);
                    const uint32_t hm = tm_all & range;
                    if (!hm) break;
                    const unsigned f = (unsigned)cw_ffs(hm) - 1u;
                    if (!((closed_all >> f) & 1u)) break;
                    chain |= 1u << f;
                    ins |= range & ((2u << f) - 1u);
                    const uint32_t de = cw_shfl(my_de, f);
                    const uint32_t e = base + de;
                    spec_on = de - f - 4u < 28u;
                    anchor = e;
                    if (e >= mfl1) {
                        done = true;
                        break;
                    }
                    sbase = e + 1u;
                    const bool in2 = de < 34u;
                    ins |= in2 ? 1u << ((de - 2u) & 31u) : 0u;
                    pend = in2 ? pend : e - 2u;
                    i0 = de;
                    if (de < 32u) continue;
                    restart = true;
                    break;
                }
                if (chain) {
                    const uint32_t cnt = (uint32_t)cw_popc(chain), below = chain & lower;
                    cw_lz4_q_reserve(q, qh, tailc, cnt, lane, qlen);
                    // a record's anchor is the end of the chain match before it (the carried-in anchor for the first)
                    const uint32_t prev_de = cw_shfl(my_de, below ? (unsigned)(31 - cw_clz(below)) : lane);
                    if ((chain >> lane) & 1u) {
                        const uint32_t k = (qh + (uint32_t)cw_popc(below)) & (qlen - 1u);
                        q->rec[k][0] = (below ? base + prev_de : anchor0) | (pos << 16);
                        q->rec[k][1] = t | ((spec & 0x7fffffffu) << 16);
                    }
                    qh += cnt;
                }
            }
            if (!restart && !done) for (;;) {
                const uint32_t range = ~((1u << i0) - 1u);
                uint32_t hitm = tm_all & range;
                uint32_t cand = t, sp = spec;
                if (dyn_all & range) {
                    const uint32_t cm = M & (ins | range) & lower;
                    bool hit = tmatch;
                    if (cm) {
                        const int jl = 31 - cw_clz(cm);
                        hit = valid && ((E >> jl) & 1u);
                        cand = unit ? base + (uint32_t)jl : sbase + (uint32_t)cw_lz4_S(j0 + jl);
                        sp = 0x80000000u;  // candidate inside the window: nothing speculated
                    }
                    hitm = cw_ballot(hit) & range;
                }
                if (!hitm) {  // no (further) match in this window: the rest of its valid lanes were probed;
                    ins |= range & vmask;  // an invalid lane in range is the parser's end-of-input rule
                    done = (range & ~vmask) != 0u;
                    break;
                }
                const unsigned f = (unsigned)cw_ffs(hitm) - 1u;
                ins |= range & ((2u << f) - 1u);
                const uint32_t cur = unit ? base + f : cw_shfl(pos, f);
                sp = cw_shfl(sp, f);
                uint32_t fwd = sp & 0x7fffffffu;
                if (sp & 0x80000000u) {
                    // ---- forward extension by the whole warp: 32 lanes x 4 bytes per step, continuing at fwd ----
                    const uint32_t cnd = cw_shfl(cand, f);
                    for (;;) {
                        const uint32_t pa = cur + 4u + fwd + 4u * lane, pb = cnd + 4u + fwd + 4u * lane;
                        const uint32_t d = cw_lz4_src32<DIRECT>(s, pa, nw) ^ cw_lz4_src32<DIRECT>(s, pb, nw);
                        uint32_t eq = d ? (uint32_t)(cw_ffs(d) - 1) >> 3 : 4u;  // equal leading bytes
                        eq = cw_umin(eq, pa < matchlimit ? matchlimit - pa : 0u);
                        const uint32_t stop = cw_ballot(eq < 4u);
                        if (stop) {
                            const unsigned L = (unsigned)cw_ffs(stop) - 1u;
                            fwd += 4u * L + cw_shfl(eq, L);
                            break;
                        }
                        fwd += 128u;
                    }
                }
                const uint32_t e = cur + 4u + fwd;  // end of the match = next parser position
                spec_on = fwd < 28u;
                // everything else about this sequence (backward extension, lengths, bytes) is the emitter's job;
                // the matching lane itself holds the candidate, so it writes the record
                cw_lz4_q_push(q, qh, tailc, anchor | (cur << 16), cand | (fwd << 16), lane, f, qlen);
                anchor = e;
                if (e >= mfl1) {  // "if (ip > mflimit) break" -> last literals
                    done = true;
                    break;
                }
                // parser: table[hash(e-2)] = e-2, then immediate re-test at e, then a fresh search at e+1
                sbase = e + 1u;
                const uint32_t de = e - base;  // lane index of the next parser position (unit windows)
                const bool in2 = unit && de < 34u, in0 = unit && de < 32u;
                ins |= in2 ? 1u << ((de - 2u) & 31u) : 0u;
                pend = in2 ? pend : e - 2u;
                i0 = de;
                if (in0) continue;
                restart = true;
                break;
            }
            if (qh - pub >= cw_umin(CW_LZ4_PUB_BATCH, qlen >> 1)) {  // publishing costs a fence: not after every window (a ring of 16: every 8)
                cw_lz4_q_publish(q, qh, lane);
                pub = qh;
            }
            if (done) break;
            // ---------------- phase C: every touched slot gets its final value ----------------
            if (twins) {
                // slot owners first put the old entry back, then the highest inserted twin of each slot writes
                if (owner) table[h] = (uint16_t)t;
                cw_syncwarp();
                if (((ins >> lane) & 1u) && !(M & ins & ~((2u << lane) - 1u))) table[h] = (uint16_t)pos;
            } else {
                if (valid) table[h] = (uint16_t)(((ins >> lane) & 1u) ? pos : t);
            }
            cw_syncwarp();
            if (pend != CW_LZ4_NONE) {
                if (lane == 0) table[cw_lz4_hash<HASHLOG>(cw_lz4_src32<DIRECT>(s, pend, nw))] = (uint16_t)pend;
                cw_syncwarp();
            }
            if (restart) j0 = -1;                                       // lane 0 = re-test at e
            else if (unit) j0 = (int32_t)(base + 32u) - (int32_t)sbase;  // S(j) = j while steps are 1
            else j0 += 32;
        }
    }
    // ---- end of chunk: the final literal run [anchor, n) ----
    cw_lz4_q_push(q, qh, tailc, anchor | (CW_LZ4_REC_END << 16), 0u, lane, 0u, qlen);
    cw_lz4_q_publish(q, qh, lane);
}
