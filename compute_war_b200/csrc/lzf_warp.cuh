// lzf_warp.cuh -- LZF compression of ONE chunk by ONE warp, byte-identical to the reference.
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
// As in lz4_warp.cuh a warp evaluates 32 consecutive probe positions at once and then replays the
// parser's control flow over the window with an "inserted lanes" mask.
#pragma once
#include "cw_device.cuh"
#include "lz4_warp.cuh"  // cw_lz4_read32, cw_umin

#define CW_LZF_BUCKETS 8192u
#define CW_LZF_NONE 0xffffu
#define CW_LZF_MAX_OFF 8192u
#define CW_LZF_MAX_REF 264u

struct CwLzfScratch {
    uint16_t head[CW_LZF_BUCKETS];  // newest inserted position per bucket (slot & 8191), NONE if empty
    uint16_t next[CW_LZF_BUCKETS];  // ring: previous position in the same bucket, indexed by pos & 8191
};

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

// Compress chunk s[0..n) into out (capacity out_len, the reference passes n-1).  scr: shared-memory scratch,
// contents irrelevant on entry.  Returns bytes written, 0 = does not fit (same value in every lane).
CW_DEV uint32_t cw_lzf_compress_warp(const uint8_t *s, uint32_t n, CwLzfScratch *scr, uint8_t *out, uint32_t out_len)
{
    const unsigned lane = cw_lane();
    const uint32_t lower = (1u << lane) - 1u;
    if (n == 0u || out_len == 0u) return 0u;
    {
        uint32_t *h32 = reinterpret_cast<uint32_t *>(scr->head);
        for (uint32_t i = lane; i < CW_LZF_BUCKETS / 2u; i += 32u) h32[i] = 0xffffffffu;
        cw_syncwarp();
    }
    uint32_t op = 1;      // one byte reserved for the first run header
    uint32_t anchor = 0;  // first literal of the current stretch
    const uint32_t limit = n > 2u ? n - 2u : 0u;  // positions < limit are probed
    uint32_t base = 0;
    bool ended_by_match = false;
    while (base < limit) {
        // ---------------- phase A ----------------
        const uint32_t pos = base + lane;
        const bool valid = pos < limit;
        const uint32_t hv = cw_lzf_hv(s, valid ? pos : 0u);
        const uint32_t slot = cw_lzf_slot(hv);
        const uint32_t b = slot & (CW_LZF_BUCKETS - 1u);
        const uint32_t head_old = scr->head[b];
        uint32_t q = valid ? head_old : CW_LZF_NONE, hq = 0;
        while (q != CW_LZF_NONE) {
            if (pos - q > CW_LZF_MAX_OFF) {  // too far to match, and its ring entry may be stale
                q = CW_LZF_NONE;
                break;
            }
            hq = cw_lzf_hv(s, q);
            if (cw_lzf_slot(hq) == slot) break;
            q = scr->next[q & (CW_LZF_BUCKETS - 1u)];
        }
        const bool tmatch = valid && q != CW_LZF_NONE && q > 0u && (pos - q - 1u) < CW_LZF_MAX_OFF && hq == hv;
        const uint32_t own = 1u << lane;
        // lanes that share a bucket inside the window ("twins") are detected by tagging the bucket head and reading
        // it back (MATCH.ANY costs several hundred cycles on sm_100); the chain walk above is done by then, and
        // phase C rewrites every tagged head
        cw_syncwarp();
        if (valid) scr->head[b] = (uint16_t)(0xffe0u | lane);
        cw_syncwarp();
        const bool owner = valid && scr->head[b] == (0xffe0u | lane);
        uint32_t M13 = own, M16 = own, E = own;
        const uint32_t twins = cw_ballot(valid && !owner);
        if (twins) {
            M13 = cw_match_any(valid ? b : 0x10000u + lane);
            M16 = cw_match_any(valid ? slot : 0x10000u + lane);
            E = cw_match_any(valid ? hv : 0x1000000u + lane);
        }
        // ---------------- phase B ----------------
        uint32_t ins = 0, i0 = 0, pend0 = CW_LZF_NONE, pend1 = CW_LZF_NONE, e = 0;
        bool restart = false;
        for (;;) {
            const uint32_t range = ~((1u << i0) - 1u);
            const uint32_t cm = M16 & (ins | range) & lower;
            bool hit = tmatch;
            uint32_t cand = q;
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
            // ---- literals [anchor, cur): space checks of the sequential encoder, then emit ----
            const uint32_t lit = cur - anchor;
            if (lit) {
                const uint32_t k = lit - 1u;
                if (op + k + (k >> 5) >= out_len) return 0u;  // "if (op >= out_end) return 0" at the last literal
            }
            const uint32_t opn = op + lit + (lit >> 5);
            const uint32_t op2 = opn - ((lit & 31u) == 0u ? 1u : 0u);  // an empty run gives its header byte back
            if (op2 + 4u >= out_len) return 0u;
            cw_lzf_put_literals(out, op, s, anchor, lit, lane);
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
            const uint32_t code = len - 2u, off = cur - ref - 1u;
            if (lane == 0) {
                if (code < 7u) {
                    out[op2] = (uint8_t)((off >> 8) + (code << 5));
                    out[op2 + 1u] = (uint8_t)off;
                } else {
                    out[op2] = (uint8_t)((off >> 8) + (7u << 5));
                    out[op2 + 1u] = (uint8_t)(code - 7u);
                    out[op2 + 2u] = (uint8_t)off;
                }
            }
            op = op2 + (code < 7u ? 2u : 3u) + 1u;  // + the next run's reserved header
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
        if (owner) scr->head[b] = (uint16_t)head_old;
        cw_syncwarp();
        if (valid && ((ins >> lane) & 1u)) {
            const uint32_t grp = M13 & ins;
            const uint32_t lo = grp & lower;
            scr->next[pos & (CW_LZF_BUCKETS - 1u)] = (uint16_t)(lo ? base + (uint32_t)(31 - cw_clz(lo)) : head_old);
            if (!(grp & ~((2u << lane) - 1u))) scr->head[b] = (uint16_t)pos;
        }
        cw_syncwarp();
        if (pend0 != CW_LZF_NONE || pend1 != CW_LZF_NONE) {
            if (lane == 0) {
                for (int t = 0; t < 2; t++) {
                    const uint32_t p = t ? pend1 : pend0;
                    if (p == CW_LZF_NONE) continue;
                    const uint32_t pb = cw_lzf_slot(cw_lzf_hv(s, p)) & (CW_LZF_BUCKETS - 1u);
                    scr->next[p & (CW_LZF_BUCKETS - 1u)] = scr->head[pb];
                    scr->head[pb] = (uint16_t)p;
                }
            }
            cw_syncwarp();
        }
        base = restart ? e : base + 32u;
    }
    // ---- end of input: remaining literals of the main loop, the 3-byte slack check, then the tail ----
    {
        const uint32_t lm = (!ended_by_match && limit > anchor) ? limit - anchor : 0u;  // literals emitted by the loop
        if (lm) {
            const uint32_t k = lm - 1u;
            if (op + k + (k >> 5) >= out_len) return 0u;
        }
        if (op + lm + (lm >> 5) + 3u > out_len) return 0u;
        const uint32_t lit = n - anchor;
        cw_lzf_put_literals(out, op, s, anchor, lit, lane);
        op = op + lit + (lit >> 5);
        if ((lit & 31u) == 0u) op -= 1u;
    }
    return op;
}
