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
//            reading it back -- MATCH.ANY, which costs several hundred cycles on sm_100, is only issued
//            for windows that really contain twins.
//   phase B (uniform, registers only): replay the parser's control flow over the window.  `ins` is the
//            set of window lanes whose position the sequential parser has inserted so far; a lane's true
//            candidate is the highest inserted twin below it, else the table entry read in phase A.  The
//            first hitting lane is the parser's next match; the warp extends it (28 lanes x 4 bytes
//            forward and 4 lanes x 1 byte backward in one step), the sequence is *recorded* (lane k
//            keeps the k-th sequence of the window), the parser's post-match insert of position e-2 is
//            added to `ins`, and the replay resumes at lane e (positions inside a match are never probed,
//            exactly as in the reference).
//   phase C: every lane that owns its slot writes the slot's final value: its position if the parser
//            inserted it, else the old entry.
//   phase D: the recorded sequences are written out in parallel, one lane per sequence; long literal
//            runs are copied by the whole warp with 32-bit stores.
// After 64 failed probes the reference widens its stride (step = searchNb++ >> 6); lane positions follow
// the closed form of that schedule, so incompressible data is skipped as quickly as the CPU does.
//
// Memory: the chunk (followed by >= CW_LZ4_PAD readable bytes) and the table live in shared memory, staged
// by the caller; output bytes go straight to global memory.
#pragma once
#include "cw_device.cuh"

#define CW_LZ4_NONE 0xffffffffu
#define CW_LZ4_PAD 144u       // readable bytes after the chunk: the forward compare reads up to 131 bytes past n-12
#define CW_LZ4_LIT_SELF 16u   // literal runs up to this long are copied by the recording lane itself
#define CW_LZ4_FW_LANES 28u   // lanes 0..27 compare forward, lanes 28..31 backward
#define CW_LZ4_TAG 0x8000u    // slot tag used for twin detection (chunks <= 32768 bytes: positions < 0x8000)

CW_DEV uint32_t cw_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// unaligned little-endian 32-bit read from the staged chunk (base 4-byte aligned)
CW_DEV uint32_t cw_lz4_read32(const uint8_t *s, uint32_t pos)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
    const uint32_t i = pos >> 2;
    return cw_funnel_r(w[i], w[i + 1], (pos & 3u) * 8u);
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
    while (r >= 255u) {
        out[op++] = 255u;
        r -= 255u;
    }
    out[op++] = (uint8_t)r;
    return op;
}
CW_DEV uint32_t cw_lz4_ext_bytes(uint32_t len) { return len >= 15u ? (len - 15u) / 255u + 1u : 0u; }

// whole-warp copy of len literal bytes s[src..] -> out[dst..]; 32-bit stores on the aligned body
CW_DEV void cw_lz4_copy_warp(uint8_t *out, uint32_t dst, const uint8_t *s, uint32_t src, uint32_t len, unsigned lane)
{
    if (len < 64u) {
        for (uint32_t x = lane; x < len; x += 32u) out[dst + x] = s[src + x];
        return;
    }
    const uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(out + dst) & 3u)) & 3u;
    if (lane < head) out[dst + lane] = s[src + lane];
    const uint32_t words = (len - head) >> 2;
    uint32_t *o32 = reinterpret_cast<uint32_t *>(out + dst + head);
    for (uint32_t w = lane; w < words; w += 32u) o32[w] = cw_lz4_read32(s, src + head + 4u * w);
    const uint32_t done = head + 4u * words;
    if (lane < len - done) out[dst + done + lane] = s[src + done + lane];
}

// Compress chunk s[0..n) (n <= 65536).  table: zeroed u16[1 << HASHLOG] in shared memory.  Returns the
// number of bytes written to out (same value in every lane).  out must hold LZ4_COMPRESSBOUND(n) bytes.
template <int HASHLOG>
CW_DEV uint32_t cw_lz4_compress_warp(const uint8_t *s, uint32_t n, uint16_t *table, uint8_t *out)
{
    const unsigned lane = cw_lane();
    const uint32_t own = 1u << lane;
    const uint32_t lower = own - 1u;
    const bool tag_ok = n <= 32768u;  // positions fit below CW_LZ4_TAG
    uint32_t op = 0, anchor = 0;

    if (n >= 13u) {
        const uint32_t mfl1 = n - 11u;       // mflimit + 1: a probe happens only if the NEXT probe position is <= this
        const uint32_t matchlimit = n - 5u;  // matches never cover the last 5 bytes
        uint32_t sbase = 1;                  // first position of the current search
        int32_t j0 = 0;                      // search index of lane 0 in the next window
        bool done = false;
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
            const uint32_t v = cw_lz4_read32(s, valid ? pos : 0u);
            const uint32_t h = cw_lz4_hash<HASHLOG>(v);
            const uint32_t t = table[h];
            // twin detection: tag the slot, read it back; the lane whose tag survives owns the slot
            uint32_t M = own;
            bool owner = valid;
            uint32_t twins;
            if (tag_ok) {
                cw_syncwarp();  // every lane has read its old entry before any lane tags a slot
                if (valid) table[h] = (uint16_t)(CW_LZ4_TAG | lane);
                cw_syncwarp();
                owner = valid && table[h] == (CW_LZ4_TAG | lane);
                twins = cw_ballot(valid && !owner);
            } else {
                M = cw_match_any(valid ? h : 0x10000u + lane);
                twins = cw_ballot(M != own);
                owner = valid && !(M & ~((2u << lane) - 1u));  // highest lane of the group
            }
            const bool tmatch = valid && cw_lz4_read32(s, t) == v;
            const uint32_t tm_all = cw_ballot(tmatch);
            const uint32_t inv_all = cw_ballot(!valid);
            uint32_t E = own, dyn_all = 0;
            if (twins) {
                if (tag_ok) M = cw_match_any(valid ? h : 0x10000u + lane);
                E = cw_match_any(v);
                dyn_all = cw_ballot(valid && (M & lower) != 0u);  // lanes that may have an in-window candidate
            }
            // ---------------- phase B ----------------
            uint32_t ins = 0, i0 = 0, pend = CW_LZ4_NONE, nrec = 0;
            uint32_t r_src = 0, r_lit = 0, r_off = 0, r_mcode = 0, r_op = 0;  // lane k: k-th sequence of this window
            bool restart = false;
            for (;;) {
                const uint32_t range = ~((1u << i0) - 1u);
                uint32_t hitm = tm_all & range;
                uint32_t cand = t;
                if (dyn_all & range) {
                    const uint32_t cm = M & (ins | range) & lower;
                    bool hit = tmatch;
                    if (cm) {
                        const int jl = 31 - cw_clz(cm);
                        hit = valid && ((E >> jl) & 1u);
                        cand = unit ? base + (uint32_t)jl : sbase + (uint32_t)cw_lz4_S(j0 + jl);
                    }
                    hitm = cw_ballot(hit) & range;
                }
                const uint32_t invm = inv_all & range;
                const int f = cw_ffs(hitm) - 1, iv = cw_ffs(invm) - 1;
                if (iv >= 0 && (f < 0 || iv < f)) {  // the parser's end-of-input rule fires first
                    done = true;
                    break;
                }
                if (f < 0) {  // rest of the window probed, no match
                    ins |= range;
                    break;
                }
                ins |= range & ((2u << f) - 1u);
                const uint32_t cur = unit ? base + (uint32_t)f : cw_shfl(pos, (unsigned)f);
                const uint32_t cnd = cw_shfl(cand, (unsigned)f);
                // ---- extend: lanes 0..27 compare 4 bytes each forward from cur+4, lanes 28..31 one byte each
                // backward; one branch-free instruction stream for both roles ----
                uint32_t fwd = 0, back = 0;
                {
                    const uint32_t maxb = cw_umin(cur - anchor, cnd);
                    const bool fl = lane < CW_LZ4_FW_LANES;
                    const uint32_t xb = lane - (CW_LZ4_FW_LANES - 1u);  // 1..4 on the backward lanes
                    const bool bl = !fl && xb <= maxb;                  // backward lane with a byte to look at
                    uint32_t pa = fl ? cur + 4u + 4u * lane : (bl ? cur - xb : 0u);
                    uint32_t pb = fl ? cnd + 4u + 4u * lane : (bl ? cnd - xb : 0u);
                    uint32_t d = cw_lz4_read32(s, pa) ^ cw_lz4_read32(s, pb);
                    uint32_t eq = d ? (uint32_t)(cw_ffs(d) - 1) >> 3 : 4u;  // equal leading bytes
                    const uint32_t room = pa < matchlimit ? matchlimit - pa : 0u;
                    eq = fl ? cw_umin(eq, room) : eq;
                    const bool stopl = fl ? eq < 4u : !(bl && eq != 0u);
                    const uint32_t stop = cw_ballot(stopl);
                    uint32_t fstop = stop & ((1u << CW_LZ4_FW_LANES) - 1u);
                    const uint32_t bstop = stop >> CW_LZ4_FW_LANES;
                    while (!fstop) {  // more than 112 equal bytes: keep going, all 32 lanes forward
                        fwd += fwd ? 128u : 4u * CW_LZ4_FW_LANES;
                        pa = cur + 4u + fwd + 4u * lane;
                        pb = cnd + 4u + fwd + 4u * lane;
                        d = cw_lz4_read32(s, pa) ^ cw_lz4_read32(s, pb);
                        eq = d ? (uint32_t)(cw_ffs(d) - 1) >> 3 : 4u;
                        eq = cw_umin(eq, pa < matchlimit ? matchlimit - pa : 0u);
                        fstop = cw_ballot(eq < 4u);
                    }
                    {
                        const unsigned L = (unsigned)cw_ffs(fstop) - 1u;
                        fwd += 4u * L + cw_shfl(eq, L);
                    }
                    if (bstop) {
                        back = (uint32_t)cw_ffs(bstop) - 1u;
                    } else {  // 4 equal bytes before the match and room for more: whole warp, byte-wise
                        back = 4u;
                        for (;;) {
                            const uint32_t x = back + lane + 1u;
                            const bool ok = x <= maxb && s[cur - cw_umin(x, cur)] == s[cnd - cw_umin(x, cnd)];
                            const uint32_t bs = cw_ballot(!ok);
                            if (bs) {
                                back += (uint32_t)cw_ffs(bs) - 1u;
                                break;
                            }
                            back += 32u;
                        }
                    }
                }
                const uint32_t e = cur + 4u + fwd;  // end of the match = next parser position
                // ---- record the sequence: token, literal length, literals, offset, match length ----
                {
                    const uint32_t lit = cur - back - anchor;
                    const uint32_t mcode = back + fwd;  // match length - 4
                    if (lane == nrec) {
                        r_src = anchor;
                        r_lit = lit;
                        r_off = cur - cnd;
                        r_mcode = mcode;
                        r_op = op;
                    }
                    nrec++;
                    op += 3u + lit;
                    if ((lit | mcode) >= 15u) op += cw_lz4_ext_bytes(lit) + cw_lz4_ext_bytes(mcode);  // rare in text
                }
                anchor = e;
                if (e >= mfl1) {  // "if (ip > mflimit) break" -> last literals
                    done = true;
                    break;
                }
                // parser: table[hash(e-2)] = e-2, then immediate re-test at e, then a fresh search at e+1
                sbase = e + 1u;
                const uint32_t de = e - base;  // lane index of the next parser position (unit windows)
                if (unit && de < 34u) ins |= 1u << (de - 2u);
                else pend = e - 2u;
                if (unit && de < 32u) {
                    i0 = de;
                    continue;
                }
                restart = true;
                break;
            }
            // ---------------- phase D: write the recorded sequences, one lane each ----------------
            if (nrec) {
                const bool mine = lane < nrec;
                if (mine) {
                    uint32_t p = r_op;
                    out[p++] = (uint8_t)((cw_umin(r_lit, 15u) << 4) | cw_umin(r_mcode, 15u));
                    if (r_lit >= 15u) p = cw_lz4_put_ext1(out, p, r_lit - 15u);
                    if (r_lit <= CW_LZ4_LIT_SELF)
                        for (uint32_t x = 0; x < r_lit; x++) out[p + x] = s[r_src + x];
                    p += r_lit;
                    out[p] = (uint8_t)r_off;
                    out[p + 1u] = (uint8_t)(r_off >> 8);
                    if (r_mcode >= 15u) cw_lz4_put_ext1(out, p + 2u, r_mcode - 15u);
                }
                uint32_t big = cw_ballot(mine && r_lit > CW_LZ4_LIT_SELF);
                while (big) {
                    const unsigned k = (unsigned)cw_ffs(big) - 1u;
                    big &= big - 1u;
                    const uint32_t lit = cw_shfl(r_lit, k);
                    const uint32_t dst = cw_shfl(r_op, k) + 1u + cw_lz4_ext_bytes(lit);
                    cw_lz4_copy_warp(out, dst, s, cw_shfl(r_src, k), lit, lane);
                }
            }
            if (done) break;
            // ---------------- phase C: every touched slot gets its final value ----------------
            if (twins) {
                // slot owners first put the old entry back, then the highest inserted twin of each slot writes
                if (tag_ok && owner) table[h] = (uint16_t)t;
                cw_syncwarp();
                if (((ins >> lane) & 1u) && !(M & ins & ~((2u << lane) - 1u))) table[h] = (uint16_t)pos;
            } else if (tag_ok) {
                if (valid) table[h] = (uint16_t)(((ins >> lane) & 1u) ? pos : t);
            } else {
                if ((ins >> lane) & 1u) table[h] = (uint16_t)pos;
            }
            cw_syncwarp();
            if (pend != CW_LZ4_NONE) {
                if (lane == 0) table[cw_lz4_hash<HASHLOG>(cw_lz4_read32(s, pend))] = (uint16_t)pend;
                cw_syncwarp();
            }
            if (restart) j0 = -1;                                       // lane 0 = re-test at e
            else if (unit) j0 = (int32_t)(base + 32u) - (int32_t)sbase;  // S(j) = j while steps are 1
            else j0 += 32;
        }
    }
    // ---- last literals ----
    {
        const uint32_t lit = n - anchor;
        if (lane == 0) {
            out[op] = (uint8_t)(cw_umin(lit, 15u) << 4);
            if (lit >= 15u) cw_lz4_put_ext1(out, op + 1u, lit - 15u);
        }
        op += 1u + cw_lz4_ext_bytes(lit);
        cw_lz4_copy_warp(out, op, s, anchor, lit, lane);
        op += lit;
    }
    return op;
}
