// lz4_warp.cuh -- LZ4 block compression of ONE chunk by ONE warp, byte-identical to the reference.
//
// Replaces the lz4 lambda  LZ4_compress_default(s, d, l, 2*l)  (HashAndCompress.cpp:349-355; also
// experiment.cpp:249).  The reference links a prebuilt liblz4.a 1.8.2 whose source is not in its tree
// (header: src/compression_perf/include/lz4/lz4.h:94-96,126-156); the parse reproduced here is that
// library's greedy single-pass parse for inputs < 65547 bytes (13-bit hash, u16 position table,
// acceleration 1) -- SURVEY.md Appendix A.2 -- and it is accepted only through byte-equality with the
// vendored archive (tests/golden, oracle/).
//
// The sequential parser probes one position at a time: h = hash(read32(p)); cand = table[h];
// table[h] = p; match iff read32(cand) == read32(p).  Here a warp evaluates the next 32 probe positions
// of the parser at once ("window"):
//   phase A (all lanes, independent): read32, hash, table look-up, candidate compare, plus two
//            MATCH.ANY masks: lanes with the same hash (they would overwrite each other's table slot) and
//            lanes with the same 4 bytes.
//   phase B (uniform, in registers): replay the parser's control flow over the window.  `ins` is the set
//            of window lanes whose position the sequential parser has inserted so far; a lane's true
//            candidate is the highest inserted lane below it with the same hash, else the table entry read
//            in phase A.  The first hitting lane is the parser's next match; it is extended backwards
//            and forwards by the whole warp (32 x 4 bytes per step), the sequence is emitted, the parser's
//            post-match insert of position e-2 is added to `ins`, and the replay resumes at lane e
//            (positions inside the match are never probed, exactly as in the reference).
//   phase C: lanes in `ins` commit table[h] = pos (highest lane wins per slot).
// After 64 failed probes the reference widens its stride (step = searchNb++ >> 6); lane positions follow
// the closed form of that schedule, so incompressible data is skipped as quickly as the CPU does.
//
// Memory: the chunk and the 16 KB table live in shared memory (staged by the caller); output bytes go
// straight to global memory.
#pragma once
#include "cw_device.cuh"

#define CW_LZ4_HASH_ENTRIES 8192u
#define CW_LZ4_NONE 0xffffffffu

CW_DEV uint32_t cw_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// unaligned little-endian 32-bit read from the staged chunk (base 4-byte aligned, >= 8 bytes of padding)
CW_DEV uint32_t cw_lz4_read32(const uint8_t *s, uint32_t pos)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s);
    const uint32_t i = pos >> 2;
    return cw_funnel_r(w[i], w[i + 1], (pos & 3u) * 8u);
}
CW_DEV uint32_t cw_lz4_hash(uint32_t v) { return (v * 2654435761u) >> 19; }

// probe schedule of one search: probe j sits S(j) bytes after the search start and
// the next probe is step(j) further.  j = -1 is the immediate re-test right after a match.
CW_DEV uint32_t cw_lz4_F(uint32_t m)
{
    const uint32_t q = m >> 6, r = m & 63u;
    return 32u * q * (q - 1u) + q * r;  // sum_{x<m} (x >> 6)
}
CW_DEV int32_t cw_lz4_S(int32_t j) { return j <= 0 ? j : 1 + (int32_t)cw_lz4_F(63u + (uint32_t)j); }
CW_DEV uint32_t cw_lz4_step(int32_t j) { return j <= 0 ? 1u : (63u + (uint32_t)j) >> 6; }

// 255-run length extension shared by literal and match lengths: writes the bytes for value r (>= 0)
CW_DEV uint32_t cw_lz4_put_ext(uint8_t *out, uint32_t op, uint32_t r, unsigned lane)
{
    const uint32_t full = r / 255u;
    for (uint32_t x = lane; x < full; x += 32u) out[op + x] = 255u;
    if (lane == 0) out[op + full] = (uint8_t)(r - full * 255u);
    return op + full + 1u;
}

// Compress chunk s[0..n) (n <= 65536).  table: zeroed u16[8192] in shared memory.  Returns the number of
// bytes written to out (same value in every lane).  out must hold LZ4_COMPRESSBOUND(n) bytes.
CW_DEV uint32_t cw_lz4_compress_warp(const uint8_t *s, uint32_t n, uint16_t *table, uint8_t *out)
{
    const unsigned lane = cw_lane();
    const uint32_t lower = (1u << lane) - 1u;
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
            const int32_t j = j0 + (int32_t)lane;
            const uint32_t pos = sbase + (uint32_t)cw_lz4_S(j);
            const uint32_t base = sbase + (uint32_t)cw_lz4_S(j0);  // position of lane 0
            const bool valid = pos + cw_lz4_step(j) <= mfl1;
            const uint32_t v = cw_lz4_read32(s, valid ? pos : 0u);
            const uint32_t h = cw_lz4_hash(v);
            const uint32_t M = cw_match_any(h);
            const uint32_t E = cw_match_any(v);
            const uint32_t t = table[h];
            const bool tmatch = cw_lz4_read32(s, t) == v;
            // ---------------- phase B ----------------
            uint32_t ins = 0, i0 = 0, pend = CW_LZ4_NONE;
            bool restart = false;
            for (;;) {
                const uint32_t range = ~((1u << i0) - 1u);
                const uint32_t cm = M & (ins | range) & lower;
                bool hit;
                uint32_t cand;
                if (cm) {
                    const int jl = 31 - cw_clz(cm);
                    hit = (E >> jl) & 1u;
                    cand = unit ? base + (uint32_t)jl : sbase + (uint32_t)cw_lz4_S(j0 + jl);
                } else {
                    hit = tmatch;
                    cand = t;
                }
                const bool act = lane >= i0;
                const uint32_t hitm = cw_ballot(act && valid && hit);
                const uint32_t invm = cw_ballot(act && !valid);
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
                uint32_t cur = cw_shfl(pos, (unsigned)f);
                uint32_t cnd = cw_shfl(cand, (unsigned)f);

                // ---- extend: forward from cur+4 (32 lanes x 4 bytes) and backward (32 lanes x 1 byte) ----
                uint32_t fwd = 0, back = 0;
                {
                    const uint32_t maxb = cw_umin(cur - anchor, cnd);
                    // first steps of both directions are issued together so their smem latencies overlap
                    uint32_t x0 = 4u * lane;
                    uint32_t pa = cur + 4u + x0, pb = cnd + 4u + x0;
                    uint32_t da = cw_lz4_read32(s, cw_umin(pa, n)) ^ cw_lz4_read32(s, cw_umin(pb, n));
                    const uint32_t xb = lane + 1u;
                    const bool bok = xb <= maxb && s[cur - cw_umin(xb, cur)] == s[cnd - cw_umin(xb, cnd)];
                    for (;;) {
                        uint32_t eq = da ? (uint32_t)(cw_ffs(da) - 1) >> 3 : 4u;
                        const uint32_t room = pa < matchlimit ? matchlimit - pa : 0u;
                        eq = cw_umin(eq, room);
                        const uint32_t stop = cw_ballot(eq < 4u);
                        if (stop) {
                            const unsigned L = (unsigned)cw_ffs(stop) - 1u;
                            fwd += 4u * L + cw_shfl(eq, L);
                            break;
                        }
                        fwd += 128u;
                        pa += 128u;
                        pb += 128u;
                        da = cw_lz4_read32(s, cw_umin(pa, n)) ^ cw_lz4_read32(s, cw_umin(pb, n));
                    }
                    uint32_t bstop = cw_ballot(!bok);
                    if (bstop) {
                        back = (uint32_t)cw_ffs(bstop) - 1u;
                    } else {
                        back = 32u;
                        for (;;) {
                            const uint32_t x = back + lane + 1u;
                            const bool ok = x <= maxb && s[cur - cw_umin(x, cur)] == s[cnd - cw_umin(x, cnd)];
                            bstop = cw_ballot(!ok);
                            if (bstop) {
                                back += (uint32_t)cw_ffs(bstop) - 1u;
                                break;
                            }
                            back += 32u;
                        }
                    }
                }
                const uint32_t e = cur + 4u + fwd;  // end of the match = next parser position
                // ---- emit: token, literal length, literals, offset, match length ----
                {
                    const uint32_t lit = cur - back - anchor;
                    const uint32_t mcode = back + fwd;  // match length - 4
                    const uint32_t off = cur - cnd;
                    if (lane == 0) out[op] = (uint8_t)((cw_umin(lit, 15u) << 4) | cw_umin(mcode, 15u));
                    op += 1u;
                    if (lit >= 15u) op = cw_lz4_put_ext(out, op, lit - 15u, lane);
                    for (uint32_t x = lane; x < lit; x += 32u) out[op + x] = s[anchor + x];
                    op += lit;
                    if (lane < 2u) out[op + lane] = (uint8_t)(off >> (8u * lane));
                    op += 2u;
                    if (mcode >= 15u) op = cw_lz4_put_ext(out, op, mcode - 15u, lane);
                }
                anchor = e;
                if (e >= mfl1) {  // "if (ip > mflimit) break" -> last literals
                    done = true;
                    break;
                }
                // parser: table[hash(e-2)] = e-2, then immediate re-test at e, then a fresh search at e+1
                sbase = e + 1u;
                if (unit && e - 2u <= base + 31u) ins |= 1u << (e - 2u - base);
                else pend = e - 2u;
                if (unit && e <= base + 31u) {
                    i0 = e - base;
                    continue;
                }
                restart = true;
                break;
            }
            if (done) break;
            // ---------------- phase C: commit the parser's table writes ----------------
            if (((ins >> lane) & 1u) && !(M & ins & ~((2u << lane) - 1u))) table[h] = (uint16_t)pos;
            cw_syncwarp();
            if (pend != CW_LZ4_NONE) {
                if (lane == 0) table[cw_lz4_hash(cw_lz4_read32(s, pend))] = (uint16_t)pend;
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
        if (lane == 0) out[op] = (uint8_t)(cw_umin(lit, 15u) << 4);
        op += 1u;
        if (lit >= 15u) op = cw_lz4_put_ext(out, op, lit - 15u, lane);
        for (uint32_t x = lane; x < lit; x += 32u) out[op + x] = s[anchor + x];
        op += lit;
    }
    return op;
}
