/*
 * cw_synth.h -- definition of the synthetic mixed-entropy workload (BASELINE.json configs[4]:
 * "16 GB synthetic mixed-entropy buffer").
 *
 * The reference has no such generator (its inputs are corpus files, HashAndCompress.cpp:188-218);
 * this is workload synthesis, not part of the hash/compress path.  The stream is a pure function of
 * (seed, chunk index, chunk size) built from integer arithmetic only, so the CUDA generator
 * (csrc/synth.cu) and the CPU one (oracle/cw_oracle.c: cwo_synth_fill) emit identical bytes and a
 * multi-GPU run can create its own shard in HBM without a host transfer.
 *
 * Chunk i draws a class from an 8-slot table {P,T,T,T,R,R,X,X}:
 *   P  pattern : 64-byte random pattern repeated, one byte replaced every 61 bytes
 *   T  text    : space-separated lower-case "words" from a 4096-word vocabulary with a skewed
 *                (min-of-two-uniforms) rank distribution, occasional ", " / ".\n"
 *   R  records : 128-byte records, each a copy of the previous one with a 4-byte counter bumped
 *                and the last 32 bytes refreshed (spreadsheet / database page like)
 *   X  random  : raw splitmix64 output (incompressible)
 */
#ifndef CW_SYNTH_H
#define CW_SYNTH_H
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CW_HD __host__ __device__ __forceinline__
#else
#define CW_HD static inline
#endif

#define CW_SYNTH_DEFAULT_SEED 0xC0FFEEull

CW_HD uint64_t cw_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* byte sink that packs 8 bytes per store; dst must be 8-byte aligned, n a multiple of 8 */
typedef struct {
    uint64_t *p;
    uint64_t acc;
    unsigned k;
    size_t left;
} cw_sink;

CW_HD void cw_sink_put(cw_sink *s, unsigned byte)
{
    if (!s->left) return;
    s->acc |= (uint64_t)(byte & 0xffu) << (8u * s->k);
    s->left--;
    if (++s->k == 8u) {
        *s->p++ = s->acc;
        s->acc = 0;
        s->k = 0;
    }
}

CW_HD unsigned cw_synth_class(uint64_t seed, uint64_t idx)
{
    /* slots: 0=P 1,2,3=T 4,5=R 6,7=X */
    unsigned slot = (unsigned)(cw_splitmix64(seed ^ (idx * 0xD1342543DE82EF95ull)) >> 61);
    return slot == 0 ? 0u : (slot <= 3 ? 1u : (slot <= 5 ? 2u : 3u));
}

/* Fill one chunk.  chunk must be a multiple of 8 and dst 8-byte aligned. */
CW_HD void cw_synth_chunk(uint8_t *dst, size_t chunk, uint64_t idx, uint64_t seed)
{
    cw_sink s;
    s.p = (uint64_t *)dst;
    s.acc = 0;
    s.k = 0;
    s.left = chunk;
    const unsigned cls = cw_synth_class(seed, idx);
    uint64_t st = cw_splitmix64(seed + 0x632BE59BD9B4E019ull * (idx + 1));
    if (cls == 3u) {
        for (size_t o = 0; o < chunk; o += 8) {
            st = cw_splitmix64(st);
            *s.p++ = st;
        }
        return;
    }
    if (cls == 0u) {
        uint64_t pat[8];
        for (int j = 0; j < 8; j++) {
            st = cw_splitmix64(st);
            pat[j] = st;
        }
        unsigned until = 0;
        for (size_t o = 0; o < chunk; o++) {
            unsigned b = (unsigned)(pat[(o >> 3) & 7] >> (8 * (o & 7))) & 0xffu;
            if (until == 0) {
                b = (unsigned)(cw_splitmix64(st ^ o) & 0xffu);
                until = 61;
            }
            until--;
            cw_sink_put(&s, b);
        }
        return;
    }
    if (cls == 1u) {
        const uint64_t letters_lo = 0x6c64726873696e6full; /* "onishrdl" */
        const uint64_t letters_hi = 0x656174656f617465ull; /* "etaoetae" */
        while (s.left) {
            st = cw_splitmix64(st);
            unsigned r1 = (unsigned)st & 4095u, r2 = (unsigned)(st >> 12) & 4095u;
            unsigned w = r1 < r2 ? r1 : r2;
            uint64_t h = cw_splitmix64(0xA0761D6478BD642Full * (w + 1));
            unsigned len = 2u + (unsigned)(h & 7u);
            for (unsigned k = 0; k < len; k++) {
                unsigned c = (unsigned)(h >> (3 + 5 * k)) & 31u;
                unsigned ch;
                if (c < 8u) ch = (unsigned)(letters_hi >> (8 * c)) & 0xffu;
                else if (c < 16u) ch = (unsigned)(letters_lo >> (8 * (c - 8))) & 0xffu;
                else ch = 'a' + (c - 16u) + (c >= 26u ? 4u : 0u);
                cw_sink_put(&s, ch);
            }
            unsigned p = (unsigned)(st >> 24) & 31u;
            if (p == 0u) {
                cw_sink_put(&s, '.');
                cw_sink_put(&s, '\n');
            } else if (p < 3u) {
                cw_sink_put(&s, ',');
                cw_sink_put(&s, ' ');
            } else {
                cw_sink_put(&s, ' ');
            }
        }
        return;
    }
    /* cls == 2: records */
    {
        uint64_t rec[12]; /* first 96 bytes of the record image, persistent part */
        for (int j = 0; j < 12; j++) {
            st = cw_splitmix64(st);
            /* printable-ish digits/upper case so that byte values cluster */
            rec[j] = (st & 0x0f0f0f0f0f0f0f0full) | 0x3030303030303030ull;
        }
        uint32_t counter = (uint32_t)st;
        while (s.left) {
            rec[0] = (rec[0] & 0xffffffff00000000ull) | counter++;
            for (unsigned o = 0; o < 96u; o++) cw_sink_put(&s, (unsigned)(rec[o >> 3] >> (8 * (o & 7))));
            for (unsigned j = 0; j < 4u; j++) {
                st = cw_splitmix64(st);
                for (unsigned o = 0; o < 8u; o++) cw_sink_put(&s, (unsigned)(st >> (8 * o)));
            }
            /* one persistent field drifts every record */
            st = cw_splitmix64(st);
            rec[1 + (st % 11u)] ^= (st >> 8) & 0x0000000f0f000000ull;
        }
    }
}

#endif /* CW_SYNTH_H */
