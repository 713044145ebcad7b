/*
 * lz4_fast_model.c -- scalar model of the window-parallel LZ4 "fast mode" parse (csrc/lz4_fast.cuh), used to
 * choose its parameters (table size, in-window candidates, inserts inside matches) by compressed size against
 * the restated reference parse (oracle/cw_oracle.c) BEFORE spending GPU time.  Development tool only; links
 * oracle/liboracle.so for the reference sizes and the decoder.
 *
 *   gcc -O2 -o /tmp/lz4_fast_model tools/lz4_fast_model.c -Ioracle -Loracle -loracle -Wl,-rpath,$PWD/oracle
 *   /tmp/lz4_fast_model <chunk> <hashlog> <inwin 0|1> <insert_in_match 0|1> <back 0|1> <cap> files...
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cw_oracle.h"

static uint32_t rd32(const uint8_t *s, uint32_t p)
{
    uint32_t v;
    memcpy(&v, s + p, 4);
    return v;
}

static int SW = 1, RETEST = 0, POST = 0, HL = 13, INWIN = 0, INS_IN_MATCH = 1, BACK = 1, W = 32, QUAD = 0, PER = 2;

static uint32_t put_ext(uint8_t *out, uint32_t op, uint32_t r)
{
    while (r >= 255) {
        out[op++] = 255;
        r -= 255;
    }
    out[op++] = (uint8_t)r;
    return op;
}

static uint32_t model_compress(const uint8_t *s, uint32_t n, uint8_t *out)
{
    static uint16_t table[1 << 16];
    memset(table, 0, sizeof(uint16_t) << HL);
    uint32_t op = 0, anchor = 0, ip = 0;
    if (n >= 13) {
        const uint32_t last = n - 12, matchlimit = n - 5;
        for (uint32_t sbase = 0; sbase <= last; sbase += W * SW) {
            static uint32_t cand[16][128], v[16][128], h[16][128];
            static int ok[16][128], live[16];
            if (!QUAD && ip >= sbase + W * SW) continue;  // whole group inside a match
            // phase 1: look-ups and inserts of every live window of the group, window by window
            for (int j = 0; j < SW; j++) {
                const uint32_t base = sbase + j * W;
                live[j] = base <= last && (QUAD || ip < base + W);  // windows inside a match when the group starts are skipped
                if (!live[j]) continue;
                for (int l = 0; l < W; l++) {
                    const uint32_t p = base + l;
                    ok[j][l] = 0;
                    if (p > last) continue;
                    v[j][l] = rd32(s, p);
                    h[j][l] = (v[j][l] * 2654435761u) >> (32 - HL);
                    cand[j][l] = table[h[j][l]];
                }
                if (QUAD) {  // owner = minimal (p & 3, p >> 2): four store rounds k = 3..0, lowest lane wins a round
                    for (int k = 3; k >= 0; k--)
                        for (int l = W - 4 + k; l >= 0; l -= 4)
                            if (base + l <= last && base + l != 0) table[h[j][l]] = (uint16_t)(base + l);
                } else
                for (int l = W - 1; l >= 0; l--)
                    if (base + l <= last && base + l != 0) table[h[j][l]] = (uint16_t)(base + l);  // lowest position owns the slot
                for (int l = 0; l < W; l++) {
                    const uint32_t p = base + l;
                    if (p > last || p == 0) continue;
                    const uint32_t w = table[h[j][l]];
                    if (w < p) cand[j][l] = w;
                    if (PER && p >= (uint32_t)PER && rd32(s, p - PER) == v[j][l]) cand[j][l] = p - PER;
                    ok[j][l] = rd32(s, cand[j][l]) == v[j][l];
                }
            }
            // phase 2: greedy walk
            for (int j = 0; j < SW; j++) {
                const uint32_t base = sbase + j * W;
                if (!live[j]) continue;
                for (int l = ip > base ? (int)(ip - base) : 0; l < W; l++) {
                    if (!ok[j][l]) continue;
                    uint32_t cur = base + l, c = cand[j][l];
                    uint32_t e = cur + 4, ce = c + 4;
                    while (e < matchlimit && s[e] == s[ce]) e++, ce++;
                    if (BACK)
                        while (cur > anchor && c > 0 && s[cur - 1] == s[c - 1]) cur--, c--;
                    const uint32_t lit = cur - anchor, mcode = e - cur - 4;
                    out[op++] = (uint8_t)(((lit < 15 ? lit : 15) << 4) | (mcode < 15 ? mcode : 15));
                    if (lit >= 15) op = put_ext(out, op, lit - 15);
                    memcpy(out + op, s + anchor, lit);
                    op += lit;
                    out[op++] = (uint8_t)(cur - c);
                    out[op++] = (uint8_t)((cur - c) >> 8);
                    if (mcode >= 15) op = put_ext(out, op, mcode - 15);
                    anchor = ip = e;
                    l = (int)(e - base) - 1;
                }
            }
        }
    }
    const uint32_t lit = n - anchor;
    out[op++] = (uint8_t)((lit < 15 ? lit : 15) << 4);
    if (lit >= 15) op = put_ext(out, op, lit - 15);
    memcpy(out + op, s + anchor, lit);
    return op + lit;
}

int main(int argc, char **argv)
{
    if (argc < 8) return 2;
    const uint32_t chunk = (uint32_t)atoi(argv[1]);
    HL = atoi(argv[2]);
    INWIN = atoi(argv[3]);
    INS_IN_MATCH = atoi(argv[4]);
    BACK = atoi(argv[5]);
    W = atoi(argv[6]);
    RETEST = getenv("RETEST") ? atoi(getenv("RETEST")) : 0;
    SW = getenv("SW") ? atoi(getenv("SW")) : 1;
    QUAD = getenv("QUAD") ? atoi(getenv("QUAD")) : 0;
    PER = getenv("PER") ? atoi(getenv("PER")) : 2;
    POST = getenv("POST") ? atoi(getenv("POST")) : 0;
    uint8_t *out = malloc(2 * chunk + 64), *ref = malloc(2 * chunk + 64), *dec = malloc(chunk);
    unsigned long long T = 0, R = 0;
    double worst = 0;
    for (int a = 7; a < argc; a++) {
        FILE *f = fopen(argv[a], "rb");
        if (!f) continue;
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        uint8_t *d = malloc(sz + 1);
        if (fread(d, 1, sz, f) != (size_t)sz) return 3;
        fclose(f);
        unsigned long long t = 0, r = 0;
        for (long i = 0; i + chunk <= sz; i += chunk) {
            const uint32_t m = model_compress(d + i, chunk, out);
            if (cwo_lz4_decompress(out, (int)m, dec, (int)chunk) != (int)chunk || memcmp(dec, d + i, chunk)) {
                printf("ROUNDTRIP FAILURE %s chunk %ld\n", argv[a], i / chunk);
                return 1;
            }
            t += m;
            if (getenv("DUMP")) {
                FILE *o = fopen(getenv("DUMP"), "wb");
                fwrite(out, 1, m, o);
                fclose(o);
            }
            r += (unsigned)cwo_lz4_compress(d + i, (int)chunk, ref, (int)(2 * chunk + 64));
        }
        if (r) {
            const char *b = strrchr(argv[a], '/');
            printf("%-16s model %9llu ref %9llu  %.4f\n", b ? b + 1 : argv[a], t, r, (double)t / r);
            if ((double)t / r > worst) worst = (double)t / r;
        }
        T += t;
        R += r;
        free(d);
    }
    printf("TOTAL chunk %u HL %d inwin %d insm %d back %d W %d: %.4f  worst file %.4f\n", chunk, HL, INWIN, INS_IN_MATCH, BACK, W,
           (double)T / R, worst);
    return 0;
}
