/*
 * ref_cli.c -- driver around the REFERENCE's own per-chunk primitives (test infrastructure only).
 *
 * Built by oracle/Makefile into oracle/_ref/ref_cli (git-ignored, shipped to the GPU box) from:
 *   - /root/reference/reference_code/skein/Optimized_64bit/{skein.c,skein_block.c}  (compiled in place)
 *   - /root/reference/src/compression_perf/lib/{liblz4.a,liblzf.a}                   (vendored, non-PIC
 *     => this must be a -no-pie executable, it cannot be a shared object; SURVEY.md hard part F)
 *   - OpenSSL SHA256()                                                               (hash.cpp:35)
 * Nothing here is copied from the reference; it only calls the reference's public entry points the way
 * HashAndCompress.cpp:121-134 (Skein Init/Update/Final per block), :342-355 (lzf / lz4 lambdas) and
 * experiment.cpp:110,118,249,256 do.
 *
 *   ref_cli hash       <skein512|skein256-128|sha256> <chunk> <in> <out.digests>
 *   ref_cli compress   <lz4|lzf> <chunk> <in> <out.blocks> <out.lens>      (stride = bound(chunk))
 *   ref_cli decompress <lz4|lzf> <chunk> <in.blocks> <in.lens> <out>
 *   ref_cli bench      <hash> <comp> <chunk> <n_chunks> <threads> <first_chunk> [seed]
 *                      synthetic stream (include/cw_synth.h), compress-then-hash per chunk like
 *                      ProcessBlock (HashAndCompress.cpp:231-261); prints one JSON line
 *   ref_cli version
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <openssl/sha.h>

#include "lz4/lz4.h"
#include "lzf/lzf.h"
#include "skein.h"

#include "../include/cw_synth.h"

enum { H_SK512, H_SK256, H_SHA };
enum { C_LZ4, C_LZF };

static int parse_hash(const char *s)
{
    if (!strcmp(s, "skein512")) return H_SK512;
    if (!strcmp(s, "skein256-128") || !strcmp(s, "skein")) return H_SK256;
    if (!strcmp(s, "sha256")) return H_SHA;
    fprintf(stderr, "bad hash alg %s\n", s);
    exit(2);
}
static int parse_comp(const char *s)
{
    if (!strcmp(s, "lz4")) return C_LZ4;
    if (!strcmp(s, "lzf")) return C_LZF;
    fprintf(stderr, "bad comp alg %s\n", s);
    exit(2);
}
static int digest_bytes(int h) { return h == H_SK512 ? 64 : h == H_SK256 ? 16 : 32; }
static size_t bound(int c, size_t chunk) { return c == C_LZ4 ? (size_t)LZ4_COMPRESSBOUND(chunk) : chunk; }

static void hash_one(int h, const uint8_t *p, size_t n, uint8_t *out)
{
    if (h == H_SK512) {
        Skein_512_Ctxt_t c;
        Skein_512_Init(&c, 512);
        Skein_512_Update(&c, p, n);
        Skein_512_Final(&c, out);
    } else if (h == H_SK256) {
        Skein_256_Ctxt_t c;
        Skein_256_Init(&c, 128);
        Skein_256_Update(&c, p, n);
        Skein_256_Final(&c, out);
    } else
        SHA256(p, n, out);
}

/* liblzf.a was built with INIT_HTAB 0: its 256 KB table is uninitialised stack.  Scrub the region the
 * callee's frame will occupy so that "compat" is defined against a zeroed table. */
static __attribute__((noinline)) void scrub_stack(void)
{
    volatile uint8_t pad[0x50000];
    memset((void *)pad, 0, sizeof pad);
    __asm__ volatile("" ::"r"(pad) : "memory");
}
static __attribute__((noinline)) unsigned lzf_zeroed(const uint8_t *p, size_t n, uint8_t *out, size_t cap)
{
    scrub_stack();
    return lzf_compress(p, (unsigned)n, out, (unsigned)cap);
}

static uint32_t compress_one(int c, const uint8_t *p, size_t n, uint8_t *out, size_t cap)
{
    if (c == C_LZ4) return (uint32_t)LZ4_compress_default((const char *)p, (char *)out, (int)n, (int)cap);
    return lzf_zeroed(p, n, out, n - 1);
}

static uint8_t *slurp(const char *path, size_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        perror(path);
        exit(1);
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *b = (uint8_t *)malloc((size_t)n + 16);
    if (n && fread(b, 1, (size_t)n, f) != (size_t)n) {
        perror("read");
        exit(1);
    }
    fclose(f);
    *len = (size_t)n;
    return b;
}
static void spit(const char *path, const void *p, size_t n)
{
    FILE *f = fopen(path, "wb");
    if (!f || (n && fwrite(p, 1, n, f) != n)) {
        perror(path);
        exit(1);
    }
    fclose(f);
}

typedef struct {
    int h, c, cpu;
    const uint8_t *src;
    size_t chunk, lo, hi, stride;
    uint8_t *dig, *dst;
    uint32_t *lens;
} job_t;

static void *worker(void *a)
{
    job_t *j = (job_t *)a;
    if (j->cpu >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(j->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set); /* taskset idea of run_tests:14-30 */
    }
    const int D = digest_bytes(j->h);
    for (size_t i = j->lo; i < j->hi; i++) {
        /* order of ProcessBlock: compress, then hash */
        j->lens[i] = compress_one(j->c, j->src + i * j->chunk, j->chunk, j->dst + i * j->stride, j->stride);
        hash_one(j->h, j->src + i * j->chunk, j->chunk, j->dig + i * D);
    }
    return NULL;
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: ref_cli hash|compress|decompress|bench|version ...\n");
        return 2;
    }
    if (!strcmp(argv[1], "version")) {
        printf("lz4 %s, lzf api 0x%x, skein v%d (round-1 constants)\n", LZ4_versionString(), LZF_VERSION,
               SKEIN_VERSION);
        return 0;
    }
    if (!strcmp(argv[1], "hash") && argc == 6) {
        int h = parse_hash(argv[2]);
        size_t chunk = strtoull(argv[3], 0, 10), len;
        uint8_t *in = slurp(argv[4], &len);
        size_t n = len / chunk; /* only full chunks, file.cpp:18-33 */
        int D = digest_bytes(h);
        uint8_t *out = (uint8_t *)malloc(n * D + 1);
        for (size_t i = 0; i < n; i++) hash_one(h, in + i * chunk, chunk, out + i * D);
        spit(argv[5], out, n * D);
        return 0;
    }
    if (!strcmp(argv[1], "compress") && argc == 7) {
        int c = parse_comp(argv[2]);
        size_t chunk = strtoull(argv[3], 0, 10), len;
        uint8_t *in = slurp(argv[4], &len);
        size_t n = len / chunk, stride = bound(c, chunk);
        uint8_t *out = (uint8_t *)calloc(n * stride + 1, 1);
        uint32_t *lens = (uint32_t *)calloc(n + 1, 4);
        for (size_t i = 0; i < n; i++) lens[i] = compress_one(c, in + i * chunk, chunk, out + i * stride, stride);
        spit(argv[5], out, n * stride);
        spit(argv[6], lens, n * 4);
        return 0;
    }
    if (!strcmp(argv[1], "decompress") && argc == 7) {
        int c = parse_comp(argv[2]);
        size_t chunk = strtoull(argv[3], 0, 10), blen, llen;
        uint8_t *blocks = slurp(argv[4], &blen);
        uint32_t *lens = (uint32_t *)slurp(argv[5], &llen);
        size_t n = llen / 4, stride = n ? blen / n : 0;
        uint8_t *out = (uint8_t *)calloc(n * chunk + 1, 1);
        int bad = 0;
        for (size_t i = 0; i < n; i++) {
            long got;
            if (c == C_LZ4)
                got = LZ4_decompress_safe((const char *)blocks + i * stride, (char *)out + i * chunk, (int)lens[i],
                                          (int)chunk);
            else if (lens[i] == 0) { /* "did not fit": the caller stored the chunk raw */
                memcpy(out + i * chunk, blocks + i * stride, chunk);
                got = (long)chunk;
            } else
                got = (long)lzf_decompress(blocks + i * stride, lens[i], out + i * chunk, (unsigned)chunk);
            if (got != (long)chunk) {
                fprintf(stderr, "chunk %zu: decoder returned %ld\n", i, got);
                bad = 1;
            }
        }
        spit(argv[6], out, n * chunk);
        return bad;
    }
    if (!strcmp(argv[1], "bench") && argc >= 8) {
        int h = parse_hash(argv[2]), c = parse_comp(argv[3]);
        size_t chunk = strtoull(argv[4], 0, 10), n = strtoull(argv[5], 0, 10);
        int threads = atoi(argv[6]);
        uint64_t first = strtoull(argv[7], 0, 10);
        uint64_t seed = argc > 8 ? strtoull(argv[8], 0, 0) : CW_SYNTH_DEFAULT_SEED;
        int reps = argc > 9 ? atoi(argv[9]) : 1;
        size_t stride = bound(c, chunk);
        int D = digest_bytes(h);
        uint8_t *src = (uint8_t *)malloc(n * chunk);
        uint8_t *dst = (uint8_t *)malloc(n * stride);
        uint8_t *dig = (uint8_t *)malloc(n * D);
        uint32_t *lens = (uint32_t *)malloc(n * 4);
        for (size_t i = 0; i < n; i++) cw_synth_chunk(src + i * chunk, chunk, first + i, seed);
        memset(dst, 0, n * stride); /* fault the pages in before timing */
        if (threads < 1) threads = 1;
        pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
        job_t *jobs = (job_t *)malloc(sizeof(job_t) * threads);
        cpu_set_t avail;
        sched_getaffinity(0, sizeof avail, &avail);
        int ncpu = CPU_COUNT(&avail);
        int *cpus = (int *)malloc(sizeof(int) * (ncpu > 0 ? ncpu : 1));
        for (int i = 0, k = 0; i < CPU_SETSIZE && k < ncpu; i++)
            if (CPU_ISSET(i, &avail)) cpus[k++] = i;
        for (int rep = 0; rep < reps; rep++) {
            struct timespec t0, t1;
            clock_gettime(CLOCK_MONOTONIC, &t0);
            for (int t = 0; t < threads; t++) {
                job_t j = {h, c, ncpu > 0 ? cpus[t % ncpu] : -1, src, chunk, n * t / threads, n * (t + 1) / threads,
                           stride, dig, dst, lens};
                jobs[t] = j;
                pthread_create(&tid[t], NULL, worker, &jobs[t]);
            }
            for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
            clock_gettime(CLOCK_MONOTONIC, &t1);
            double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
            uint64_t csum = 0, dsum = 0;
            for (size_t i = 0; i < n; i++) csum += lens[i] ? lens[i] : chunk;
            for (size_t i = 0; i < n * (size_t)D; i++) dsum = dsum * 1099511628211ull + dig[i];
            printf("{\"impl\": \"reference\", \"kind\": \"reference\", \"hash\": \"%s\", \"comp\": \"%s\", "
                   "\"chunk\": %zu, \"n_chunks\": %zu, \"threads\": %d, \"host_cpus\": %d, \"seconds\": %.6f, "
                   "\"gbps\": %.6f, \"ratio\": %.6f, \"comp_bytes\": %llu, \"digest_fnv\": \"%016llx\"}\n",
                   argv[2], argv[3], chunk, n, threads, ncpu, sec, (double)(n * chunk) / sec / 1e9,
                   (double)csum / (double)(n * chunk), (unsigned long long)csum, (unsigned long long)dsum);
            fflush(stdout);
        }
        return 0;
    }
    fprintf(stderr, "bad arguments\n");
    return 2;
}
