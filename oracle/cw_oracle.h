/*
 * cw_oracle.h -- CPU ORACLE for the chunked hash-and-compress path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the algorithms the
 * reference (ankitkumar93/compute-war) runs per chunk.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and there only as the checker -- never as the product path.  The
 * product (compute_war_b200/libcomputewar.so) has no CPU fallback.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   Skein-512 / Skein-256 : pinned by the reference's KAT files (tests/golden/skein_kat.json,
 *                           extracted from reference_code/skein/KAT_MCT/) and by the reference's own
 *                           source compiled into oracle/_ref/libskref.so.
 *   SHA-256               : pinned by FIPS 180-4 known answers and OpenSSL SHA256() (the call the
 *                           reference makes, src/hashing_perf/hash.cpp:35).
 *   LZ4 / LZF compressors : source NOT in the reference tree (prebuilt liblz4.a 1.8.2 / liblzf.a);
 *                           pinned by outputs of those vendored archives run here
 *                           (oracle/_ref/ref_cli, fixtures in tests/golden/).
 */
#ifndef CW_ORACLE_H
#define CW_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* algorithm selectors: numerically identical to include/compute_war.h */
enum { CWO_HASH_SKEIN512 = 0, CWO_HASH_SKEIN256_128 = 1, CWO_HASH_SHA256 = 2 };
enum { CWO_COMP_LZ4 = 0, CWO_COMP_LZF = 1 };

/* Skein v1.1 (NIST round-1) sequential hashing.  state_bits in {256,512}; msg length in BITS
 * (partial final byte handled as SHA3api_ref.c:59-83); out gets (hash_bits+7)/8 bytes. */
void cwo_skein(unsigned state_bits, unsigned hash_bits, const uint8_t *msg, uint64_t msg_bits,
               uint8_t *out);
/* SHA-256 (FIPS 180-4) == OpenSSL SHA256() as called at src/hashing_perf/hash.cpp:35 */
void cwo_sha256(const uint8_t *msg, size_t len, uint8_t out[32]);

/* LZ4 block compressor with the exact parse of LZ4 1.8.2 LZ4_compress_default (acceleration 1,
 * fresh table) for inputs < 65547 bytes -- call sites HashAndCompress.cpp:349-355, experiment.cpp:249.
 * Returns bytes written, 0 if dst_cap is too small. */
int cwo_lz4_compress(const uint8_t *src, int n, uint8_t *dst, int dst_cap);
/* The ENGINE's own fast-mode LZ4 parse (csrc/lz4_fast.cuh), restated in scalar form -- not a reference algorithm: the
 * reference constrains fast mode only through its decoder and its size (see cw_oracle.c).  hashlog: table bits,
 * cwo_lz4_fast_hashlog(chunk) = what the kernels pick.  Returns bytes written, 0 if dst_cap < bound. */
int cwo_lz4_fast_hashlog(size_t chunk);
int cwo_lz4_fast_compress(const uint8_t *src, int n, uint8_t *dst, int dst_cap, int hashlog);
/* the LZF flavour of the same parse; returns bytes written, 0 when the block would take more than dst_cap bytes (the
 * engine passes n - 1 like the reference's lzf_compress call) */
int cwo_lzf_fast_hashlog(size_t chunk);
int cwo_lzf_fast_compress(const uint8_t *src, int n, uint8_t *dst, int dst_cap, int hashlog);
/* LZ4 block decoder, bounds-checked; returns decoded size or <0 on malformed input
 * (role of LZ4_decompress_safe, experiment.cpp:256). */
int cwo_lz4_decompress(const uint8_t *src, int n, uint8_t *dst, int dst_cap);

/* liblzf lzf_compress as built in the reference (HLOG 16, VERY_FAST, zeroed table) -- call sites
 * HashAndCompress.cpp:342-348, experiment.cpp:110.  Returns bytes written, 0 = does not fit. */
unsigned cwo_lzf_compress(const uint8_t *in, unsigned in_len, uint8_t *out, unsigned out_len);
/* lzf_decompress (experiment.cpp:118); returns decoded size, 0 on error. */
unsigned cwo_lzf_decompress(const uint8_t *in, unsigned in_len, uint8_t *out, unsigned out_len);

/* digest size in bytes for a hash selector */
int cwo_digest_bytes(int hash_alg);
/* worst-case compressed size of one chunk (== cw_compress_bound) */
size_t cwo_compress_bound(int comp_alg, size_t chunk);

/* Per-chunk drivers mirroring doHashing / doCompression / ProcessBlock
 * (HashAndCompress.cpp:121-134, :342-355, :231-261).  src = n*chunk contiguous bytes. */
void cwo_hash_chunks(int hash_alg, const uint8_t *src, size_t chunk, size_t n, uint8_t *digests);
/* LZ4: dst capacity per chunk = stride; LZF: out_len = chunk-1 as in the reference (lens[i]==0 => raw) */
void cwo_compress_chunks(int comp_alg, const uint8_t *src, size_t chunk, size_t n, uint8_t *dst,
                         size_t stride, uint32_t *lens);
/* compress-then-hash every chunk (order of ProcessBlock) */
void cwo_hash_compress_chunks(int hash_alg, int comp_alg, const uint8_t *src, size_t chunk, size_t n,
                              uint8_t *digests, uint8_t *dst, size_t stride, uint32_t *lens);
/* decode chunk i of a strided block array back; returns 0 if every chunk round-trips to src */
int cwo_verify_roundtrip(int comp_alg, const uint8_t *src, size_t chunk, size_t n, const uint8_t *blocks,
                         size_t stride, const uint32_t *lens);
/* multi-threaded ProcessBlock loop over contiguous chunk ranges (pthreads); returns seconds of the
 * processing phase only, like HashAndCompress.cpp:397-404 */
double cwo_hash_compress_mt(int hash_alg, int comp_alg, const uint8_t *src, size_t chunk, size_t n,
                            uint8_t *digests, uint8_t *dst, size_t stride, uint32_t *lens, int threads);

/* synthetic mixed-entropy workload (definition in include/cw_synth.h) */
unsigned cwo_synth_class(uint64_t seed, uint64_t idx);
void cwo_synth_fill(uint8_t *dst, size_t chunk, uint64_t first_chunk, size_t n, uint64_t seed);

/* The digest randomness study, Hasher::LogResultsInternal + LogByteResults (src/hashing_randomness/hash.cpp:30-101),
 * over the first n_bytes bytes of n digests of D bytes: avg[b], stddev[b] = the two numbers of the reference's log line
 * for byte b; returns "Average std dev". */
double cwo_digest_byte_stats(const uint8_t *digests, size_t D, size_t n, size_t n_bytes, double *avg, double *stddev);
/* first[i] = smallest j with digest j == digest i (definition of the dedup map; the reference has no counterpart);
 * returns the number of distinct digests */
size_t cwo_dedup_first(const uint8_t *digests, size_t D, size_t n, uint32_t *first);

#ifdef __cplusplus
}
#endif
#endif
