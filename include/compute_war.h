/*
 * compute_war.h -- C ABI of the B200-native chunked hash-and-compress engine (libcomputewar.so).
 *
 * This is the drop-in boundary for the per-chunk path of ankitkumar93/compute-war.  Every entry point
 * names the reference interface it replaces (paths relative to the reference repository).  Plain
 * pointers and sizes only; no C++ / torch types.  All functions return CW_OK (0) or a negative
 * cw_status; they never abort (the reference tools assert/abort, shared.h:30-41).
 *
 * Data layout shared by every call:
 *   src      n_chunks * chunk_bytes contiguous input bytes; chunk i = src[i*chunk_bytes ...).
 *            Only whole chunks exist at this level -- callers drop a file's tail like
 *            File::ReadAllBlocks does (src/hashing_perf/file.cpp:18-33).
 *   digests  n_chunks * cw_digest_bytes(alg), digest i at digests + i*D
 *            (doSkeinHashing packs digests at singleHashSize stride, HashAndCompress.cpp:121-134).
 *   dst      compressed block i at dst + i*dst_stride, dst_stride >= cw_compress_bound()
 *            (ProcessBlock gives every slot 2*blockSize of scratch, HashAndCompress.cpp:237-253).
 *   lens     u32 per chunk = the codec's own return value: LZ4 bytes written (> 0);
 *            LZF bytes written or 0 = "does not fit in chunk_bytes-1" => the caller stores the chunk raw
 *            (lzf.h:59-64, call site HashAndCompress.cpp:346).
 *
 * There is NO CPU fallback: without a CUDA device cw_init fails with CW_ERR_NO_DEVICE.
 */
#ifndef COMPUTE_WAR_H
#define COMPUTE_WAR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CW_VERSION 100

typedef struct cw_ctx cw_ctx;

typedef enum {
    CW_OK = 0,
    CW_ERR_ARG = -1,       /* bad pointer / size / selector */
    CW_ERR_NO_DEVICE = -2, /* no CUDA device: the engine has no CPU path */
    CW_ERR_CUDA = -3,      /* a CUDA call failed; see cw_last_error() */
    CW_ERR_NOMEM = -4,
    CW_ERR_UNSUPPORTED = -5 /* e.g. chunk_bytes > CW_MAX_CHUNK */
} cw_status;

/* --hash-alg of HashAndCompress.cpp:361-370.  The reference tools run Skein-256 with a 128-bit digest
 * (hash.h:16-17, HashAndCompress.cpp:123-129); Skein-512/512 is the north-star variant
 * (reference_code/skein/Optimized_64bit/skein.c:226-408).  All Skein is v1.1 (round-1 constants,
 * skein.h:196,285-292).  SHA-256 == OpenSSL SHA256() (hash.cpp:35). */
typedef enum { CW_HASH_SKEIN512 = 0, CW_HASH_SKEIN256_128 = 1, CW_HASH_SHA256 = 2 } cw_hash_alg;

/* --comp-alg of HashAndCompress.cpp:342-355 */
typedef enum { CW_COMP_LZ4 = 0, CW_COMP_LZF = 1 } cw_comp_alg;

/* COMPAT: streams byte-identical to the reference's LZ4_compress_default (liblz4 1.8.2) /
 *         lzf_compress (liblzf, zeroed table).
 * FAST  : any valid stream that the reference decoders (LZ4_decompress_safe / lzf_decompress,
 *         experiment.cpp:118,256) restore bit-exactly, ratio within 1 % of the reference.  Both codecs run a
 *         position-parallel parse by teams of warps (csrc/lz4_fast.cuh), 1.9-3.7 x (LZ4) / 5-7 x (LZF) the
 *         compat kernels' pace; the output is a deterministic function of the chunk.  The LZF length
 *         convention is unchanged: 0 = the block would not fit in chunk - 1 bytes, keep the chunk raw. */
typedef enum { CW_MODE_COMPAT = 0, CW_MODE_FAST = 1 } cw_mode;

#define CW_MAX_CHUNK 65536u /* LZ4's u16-table path covers inputs < 65547 bytes */

/* ---- lifecycle: replaces initializeGpu() (HashAndCompress.cpp:95-98) and the shutdown TODO (:182) ---- */
int cw_init(cw_ctx **out, int device);
void cw_shutdown(cw_ctx *ctx);
const char *cw_last_error(const cw_ctx *ctx); /* text of the last CW_ERR_CUDA on this context */
const char *cw_status_string(int status);
int cw_device_count(void);

/* ---- sizes ---- */
size_t cw_digest_bytes(cw_hash_alg alg); /* kHashSizeBytesSkein / kHashSizeBytesSHA, HashAndCompress.h */
/* worst-case bytes of one compressed chunk: LZ4_COMPRESSBOUND (lz4.h:156); chunk_bytes for LZF, which is
 * called with out_len = chunk_bytes-1 */
size_t cw_compress_bound(cw_comp_alg alg, size_t chunk_bytes);

/* ---- host-buffer calls: the reference's two function slots and ProcessBlock ----
 * Blocking.  Buffers are the caller's (the library never frees or keeps them); pinned host memory makes
 * the internal H2D/D2H copies asynchronous and overlapped with the kernels.  Ordinary (pageable) memory works too:
 * the library stages it through pinned buffers of its own with a few helper threads (CW_COPY_THREADS, default
 * min(8, host threads / 2)), at roughly a third of the pinned rate. */

/* Page-locks (and later releases) a caller-owned buffer so that the copies of the host-buffer calls run asynchronously at
 * full PCIe rate; the reference allocates its buffers with malloc / std::string (HashAndCompress.cpp:233-239), which
 * works here too but is staged through the library's pinned buffers by helper threads.  Registration is slow (about a second per GiB):
 * do it once for buffers that are reused.  No counterpart in the reference. */
int cw_host_register(cw_ctx *ctx, void *ptr, size_t bytes);
int cw_host_unregister(cw_ctx *ctx, void *ptr);

/* doHashing(const char* src, char* dst, int count): HashAndCompress.cpp:119, doSkeinHashing :121-134,
 * HashBlockSkein256 / HashBlockSHA256 (src/hashing_perf/hash.cpp:5-46) */
int cw_hash(cw_ctx *ctx, cw_hash_alg alg, const void *src, size_t chunk_bytes, size_t n_chunks, void *digests);

/* Skein_512_Init / Skein_256_Init(ctx, hashBitLen) + Update + Final (reference_code/skein/Optimized_64bit/skein.c:226-266,
 * :31-71, :329-408) for any digest length: state_bits 512 or 256, hash_bits a multiple of 8 up to state_bits.  The two
 * selectors of cw_hash are the pairs (512, 512) and (256, 128); this entry exists so that every byte-aligned vector of
 * the reference's KAT files (Skein-512/384, Skein-256/256, Skein-256/224, KAT_MCT/) runs on the GPU.  Host buffers;
 * digest i (hash_bits / 8 bytes) goes to digests + i * hash_bits / 8. */
int cw_hash_skein(cw_ctx *ctx, unsigned state_bits, unsigned hash_bits, const void *src, size_t chunk_bytes,
                  size_t n_chunks, void *digests);

/* doCompression(const char* s, char* d, size_t l): HashAndCompress.cpp:111, lambdas :342-355;
 * experiment.cpp:110 (lzf), :249 (lz4) -- applied to n_chunks chunks at once */
int cw_compress(cw_ctx *ctx, cw_comp_alg alg, cw_mode mode, const void *src, size_t chunk_bytes, size_t n_chunks,
                void *dst, size_t dst_stride, uint32_t *lens);

/* ProcessBlock(const std::string&): HashAndCompress.cpp:231-261 -- compress and hash every chunk of the
 * batch in ONE fused pass over the input */
int cw_hash_compress(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src,
                     size_t chunk_bytes, size_t n_chunks, void *digests, void *dst, size_t dst_stride,
                     uint32_t *lens);

/* ProcessBlock with PACKED output: the blocks of the batch back to back in `packed` (a chunk LZF could not shrink,
 * lens[i] == 0, is stored raw -- what the reference's caller does with lzf_compress's 0, lzf.h:59-64), offsets[i] =
 * start of block i, offsets[n_chunks] = total bytes (u64, n_chunks + 1 entries), lens / digests as in
 * cw_hash_compress.  The reference keeps `2 * readSize` bytes of scratch per work item and throws the blocks away
 * (HashAndCompress.cpp:233-239,259-260); a storage front-end keeps them, and copying whole strides back costs 1.5 x
 * the PCIe bytes of the blocks themselves.  packed_cap: capacity of `packed`; n_chunks * cw_compress_bound() always
 * suffices, the real need is the compressed size.  CW_ERR_ARG if the blocks do not fit. */
int cw_hash_compress_packed(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src,
                            size_t chunk_bytes, size_t n_chunks, void *digests, void *packed, size_t packed_cap,
                            uint64_t *offsets, uint32_t *lens);

/* ---- device-buffer calls: same contracts, every pointer is device memory on ctx's GPU ----
 * Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the context's own stream).  This is
 * what a storage front-end that already holds its write buffers in HBM calls, and what bench.py times
 * for the device-resident figure. */
int cw_hash_dev(cw_ctx *ctx, cw_hash_alg alg, const void *d_src, size_t chunk_bytes, size_t n_chunks,
                void *d_digests, void *stream);
int cw_compress_dev(cw_ctx *ctx, cw_comp_alg alg, cw_mode mode, const void *d_src, size_t chunk_bytes,
                    size_t n_chunks, void *d_dst, size_t dst_stride, uint32_t *d_lens, void *stream);
int cw_hash_compress_dev(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *d_src,
                         size_t chunk_bytes, size_t n_chunks, void *d_digests, void *d_dst, size_t dst_stride,
                         uint32_t *d_lens, void *stream);
int cw_sync(cw_ctx *ctx, void *stream);

/* ---- decompression: LZ4_decompress_safe / lzf_decompress as the reference calls them after each compression
 * (experiment.cpp:256, :118; decls lz4.h:141-149, lzf.h:98-100), over a strided block array.
 * out_lens[i] = decoded bytes (chunk_bytes on success), 0 for a malformed block.  An LZF block of length 0 is a chunk
 * the caller stored raw (lzf.h:59-64): it is copied.  SURVEY.md 8(f) row 2. ---- */
int cw_decompress(cw_ctx *ctx, cw_comp_alg alg, const void *blocks, size_t stride, const uint32_t *lens, size_t chunk_bytes,
                  size_t n_chunks, void *out, uint32_t *out_lens);
int cw_decompress_dev(cw_ctx *ctx, cw_comp_alg alg, const void *d_blocks, size_t stride, const uint32_t *d_lens,
                      size_t chunk_bytes, size_t n_chunks, void *d_out, uint32_t *d_out_lens, void *stream);

/* ---- packed output (device buffers): compacts a strided block array into one contiguous stream.
 * offsets[i] (u64, n_chunks + 1 entries) = start of block i in `packed`; offsets[n_chunks] = total bytes.
 * A block of length 0 (lzf_compress "did not fit", lzf.h:59-64) is stored raw: chunk_bytes copied from d_src (required
 * then).  If packed_cap is too small the offsets are still produced (so the caller learns the size) and blocks that
 * do not fit are skipped.  The reference discards its blocks (HashAndCompress.cpp:259-260); SURVEY.md 8(f) row 2.
 * The single-stream use of one context is assumed (the scratch buffer belongs to the context). ---- */
int cw_pack_dev(cw_ctx *ctx, const void *d_blocks, size_t stride, const uint32_t *d_lens, const void *d_src,
                size_t chunk_bytes, size_t n_chunks, void *d_packed, size_t packed_cap, uint64_t *d_offsets, void *stream);

/* ---- digest post-processing: what the digests are computed for ----
 * cw_digest_hist_dev: hist[b*256 + v] (u64, digest_bytes*256 entries) = number of digests whose byte b has value v --
 *   the table Hasher::LogResultsInternal accumulates (src/hashing_randomness/hash.cpp:30-58).
 * cw_digest_byte_stats: the per-byte figures the reference prints from that table (hash.cpp:60-101), for the first
 *   n_bytes digest bytes (the reference uses 16): avg[b] = total/255, stddev[b] = sqrt(sum (count-avg)^2 / 254) over the
 *   values 0..254 plus 255 only when it occurred (its map is pre-filled with `bucket < UINT8_MAX`), *avg_stddev = their
 *   mean ("Average std dev").  Pure arithmetic on 256*n_bytes numbers; runs on the host.
 * cw_digest_stats: both steps for digests in host memory.
 * cw_dedup_dev: first[i] = smallest j with digest j == digest i (first[i] == i <=> a deduplicating store keeps chunk
 *   i), *n_unique = number of distinct digests.  digest_bytes % 4 == 0, n < 2^31. */
int cw_digest_hist_dev(cw_ctx *ctx, const void *d_digests, size_t digest_bytes, size_t n, uint64_t *d_hist, void *stream);
int cw_digest_byte_stats(const uint64_t *hist, size_t n_bytes, double *avg, double *stddev, double *avg_stddev);
int cw_digest_stats(cw_ctx *ctx, const void *digests, size_t digest_bytes, size_t n, size_t n_bytes, double *avg,
                    double *stddev, double *avg_stddev);
int cw_dedup_dev(cw_ctx *ctx, const void *d_digests, size_t digest_bytes, size_t n, uint32_t *d_first, uint64_t *d_n_unique,
                 void *stream);

/* ---- asynchronous offload: HashOffload::{Reset,Enqueue,Start,Complete} (HashOffload.h:15-64) and the
 * consumer loop hashing_offload_entry_point (HashAndCompress.cpp:160-183) ----
 * cw_submit = Reset+Enqueue+Start: queues one fused batch (host buffers) and returns at once;
 * cw_wait   = Complete: blocks until that batch's outputs are in the caller's buffers.
 * on_complete mirrors HashOffload's onComplete callback; it runs on a library thread. */
typedef uint64_t cw_ticket;
typedef void (*cw_complete_fn)(cw_ticket ticket, int status, void *user);
int cw_submit(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src, size_t chunk_bytes,
              size_t n_chunks, void *digests, void *dst, size_t dst_stride, uint32_t *lens,
              cw_complete_fn on_complete, void *user, cw_ticket *ticket);
int cw_wait(cw_ctx *ctx, cw_ticket ticket);

/* ---- workload synthesis on the device (include/cw_synth.h); not part of the reference path ---- */
int cw_synth_dev(cw_ctx *ctx, void *d_dst, size_t chunk_bytes, uint64_t first_chunk, size_t n_chunks, uint64_t seed,
                 void *stream);

/* ---- introspection for the bench harness ---- */
/* kernels launched by this context since creation (bench.py's gpu_launches = difference around the
 * timed region) */
uint64_t cw_launch_count(const cw_ctx *ctx);
/* small-batch front-end: host-buffer calls (cw_hash_compress, cw_submit) of up to 4 MiB that arrive from several
 * threads -- the reference's --c-threads workers each hand over one readBlockFactor x 4 KiB work item at a time
 * (HashAndCompress.cpp:160-183,263-272) -- are gathered into shared batches of one copy in / one launch / one set of
 * copies out.  batches = launches made for such calls, calls = calls served by them. */
void cw_batch_stats(const cw_ctx *ctx, uint64_t *batches, uint64_t *calls);
/* name + duration of the dominant kernel: the library records CUDA events around the kernel launches of
 * the *_dev calls when profiling is on; cw_prof_read returns the accumulated milliseconds and launches */
int cw_prof_enable(cw_ctx *ctx, int on);
int cw_prof_read(cw_ctx *ctx, double *kernel_ms, uint64_t *launches); /* synchronises ctx's stream */

#ifdef __cplusplus
}
#endif
#endif /* COMPUTE_WAR_H */
