// launch.h -- internal (C++) interface between the C-ABI layer (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct CwSkeinIV;

struct CwDeviceInfo {
    int sm_count;
    size_t smem_optin;  // max dynamic shared memory per block
};

// hash_kernels.cu
cudaError_t cw_launch_hash(int alg, const CwSkeinIV &iv, const uint8_t *d_src, uint32_t chunk, uint64_t n,
                           uint8_t *d_digests, cudaStream_t st);
void cw_make_skein_iv(CwSkeinIV *iv);
// chaining values for Skein-<state_bits> with a digest of hash_bits; alg 3 of cw_launch_hash = Skein-256, whole 32-byte block
void cw_make_skein_iv_bits(CwSkeinIV *iv, unsigned state_bits, unsigned hash_bits);

// lz4_kernels.cu
// work: device u64 zeroed by the caller on `st` before the launch (dynamic chunk queue)
cudaError_t cw_launch_lz4(const CwDeviceInfo &dev, bool fast, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                          uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st);
cudaError_t cw_lz4_configure(const CwDeviceInfo &dev);
// fast mode: teams (<= slots) and compress warps (<= max_warps) of one CTA (CW_LZ4F_TEAM forces a team size)
void cw_lz4f_shape(int slots, int max_warps, int *teams, int *warps);

// synth.cu
cudaError_t cw_launch_synth(uint8_t *d_dst, uint32_t chunk, uint64_t first_chunk, uint64_t n, uint64_t seed,
                            cudaStream_t st);

// lzf_kernels.cu
cudaError_t cw_launch_lzf(const CwDeviceInfo &dev, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                          uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st);
cudaError_t cw_lzf_configure(const CwDeviceInfo &dev);
// lz4_kernels.cu: LZF fast mode runs the team kernel of lz4_fast.cuh
cudaError_t cw_launch_lzf_fast(const CwDeviceInfo &dev, const uint8_t *d_src, uint32_t chunk, uint64_t n, uint8_t *d_dst,
                               uint64_t stride, uint32_t *d_lens, unsigned long long *work, cudaStream_t st);

// fused_kernels.cu -- ProcessBlock: digest + compressed block of every chunk in one pass
// work: two device u64 counters zeroed by the caller
cudaError_t cw_launch_fused(const CwDeviceInfo &dev, int halg, int calg, bool fast, const CwSkeinIV &iv, const uint8_t *d_src,
                            uint32_t chunk, uint64_t n, uint8_t *d_digests, uint8_t *d_dst, uint64_t stride,
                            uint32_t *d_lens, unsigned long long *work, cudaStream_t st);
cudaError_t cw_fused_configure(const CwDeviceInfo &dev);

// decode_kernels.cu
cudaError_t cw_launch_decode(const CwDeviceInfo &dev, int alg, const uint8_t *d_blocks, uint64_t stride, const uint32_t *d_lens,
                             uint32_t chunk, uint64_t n, uint8_t *d_out, uint32_t *d_out_lens, unsigned long long *work,
                             cudaStream_t st);
cudaError_t cw_decode_configure(const CwDeviceInfo &dev);

// pack_kernels.cu -- d_offsets: u64[n + 1]; d_scratch: u64[ceil(n / 1024) + 1]
cudaError_t cw_launch_pack(const uint8_t *d_blocks, uint64_t stride, const uint32_t *d_lens, const uint8_t *d_src, uint32_t chunk,
                           uint64_t n, uint8_t *d_packed, uint64_t packed_cap, unsigned long long *d_offsets,
                           unsigned long long *d_scratch, cudaStream_t st);

// digest_kernels.cu -- d_hist: u64[D * 256] (zeroed by the launch); d_table: u32[table_slots], table_slots a power of
// two >= 2 * n; D % 4 == 0 for the dedup kernels
cudaError_t cw_launch_digest_hist(const CwDeviceInfo &dev, const uint8_t *d_digests, uint32_t D, uint64_t n,
                                  unsigned long long *d_hist, cudaStream_t st);
cudaError_t cw_launch_dedup(const uint8_t *d_digests, uint32_t D, uint64_t n, uint32_t *d_table, uint32_t table_slots,
                            uint32_t *d_first, unsigned long long *d_n_unique, cudaStream_t st);

// all per-kernel attribute set-up (opt-in shared memory sizes); called once from cw_init
cudaError_t cw_configure_kernels(const CwDeviceInfo &dev);
