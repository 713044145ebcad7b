// fused_kernels.cu -- ProcessBlock (HashAndCompress.cpp:231-261): compressed block + digest of every chunk.
// INTERIM (round-1 first GPU run): compress kernel followed by the hash kernel on the same stream; the
// single-pass warp-specialised kernel replaces this once the stand-alone kernels are measured.
#include "hash_kernels.cuh"
#include "launch.h"

cudaError_t cw_launch_fused(const CwDeviceInfo &dev, int halg, int calg, const CwSkeinIV &iv, const uint8_t *d_src,
                            uint32_t chunk, uint64_t n, uint8_t *d_digests, uint8_t *d_dst, uint64_t stride,
                            uint32_t *d_lens, unsigned long long *work, cudaStream_t st)
{
    cudaError_t e = calg == 0 ? cw_launch_lz4(dev, d_src, chunk, n, d_dst, stride, d_lens, work, st)
                              : cw_launch_lzf(dev, d_src, chunk, n, d_dst, stride, d_lens, work, st);
    if (e != cudaSuccess) return e;
    return cw_launch_hash(halg, iv, d_src, chunk, n, d_digests, st);
}

cudaError_t cw_fused_configure(const CwDeviceInfo &) { return cudaSuccess; }

cudaError_t cw_configure_kernels(const CwDeviceInfo &dev)
{
    cudaError_t e = cw_lz4_configure(dev);
    if (e == cudaSuccess) e = cw_lzf_configure(dev);
    if (e == cudaSuccess) e = cw_fused_configure(dev);
    return e;
}
