// lzf_kernels.cu -- placeholder until the LZF warp kernel lands (next milestone); fails loudly.
#include "launch.h"
cudaError_t cw_launch_lzf(const CwDeviceInfo &, const uint8_t *, uint32_t, uint64_t, uint8_t *, uint64_t, uint32_t *,
                          unsigned long long *, cudaStream_t)
{
    return cudaErrorNotSupported;
}
cudaError_t cw_lzf_configure(const CwDeviceInfo &) { return cudaSuccess; }
