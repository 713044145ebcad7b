// synth.cu -- on-device generator of the synthetic mixed-entropy workload (include/cw_synth.h).
// Workload synthesis for the bench harness (BASELINE.json configs[4]); not part of the reference path.
#include "../../include/cw_synth.h"
#include "launch.h"

__global__ void __launch_bounds__(128) cw_synth_kernel(uint8_t *dst, uint32_t chunk, uint64_t first, uint64_t n, uint64_t seed)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    cw_synth_chunk(dst + g * chunk, chunk, first + g, seed);
}

cudaError_t cw_launch_synth(uint8_t *d_dst, uint32_t chunk, uint64_t first_chunk, uint64_t n, uint64_t seed, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    cw_synth_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_dst, chunk, first_chunk, n, seed);
    return cudaGetLastError();
}
