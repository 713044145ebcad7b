// skein_probe.cu -- development probe: Skein-512 of one short message through both loop forms of cw_skein_chunk
// (PREFETCH on / off) on the GPU, against the same source compiled for the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -o gpurun_out/skein_probe tools/skein_probe.cu
#include <cstdio>
#include <cstring>
#include "compute_war_b200/csrc/cw_device.cuh"
#include "compute_war_b200/csrc/skein.cuh"

struct IV { uint64_t iv[8]; };
template <bool PF, bool RESTRICT>
__global__ void k(const __grid_constant__ IV iv, const uint8_t *src, uint32_t len, uint8_t *out)
{
    uint64_t X[8];
    cw_skein_chunk<8, 64, false, true, PF>(iv.iv, src, len, X);
    for (int i = 0; i < 64; i++) out[i] = (uint8_t)(X[i >> 3] >> (8 * (i & 7)));
}
template <bool PF>
__global__ void kr(const __grid_constant__ IV iv, const uint8_t *__restrict__ src, uint32_t len, uint8_t *__restrict__ out)
{
    uint64_t X[8];
    cw_skein_chunk<8, 64, false, true, PF>(iv.iv, src, len, X);
    for (int i = 0; i < 64; i++) out[i] = (uint8_t)(X[i >> 3] >> (8 * (i & 7)));
}
int main()
{
    IV iv;
    cw_skein_iv<8>(iv.iv, 512);
    uint8_t msg[256], *d;
    for (int i = 0; i < 256; i++) msg[i] = (uint8_t)(i * 7 + 3);
    cudaMalloc(&d, 1024);
    int bad[4] = {0, 0, 0, 0};
    for (uint32_t len = 1; len <= 130; len++) {
        uint64_t X[8];
        cw_skein_chunk<8, 64, false, true, false>(iv.iv, msg, len, X);
        uint8_t ref[64], got[64];
        for (int i = 0; i < 64; i++) ref[i] = (uint8_t)(X[i >> 3] >> (8 * (i & 7)));
        for (int v = 0; v < 4; v++) {
            cudaMemcpy(d, msg, 256, cudaMemcpyHostToDevice);
            uint8_t *o = (v & 2) ? d + len : d + 512;  // output right behind the message (same line) or far away
            if (v & 1) kr<false><<<1, 1>>>(iv, d, len, o);
            else kr<true><<<1, 1>>>(iv, d, len, o);
            cudaMemcpy(got, o, 64, cudaMemcpyDeviceToHost);
            if (memcmp(got, ref, 64)) {
                if (bad[v]++ < 12) printf("variant %d (prefetch %d, out %s) len %u differs\n", v, !(v & 1), (v & 2) ? "adjacent" : "far", len);
            }
        }
    }
    printf("bad per variant: pf/far %d  nopf/far %d  pf/adjacent %d  nopf/adjacent %d  (%s)\n", bad[0], bad[1], bad[2], bad[3],
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
