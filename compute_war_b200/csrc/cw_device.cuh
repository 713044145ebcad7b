// cw_device.cuh -- small device-side vocabulary shared by the kernels.
//
// The per-chunk algorithms (skein.cuh, sha256.cuh, lz4_warp.cuh, lzf_warp.cuh) are written against the
// handful of names below so that the SAME source also compiles as plain C++ for the lane-accurate
// warp emulator in tests/simt/ (CW_EMULATE defined there).  That emulator is test infrastructure: it lets
// the build container, which has no GPU, check the warp algorithms bit-for-bit against the oracle before
// GPU time is spent.  The product only ever compiles the __CUDACC__ branch.
#pragma once
#include <stdint.h>

#if !defined(__CUDACC__)
// host-only build of the per-chunk algorithms (tests/simt/): scalar code compiles as is, warp code runs on
// the fibre-based warp emulator
#define CW_DEV static inline
#define CW_DEV_NI static
#define CW_HOSTDEV static inline
#if defined(CW_EMULATE)
#include "warp_emu.h"  // provides cw_lane(), cw_ballot(), cw_shfl(), cw_match_any(), cw_syncwarp(), ...
#endif
#else
#define CW_DEV __device__ __forceinline__
#define CW_DEV_NI __device__
#define CW_HOSTDEV __host__ __device__ __forceinline__
#define CW_FULL 0xffffffffu
CW_DEV unsigned cw_lane() { return threadIdx.x & 31u; }
CW_DEV unsigned cw_ballot(bool p) { return __ballot_sync(CW_FULL, p); }
CW_DEV uint32_t cw_shfl(uint32_t v, unsigned src) { return __shfl_sync(CW_FULL, v, src); }
CW_DEV unsigned cw_match_any(uint32_t v) { return __match_any_sync(CW_FULL, v); }
CW_DEV uint32_t cw_reduce_or(uint32_t v) { return __reduce_or_sync(CW_FULL, v); }  // REDUX.OR
CW_DEV void cw_syncwarp() { __syncwarp(); }
CW_DEV int cw_ffs(uint32_t v) { return __ffs(v); }       // 1-based, 0 if none
CW_DEV int cw_clz(uint32_t v) { return __clz(v); }
CW_DEV int cw_popc(uint32_t v) { return __popc(v); }
CW_DEV uint32_t cw_brev(uint32_t v) { return __brev(v); }
CW_DEV uint32_t cw_funnel_r(uint32_t lo, uint32_t hi, unsigned sh) { return __funnelshift_r(lo, hi, sh); }
// read-only global loads
CW_DEV uint32_t cw_ldg32(const uint32_t *p) { return __ldg(p); }
CW_DEV uint32_t cw_ldg8(const uint8_t *p) { return __ldg(p); }
// inter-warp communication through shared memory (parser -> emitter queues)
CW_DEV uint32_t cw_ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
CW_DEV void cw_st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
CW_DEV void cw_fence_cta() { __threadfence_block(); }
CW_DEV void cw_spin() { __nanosleep(40); }
#endif
