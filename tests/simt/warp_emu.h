// warp_emu.h -- lane-accurate emulation of CUDA warps on the CPU (TEST INFRASTRUCTURE ONLY).
//
// The build container has no GPU.  To check the warp-cooperative kernels (csrc/lz4_warp.cuh,
// csrc/lzf_warp.cuh) bit-for-bit against the oracle before spending GPU time, the same device source is
// compiled as plain C++ and run here: every warp is 32 fibres (ucontext) executing the per-lane code, and
// every warp collective (ballot / shfl / match_any / syncwarp) is a rendezvous of the 32 fibres of that warp.
// Several warps can run together (a CTA with specialised warps talking through "shared memory"); a fibre
// that spins on a flag yields with cw_spin().  A collective reached by only part of a warp (divergence
// bug) or a set of warps that all spin forever is reported as a deadlock instead of hanging.
// Nothing in the product links against this file.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

namespace cwemu {

struct Warp {
    uint32_t slot[32];     // values deposited for the current collective
    uint32_t snap[2][32];  // published values of the last two collectives
    int arrived;
    uint64_t generation;
};

struct Block {
    ucontext_t sched;
    int nwarps;
    Warp *warps;
    ucontext_t *ctx;  // nwarps * 32
    char **stack;
    bool *finished;
    int cur;  // running fibre = warp * 32 + lane
    uint64_t progress;  // bumped by every completed collective / finished fibre / explicit progress mark
    void (*body)(void *, int warp);
    void *arg;
};

extern thread_local Block *g_block;
// Which of several stores of one warp to the same address survives is undefined in CUDA; the emulator can run the lanes
// in either order so that algorithms that tolerate any winner are checked with both extremes.
extern int g_descending_lanes;

inline void yield_fibre()
{
    Block *b = g_block;
    swapcontext(&b->ctx[b->cur], &b->sched);
}

// rendezvous of the calling fibre's warp: every lane deposits v; returns the snapshot array of this collective
inline const uint32_t *collective(uint32_t v)
{
    Block *b = g_block;
    Warp *w = &b->warps[b->cur >> 5];
    const int me = b->cur & 31;
    w->slot[me] = v;
    const uint64_t gen = w->generation;
    if (++w->arrived == 32) {
        memcpy(w->snap[gen & 1], w->slot, sizeof w->slot);
        w->arrived = 0;
        w->generation = gen + 1;
        b->progress++;
    } else {
        while (w->generation == gen) yield_fibre();
    }
    return w->snap[gen & 1];
}

// run nwarps warps; body(arg, warp) is executed by each of the 32 lanes of every warp
void run_block(int nwarps, void (*body)(void *, int), void *arg);
// single warp convenience
void run_warp(void (*body)(void *), void *arg);

}  // namespace cwemu

static inline unsigned cw_lane() { return (unsigned)(cwemu::g_block->cur & 31); }
static inline unsigned cw_ballot(bool p)
{
    const uint32_t *s = cwemu::collective(p ? 1u : 0u);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (s[i] & 1u) << i;
    return m;
}
static inline uint32_t cw_shfl(uint32_t v, unsigned src) { return cwemu::collective(v)[src & 31u]; }
static inline unsigned cw_match_any(uint32_t v)
{
    const uint32_t *s = cwemu::collective(v);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(s[i] == v) << i;
    return m;
}
static inline uint32_t cw_reduce_or(uint32_t v)
{
    const uint32_t *s = cwemu::collective(v);
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) m |= s[i];
    return m;
}
static inline void cw_syncwarp() { (void)cwemu::collective(0); }
static inline int cw_ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int cw_clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int cw_popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t cw_brev(uint32_t v)
{
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline uint32_t cw_funnel_r(uint32_t lo, uint32_t hi, unsigned sh)
{
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
}
static inline uint32_t cw_ldg32(const uint32_t *p) { return *p; }
static inline uint32_t cw_ldg8(const uint8_t *p) { return *p; }
// inter-warp communication through shared memory
static inline uint32_t cw_ld_vol(const uint32_t *p) { return *(const volatile uint32_t *)p; }
static inline void cw_st_vol(uint32_t *p, uint32_t v)
{
    *(volatile uint32_t *)p = v;
    cwemu::g_block->progress++;
}
static inline void cw_fence_cta() { __sync_synchronize(); }
static inline void cw_spin() { cwemu::yield_fibre(); }
