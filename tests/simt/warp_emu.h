// warp_emu.h -- lane-accurate emulation of one CUDA warp on the CPU (TEST INFRASTRUCTURE ONLY).
//
// The build container has no GPU.  To check the warp-cooperative kernels (csrc/lz4_warp.cuh,
// csrc/lzf_warp.cuh) bit-for-bit against the oracle before spending GPU time, the same device source is
// compiled as plain C++ and run here: 32 fibres (ucontext) execute the per-lane code, and every warp
// collective (ballot / shfl / match_any / syncwarp) is a rendezvous of all 32 fibres.  A collective reached
// by only part of the warp (divergence bug) is reported as a deadlock instead of hanging.
// Nothing in the product links against this file.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

namespace cwemu {

struct Warp {
    ucontext_t sched;
    ucontext_t ctx[32];
    char *stack[32];
    bool finished[32];
    int cur;                 // running lane
    uint32_t slot[32];       // values deposited for the current collective
    uint32_t snap[2][32];    // published values of the last two collectives
    int arrived;
    uint64_t generation;
    void (*body)(void *);
    void *arg;
};

extern thread_local Warp *g_warp;

inline void yield_lane()
{
    Warp *w = g_warp;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

// rendezvous: every live lane deposits v; returns the snapshot array of this collective
inline const uint32_t *collective(uint32_t v)
{
    Warp *w = g_warp;
    const int me = w->cur;
    w->slot[me] = v;
    const uint64_t gen = w->generation;
    if (++w->arrived == 32) {
        memcpy(w->snap[gen & 1], w->slot, sizeof w->slot);
        w->arrived = 0;
        w->generation = gen + 1;
    } else {
        while (w->generation == gen) yield_lane();
    }
    return w->snap[gen & 1];
}

void run_warp(void (*body)(void *), void *arg);

}  // namespace cwemu

static inline unsigned cw_lane() { return (unsigned)cwemu::g_warp->cur; }
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
static inline void cw_syncwarp() { (void)cwemu::collective(0); }
static inline int cw_ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int cw_clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int cw_popc(uint32_t v) { return __builtin_popcount(v); }
static inline uint32_t cw_funnel_r(uint32_t lo, uint32_t hi, unsigned sh)
{
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
}
