// role_map.h -- which role a physical warp of a warp-specialised CTA plays (host side builds the map, kernels read it).
#pragma once
#include <stdint.h>

// Which role a physical warp plays.  Warp w issues from scheduler w % 4.  A hash warp is the one role that wants issue
// slots all the time, the emitters mostly wait, the parsers are what the pass waits for: hash warps are therefore put
// on schedulers of their own as far as possible, with the emitters for company, and the parsers share the others
// (compat: scheduler 3 = {hash, emitter, emitter}, 3 parsers on each of the other three).
struct CwRoleMap {
    uint8_t logical[32];  // logical index: [0, parsers) parser, [parsers, parsers + emitters) emitter, then hash
};

// hs / es: scheduler of every hash warp / emitter warp (nullptr: roles in warp order)
static inline CwRoleMap make_role_map(int parsers, int emitters, int hash_warps, const int *hs, const int *es)
{
    CwRoleMap m;
    const int W = parsers + emitters + hash_warps;
    int role[32];
    for (int i = 0; i < 32; i++) role[i] = -1;
    auto pick_any = [&](int s) {  // highest free physical warp of scheduler s, else of the next scheduler down
        for (int k = 0; k < 4; k++)
            for (int p = W - 1; p >= 0; p--)
                if (p % 4 == (s + 4 - k) % 4 && role[p] < 0) return p;
        return -1;
    };
    if (hs && es) {
        for (int i = 0; i < hash_warps; i++) role[pick_any(hs[i])] = parsers + emitters + i;
        for (int i = 0; i < emitters; i++) role[pick_any(es[i])] = parsers + i;
    }
    int next = 0;
    for (int p = 0; p < W; p++) {
        if (!(hs && es)) role[p] = p;
        else if (role[p] < 0) role[p] = next++;
    }
    for (int i = 0; i < 32; i++) m.logical[i] = (uint8_t)(i < W ? role[i] : 0);
    return m;
}

