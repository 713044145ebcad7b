// warp_emu.cpp -- scheduler of the CPU warp emulator (test infrastructure, see warp_emu.h)
#include "warp_emu.h"

namespace cwemu {

thread_local Warp *g_warp = nullptr;

static void trampoline()
{
    Warp *w = g_warp;
    w->body(w->arg);
    w->finished[w->cur] = true;
    swapcontext(&w->ctx[w->cur], &w->sched);
}

void run_warp(void (*body)(void *), void *arg)
{
    static const size_t STACK = 256 * 1024;
    Warp *w = new Warp();
    memset(w->finished, 0, sizeof w->finished);
    w->arrived = 0;
    w->generation = 0;
    w->body = body;
    w->arg = arg;
    g_warp = w;
    for (int i = 0; i < 32; i++) {
        w->stack[i] = (char *)malloc(STACK);
        getcontext(&w->ctx[i]);
        w->ctx[i].uc_stack.ss_sp = w->stack[i];
        w->ctx[i].uc_stack.ss_size = STACK;
        w->ctx[i].uc_link = &w->sched;
        makecontext(&w->ctx[i], trampoline, 0);
    }
    int live = 32;
    while (live) {
        const uint64_t gen0 = w->generation;
        const int live0 = live;
        for (int i = 0; i < 32; i++) {
            if (w->finished[i]) continue;
            w->cur = i;
            swapcontext(&w->sched, &w->ctx[i]);
            if (w->finished[i]) live--;
        }
        if (live && live == live0 && w->generation == gen0) {
            fprintf(stderr, "warp_emu: deadlock -- %d lanes wait at a collective the others never reach "
                            "(divergent warp collective)\n", w->arrived);
            abort();
        }
        if (live && live != 32 && live != 0 && w->arrived + (32 - live) == 32 && w->generation == gen0) {
            fprintf(stderr, "warp_emu: %d lanes exited while others wait at a collective\n", 32 - live);
            abort();
        }
    }
    for (int i = 0; i < 32; i++) free(w->stack[i]);
    delete w;
    g_warp = nullptr;
}

}  // namespace cwemu
