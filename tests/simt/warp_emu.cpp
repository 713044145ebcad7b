// warp_emu.cpp -- scheduler of the CPU warp emulator (test infrastructure, see warp_emu.h)
#include "warp_emu.h"

namespace cwemu {

thread_local Block *g_block = nullptr;
int g_descending_lanes = 0;  // 1: fibres of a warp run from lane 31 down, so of several same-address stores lane 0's survives

static void trampoline()
{
    Block *b = g_block;
    const int me = b->cur;
    b->body(b->arg, me >> 5);
    b->finished[me] = true;
    b->progress++;
    swapcontext(&b->ctx[me], &b->sched);
}

void run_block(int nwarps, void (*body)(void *, int), void *arg)
{
    static const size_t STACK = 256 * 1024;
    const int nf = nwarps * 32;
    Block *b = new Block();
    b->nwarps = nwarps;
    b->warps = new Warp[nwarps]();
    b->ctx = new ucontext_t[nf];
    b->stack = new char *[nf];
    b->finished = new bool[nf]();
    b->progress = 0;
    b->body = body;
    b->arg = arg;
    g_block = b;
    for (int i = 0; i < nf; i++) {
        b->stack[i] = (char *)malloc(STACK);
        getcontext(&b->ctx[i]);
        b->ctx[i].uc_stack.ss_sp = b->stack[i];
        b->ctx[i].uc_stack.ss_size = STACK;
        b->ctx[i].uc_link = &b->sched;
        makecontext(&b->ctx[i], trampoline, 0);
    }
    int live = nf, idle_rounds = 0;
    while (live) {
        const uint64_t p0 = b->progress;
        for (int k = 0; k < nf; k++) {
            const int i = g_descending_lanes ? (k & ~31) + 31 - (k & 31) : k;
            if (b->finished[i]) continue;
            b->cur = i;
            swapcontext(&b->sched, &b->ctx[i]);
            if (b->finished[i]) live--;
        }
        if (live && b->progress == p0) {
            if (++idle_rounds > 64) {
                fprintf(stderr, "warp_emu: deadlock -- %d fibres alive, none makes progress (divergent warp collective "
                                "or warps spinning on each other)\n", live);
                abort();
            }
        } else
            idle_rounds = 0;
    }
    for (int i = 0; i < nf; i++) free(b->stack[i]);
    delete[] b->warps;
    delete[] b->ctx;
    delete[] b->stack;
    delete[] b->finished;
    delete b;
    g_block = nullptr;
}

struct OneWarp {
    void (*body)(void *);
    void *arg;
};
static void one_warp_body(void *a, int) { ((OneWarp *)a)->body(((OneWarp *)a)->arg); }

void run_warp(void (*body)(void *), void *arg)
{
    OneWarp w{body, arg};
    run_block(1, one_warp_body, &w);
}

}  // namespace cwemu
