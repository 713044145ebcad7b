// emu_codecs.cpp -- runs the product's warp-cooperative compressors (csrc/lz4_warp.cuh, csrc/lzf_warp.cuh)
// on the CPU warp emulator.  Test infrastructure only; exports a C entry point per codec for ctypes.
#define CW_EMULATE 1
#include <vector>
#include "warp_emu.h"
#include "../../compute_war_b200/csrc/lz4_warp.cuh"
#if __has_include("../../compute_war_b200/csrc/lzf_warp.cuh")
#include "../../compute_war_b200/csrc/lzf_warp.cuh"
#define HAVE_LZF 1
#endif

struct Job {
    const uint8_t *s;
    uint32_t n;
    uint16_t *table;
    void *aux;
    uint8_t *out;
    uint32_t out_cap;
    uint32_t result;
    int mode;
};

static void lz4_body(void *a)
{
    Job *j = (Job *)a;
    uint32_t r = j->mode == 12 ? cw_lz4_compress_warp<12>(j->s, j->n, j->table, j->out)
                               : cw_lz4_compress_warp<13>(j->s, j->n, j->table, j->out);
    if (cw_lane() == 0) j->result = r;
}

extern "C" uint32_t emu_lz4_compress_hl(const uint8_t *src, uint32_t n, uint8_t *out, int hashlog)
{
    std::vector<uint32_t> staged((n + CW_LZ4_PAD + 3) / 4 + 4, 0);  // 4-byte aligned, zero padded like the kernel's smem
    memcpy(staged.data(), src, n);
    std::vector<uint16_t> table(1u << hashlog, 0);
    Job j{(const uint8_t *)staged.data(), n, table.data(), nullptr, out, 0, 0, hashlog};
    cwemu::run_warp(lz4_body, &j);
    return j.result;
}

extern "C" uint32_t emu_lz4_compress(const uint8_t *src, uint32_t n, uint8_t *out) { return emu_lz4_compress_hl(src, n, out, 13); }

#ifdef HAVE_LZF
static void lzf_body(void *a)
{
    Job *j = (Job *)a;
    uint32_t r = cw_lzf_compress_warp(j->s, j->n, (CwLzfScratch *)j->aux, j->out, j->out_cap);
    if (cw_lane() == 0) j->result = r;
}
extern "C" uint32_t emu_lzf_compress(const uint8_t *src, uint32_t n, uint8_t *out, uint32_t out_cap)
{
    std::vector<uint32_t> staged((n + CW_LZ4_PAD + 3) / 4 + 4, 0);
    memcpy(staged.data(), src, n);
    std::vector<uint8_t> scratch(sizeof(CwLzfScratch), 0);
    Job j{(const uint8_t *)staged.data(), n, nullptr, scratch.data(), out, out_cap, 0, 0};
    cwemu::run_warp(lzf_body, &j);
    return j.result;
}
#endif
