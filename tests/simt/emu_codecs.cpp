// emu_codecs.cpp -- runs the product's warp-cooperative compressors (csrc/lz4_warp.cuh, csrc/lzf_warp.cuh)
// on the CPU warp emulator.  Test infrastructure only; exports a C entry point per codec for ctypes.
#define CW_EMULATE 1
#include <sys/mman.h>
#include <vector>
#include "warp_emu.h"
#include "../../compute_war_b200/csrc/lz4_warp.cuh"
#include "../../compute_war_b200/csrc/lz4_fast.cuh"
#include "../../compute_war_b200/csrc/decode_warp.cuh"
#if __has_include("../../compute_war_b200/csrc/lzf_warp.cuh")
#include "../../compute_war_b200/csrc/lzf_warp.cuh"
#define HAVE_LZF 1
#endif

// ring length the parser / emitter pairs below run with (the kernels use 64, or 32 where shared memory is short)
static uint32_t g_qlen = CW_LZ4_QLEN;
extern "C" void emu_set_store_winner(int lowest_lane) { cwemu::g_descending_lanes = lowest_lane ? 1 : 0; }
extern "C" void emu_set_qlen(uint32_t qlen) { g_qlen = qlen == 32u ? 32u : CW_LZ4_QLEN; }

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

// LZ4: a two-warp CTA -- warp 0 parses and pushes records, warp 1 is the emitter (csrc/lz4_warp.cuh)
struct Lz4Job {
    const uint8_t *s;
    uint32_t n;
    uint16_t *table;
    CwLz4Queue *q;
    uint8_t *out;
    uint32_t lens[1];
    int hashlog;
    int direct;
    int fast;
};

static void lz4_block_body(void *a, int warp)
{
    Lz4Job *j = (Lz4Job *)a;
    const unsigned lane = cw_lane();
    if (warp == 0) {
        uint32_t qh = 0;
        if (j->direct) {
            if (j->hashlog == 12) cw_lz4_parse_warp<12, true>(j->s, j->n, j->table, j->q, qh, g_qlen);
            else cw_lz4_parse_warp<13, true>(j->s, j->n, j->table, j->q, qh, g_qlen);
        } else if (j->hashlog == 12) cw_lz4_parse_warp<12, false>(j->s, j->n, j->table, j->q, qh, g_qlen);
        else cw_lz4_parse_warp<13, false>(j->s, j->n, j->table, j->q, qh, g_qlen);
        cw_lz4_q_drain(j->q, qh, lane);
        if (lane == 0) {
            cw_fence_cta();
            cw_st_vol(&j->q->done, 1u);
        }
    } else {
        for (;;) {
            uint32_t d = 0;
            if (lane == 0) d = cw_ld_vol(&j->q->done);
            d = cw_shfl(d, 0);
            const uint32_t cnt = j->direct ? cw_lz4_emit_batch<true>(j->q, j->s, j->n, j->out, 0, j->lens, g_qlen)
                                           : cw_lz4_emit_batch<false>(j->q, j->s, j->n, j->out, 0, j->lens, g_qlen);
            if (d && cnt == 0) break;
            if (!cnt) cw_spin();
        }
    }
}

extern "C" uint32_t emu_lz4_compress_hl(const uint8_t *src, uint32_t n, uint8_t *out, int hashlog)
{
    std::vector<uint32_t> staged((n + CW_LZ4_PAD + 3) / 4 + 4, 0);  // 4-byte aligned, zero padded like the kernel's smem
    memcpy(staged.data(), src, n);
    std::vector<uint16_t> table(1u << hashlog, 0);
    CwLz4Queue q;
    memset(&q, 0, sizeof q);
    Lz4Job j{(const uint8_t *)staged.data(), n, table.data(), &q, out, {0}, hashlog, 0, 0};
    cwemu::run_block(2, lz4_block_body, &j);
    return j.lens[0];
}

// fast mode (csrc/lz4_fast.cuh): a team of nteam warps parses and emits one chunk
struct Lz4fJob {
    const uint8_t *s;
    uint32_t n;
    uint16_t *table;
    CwLz4fTeam *team;
    CwLz4fScratch *scr;  // one per warp
    uint8_t *out;
    uint32_t len;
    int hashlog, direct, nteam, codec;
};
static void lz4f_body(void *a, int warp)
{
    Lz4fJob *j = (Lz4fJob *)a;
    uint32_t r, seen = 0;
    const uint32_t rank = (uint32_t)warp, nt = (uint32_t)j->nteam;
    if (j->codec == CW_F_LZF) {  // the LZF flavour: a block may take n - 1 bytes
        const uint32_t cap = j->n ? j->n - 1u : 0u;
        if (j->direct) r = cw_lz4f_team_chunk<CW_F_LZF, 14, true>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen, cap);
        else if (j->hashlog == 13) r = cw_lz4f_team_chunk<CW_F_LZF, 13, false>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen, cap);
        else r = cw_lz4f_team_chunk<CW_F_LZF, 14, false>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen, cap);
    } else if (j->direct) {
        if (j->hashlog == 14) r = cw_lz4f_team_chunk<CW_F_LZ4, 14, true>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen);
        else r = cw_lz4f_team_chunk<CW_F_LZ4, 13, true>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen);
    } else if (j->hashlog == 12) r = cw_lz4f_team_chunk<CW_F_LZ4, 12, false>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen);
    else if (j->hashlog == 13) r = cw_lz4f_team_chunk<CW_F_LZ4, 13, false>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen);
    else r = cw_lz4f_team_chunk<CW_F_LZ4, 14, false>(j->s, j->n, j->table, j->team, j->scr + warp, j->out, rank, nt, seen);
    if (r != CW_LZ4_NONE && cw_lane() == 0) j->len = r;
}
// hashlog 0 = what the kernels pick for the chunk size; nteam = warps per chunk
static uint32_t emu_fast(int codec, const uint8_t *src, uint32_t n, uint8_t *out, int hashlog, int direct, int nteam);
extern "C" uint32_t emu_lz4_fast(const uint8_t *src, uint32_t n, uint8_t *out, int hashlog, int direct, int nteam)
{
    return emu_fast(CW_F_LZ4, src, n, out, hashlog, direct, nteam);
}
// LZF flavour: hashlog 13 or 14 (0 = what the kernels pick); direct runs the 14-bit in-place variant
extern "C" uint32_t emu_lzf_fast(const uint8_t *src, uint32_t n, uint8_t *out, int hashlog, int direct, int nteam)
{
    return emu_fast(CW_F_LZF, src, n, out, hashlog, direct, nteam);
}
static uint32_t emu_fast(int codec, const uint8_t *src, uint32_t n, uint8_t *out, int hashlog, int direct, int nteam)
{
    if (!hashlog) hashlog = codec == CW_F_LZF ? cw_lzff_hashlog(n) : cw_lz4f_hashlog(n);
    if (codec == CW_F_LZF && direct) hashlog = 14;
    if (nteam < 1) nteam = 1;
    std::vector<uint16_t> table(1u << hashlog, 0);
    std::vector<CwLz4fScratch> scr(nteam);
    memset(scr.data(), 0xA5, sizeof(CwLz4fScratch) * nteam);
    CwLz4fTeam team;
    memset(&team, 0, sizeof team);
    team.ip = 1;
    if (!direct) {
        std::vector<uint32_t> staged((n + CW_LZ4_PAD + 3) / 4 + 4, 0);
        memcpy(staged.data(), src, n);
        Lz4fJob j{(const uint8_t *)staged.data(), n, table.data(), &team, scr.data(), out, 0xffffffffu, hashlog, 0, nteam, codec};
        cwemu::run_block(nteam, lz4f_body, &j);
        return j.len;
    }
    const size_t page = 4096, span = (n + page - 1) / page * page + page;  // chunk ends at a PROT_NONE guard page
    uint8_t *map = (uint8_t *)mmap(nullptr, span + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    mprotect(map + span, page, PROT_NONE);
    uint8_t *exact = map + span - n;
    memcpy(exact, src, n);
    Lz4fJob j{(const uint8_t *)exact, n, table.data(), &team, scr.data(), out, 0xffffffffu, hashlog, 1, nteam, codec};
    cwemu::run_block(nteam, lz4f_body, &j);
    munmap(map, span + page);
    return j.len;
}

// DIRECT mode: the chunk is read in place, nothing readable behind it (n must be a multiple of 4)
extern "C" uint32_t emu_lz4_compress_direct(const uint8_t *src, uint32_t n, uint8_t *out, int hashlog)
{
    // the chunk ends exactly at a PROT_NONE guard page: any read past n faults
    const size_t page = 4096, span = (n + page - 1) / page * page + page;
    uint8_t *map = (uint8_t *)mmap(nullptr, span + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    mprotect(map + span, page, PROT_NONE);
    uint8_t *exact = map + span - n;
    memcpy(exact, src, n);
    std::vector<uint16_t> table(1u << hashlog, 0);
    CwLz4Queue q;
    memset(&q, 0, sizeof q);
    Lz4Job j{(const uint8_t *)exact, n, table.data(), &q, out, {0}, hashlog, 1, 0};
    cwemu::run_block(2, lz4_block_body, &j);
    munmap(map, span + page);
    return j.lens[0];
}
extern "C" uint32_t emu_lz4_compress(const uint8_t *src, uint32_t n, uint8_t *out) { return emu_lz4_compress_hl(src, n, out, 13); }

#ifdef HAVE_LZF
// LZF: a two-warp CTA -- warp 0 parses, warp 1 emits (csrc/lzf_warp.cuh)
struct LzfJob {
    const uint8_t *s;
    uint32_t n;
    uint16_t *tab;
    CwLz4Queue *q;
    uint8_t *out;
    uint32_t out_cap;
    uint32_t lens[1];
    int tagged;
};
static void lzf_block_body(void *a, int warp)
{
    LzfJob *j = (LzfJob *)a;
    const unsigned lane = cw_lane();
    if (warp == 0) {
        uint32_t qh = 0;
        if (j->tagged == 12 || j->tagged == 25) cw_lzf_table_init<12>(j->tab);
        else if (j->tagged == 24) cw_lzf_table_init<11>(j->tab);
        else if (j->tagged == 13) cw_lzf_table_init<13>(j->tab);
        else cw_lzf_table_init<0>(j->tab);
        cw_syncwarp();
        if (j->tagged == 12) cw_lzf_parse_warp<12, 12>(j->s, j->n, j->tab, j->q, qh, g_qlen);
        else if (j->tagged == 25) cw_lzf_parse_warp<12, 13>(j->s, j->n, j->tab, j->q, qh, g_qlen);
        else if (j->tagged == 24) cw_lzf_parse_warp<11, 13>(j->s, j->n, j->tab, j->q, qh, g_qlen);
        else if (j->tagged == 13) cw_lzf_parse_warp<13, 13>(j->s, j->n, j->tab, j->q, qh, g_qlen);
        else cw_lzf_parse_warp<0, 13>(j->s, j->n, j->tab, j->q, qh, g_qlen);
        cw_lz4_q_drain(j->q, qh, lane);
        if (lane == 0) {
            cw_fence_cta();
            cw_st_vol(&j->q->done, 1u);
        }
    } else {
        for (;;) {
            uint32_t d = 0;
            if (lane == 0) d = cw_ld_vol(&j->q->done);
            d = cw_shfl(d, 0);
            const uint32_t cnt = cw_lzf_emit_batch(j->q, j->s, j->n, j->out_cap, j->out, 0, j->lens, g_qlen);
            if (d && cnt == 0) break;
            if (!cnt) cw_spin();
        }
    }
}
// tagged: table format of lzf_warp.cuh (12, 13, 25 or 0), -1 = what the kernels pick
extern "C" uint32_t emu_lzf_compress_mode(const uint8_t *src, uint32_t n, uint8_t *out, uint32_t out_cap, int tagged)
{
    if (n == 0) return 0;
    std::vector<uint32_t> staged((n + CW_LZ4_PAD + 3) / 4 + 4, 0);
    memcpy(staged.data(), src, n);
    std::vector<uint8_t> scratch(1, 0xA5);  // sized below; contents irrelevant on entry
    CwLz4Queue q;
    memset(&q, 0, sizeof q);
    if (tagged < 0) tagged = cw_lzf_table_format(n);
    scratch.assign(cw_lzf_table_bytes(tagged), 0xA5);  // exactly what the kernels give the format
    LzfJob j{(const uint8_t *)staged.data(), n, (uint16_t *)scratch.data(), &q, out, out_cap, {0}, tagged};
    cwemu::run_block(2, lzf_block_body, &j);
    return j.lens[0];
}
extern "C" uint32_t emu_lzf_compress(const uint8_t *src, uint32_t n, uint8_t *out, uint32_t out_cap)
{
    return emu_lzf_compress_mode(src, n, out, out_cap, -1);
}
#endif

// decoders (csrc/decode_warp.cuh): one warp per block
struct DecJob {
    const uint8_t *in;
    uint32_t len;
    uint8_t *out;
    uint32_t cap, result;
    int alg;
};
static void dec_body(void *a)
{
    DecJob *j = (DecJob *)a;
    uint32_t r = j->alg == 0 ? cw_lz4_decode_warp(j->in, j->len, j->out, j->cap) : cw_lzf_decode_warp(j->in, j->len, j->out, j->cap);
    if (cw_lane() == 0) j->result = r;
}
extern "C" uint32_t emu_decode(int alg, const uint8_t *in, uint32_t len, uint8_t *out, uint32_t cap)
{
    DecJob j{in, len, out, cap, 0, alg};
    cwemu::run_warp(dec_body, &j);
    return j.result;
}
