// decode_warp.cuh -- LZ4 / LZF block DEcompression of one chunk by one warp.
//
// Device-side counterpart of the reference's decoders, which the tools call right after compressing
// (experiment.cpp:256 LZ4_decompress_safe, :118 lzf_decompress; decls lz4.h:141-149, lzf.h:98-100).
// the step that follows the path on a real storage read, and what lets the
// bench verify every block of a 2 GiB shard on the device instead of sampling.
//
// Both formats are a serial stream of (literal run, back-reference) pairs; the control bytes are read by all
// lanes (shared-memory broadcast), the copies are done 32 bytes per step by the warp.  A back-reference closer
// than its length is a periodic pattern: byte x comes from (x mod offset), which only touches bytes that were
// complete before this sequence started.  Bounds-checked like the "safe" decoders: malformed input returns 0.
#pragma once
#include "cw_device.cuh"

// in: compressed block (len bytes) in shared memory; out: chunk buffer (cap bytes) in shared memory.
// Returns the number of bytes produced, 0 on malformed input (same value in every lane).
CW_DEV uint32_t cw_lz4_decode_warp(const uint8_t *in, uint32_t len, uint8_t *out, uint32_t cap)
{
    const unsigned lane = cw_lane();
    uint32_t ip = 0, op = 0;
    if (len == 0) return 0;
    for (;;) {
        if (ip >= len) return 0;
        const uint32_t token = in[ip++];
        uint32_t lit = token >> 4;
        if (lit == 15u) {
            uint32_t b;
            do {
                if (ip >= len) return 0;
                b = in[ip++];
                lit += b;
            } while (b == 255u);
        }
        if (lit > len - ip || lit > cap - op) return 0;
        for (uint32_t x = lane; x < lit; x += 32u) out[op + x] = in[ip + x];
        ip += lit;
        op += lit;
        if (ip == len) break;  // the last sequence has literals only
        if (len - ip < 2u) return 0;
        const uint32_t off = in[ip] | ((uint32_t)in[ip + 1u] << 8);
        ip += 2;
        if (off == 0u || off > op) return 0;
        uint32_t ml = token & 15u;
        if (ml == 15u) {
            uint32_t b;
            do {
                if (ip >= len) return 0;
                b = in[ip++];
                ml += b;
            } while (b == 255u);
        }
        ml += 4u;
        if (ml > cap - op) return 0;
        cw_syncwarp();  // the literals just written may be part of the match source
        const uint32_t from = op - off;
        for (uint32_t x = lane; x < ml; x += 32u) out[op + x] = out[from + (off >= ml ? x : x % off)];
        op += ml;
        cw_syncwarp();
    }
    cw_syncwarp();
    return op;
}

CW_DEV uint32_t cw_lzf_decode_warp(const uint8_t *in, uint32_t len, uint8_t *out, uint32_t cap)
{
    const unsigned lane = cw_lane();
    uint32_t ip = 0, op = 0;
    if (len == 0) return 0;
    do {
        const uint32_t ctrl = in[ip++];
        if (ctrl < 32u) {  // literal run of ctrl + 1 bytes
            const uint32_t run = ctrl + 1u;
            if (run > cap - op || run > len - ip) return 0;
            if (lane < run) out[op + lane] = in[ip + lane];
            ip += run;
            op += run;
        } else {  // back reference: 3..264 bytes, 1..8192 back
            uint32_t l = ctrl >> 5;
            if (ip >= len) return 0;
            if (l == 7u) {
                l += in[ip++];
                if (ip >= len) return 0;
            }
            const uint32_t off = ((ctrl & 0x1fu) << 8) + in[ip++] + 1u;
            const uint32_t ml = l + 2u;
            if (ml > cap - op || off > op) return 0;
            cw_syncwarp();
            const uint32_t from = op - off;
            for (uint32_t x = lane; x < ml; x += 32u) out[op + x] = out[from + (off >= ml ? x : x % off)];
            op += ml;
        }
        cw_syncwarp();
    } while (ip < len);
    return op;
}
