// hashandcompress.cpp -- host tool with the command line of the reference's `hashAndCompress`
// (src/hashandcompress/HashAndCompress.cpp:290-415), driving the B200 engine through the C ABI.
//
// Same flags (--c-threads/-c, --gpu-offload/-g, --read-blocks/-r, --hash-blocks/-G, --comp-alg/-C,
// --hash-alg/-H, positional input files; HashAndCompress.cpp:301-323) and the same result line
// `hash|comp|ms|MBps` (:406-412).  Differences, all deliberate:
//   * files are cut into whole chunks of --chunk bytes (default 4096 = blockSize, HashAndCompress.h) and
//     the tail is dropped (ReadFile :197-215), but the bytes are kept as bytes -- the reference pushes
//     std::string(rawData) and so truncates binary data at the first NUL (:213);
//   * every chunk is compressed (the reference re-compresses the first slot of a read block, :245-246);
//   * --hash-alg also accepts skein512 and sha256 (= what sha256mb intended, :136-158);
//   * the whole batch is one cw_hash_compress call (or --c-threads concurrent cw_submit batches when
//     --gpu-offload 1), instead of one call per 4 KB block;
//   * --devices N shards the chunks over N GPUs by contiguous range (one context + host thread per GPU);
//   * --randomness appends the per-byte digest distribution lines of the reference's hashing_randomness tool
//     (src/hashing_randomness/hash.cpp:74-101);
//   * --dump PREFIX writes PREFIX.digests / PREFIX.lens (u32 LE) / PREFIX.blocks (blocks back to back) so a
//     run can be compared with another one byte for byte (the tests do; the oracle lives in tests/, not here).
#include <getopt.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/compute_war.h"

static void usage(const char *argv0, const char *msg = nullptr)
{
    if (msg) fprintf(stderr, "%s\n", msg);
    fprintf(stderr,
            "usage: %s [options] file...\n"
            "  -c, --c-threads N     concurrent offload batches when --gpu-offload 1 (default 8)\n"
            "  -g, --gpu-offload B   1: asynchronous cw_submit/cw_wait batches; 0: one blocking call (default 0)\n"
            "  -r, --read-blocks N   read blocking factor (accepted for compatibility; batches are whole files)\n"
            "  -G, --hash-blocks N   ignored, as in the reference\n"
            "  -C, --comp-alg A      lzf | lz4 (default lz4)\n"
            "  -H, --hash-alg A      skein (= skein256-128) | skein512 | sha256mb | sha256 (default skein)\n"
            "  -k, --chunk BYTES     chunk size (default 4096)\n"
            "  -m, --mode M          compat | fast (default compat)\n"
            "  -d, --device N        first CUDA device (default 0)\n"
            "  -D, --devices N       shard the chunks over N GPUs starting at --device: contiguous chunk ranges, one\n"
            "                        context and host thread per GPU, no exchange between them (default 1)\n"
            "  -i, --iters N         repeat the timed pass N times (default 1)\n"
            "  -P, --no-pin          leave the host buffers pageable, as the reference's malloc'd buffers are (default: page-lock\n"
            "                        them once before the timed passes, cw_host_register)\n"
            "  -o, --dump PREFIX     write PREFIX.digests, PREFIX.lens, PREFIX.blocks\n"
            "  -R, --randomness      also print the digest byte distribution lines of the reference's hashing_randomness\n"
            "                        tool (byte|alg|avg|stddev|, then \"Average std dev\") for the first 16 digest bytes\n",
            argv0);
    exit(msg ? 2 : 0);
}

int main(int argc, char **argv)
{
    int threads = 8, offload = 0, device = 0, iters = 1, randomness = 0, devices = 1;
    std::string dump;
    int pin = 1;
    size_t chunk = 4096;
    std::string comp = "lz4", hash = "skein", mode = "compat";
    static const option opts[] = {{"help", no_argument, 0, 'h'},          {"c-threads", required_argument, 0, 'c'},
                                  {"gpu-offload", required_argument, 0, 'g'}, {"read-blocks", required_argument, 0, 'r'},
                                  {"hash-blocks", required_argument, 0, 'G'}, {"comp-alg", required_argument, 0, 'C'},
                                  {"hash-alg", required_argument, 0, 'H'},    {"chunk", required_argument, 0, 'k'},
                                  {"mode", required_argument, 0, 'm'},        {"device", required_argument, 0, 'd'},
                                  {"iters", required_argument, 0, 'i'},       {"randomness", no_argument, 0, 'R'},       {"devices", required_argument, 0, 'D'},
                                  {"dump", required_argument, 0, 'o'},       {"no-pin", no_argument, 0, 'P'},
                                  {0, 0, 0, 0}};
    int o;
    while ((o = getopt_long(argc, argv, "hc:g:r:G:C:H:k:m:d:i:RD:o:P", opts, nullptr)) != -1) {
        switch (o) {
        case 'h': usage(argv[0]); break;
        case 'c': threads = atoi(optarg); break;
        case 'g': offload = atoi(optarg) != 0 || !strcmp(optarg, "true"); break;
        case 'r': break;
        case 'G': break;
        case 'C': comp = optarg; break;
        case 'H': hash = optarg; break;
        case 'k': chunk = strtoull(optarg, nullptr, 10); break;
        case 'm': mode = optarg; break;
        case 'd': device = atoi(optarg); break;
        case 'i': iters = atoi(optarg); break;
        case 'R': randomness = 1; break;
        case 'D': devices = atoi(optarg); break;
        case 'o': dump = optarg; break;
        case 'P': pin = 0; break;
        default: usage(argv[0], "bad option");
        }
    }
    cw_comp_alg calg;
    if (comp == "lzf") calg = CW_COMP_LZF;
    else if (comp == "lz4") calg = CW_COMP_LZ4;
    else usage(argv[0], "invalid compression algorithm specified; please use either \"lzf\" or \"lz4\"");
    cw_hash_alg halg;
    if (hash == "skein" || hash == "skein256-128") halg = CW_HASH_SKEIN256_128;
    else if (hash == "skein512") halg = CW_HASH_SKEIN512;
    else if (hash == "sha256mb" || hash == "sha256") halg = CW_HASH_SHA256;
    else usage(argv[0], "invalid hashing algorithm specified; please use \"skein\", \"skein512\" or \"sha256mb\"");
    cw_mode cmode = mode == "fast" ? CW_MODE_FAST : CW_MODE_COMPAT;
    if (chunk == 0 || chunk > CW_MAX_CHUNK) usage(argv[0], "chunk must be 1..65536");
    if (threads < 1) threads = 1;

    // ReadAllFiles: whole chunks only, untimed (HashAndCompress.cpp:392)
    std::vector<uint8_t> data;
    for (int i = optind; i < argc; i++) {
        std::ifstream fs(argv[i], std::ifstream::binary);
        if (!fs) {
            fprintf(stderr, "Unable to open file: %s\n", argv[i]);
            continue;
        }
        std::vector<uint8_t> f((std::istreambuf_iterator<char>(fs)), std::istreambuf_iterator<char>());
        size_t keep = f.size() / chunk * chunk;
        data.insert(data.end(), f.begin(), f.begin() + keep);
    }
    const size_t n = data.size() / chunk;
    if (n == 0) {
        fprintf(stderr, "no full chunks to process\n");
        return 1;
    }

    // one context per GPU; GPU g of G owns chunks [g*n/G, (g+1)*n/G) -- the reference shards its blocks over worker
    // threads through a queue (HashAndCompress.cpp:263-272), here the ranges are fixed and the outputs land in place
    if (devices < 1) devices = 1;
    if ((size_t)devices > n) devices = (int)n;
    std::vector<cw_ctx *> ctxs(devices, nullptr);
    int rc = CW_OK;
    for (int g = 0; g < devices; g++) {
        rc = cw_init(&ctxs[g], device + g);
        if (rc != CW_OK) {
            fprintf(stderr, "cw_init(device %d): %s %s\n", device + g, cw_status_string(rc), cw_last_error(nullptr));
            return 1;
        }
    }
    cw_ctx *ctx = ctxs[0];
    const size_t D = cw_digest_bytes(halg), stride = cw_compress_bound(calg, chunk);
    std::vector<uint8_t> digests(n * D), blocks(n * stride);
    std::vector<uint32_t> lens(n);

    if (pin) {  // untimed, like the reference's ReadAllFiles: the timed passes then copy at full PCIe rate
        int r1 = cw_host_register(ctx, data.data(), data.size()), r2 = cw_host_register(ctx, blocks.data(), blocks.size());
        int r3 = cw_host_register(ctx, digests.data(), digests.size()), r4 = cw_host_register(ctx, lens.data(), n * sizeof(uint32_t));
        if (r1 != CW_OK || r2 != CW_OK || r3 != CW_OK || r4 != CW_OK)
            fprintf(stderr, "warning: could not page-lock the host buffers (%s); continuing with pageable memory\n", cw_last_error(ctx));
    }
    auto pass_range = [&](cw_ctx *c, size_t first, size_t cnt) -> int {
        const uint8_t *src = data.data() + first * chunk;
        uint8_t *dg = digests.data() + first * D, *bl = blocks.data() + first * stride;
        uint32_t *ln = lens.data() + first;
        if (!offload) return cw_hash_compress(c, halg, calg, cmode, src, chunk, cnt, dg, bl, stride, ln);
        // HashOffload-style: `threads` batches in flight (Reset/Enqueue/Start ... Complete)
        std::vector<cw_ticket> tk(threads);
        int bad = CW_OK;
        for (int t = 0; t < threads; t++) {
            size_t lo = cnt * t / threads, hi = cnt * (t + 1) / threads;
            int r = cw_submit(c, halg, calg, cmode, src + lo * chunk, chunk, hi - lo, dg + lo * D, bl + lo * stride, stride,
                              ln + lo, nullptr, nullptr, &tk[t]);
            if (r != CW_OK) bad = r;
        }
        for (int t = 0; t < threads; t++) {
            int r = cw_wait(c, tk[t]);
            if (r != CW_OK) bad = r;
        }
        return bad;
    };
    auto pass = [&]() -> int {
        if (devices == 1) return pass_range(ctx, 0, n);
        std::vector<int> rcs(devices, CW_OK);
        std::vector<std::thread> th;
        for (int g = 0; g < devices; g++) {
            const size_t lo = n * g / devices, hi = n * (g + 1) / devices;
            th.emplace_back([&, g, lo, hi] { rcs[g] = pass_range(ctxs[g], lo, hi - lo); });
        }
        for (auto &t : th) t.join();
        for (int g = 0; g < devices; g++)
            if (rcs[g] != CW_OK) {
                ctx = ctxs[g];  // report this context's error text
                return rcs[g];
            }
        return CW_OK;
    };

    rc = pass();  // warm-up: context, module load, pipeline buffers
    if (rc != CW_OK) {
        fprintf(stderr, "engine: %s (%s)\n", cw_status_string(rc), cw_last_error(ctx));
        return 1;
    }
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < iters && rc == CW_OK; i++) rc = pass();
    auto t1 = std::chrono::high_resolution_clock::now();
    if (rc != CW_OK) {
        fprintf(stderr, "engine: %s (%s)\n", cw_status_string(rc), cw_last_error(ctx));
        return 1;
    }
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count() / iters;
    uint64_t csum = 0;
    for (size_t i = 0; i < n; i++) csum += lens[i] ? lens[i] : chunk;
    double mbps = (double)(n * chunk) * 1000.0 / (ms * 1024.0 * 1024.0);
    // the reference's line (HashAndCompress.cpp:409-412) ...
    printf("%s|%s|%llu|%llu\n", hash.c_str(), comp.c_str(), (unsigned long long)(ms + 0.5), (unsigned long long)mbps);
    // ... plus what it throws away
    fprintf(stderr, "chunks=%zu chunk=%zu gpus=%d ms=%.3f MBps=%.1f ratio=%.4f\n", n, chunk, devices, ms, mbps, (double)csum / (double)(n * chunk));
    if (!dump.empty()) {
        std::ofstream(dump + ".digests", std::ios::binary).write((const char *)digests.data(), (std::streamsize)digests.size());
        std::ofstream(dump + ".lens", std::ios::binary).write((const char *)lens.data(), (std::streamsize)(n * sizeof(uint32_t)));
        std::ofstream bo(dump + ".blocks", std::ios::binary);
        for (size_t i = 0; i < n; i++)
            bo.write((const char *)(lens[i] ? blocks.data() + i * stride : data.data() + i * chunk), lens[i] ? lens[i] : chunk);
    }
    if (randomness) {
        // Hasher::LogResultsInternal / LogByteResults, src/hashing_randomness/hash.cpp:30-101
        const size_t nb = 16;
        double avg[16], sd[16], tot = 0;
        rc = cw_digest_stats(ctx, digests.data(), D, n, nb, avg, sd, &tot);
        if (rc != CW_OK) {
            fprintf(stderr, "engine: %s (%s)\n", cw_status_string(rc), cw_last_error(ctx));
            return 1;
        }
        for (size_t b = 0; b < nb; b++) std::cout << b << "|" << hash << "|" << avg[b] << "|" << sd[b] << "|" << std::endl;
        std::cout << "Average std dev: " << tot << std::endl;
    }
    if (pin) {
        cw_host_unregister(ctxs[0], data.data());
        cw_host_unregister(ctxs[0], blocks.data());
        cw_host_unregister(ctxs[0], digests.data());
        cw_host_unregister(ctxs[0], lens.data());
    }
    for (cw_ctx *c : ctxs) cw_shutdown(c);
    return 0;
}
