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
//   * --work-items runs the reference's own calling pattern instead: --c-threads worker threads pop work items of
//     --read-blocks chunks from a shared counter and call cw_hash_compress once per item (PopAndProcessBlocks /
//     ProcessBlock, HashAndCompress.cpp:231-272); the engine's small-batch front-end gathers them into shared launches;
//   * --experiment prints the per-block lines of the reference's compression study, `alg|csize|comp_us|decomp_us|file|block`
//     (src/compression_perf/src/experiment.cpp:114,122,253,260; blocks numbered from 1, :93,510), and --hash-perf those of
//     its hashing study, `file|idx|Skein256|us|` / `file|idx|Sha256|us|` (src/hashing_perf/hash.cpp:18-23,38-43; the index
//     advances after each line, test.cpp:19-23), so that the reference's notebooks read this engine's logs.  Times are
//     the device time of the file's one batch (CUDA events around its kernels) divided by its blocks, in microseconds;
//   * --dump PREFIX writes PREFIX.digests / PREFIX.lens (u32 LE) / PREFIX.blocks (blocks back to back) so a
//     run can be compared with another one byte for byte (the tests do; the oracle lives in tests/, not here).
#include <getopt.h>

#include <atomic>
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
            "  -r, --read-blocks N   read blocking factor: chunks per work item with --work-items (default 8); otherwise batches\n"
            "                        are whole files\n"
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
            "  -w, --work-items      the reference's calling pattern: --c-threads threads, one cw_hash_compress call per work\n"
            "                        item of --read-blocks chunks (default 8 = 32 KiB at the default chunk size)\n"
            "  -E, --experiment      per-block lines of the reference's `experiment` tool for --comp-alg (lz4 | lzf | both):\n"
            "                        alg|csize|comp_us|decomp_us|file|block\n"
            "  -Q, --hash-perf       per-block lines of the reference's hashing_perf tool: file|idx|Skein256|us| and\n"
            "                        file|idx|Sha256|us|\n"
            "  -o, --dump PREFIX     write PREFIX.digests, PREFIX.lens, PREFIX.blocks\n"
            "  -R, --randomness      also print the digest byte distribution lines of the reference's hashing_randomness\n"
            "                        tool (byte|alg|avg|stddev|, then \"Average std dev\") for the first 16 digest bytes\n",
            argv0);
    exit(msg ? 2 : 0);
}

int main(int argc, char **argv)
{
    int threads = 8, offload = 0, device = 0, iters = 1, randomness = 0, devices = 1, work_items = 0, experiment = 0, hash_perf = 0;
    size_t read_blocks = 8;
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
                                  {"work-items", no_argument, 0, 'w'},       {"experiment", no_argument, 0, 'E'},
                                  {"hash-perf", no_argument, 0, 'Q'},        {0, 0, 0, 0}};
    int o;
    while ((o = getopt_long(argc, argv, "hc:g:r:G:C:H:k:m:d:i:RD:o:PwEQ", opts, nullptr)) != -1) {
        switch (o) {
        case 'h': usage(argv[0]); break;
        case 'c': threads = atoi(optarg); break;
        case 'g': offload = atoi(optarg) != 0 || !strcmp(optarg, "true"); break;
        case 'r': read_blocks = strtoull(optarg, nullptr, 10); break;
        case 'w': work_items = 1; break;
        case 'E': experiment = 1; break;
        case 'Q': hash_perf = 1; break;
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
    else if (comp == "lz4" || (experiment && comp == "both")) calg = CW_COMP_LZ4;
    else usage(argv[0], "invalid compression algorithm specified; please use either \"lzf\" or \"lz4\"");
    cw_hash_alg halg;
    if (hash == "skein" || hash == "skein256-128") halg = CW_HASH_SKEIN256_128;
    else if (hash == "skein512") halg = CW_HASH_SKEIN512;
    else if (hash == "sha256mb" || hash == "sha256") halg = CW_HASH_SHA256;
    else usage(argv[0], "invalid hashing algorithm specified; please use \"skein\", \"skein512\" or \"sha256mb\"");
    cw_mode cmode = mode == "fast" ? CW_MODE_FAST : CW_MODE_COMPAT;
    if (chunk == 0 || chunk > CW_MAX_CHUNK) usage(argv[0], "chunk must be 1..65536");
    if (threads < 1) threads = 1;

    if (experiment || hash_perf) {
        cw_ctx *ectx = nullptr;
        int erc = cw_init(&ectx, device);
        if (erc != CW_OK) {
            fprintf(stderr, "cw_init(device %d): %s %s\n", device, cw_status_string(erc), cw_last_error(nullptr));
            return 1;
        }
        cw_prof_enable(ectx, 1);
        for (int i = optind; i < argc; i++) {
            std::ifstream fs(argv[i], std::ifstream::binary);
            if (!fs) {
                fprintf(stderr, "Unable to open file: %s\n", argv[i]);
                continue;
            }
            std::vector<uint8_t> f((std::istreambuf_iterator<char>(fs)), std::istreambuf_iterator<char>());
            const size_t nb = f.size() / chunk;  // only full blocks (experiment.cpp:102, file.cpp:18-33)
            if (nb == 0) continue;
            double ms = 0;
            uint64_t launches = 0;
            if (experiment) {
                std::vector<std::string> algs;
                if (comp == "both" || comp == "lzf") algs.push_back("lzf");  // the reference runs lzf before lz4
                if (comp == "both" || comp == "lz4") algs.push_back("lz4");
                std::vector<std::vector<uint32_t>> csz(algs.size());
                std::vector<double> cus(algs.size()), dus(algs.size());
                for (size_t a = 0; a < algs.size() && erc == CW_OK; a++) {
                    const cw_comp_alg ca = algs[a] == "lzf" ? CW_COMP_LZF : CW_COMP_LZ4;
                    const size_t st = cw_compress_bound(ca, chunk);
                    std::vector<uint8_t> blk(nb * st), back(nb * chunk);
                    std::vector<uint32_t> ln(nb), bl(nb);
                    cw_prof_read(ectx, &ms, &launches);
                    erc = cw_compress(ectx, ca, cmode, f.data(), chunk, nb, blk.data(), st, ln.data());
                    cw_prof_read(ectx, &ms, &launches);
                    cus[a] = ms * 1000.0 / (double)nb;
                    if (erc != CW_OK) break;
                    for (size_t k = 0; k < nb; k++)  // lzf_compress's 0: the caller keeps the block as it is
                        if (ca == CW_COMP_LZF && ln[k] == 0) memcpy(blk.data() + k * st, f.data() + k * chunk, chunk);
                    erc = cw_decompress(ectx, ca, blk.data(), st, ln.data(), chunk, nb, back.data(), bl.data());
                    cw_prof_read(ectx, &ms, &launches);
                    dus[a] = ms * 1000.0 / (double)nb;
                    if (erc == CW_OK && memcmp(back.data(), f.data(), nb * chunk) != 0) {
                        fprintf(stderr, "%s: %s round trip differs\n", argv[i], algs[a].c_str());
                        return 1;
                    }
                    csz[a] = ln;
                }
                for (size_t k = 0; k < nb && erc == CW_OK; k++)
                    for (size_t a = 0; a < algs.size(); a++)
                        printf("%s|%u|%.4f|%.4f|%s|%zu\n", algs[a].c_str(), csz[a][k], cus[a], dus[a], argv[i], k + 1);
            } else {
                std::vector<uint8_t> dg(nb * 32);
                cw_prof_read(ectx, &ms, &launches);
                erc = cw_hash(ectx, CW_HASH_SKEIN256_128, f.data(), chunk, nb, dg.data());
                cw_prof_read(ectx, &ms, &launches);
                const double sk = ms * 1000.0 / (double)nb;
                if (erc == CW_OK) erc = cw_hash(ectx, CW_HASH_SHA256, f.data(), chunk, nb, dg.data());
                cw_prof_read(ectx, &ms, &launches);
                const double sh = ms * 1000.0 / (double)nb;
                for (size_t k = 0; k < nb && erc == CW_OK; k++) {
                    printf("%s|%zu|Skein256|%.4f|\n", argv[i], 2 * k, sk);
                    printf("%s|%zu|Sha256|%.4f|\n", argv[i], 2 * k + 1, sh);
                }
            }
            if (erc != CW_OK) {
                fprintf(stderr, "engine: %s (%s)\n", cw_status_string(erc), cw_last_error(ectx));
                return 1;
            }
        }
        cw_shutdown(ectx);
        return 0;
    }

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
        if (work_items) {
            // PopAndProcessBlocks (HashAndCompress.cpp:263-272): every worker thread takes the next work item and calls
            // ProcessBlock for it until none is left
            const size_t per = read_blocks ? read_blocks : 1, items = (cnt + per - 1) / per;
            std::atomic<size_t> next{0};
            std::atomic<int> bad{CW_OK};
            std::vector<std::thread> th;
            for (int t = 0; t < threads; t++)
                th.emplace_back([&] {
                    // --gpu-offload 1: HashOffload style (HashOffload.h:15-64) -- Start() many items, Complete() them
                    // later: every worker keeps up to 64 tickets in flight instead of waiting for each item
                    std::vector<cw_ticket> inflight;
                    size_t head = 0;
                    for (;;) {
                        const size_t it = next.fetch_add(1);
                        if (it >= items) break;
                        const size_t lo = it * per, m = lo + per <= cnt ? per : cnt - lo;
                        int r;
                        if (offload) {
                            cw_ticket tk = 0;
                            r = cw_submit(c, halg, calg, cmode, src + lo * chunk, chunk, m, dg + lo * D, bl + lo * stride, stride,
                                          ln + lo, nullptr, nullptr, &tk);
                            if (r == CW_OK) inflight.push_back(tk);
                            if (inflight.size() - head >= 64) {
                                const int w = cw_wait(c, inflight[head++]);
                                if (w != CW_OK) bad = w;
                            }
                        } else
                            r = cw_hash_compress(c, halg, calg, cmode, src + lo * chunk, chunk, m, dg + lo * D, bl + lo * stride,
                                                 stride, ln + lo);
                        if (r != CW_OK) bad = r;
                    }
                    for (; head < inflight.size(); head++) {
                        const int w = cw_wait(c, inflight[head]);
                        if (w != CW_OK) bad = w;
                    }
                });
            for (auto &x : th) x.join();
            return bad.load();
        }
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
    if (work_items) {
        uint64_t nbat = 0, ncalls = 0;
        cw_batch_stats(ctx, &nbat, &ncalls);
        fprintf(stderr, "work items of %zu chunks from %d threads: %llu calls served by %llu launches\n", read_blocks, threads,
                (unsigned long long)ncalls, (unsigned long long)nbat);
    }
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
