// api.cu -- the C ABI of include/compute_war.h: context, device-buffer calls, the pipelined
// host-buffer calls and the asynchronous offload queue.
//
// Reference interfaces replaced (see compute_war.h for the per-function list): initializeGpu()
// HashAndCompress.cpp:95-98, doHashing :119-134, doCompression :111,342-355, ProcessBlock :231-261,
// HashOffload HashOffload.h:15-64 + hashing_offload_entry_point HashAndCompress.cpp:160-183.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/compute_war.h"
#include "hash_kernels.cuh"
#include "launch.h"

namespace {

constexpr int kSlots = 4;                     // pipeline slots: H2D / kernel / D2H of different slices in flight
constexpr size_t kSliceBytesDefault = 64u << 20;  // input bytes per pipeline slice (CW_SLICE_MB overrides); 64 MiB x 4 slots measured best
constexpr int kWorkRing = 256;
constexpr size_t kStageSliceBytes = 16u << 20;  // slice size for callers with pageable memory (staged through pinned buffers)

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_dig = nullptr;
    uint32_t *d_lens = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_dig = 0, cap_lens = 0;
    bool busy = false;
    // packed host output (cw_hash_compress_packed): the slice's blocks back to back, their offsets, the scan scratch,
    // and the slice's packed size read back into a pinned word once the scan has run
    uint8_t *d_packed = nullptr;
    unsigned long long *d_offsets = nullptr, *d_scan = nullptr, *h_total = nullptr;
    size_t cap_packed = 0, cap_offsets = 0, cap_scan = 0;
    cudaEvent_t ready = nullptr;  // kernels of the slice done, h_total valid
    size_t lo = 0, cnt = 0;       // chunks of the slice
    unsigned long long base = 0;  // where the slice starts in the caller's packed buffer
    // pinned staging buffers for callers with pageable memory (host_pass): the slice's input on its way in, its results
    // on their way out
    uint8_t *h_in = nullptr, *h_out = nullptr, *h_dig = nullptr;
    uint32_t *h_lens = nullptr;
    unsigned long long *h_offsets = nullptr;
    size_t cap_h_in = 0, cap_h_out = 0, cap_h_dig = 0, cap_h_lens = 0, cap_h_offsets = 0;
    unsigned long long total = 0;  // packed bytes of the slice
    bool staged_out = false;  // the slice's results still sit in the staging buffers
};

struct Job {
    cw_ticket ticket;
    cw_hash_alg halg;
    cw_comp_alg calg;
    cw_mode mode;
    const void *src;
    size_t chunk, n;
    void *digests, *dst;
    size_t stride;
    uint32_t *lens;
    cw_complete_fn cb;
    void *user;
};
// ---- small-batch front-end: the reference's callers hand over one small work item at a time from many threads
// (ProcessBlock per 32 KiB work item from --c-threads workers, HashAndCompress.cpp:160-183,263-272; HashOffload.h:15-64).
// Work items of equal (hash, codec, mode, chunk size) that arrive while the GPU is busy are gathered in one pinned
// batch buffer -- every caller copies its own bytes in and, afterwards, its own results out, in parallel -- and one
// copy in / one launch / one set of copies out serves them all. ----
constexpr size_t kBatchBytes = 8u << 20;   // input bytes per batch buffer
constexpr size_t kBatchSoftBytes = kBatchBytes / 8;  // a batch that holds this much is sealed while another buffer is free:
                                                     // callers that arrive together (14 threads x 1 MiB) are spread over
                                                     // several batches whose copies and kernels overlap, instead of all
                                                     // riding -- and waiting -- in one
constexpr int kBatchBufs = 6;
constexpr size_t kSmallCall = 8u << 20;    // host-buffer calls up to this many input bytes go through the batcher
constexpr size_t kBatchMinChunk = 512;     // smaller chunks: direct path (digest / length arrays would dominate)

struct BatchBuf {
    enum State { FREE, OPEN, SEALED, RUNNING, DONE };
    uint8_t *h_in = nullptr, *h_out = nullptr, *h_dig = nullptr;
    uint32_t *h_lens = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_dig = nullptr;
    uint32_t *d_lens = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    State state = FREE;
    std::vector<std::pair<Job, size_t>> jobs;  // cw_submit jobs riding in this batch, with their first chunk
    cw_hash_alg halg = CW_HASH_SKEIN512;
    cw_comp_alg calg = CW_COMP_LZ4;
    cw_mode mode = CW_MODE_COMPAT;
    size_t chunk = 0, stride = 0, D = 0, cap_chunks = 0;
    size_t chunks = 0;   // chunks reserved so far
    int copying_in = 0;  // callers still copying their input in
    int readers = 0;     // callers that still have to copy their results out
    int tasks_left = 0;  // pieces of the submitted jobs' copy-out still with the helper threads (CopyPool::mu)
    int rc = 0;
    uint64_t gen = 0;
};

struct Batcher {
    std::mutex mu;
    std::condition_variable cv_worker, cv_callers;
    BatchBuf bufs[kBatchBufs];
    int open = -1;
    std::thread th;
    bool started = false, stop = false, failed_alloc = false;
    std::deque<int> inflight;  // worker-private: launched, not yet complete
};

// copy-out helpers of the offload thread: a finished batch carries the results of dozens to hundreds of cw_submit jobs,
// and one thread copying them into the callers' (pageable) buffers block by block was what the whole front-end waited for
// (HashOffload style, 32 KiB items: 4.8 GB/s).  The offload thread cuts a batch's jobs into pieces, the helpers copy and
// complete them, the last piece frees the batch buffer.  The same threads stage the slices of callers with pageable
// memory into and out of pinned buffers (host_pass).
struct CopyPool {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> tasks;
    std::vector<std::thread> th;
    bool stop = false;
};



}  // namespace

struct cw_ctx {
    int device = 0;
    CwDeviceInfo dev{};
    CwSkeinIV iv{};
    cudaStream_t stream = nullptr;
    std::atomic<uint64_t> launches{0};
    // work counters of the persistent kernels: a ring of kWorkRing pairs.  Launches on different caller streams are not
    // ordered with each other, so every entry carries an event recorded behind the launch that used it; an entry is
    // handed out again only after that launch has finished
    unsigned long long *d_work = nullptr;
    cudaEvent_t work_ev[kWorkRing] = {};
    bool work_used[kWorkRing] = {};
    int work_next = 0;
    std::mutex mu;       // guards work ring, profiling lists, last_error
    std::mutex host_mu;  // one host-buffer pipeline at a time per context
    std::string last_error;
    Slot slots[kSlots];
    // profiling of the *_dev kernels
    bool prof = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms = 0;
    uint64_t prof_launches = 0;
    // async offload
    std::thread worker;
    std::mutex q_mu;
    std::condition_variable q_cv, done_cv;
    std::deque<Job> queue;
    std::deque<int> done_bufs;  // batch buffers whose launch is complete and that carry cw_submit jobs
    std::map<cw_ticket, int> finished;  // status of completed jobs nobody has waited for yet
    std::map<cw_ticket, bool> pending;  // submitted, not complete (small jobs travel in groups: no ticket order)
    cw_ticket next_ticket = 1;
    bool stop = false;
    Batcher bat;
    CopyPool pool;
    std::atomic<uint64_t> batches{0}, batched_calls{0};
};

namespace {

std::string g_init_error;  // why the last cw_init failed (cw_last_error(NULL))

int fail(cw_ctx *ctx, cudaError_t e, const char *what)
{
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return CW_ERR_CUDA;
}

#define CW_CUDA(ctx, call)                                    \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return fail(ctx, e_, #call);   \
    } while (0)

int check_common(const cw_ctx *ctx, size_t chunk, size_t n, const void *src)
{
    if (!ctx) return CW_ERR_ARG;
    if (chunk == 0 || (n && !src)) return CW_ERR_ARG;
    if (chunk > CW_MAX_CHUNK) return CW_ERR_UNSUPPORTED;
    return CW_OK;
}

// a pair of zeroed-by-the-caller work counters for one launch on stream st; work_done() must follow the launch
struct WorkSlot {
    unsigned long long *p;
    int idx;
};
WorkSlot next_work(cw_ctx *ctx)
{
    std::lock_guard<std::mutex> g(ctx->mu);
    const int i = ctx->work_next;
    ctx->work_next = (i + 1) % kWorkRing;
    if (ctx->work_used[i]) cudaEventSynchronize(ctx->work_ev[i]);  // kWorkRing launches later: normally long finished
    ctx->work_used[i] = false;
    return WorkSlot{ctx->d_work + 2 * i, i};  // two counters per launch
}
void work_done(cw_ctx *ctx, const WorkSlot &w, cudaStream_t st)
{
    std::lock_guard<std::mutex> g(ctx->mu);
    if (cudaEventRecord(ctx->work_ev[w.idx], st) == cudaSuccess) ctx->work_used[w.idx] = true;
}

struct ProfScope {
    cw_ctx *ctx;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(cw_ctx *c, cudaStream_t s) : ctx(c), st(s)
    {
        if (!ctx->prof) return;
        std::lock_guard<std::mutex> g(ctx->mu);
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->ev_pool.empty()) {
                e = ctx->ev_pool.back();
                ctx->ev_pool.pop_back();
            } else
                cudaEventCreate(&e);
            return e;
        };
        e0 = get();
        e1 = get();
        cudaEventRecord(e0, st);
    }
    ~ProfScope()
    {
        if (!e0) return;
        cudaEventRecord(e1, st);
        std::lock_guard<std::mutex> g(ctx->mu);
        ctx->prof_ev.emplace_back(e0, e1);
    }
};

int do_hash_dev(cw_ctx *ctx, cw_hash_alg alg, const void *d_src, size_t chunk, size_t n, void *d_dig, cudaStream_t st)
{
    if ((int)alg < 0 || (int)alg > 2) return CW_ERR_ARG;
    if (n == 0) return CW_OK;
    ProfScope ps(ctx, st);
    CW_CUDA(ctx, cw_launch_hash((int)alg, ctx->iv, (const uint8_t *)d_src, (uint32_t)chunk, n, (uint8_t *)d_dig, st));
    ctx->launches++;
    return CW_OK;
}

int do_compress_dev(cw_ctx *ctx, cw_comp_alg alg, cw_mode mode, const void *d_src, size_t chunk, size_t n, void *d_dst,
                    size_t stride, uint32_t *d_lens, cudaStream_t st)
{
    if ((int)mode < 0 || (int)mode > 1) return CW_ERR_ARG;
    if (stride < cw_compress_bound(alg, chunk)) return CW_ERR_ARG;
    if (n == 0) return CW_OK;
    if (alg == CW_COMP_LZ4) {
        const WorkSlot w = next_work(ctx);
        CW_CUDA(ctx, cudaMemsetAsync(w.p, 0, sizeof *w.p, st));
        ProfScope ps(ctx, st);
        CW_CUDA(ctx, cw_launch_lz4(ctx->dev, mode == CW_MODE_FAST, (const uint8_t *)d_src, (uint32_t)chunk, n, (uint8_t *)d_dst, stride, d_lens,
                                   w.p, st));
        work_done(ctx, w, st);
        ctx->launches++;
        return CW_OK;
    }
    if (alg == CW_COMP_LZF) {
        const WorkSlot w = next_work(ctx);
        CW_CUDA(ctx, cudaMemsetAsync(w.p, 0, sizeof *w.p, st));
        ProfScope ps(ctx, st);
        if (mode == CW_MODE_FAST)
            CW_CUDA(ctx, cw_launch_lzf_fast(ctx->dev, (const uint8_t *)d_src, (uint32_t)chunk, n, (uint8_t *)d_dst, stride, d_lens, w.p, st));
        else
            CW_CUDA(ctx, cw_launch_lzf(ctx->dev, (const uint8_t *)d_src, (uint32_t)chunk, n, (uint8_t *)d_dst, stride, d_lens, w.p, st));
        work_done(ctx, w, st);
        ctx->launches++;
        return CW_OK;
    }
    return CW_ERR_ARG;
}

int do_fused_dev(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *d_src, size_t chunk, size_t n,
                 void *d_dig, void *d_dst, size_t stride, uint32_t *d_lens, cudaStream_t st)
{
    if ((int)halg < 0 || (int)halg > 2 || (int)calg < 0 || (int)calg > 1 || (int)mode < 0 || (int)mode > 1) return CW_ERR_ARG;
    if (stride < cw_compress_bound(calg, chunk)) return CW_ERR_ARG;
    if (n == 0) return CW_OK;
    const WorkSlot w = next_work(ctx);
    CW_CUDA(ctx, cudaMemsetAsync(w.p, 0, 2 * sizeof *w.p, st));
    ProfScope ps(ctx, st);
    CW_CUDA(ctx, cw_launch_fused(ctx->dev, (int)halg, (int)calg, mode == CW_MODE_FAST, ctx->iv, (const uint8_t *)d_src, (uint32_t)chunk, n,
                                 (uint8_t *)d_dig, (uint8_t *)d_dst, stride, d_lens, w.p, st));
    work_done(ctx, w, st);
    ctx->launches++;
    return CW_OK;
}

int ensure(cw_ctx *ctx, void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return CW_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) {
        fail(ctx, e, "cudaMalloc(pipeline slot)");
        return CW_ERR_NOMEM;
    }
    *cap = need;
    return CW_OK;
}

// ---------------- copy helpers ----------------
int pool_threads()
{
    static const int n = [] {
        const char *e = getenv("CW_COPY_THREADS");  // 0: no helpers (the calling / offload thread copies everything itself)
        int v = e ? atoi(e) : (int)std::thread::hardware_concurrency() / 2;
        return v < 0 ? 0 : (v > 8 ? 8 : v);
    }();
    return n;
}

void copy_helper(cw_ctx *ctx)
{
    CopyPool &P = ctx->pool;
    for (;;) {
        std::function<void()> t;
        {
            std::unique_lock<std::mutex> lk(P.mu);
            P.cv.wait(lk, [&] { return P.stop || !P.tasks.empty(); });
            if (P.tasks.empty()) return;  // stop requested and nothing left
            t = std::move(P.tasks.front());
            P.tasks.pop_front();
        }
        t();
    }
}

// P.mu held
void pool_start(cw_ctx *ctx)
{
    CopyPool &P = ctx->pool;
    if (P.th.empty())
        for (int t = 0; t < pool_threads(); t++) P.th.emplace_back(copy_helper, ctx);
}

// fn(k) for every k in [0, pieces), on the helpers and the calling thread; returns when all pieces are done
void pool_for(cw_ctx *ctx, size_t pieces, const std::function<void(size_t)> &fn)
{
    if (pool_threads() < 2 || pieces < 2) {
        for (size_t k = 0; k < pieces; k++) fn(k);
        return;
    }
    // shared state: a helper that gets to its task late (everything already done, the caller gone) finds no piece left
    // and touches nothing of the caller's
    struct State {
        std::atomic<size_t> next{0};
        size_t pieces = 0, left = 0;
        std::mutex m;
        std::condition_variable c;
        const std::function<void(size_t)> *fn = nullptr;
    };
    auto st = std::make_shared<State>();
    st->pieces = st->left = pieces;
    st->fn = &fn;
    auto body = [st] {
        for (;;) {
            const size_t k = st->next.fetch_add(1);
            if (k >= st->pieces) return;
            (*st->fn)(k);
            std::lock_guard<std::mutex> lk(st->m);
            if (--st->left == 0) st->c.notify_all();
        }
    };
    CopyPool &P = ctx->pool;
    const size_t helpers = std::min<size_t>(pieces - 1, (size_t)pool_threads());
    {
        std::lock_guard<std::mutex> lk(P.mu);
        pool_start(ctx);
        for (size_t h = 0; h < helpers; h++) P.tasks.emplace_back(body);
    }
    P.cv.notify_all();
    body();
    std::unique_lock<std::mutex> lk(st->m);
    st->c.wait(lk, [&] { return st->left == 0; });
}

void pool_memcpy(cw_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    constexpr size_t kPiece = 512u << 10;  // (a 4 MiB slice of a small staged call still spreads over 8 threads)
    const size_t pieces = (bytes + kPiece - 1) / kPiece;
    pool_for(ctx, pieces, [&](size_t k) {
        const size_t lo = k * kPiece, len = std::min(kPiece, bytes - lo);
        memcpy((uint8_t *)dst + lo, (const uint8_t *)src + lo, len);
    });
}

int ensure_host(cw_ctx *ctx, void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return CW_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaHostAlloc(p, need, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        fail(ctx, e, "cudaHostAlloc(staging buffer)");
        return CW_ERR_NOMEM;
    }
    *cap = need;
    return CW_OK;
}

// true for ordinary (pageable) host memory: not page-locked by cudaHostAlloc / cw_host_register
bool is_pageable(const void *p)
{
    static const bool off = getenv("CW_NO_STAGING") != nullptr;  // A/B knob: leave pageable buffers to the driver
    if (off || !p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// Pipelined host-buffer pass: slices of 64 MiB rotate through four slots, each with its own stream, so
// the H2D copy of slice k+1, the kernels of slice k and the D2H copy of slice k-1 overlap.
// Callers with PAGEABLE buffers (the reference's tools read their files into ordinary memory): the driver stages such
// copies itself, synchronously and on one thread (6 GB/s for a whole-file call on the 16-thread box).  Instead, slices of
// 16 MiB are staged through pinned buffers of the slot by the copy helpers -- input before the slice's copy-in,
// results (only the bytes that exist) when the slot comes round again -- while the GPU works on the slices in between.
int host_pass(cw_ctx *ctx, bool want_hash, bool want_comp, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode,
              const void *src, size_t chunk, size_t n, void *digests, void *dst, size_t stride, uint32_t *lens)
{
    std::lock_guard<std::mutex> hg(ctx->host_mu);
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t D = want_hash ? cw_digest_bytes(halg) : 0;
    static const size_t slice_bytes = [] {
        const char *e = getenv("CW_SLICE_MB");
        size_t mb = e ? (size_t)atoi(e) : 0;
        return mb ? mb << 20 : kSliceBytesDefault;
    }();
    static const int n_slots = [] {
        const char *e = getenv("CW_SLOTS");
        int v = e ? atoi(e) : kSlots;
        return v < 1 ? 1 : (v > kSlots ? kSlots : v);
    }();
    const bool stage_in = is_pageable(src);
    const bool stage_out = (want_comp && (is_pageable(dst) || is_pageable(lens))) || (want_hash && is_pageable(digests));
    // staged calls: at least four slices per call where the call is big enough (a 16 MiB call in one slice would run its
    // staging, its copies and its kernel one after the other)
    size_t stage_slice = std::min(kStageSliceBytes, std::max<size_t>(2u << 20, (n * chunk / 4 + 0xfffffu) & ~(size_t)0xfffffu));
    size_t per = ((stage_in || stage_out) ? stage_slice : slice_bytes) / chunk;
    if (per < 1) per = 1;
    if (per > n) per = n;  // (a small call allocates what it needs)
    int rc = CW_OK;
    // results of a finished slice: from the staging buffers to the caller's arrays
    auto drain = [&](Slot &s) -> int {
        if (!s.busy) return CW_OK;
        s.busy = false;
        cudaError_t e = cudaEventSynchronize(s.done);
        if (e != cudaSuccess) {
            s.staged_out = false;
            return fail(ctx, e, "cudaEventSynchronize(slot)");
        }
        if (!s.staged_out) return CW_OK;
        s.staged_out = false;
        const size_t lo = s.lo, cnt = s.cnt;
        const size_t kPieceChunks = std::max<size_t>(1, (1u << 20) / chunk);  // ~1 MiB of input per piece
        pool_for(ctx, (cnt + kPieceChunks - 1) / kPieceChunks, [&](size_t k) {
            const size_t a = k * kPieceChunks, m = std::min(kPieceChunks, cnt - a);
            if (want_hash) memcpy((uint8_t *)digests + (lo + a) * D, s.h_dig + a * D, m * D);
            if (want_comp) {
                memcpy(lens + lo + a, s.h_lens + a, m * sizeof(uint32_t));
                for (size_t i = a; i < a + m; i++)  // only the bytes that exist (LZF: 0 = "store raw", nothing to copy)
                    memcpy((uint8_t *)dst + (lo + i) * stride, s.h_out + i * stride, s.h_lens[i]);
            }
        });
        return CW_OK;
    };
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess) rc = fail(ctx, e, what);
        return e == cudaSuccess;
    };
    size_t k = 0;
    for (size_t lo = 0; lo < n && rc == CW_OK; lo += per, k++) {
        const size_t cnt = (n - lo < per) ? n - lo : per;
        Slot &s = ctx->slots[k % n_slots];
        if ((rc = drain(s)) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_in, &s.cap_in, per * chunk)) != CW_OK) break;
        if (want_comp) {
            if ((rc = ensure(ctx, (void **)&s.d_out, &s.cap_out, per * stride)) != CW_OK) break;
            if ((rc = ensure(ctx, (void **)&s.d_lens, &s.cap_lens, per * sizeof(uint32_t))) != CW_OK) break;
        }
        if (want_hash && (rc = ensure(ctx, (void **)&s.d_dig, &s.cap_dig, per * D)) != CW_OK) break;
        const uint8_t *in = (const uint8_t *)src + lo * chunk;
        if (stage_in) {
            if ((rc = ensure_host(ctx, (void **)&s.h_in, &s.cap_h_in, per * chunk)) != CW_OK) break;
            pool_memcpy(ctx, s.h_in, in, cnt * chunk);
            in = s.h_in;
        }
        if (stage_out) {
            if (want_comp) {
                if ((rc = ensure_host(ctx, (void **)&s.h_out, &s.cap_h_out, per * stride)) != CW_OK) break;
                if ((rc = ensure_host(ctx, (void **)&s.h_lens, &s.cap_h_lens, per * sizeof(uint32_t))) != CW_OK) break;
            }
            if (want_hash && (rc = ensure_host(ctx, (void **)&s.h_dig, &s.cap_h_dig, per * D)) != CW_OK) break;
        }
        if (!cu(cudaMemcpyAsync(s.d_in, in, cnt * chunk, cudaMemcpyHostToDevice, s.st), "cudaMemcpyAsync(slice in)")) break;
        if (want_hash && want_comp)
            rc = do_fused_dev(ctx, halg, calg, mode, s.d_in, chunk, cnt, s.d_dig, s.d_out, stride, s.d_lens, s.st);
        else if (want_hash)
            rc = do_hash_dev(ctx, halg, s.d_in, chunk, cnt, s.d_dig, s.st);
        else
            rc = do_compress_dev(ctx, calg, mode, s.d_in, chunk, cnt, s.d_out, stride, s.d_lens, s.st);
        if (rc != CW_OK) break;
        s.busy = true;  // (from here on the slot has copies into host memory in flight)
        s.staged_out = false;
        if (want_comp) {
            if (!cu(cudaMemcpyAsync(stage_out ? s.h_out : (uint8_t *)dst + lo * stride, s.d_out, cnt * stride, cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync(blocks out)")) break;
            if (!cu(cudaMemcpyAsync(stage_out ? (void *)s.h_lens : (void *)(lens + lo), s.d_lens, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync(lengths out)")) break;
        }
        if (want_hash &&
            !cu(cudaMemcpyAsync(stage_out ? s.h_dig : (uint8_t *)digests + lo * D, s.d_dig, cnt * D, cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync(digests out)"))
            break;
        if (!cu(cudaEventRecord(s.done, s.st), "cudaEventRecord(slot)")) break;
        s.staged_out = stage_out;
        s.lo = lo;
        s.cnt = cnt;
    }
    // the slices still in flight, oldest first
    for (int i = 0; i < n_slots && rc == CW_OK; i++) rc = drain(ctx->slots[(k + i) % n_slots]);
    if (rc != CW_OK) {  // error path: nothing is delivered any more, but every copy into host memory ends before we return
        for (int i = 0; i < kSlots; i++) {
            Slot &s = ctx->slots[i];
            if (s.busy) cudaStreamSynchronize(s.st);
            s.busy = s.staged_out = false;
        }
    }
    return rc;
}

// cw_hash_compress_packed: the pipeline of host_pass with a packed device-to-host leg.  A slice's packed size is only
// known once its scan has run, so a slice takes two visits: A enqueues copy-in, the fused kernel, the pack kernels and
// the read-back of the size (8 bytes); B, issued one slice later -- while the next slice's kernels run -- waits for that
// word and enqueues copies of exactly the bytes that exist.
int host_pass_packed(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src, size_t chunk, size_t n,
                     void *digests, void *packed, size_t packed_cap, uint64_t *offsets, uint32_t *lens)
{
    std::lock_guard<std::mutex> hg(ctx->host_mu);
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t D = cw_digest_bytes(halg), stride = cw_compress_bound(calg, chunk);
    static const size_t slice_bytes = [] {
        const char *e = getenv("CW_SLICE_MB");
        size_t mb = e ? (size_t)atoi(e) : 0;
        return mb ? mb << 20 : kSliceBytesDefault;
    }();
    // callers with pageable memory: staged through pinned slot buffers by the copy helpers, like host_pass
    const bool stage_in = is_pageable(src);
    const bool stage_out = is_pageable(packed) || is_pageable(offsets) || is_pageable(lens) || is_pageable(digests);
    const size_t stage_slice = std::min(kStageSliceBytes, std::max<size_t>(2u << 20, (n * chunk / 4 + 0xfffffu) & ~(size_t)0xfffffu));
    size_t per = ((stage_in || stage_out) ? stage_slice : slice_bytes) / chunk;
    if (per < 1) per = 1;
    int rc = CW_OK;
    unsigned long long base = 0;
    auto cu = [&](cudaError_t e, const char *what) {  // (no early return: slices in flight are drained below)
        if (e != cudaSuccess) rc = fail(ctx, e, what);
        return e == cudaSuccess;
    };
    // a finished slice's offsets become offsets into the caller's buffer
    auto finish = [&](Slot &s) -> int {
        if (!s.busy) return CW_OK;
        s.busy = false;
        cudaError_t e = cudaEventSynchronize(s.done);
        if (e != cudaSuccess) return fail(ctx, e, "cudaEventSynchronize(slot)");
        if (!s.staged_out) {
            for (size_t i = 0; i < s.cnt; i++) offsets[s.lo + i] += s.base;
            return CW_OK;
        }
        s.staged_out = false;
        pool_memcpy(ctx, (uint8_t *)packed + s.base, s.h_out, s.total);
        for (size_t i = 0; i < s.cnt; i++) offsets[s.lo + i] = s.h_offsets[i] + s.base;
        memcpy(lens + s.lo, s.h_lens, s.cnt * sizeof(uint32_t));
        memcpy((uint8_t *)digests + s.lo * D, s.h_dig, s.cnt * D);
        return CW_OK;
    };
    auto stage_b = [&](Slot &s) -> int {
        cudaError_t e = cudaEventSynchronize(s.ready);
        if (e != cudaSuccess) return fail(ctx, e, "cudaEventSynchronize(slice)");
        const unsigned long long total = *s.h_total;
        if (base + total > packed_cap) {
            ctx->last_error = "cw_hash_compress_packed: packed buffer too small";
            return CW_ERR_ARG;
        }
        s.base = base;
        s.total = total;
        s.busy = true;  // (copies into host memory in flight from here on)
        s.staged_out = false;
        if (!cu(cudaMemcpyAsync(stage_out ? s.h_out : (uint8_t *)packed + base, s.d_packed, total, cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync(packed out)")) return rc;
        if (!cu(cudaMemcpyAsync(stage_out ? (void *)s.h_offsets : (void *)(offsets + s.lo), s.d_offsets, s.cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync(offsets out)")) return rc;
        if (!cu(cudaMemcpyAsync(stage_out ? (void *)s.h_lens : (void *)(lens + s.lo), s.d_lens, s.cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync(lengths out)")) return rc;
        if (!cu(cudaMemcpyAsync(stage_out ? s.h_dig : (uint8_t *)digests + s.lo * D, s.d_dig, s.cnt * D, cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync(digests out)")) return rc;
        if (!cu(cudaEventRecord(s.done, s.st), "cudaEventRecord")) return rc;
        s.staged_out = stage_out;
        base += total;
        return CW_OK;
    };
    Slot *prev = nullptr;
    size_t k = 0;
    for (size_t lo = 0; lo < n && rc == CW_OK; lo += per, k++) {
        const size_t cnt = (n - lo < per) ? n - lo : per;
        Slot &s = ctx->slots[k % kSlots];
        if ((rc = finish(s)) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_in, &s.cap_in, per * chunk)) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_out, &s.cap_out, per * stride)) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_lens, &s.cap_lens, per * sizeof(uint32_t))) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_dig, &s.cap_dig, per * D)) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_packed, &s.cap_packed, per * stride)) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_offsets, &s.cap_offsets, (per + 1) * sizeof(uint64_t))) != CW_OK) break;
        if ((rc = ensure(ctx, (void **)&s.d_scan, &s.cap_scan, ((per + 1023) / 1024 + 1) * sizeof(uint64_t))) != CW_OK) break;
        if (!s.h_total) if (!cu(cudaHostAlloc((void **)&s.h_total, sizeof(uint64_t), cudaHostAllocDefault), "cudaHostAlloc")) break;
        if (!s.ready) if (!cu(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming), "cudaEventCreateWithFlags")) break;
        s.lo = lo;
        s.cnt = cnt;
        const uint8_t *in = (const uint8_t *)src + lo * chunk;
        if (stage_in) {
            if ((rc = ensure_host(ctx, (void **)&s.h_in, &s.cap_h_in, per * chunk)) != CW_OK) break;
            pool_memcpy(ctx, s.h_in, in, cnt * chunk);
            in = s.h_in;
        }
        if (stage_out) {
            if ((rc = ensure_host(ctx, (void **)&s.h_out, &s.cap_h_out, per * stride)) != CW_OK) break;
            if ((rc = ensure_host(ctx, (void **)&s.h_lens, &s.cap_h_lens, per * sizeof(uint32_t))) != CW_OK) break;
            if ((rc = ensure_host(ctx, (void **)&s.h_dig, &s.cap_h_dig, per * D)) != CW_OK) break;
            if ((rc = ensure_host(ctx, (void **)&s.h_offsets, &s.cap_h_offsets, (per + 1) * sizeof(uint64_t))) != CW_OK) break;
        }
        if (!cu(cudaMemcpyAsync(s.d_in, in, cnt * chunk, cudaMemcpyHostToDevice, s.st), "cudaMemcpyAsync(slice in)")) break;
        if ((rc = do_fused_dev(ctx, halg, calg, mode, s.d_in, chunk, cnt, s.d_dig, s.d_out, stride, s.d_lens, s.st)) != CW_OK) break;
        if (!cu(cw_launch_pack(s.d_out, stride, s.d_lens, s.d_in, (uint32_t)chunk, cnt, s.d_packed, per * stride, s.d_offsets, s.d_scan, s.st),
                "cw_launch_pack"))
            break;
        ctx->launches += 3;
        if (!cu(cudaMemcpyAsync(s.h_total, s.d_offsets + cnt, sizeof(uint64_t), cudaMemcpyDeviceToHost, s.st), "cudaMemcpyAsync")) break;
        if (!cu(cudaEventRecord(s.ready, s.st), "cudaEventRecord")) break;
        if (prev && (rc = stage_b(*prev)) != CW_OK) break;
        prev = &s;
    }
    if (rc == CW_OK && prev) rc = stage_b(*prev);
    for (int i = 0; i < kSlots; i++) {
        Slot &s = ctx->slots[i];
        if (rc == CW_OK) rc = finish(s);
        else if (s.busy) {  // error path: let the copies into the caller's buffers end before returning
            cudaStreamSynchronize(s.st);
            s.busy = s.staged_out = false;
        }
    }
    if (rc != CW_OK) cudaDeviceSynchronize();  // slices whose visit B never ran
    if (rc == CW_OK) offsets[n] = base;
    return rc;
}


// ---------------- small-batch front-end ----------------
bool batch_alloc(cw_ctx *ctx, BatchBuf &b)
{
    if (b.h_in) return true;
    const size_t out_bytes = 2 * kBatchBytes + (kBatchBytes / kBatchMinChunk) * 32;  // stride <= chunk + chunk/255 + 16
    const size_t dig_bytes = (kBatchBytes / kBatchMinChunk) * 64, len_bytes = (kBatchBytes / kBatchMinChunk) * 4;
    bool ok = cudaHostAlloc((void **)&b.h_in, kBatchBytes, cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc((void **)&b.h_out, out_bytes, cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc((void **)&b.h_dig, dig_bytes, cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc((void **)&b.h_lens, len_bytes, cudaHostAllocDefault) == cudaSuccess &&
              cudaMalloc(&b.d_in, kBatchBytes) == cudaSuccess && cudaMalloc(&b.d_out, out_bytes) == cudaSuccess &&
              cudaMalloc(&b.d_dig, dig_bytes) == cudaSuccess && cudaMalloc(&b.d_lens, len_bytes) == cudaSuccess &&
              cudaStreamCreateWithFlags(&b.st, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) fail(ctx, cudaGetLastError(), "batch buffer allocation");
    return ok;
}

void batch_free(BatchBuf &b)
{
    if (b.h_in) cudaFreeHost(b.h_in);
    if (b.h_out) cudaFreeHost(b.h_out);
    if (b.h_dig) cudaFreeHost(b.h_dig);
    if (b.h_lens) cudaFreeHost(b.h_lens);
    if (b.d_in) cudaFree(b.d_in);
    if (b.d_out) cudaFree(b.d_out);
    if (b.d_dig) cudaFree(b.d_dig);
    if (b.d_lens) cudaFree(b.d_lens);
    if (b.done) cudaEventDestroy(b.done);
    if (b.st) cudaStreamDestroy(b.st);
    b = BatchBuf();
}

// one copy in, one launch, copies out; returns with the work enqueued on the buffer's stream
int batch_launch(cw_ctx *ctx, BatchBuf &b)
{
    CW_CUDA(ctx, cudaMemcpyAsync(b.d_in, b.h_in, b.chunks * b.chunk, cudaMemcpyHostToDevice, b.st));
    int rc = do_fused_dev(ctx, b.halg, b.calg, b.mode, b.d_in, b.chunk, b.chunks, b.d_dig, b.d_out, b.stride, b.d_lens, b.st);
    if (rc != CW_OK) return rc;
    CW_CUDA(ctx, cudaMemcpyAsync(b.h_out, b.d_out, b.chunks * b.stride, cudaMemcpyDeviceToHost, b.st));
    CW_CUDA(ctx, cudaMemcpyAsync(b.h_lens, b.d_lens, b.chunks * sizeof(uint32_t), cudaMemcpyDeviceToHost, b.st));
    CW_CUDA(ctx, cudaMemcpyAsync(b.h_dig, b.d_dig, b.chunks * b.D, cudaMemcpyDeviceToHost, b.st));
    CW_CUDA(ctx, cudaEventRecord(b.done, b.st));
    ctx->batches++;
    return CW_OK;
}

// B.mu held: the launch of buffer i is complete
void batch_done(cw_ctx *ctx, int i)
{
    Batcher &B = ctx->bat;
    B.bufs[i].state = BatchBuf::DONE;
    B.cv_callers.notify_all();
    if (!B.bufs[i].jobs.empty()) {  // cw_submit jobs: the offload thread copies their results out and completes them
        {
            std::lock_guard<std::mutex> lk(ctx->q_mu);
            ctx->done_bufs.push_back(i);
        }
        ctx->q_cv.notify_one();
    }
}

void batch_worker(cw_ctx *ctx)
{
    Batcher &B = ctx->bat;
    cudaSetDevice(ctx->device);
    std::unique_lock<std::mutex> lk(B.mu);
    bool left_wait = false;
    for (;;) {
        // a buffer is ready when it is sealed (full, or of another kind than the next caller) or simply holds what arrived
        // while the GPU was busy -- an idle worker takes whatever there is -- and nobody is still copying into it
        int ready = -1;
        for (int i = 0; i < kBatchBufs && ready < 0; i++) {  // sealed buffers first
            BatchBuf &b = B.bufs[i];
            if (b.copying_in == 0 && b.chunks > 0 && b.state == BatchBuf::SEALED) ready = i;
        }
        // (left_wait: the worker only interrupted its wait for a launch to take a sealed buffer -- the open one goes on gathering)
        for (int i = 0; i < kBatchBufs && ready < 0 && !left_wait; i++) {
            BatchBuf &b = B.bufs[i];
            if (b.copying_in == 0 && b.chunks > 0 && b.state == BatchBuf::OPEN) ready = i;
        }
        if (ready >= 0) {
            BatchBuf &b = B.bufs[ready];
            if (B.open == ready) B.open = -1;
            b.state = BatchBuf::RUNNING;
            lk.unlock();
            const int rc = batch_launch(ctx, b);
            lk.lock();
            b.rc = rc;
            if (rc == CW_OK) B.inflight.push_back(ready);
            else {
                cudaStreamSynchronize(b.st);
                batch_done(ctx, ready);
            }
            continue;
        }
        if (!B.inflight.empty()) {  // nothing to launch: finish the oldest launch
            const int i = B.inflight.front();
            BatchBuf &b = B.bufs[i];
            // CW_BATCH_POLL=1: poll the launch instead of blocking on it, and launch a buffer that is sealed meanwhile first.
            // Measured (profiles/r02_small_batch_sweep.txt): blocking 256 KiB items 4.4 -> 8.9 GiB/s, but 1 MiB items
            // 13.0 -> 10.0 and HashOffload 32 KiB items 10.1 -> 9.5: not the default
            static const bool block = getenv("CW_BATCH_POLL") == nullptr;
            cudaError_t e;
            for (;;) {
                lk.unlock();
                e = block ? cudaEventSynchronize(b.done) : cudaEventQuery(b.done);
                if (e == cudaErrorNotReady) std::this_thread::yield();
                lk.lock();
                if (e != cudaErrorNotReady) break;
                bool more = false;
                for (int k = 0; k < kBatchBufs && !more; k++) {
                    const BatchBuf &o = B.bufs[k];
                    // (sealed ones only: the open buffer goes on gathering callers while the GPU is busy, as before)
                    more = o.copying_in == 0 && o.chunks > 0 && o.state == BatchBuf::SEALED;
                }
                if (more) break;
            }
            left_wait = e == cudaErrorNotReady;
            if (left_wait) continue;  // launch what became ready, then come back
            B.inflight.pop_front();
            if (e != cudaSuccess) b.rc = fail(ctx, e, "cudaEventSynchronize(batch)");
            batch_done(ctx, i);
            continue;
        }
        if (B.stop) return;
        B.cv_worker.wait(lk);
    }
}

// reserve room for n chunks in a batch buffer of this kind; returns the buffer (its generation in *gen, first chunk in
// *first) with copying_in and readers raised, or nullptr if the batcher cannot be used.  B.mu held by the caller.
BatchBuf *batch_reserve(cw_ctx *ctx, std::unique_lock<std::mutex> &lk, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode,
                        size_t chunk, size_t n, size_t *first)
{
    Batcher &B = ctx->bat;
    if (B.failed_alloc) return nullptr;
    if (!B.started) {
        B.started = true;
        B.th = std::thread(batch_worker, ctx);
    }
    const size_t stride = cw_compress_bound(calg, chunk);
    for (;;) {
        if (B.open >= 0) {
            BatchBuf &b = B.bufs[B.open];
            bool soft_full = false;
            if (b.chunks * b.chunk >= kBatchSoftBytes)
                for (int i = 0; i < kBatchBufs && !soft_full; i++) soft_full = B.bufs[i].state == BatchBuf::FREE;
            if (b.state == BatchBuf::OPEN && b.halg == halg && b.calg == calg && b.mode == mode && b.chunk == chunk &&
                b.chunks + n <= b.cap_chunks && !soft_full) {
                *first = b.chunks;
                b.chunks += n;
                b.copying_in++;
                b.readers++;
                return &b;
            }
            b.state = BatchBuf::SEALED;  // full, or another kind of work: off it goes
            B.open = -1;
            B.cv_worker.notify_one();
        }
        int f = -1;
        for (int i = 0; i < kBatchBufs && f < 0; i++)
            if (B.bufs[i].state == BatchBuf::FREE) f = i;
        if (f < 0) {
            B.cv_callers.wait(lk);
            continue;
        }
        BatchBuf &b = B.bufs[f];
        if (!batch_alloc(ctx, b)) {
            B.failed_alloc = true;
            return nullptr;
        }
        b.state = BatchBuf::OPEN;
        b.halg = halg;
        b.calg = calg;
        b.mode = mode;
        b.chunk = chunk;
        b.stride = stride;
        b.D = cw_digest_bytes(halg);
        b.cap_chunks = kBatchBytes / chunk;
        b.chunks = 0;
        b.rc = CW_OK;
        B.open = f;  // the reservation itself happens in the branch above
    }
}

void batch_copy_out(const BatchBuf &b, size_t first, size_t n, void *digests, void *dst, size_t stride, uint32_t *lens)
{
    memcpy(digests, b.h_dig + first * b.D, n * b.D);
    memcpy(lens, b.h_lens + first, n * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++)  // only the bytes that exist (LZF: 0 = "store raw", nothing to copy)
        memcpy((uint8_t *)dst + i * stride, b.h_out + (first + i) * b.stride, b.h_lens[first + i]);
}

void batch_release(cw_ctx *ctx, BatchBuf *b)
{
    Batcher &B = ctx->bat;
    std::lock_guard<std::mutex> lk(B.mu);
    if (--b->readers == 0) {
        b->state = BatchBuf::FREE;
        b->jobs.clear();
        b->chunks = 0;
        b->gen++;
        B.cv_callers.notify_all();
    }
}

// a blocking small call: join the open batch, wait for it, take the results.  Returns CW_ERR_UNSUPPORTED if the
// batcher is not available (the caller then uses the direct pipeline).
int batch_call(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src, size_t chunk, size_t n,
               void *digests, void *dst, size_t stride, uint32_t *lens)
{
    Batcher &B = ctx->bat;
    size_t first = 0;
    BatchBuf *b;
    uint64_t gen;
    {
        std::unique_lock<std::mutex> lk(B.mu);
        b = batch_reserve(ctx, lk, halg, calg, mode, chunk, n, &first);
        if (!b) return CW_ERR_UNSUPPORTED;
        gen = b->gen;
    }
    memcpy(b->h_in + first * chunk, src, n * chunk);
    int rc;
    {
        std::unique_lock<std::mutex> lk(B.mu);
        b->copying_in--;
        B.cv_worker.notify_one();
        B.cv_callers.wait(lk, [&] { return b->gen == gen && b->state == BatchBuf::DONE; });
        rc = b->rc;
    }
    if (rc == CW_OK) batch_copy_out(*b, first, n, digests, dst, stride, lens);
    batch_release(ctx, b);
    ctx->batched_calls++;
    return rc;
}

bool batchable(size_t chunk, size_t n)
{
    static const bool off = getenv("CW_NO_BATCHER") != nullptr;  // A/B knob for the profiling scripts
    return !off && chunk >= kBatchMinChunk && n > 0 && n * chunk <= kSmallCall;
}

void complete_job(cw_ctx *ctx, const Job &j, int rc)
{
    {
        // the status is visible before the callback runs, so the callback itself may call cw_wait; a job with a
        // callback keeps no entry (its status was delivered), so callback-only callers do not grow the map
        std::lock_guard<std::mutex> lk(ctx->q_mu);
        if (!j.cb) ctx->finished[j.ticket] = rc;
        ctx->pending.erase(j.ticket);
    }
    ctx->done_cv.notify_all();
    if (j.cb) j.cb(j.ticket, rc, j.user);
}

// cw_submit of a small job: the submitting thread copies the input into the open batch (in parallel with other
// submitters) and returns; the job rides in that batch.  false: the batcher is not available.
bool batch_submit(cw_ctx *ctx, const Job &j)
{
    Batcher &B = ctx->bat;
    size_t first = 0;
    BatchBuf *b;
    {
        std::unique_lock<std::mutex> lk(B.mu);
        b = batch_reserve(ctx, lk, j.halg, j.calg, j.mode, j.chunk, j.n, &first);
        if (!b) return false;
        if (!b->jobs.empty()) b->readers--;  // the submitted jobs of a batch count as one reader: the offload thread
        b->jobs.emplace_back(j, first);
    }
    memcpy(b->h_in + first * j.chunk, j.src, j.n * j.chunk);
    {
        std::lock_guard<std::mutex> lk(B.mu);
        b->copying_in--;
    }
    B.cv_worker.notify_one();
    return true;
}

// jobs [lo, hi) of a finished batch: results to the callers' buffers, then ONE pass over the ticket maps and one wake-up
// for all of them (a notify per 32 KiB job kept 14 waiting threads busy with the mutex), then the callbacks
void finish_job_range(cw_ctx *ctx, BatchBuf &b, size_t lo, size_t hi)
{
    if (b.rc == CW_OK)
        for (size_t k = lo; k < hi; k++) {
            const Job &j = b.jobs[k].first;
            batch_copy_out(b, b.jobs[k].second, j.n, j.digests, j.dst, j.stride, j.lens);
        }
    {
        std::lock_guard<std::mutex> lk(ctx->q_mu);
        for (size_t k = lo; k < hi; k++) {
            const Job &j = b.jobs[k].first;
            if (!j.cb) ctx->finished[j.ticket] = b.rc;
            ctx->pending.erase(j.ticket);
        }
    }
    ctx->done_cv.notify_all();
    for (size_t k = lo; k < hi; k++) {
        const Job &j = b.jobs[k].first;
        if (j.cb) j.cb(j.ticket, b.rc, j.user);
    }
}

// offload thread: a batch that carries submitted jobs is complete
void finish_batch_jobs(cw_ctx *ctx, int i)
{
    BatchBuf &b = ctx->bat.bufs[i];
    const size_t nj = b.jobs.size();
    if (pool_threads() < 2 || nj < 8) {
        finish_job_range(ctx, b, 0, nj);
        ctx->batched_calls += nj;
        batch_release(ctx, &b);
        return;
    }
    CopyPool &P = ctx->pool;
    const size_t pieces = std::min<size_t>(nj / 4, (size_t)pool_threads() * 2);
    {
        std::lock_guard<std::mutex> lk(P.mu);
        pool_start(ctx);
        b.tasks_left = (int)pieces;
        for (size_t k = 0; k < pieces; k++) {
            const size_t lo = nj * k / pieces, hi = nj * (k + 1) / pieces;
            P.tasks.emplace_back([ctx, i, lo, hi] {
                BatchBuf &bb = ctx->bat.bufs[i];
                finish_job_range(ctx, bb, lo, hi);
                bool last;
                {
                    std::lock_guard<std::mutex> lk2(ctx->pool.mu);
                    last = --bb.tasks_left == 0;
                }
                if (last) {
                    ctx->batched_calls += bb.jobs.size();
                    batch_release(ctx, &bb);
                }
            });
        }
    }
    P.cv.notify_all();
}

void worker_loop(cw_ctx *ctx)
{
    for (;;) {
        Job j;
        int done_buf = -1;
        {
            std::unique_lock<std::mutex> lk(ctx->q_mu);
            ctx->q_cv.wait(lk, [&] { return ctx->stop || !ctx->queue.empty() || !ctx->done_bufs.empty(); });
            if (!ctx->done_bufs.empty()) {
                done_buf = ctx->done_bufs.front();
                ctx->done_bufs.pop_front();
            } else if (!ctx->queue.empty()) {
                j = ctx->queue.front();
                ctx->queue.pop_front();
            } else if (ctx->pending.empty())
                return;  // stop requested and everything delivered
            else {
                ctx->q_cv.wait_for(lk, std::chrono::milliseconds(1));  // jobs still riding in batches
                continue;
            }
        }
        if (done_buf >= 0) finish_batch_jobs(ctx, done_buf);
        else complete_job(ctx, j, host_pass(ctx, true, true, j.halg, j.calg, j.mode, j.src, j.chunk, j.n, j.digests, j.dst, j.stride, j.lens));
    }
}

}  // namespace

extern "C" {

int cw_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *cw_status_string(int s)
{
    switch (s) {
    case CW_OK: return "ok";
    case CW_ERR_ARG: return "bad argument";
    case CW_ERR_NO_DEVICE: return "no CUDA device (the engine has no CPU path)";
    case CW_ERR_CUDA: return "CUDA error";
    case CW_ERR_NOMEM: return "out of device memory";
    case CW_ERR_UNSUPPORTED: return "unsupported size or selector";
    }
    return "unknown";
}

int cw_init(cw_ctx **out, int device)
{
    if (!out) return CW_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return CW_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return CW_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return CW_ERR_CUDA;
    cw_ctx *ctx = new cw_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return CW_ERR_CUDA;
    }
    ctx->dev.sm_count = prop.multiProcessorCount;
    ctx->dev.smem_optin = prop.sharedMemPerBlockOptin;
    cw_make_skein_iv(&ctx->iv);
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_work, 2 * kWorkRing * sizeof(unsigned long long)) == cudaSuccess;
    for (int i = 0; ok && i < kWorkRing; i++) ok = cudaEventCreateWithFlags(&ctx->work_ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < kSlots; i++) {
        ok = cudaStreamCreateWithFlags(&ctx->slots[i].st, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->slots[i].done, cudaEventDisableTiming) == cudaSuccess;
    }
    cudaError_t ce = ok ? cw_configure_kernels(ctx->dev) : cudaGetLastError();
    if (!ok || ce != cudaSuccess) {
        g_init_error = std::string("cw_init: ") + cudaGetErrorString(ce != cudaSuccess ? ce : cudaGetLastError());
        cw_shutdown(ctx);
        return CW_ERR_CUDA;
    }
    *out = ctx;
    return CW_OK;
}

void cw_shutdown(cw_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->worker.joinable()) {
        {
            std::lock_guard<std::mutex> lk(ctx->q_mu);
            ctx->stop = true;
        }
        ctx->q_cv.notify_all();
        ctx->worker.join();
    }
    if (!ctx->pool.th.empty()) {  // (the offload thread is gone: nobody adds tasks any more)
        {
            std::lock_guard<std::mutex> lk(ctx->pool.mu);
            ctx->pool.stop = true;
        }
        ctx->pool.cv.notify_all();
        for (auto &t : ctx->pool.th) t.join();
    }
    if (ctx->bat.started) {
        {
            std::lock_guard<std::mutex> lk(ctx->bat.mu);
            ctx->bat.stop = true;
        }
        ctx->bat.cv_worker.notify_all();
        ctx->bat.th.join();
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &b : ctx->bat.bufs) batch_free(b);
    for (auto &s : ctx->slots) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.d_dig) cudaFree(s.d_dig);
        if (s.d_lens) cudaFree(s.d_lens);
        if (s.d_packed) cudaFree(s.d_packed);
        if (s.d_offsets) cudaFree(s.d_offsets);
        if (s.d_scan) cudaFree(s.d_scan);
        if (s.h_total) cudaFreeHost(s.h_total);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.h_dig) cudaFreeHost(s.h_dig);
        if (s.h_lens) cudaFreeHost(s.h_lens);
        if (s.h_offsets) cudaFreeHost(s.h_offsets);
        if (s.ready) cudaEventDestroy(s.ready);
        if (s.done) cudaEventDestroy(s.done);
        if (s.st) cudaStreamDestroy(s.st);
    }
    for (auto &p : ctx->prof_ev) {
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->d_work) cudaFree(ctx->d_work);
    for (auto e : ctx->work_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void cw_batch_stats(const cw_ctx *ctx, uint64_t *batches, uint64_t *calls)
{
    if (batches) *batches = ctx ? ctx->batches.load() : 0;
    if (calls) *calls = ctx ? ctx->batched_calls.load() : 0;
}

const char *cw_last_error(const cw_ctx *ctx) { return ctx ? ctx->last_error.c_str() : g_init_error.c_str(); }

size_t cw_digest_bytes(cw_hash_alg alg)
{
    return alg == CW_HASH_SKEIN512 ? 64 : (alg == CW_HASH_SKEIN256_128 ? 16 : (alg == CW_HASH_SHA256 ? 32 : 0));
}

size_t cw_compress_bound(cw_comp_alg alg, size_t chunk)
{
    if (alg == CW_COMP_LZ4) return chunk + chunk / 255 + 16;
    return chunk;
}

int cw_host_register(cw_ctx *ctx, void *ptr, size_t bytes)
{
    if (!ctx || !ptr || bytes == 0) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    CW_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return CW_OK;
}

int cw_host_unregister(cw_ctx *ctx, void *ptr)
{
    if (!ctx || !ptr) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    CW_CUDA(ctx, cudaHostUnregister(ptr));
    return CW_OK;
}

int cw_hash_dev(cw_ctx *ctx, cw_hash_alg alg, const void *d_src, size_t chunk, size_t n, void *d_digests, void *stream)
{
    int rc = check_common(ctx, chunk, n, d_src);
    if (rc != CW_OK) return rc;
    if (n && !d_digests) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    return do_hash_dev(ctx, alg, d_src, chunk, n, d_digests, stream ? (cudaStream_t)stream : ctx->stream);
}

int cw_compress_dev(cw_ctx *ctx, cw_comp_alg alg, cw_mode mode, const void *d_src, size_t chunk, size_t n, void *d_dst,
                    size_t stride, uint32_t *d_lens, void *stream)
{
    int rc = check_common(ctx, chunk, n, d_src);
    if (rc != CW_OK) return rc;
    if (n && (!d_dst || !d_lens)) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    return do_compress_dev(ctx, alg, mode, d_src, chunk, n, d_dst, stride, d_lens, stream ? (cudaStream_t)stream : ctx->stream);
}

int cw_hash_compress_dev(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *d_src, size_t chunk,
                         size_t n, void *d_digests, void *d_dst, size_t stride, uint32_t *d_lens, void *stream)
{
    int rc = check_common(ctx, chunk, n, d_src);
    if (rc != CW_OK) return rc;
    if (n && (!d_dst || !d_lens || !d_digests)) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    return do_fused_dev(ctx, halg, calg, mode, d_src, chunk, n, d_digests, d_dst, stride, d_lens,
                        stream ? (cudaStream_t)stream : ctx->stream);
}

int cw_sync(cw_ctx *ctx, void *stream)
{
    if (!ctx) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    CW_CUDA(ctx, cudaStreamSynchronize(stream ? (cudaStream_t)stream : ctx->stream));
    return CW_OK;
}

int cw_hash(cw_ctx *ctx, cw_hash_alg alg, const void *src, size_t chunk, size_t n, void *digests)
{
    int rc = check_common(ctx, chunk, n, src);
    if (rc != CW_OK) return rc;
    if ((int)alg < 0 || (int)alg > 2 || (n && !digests)) return CW_ERR_ARG;
    return host_pass(ctx, true, false, alg, CW_COMP_LZ4, CW_MODE_COMPAT, src, chunk, n, digests, nullptr, 0, nullptr);
}

int cw_hash_skein(cw_ctx *ctx, unsigned state_bits, unsigned hash_bits, const void *src, size_t chunk, size_t n, void *digests)
{
    int rc = check_common(ctx, chunk, n, src);
    if (rc != CW_OK) return rc;
    if ((state_bits != 512u && state_bits != 256u) || hash_bits == 0u || hash_bits % 8u || hash_bits > state_bits) return CW_ERR_ARG;
    if (n && !digests) return CW_ERR_ARG;
    if (n == 0) return CW_OK;
    // a KAT-sized convenience path: one allocation, one copy each way, the caller's thread waits.  The chunk kernels of
    // cw_hash run with the chaining value of the requested digest length and write whole output blocks; the digest is
    // the head of its block
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t D = hash_bits / 8u, B = state_bits / 8u;
    CwSkeinIV iv;
    cw_make_skein_iv_bits(&iv, state_bits, hash_bits);
    const size_t in_bytes = (n * chunk + 255) & ~(size_t)255;
    uint8_t *d_in = nullptr;
    CW_CUDA(ctx, cudaMalloc(&d_in, in_bytes + n * B));
    uint8_t *d_out = d_in + in_bytes;
    std::vector<uint8_t> blocks(n * B);
    cudaError_t e = cudaMemcpy(d_in, src, n * chunk, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cw_launch_hash(state_bits == 512u ? 0 : 3, iv, d_in, (uint32_t)chunk, n, d_out, 0);
    if (e == cudaSuccess) ctx->launches++;
    if (e == cudaSuccess) e = cudaMemcpy(blocks.data(), d_out, n * B, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    CW_CUDA(ctx, e);
    for (size_t i = 0; i < n; i++) memcpy((uint8_t *)digests + i * D, blocks.data() + i * B, D);
    return CW_OK;
}

int cw_compress(cw_ctx *ctx, cw_comp_alg alg, cw_mode mode, const void *src, size_t chunk, size_t n, void *dst,
                size_t stride, uint32_t *lens)
{
    int rc = check_common(ctx, chunk, n, src);
    if (rc != CW_OK) return rc;
    if ((int)alg < 0 || (int)alg > 1 || (int)mode < 0 || (int)mode > 1 || (n && (!dst || !lens))) return CW_ERR_ARG;
    if (stride < cw_compress_bound(alg, chunk)) return CW_ERR_ARG;
    return host_pass(ctx, false, true, CW_HASH_SKEIN512, alg, mode, src, chunk, n, nullptr, dst, stride, lens);
}

int cw_hash_compress(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src, size_t chunk, size_t n,
                     void *digests, void *dst, size_t stride, uint32_t *lens)
{
    int rc = check_common(ctx, chunk, n, src);
    if (rc != CW_OK) return rc;
    if ((int)halg < 0 || (int)halg > 2 || (int)calg < 0 || (int)calg > 1 || (int)mode < 0 || (int)mode > 1) return CW_ERR_ARG;
    if (n && (!dst || !lens || !digests)) return CW_ERR_ARG;
    if (stride < cw_compress_bound(calg, chunk)) return CW_ERR_ARG;
    if (batchable(chunk, n)) {  // small work items from many threads share batches (HashAndCompress.cpp:263-272)
        rc = batch_call(ctx, halg, calg, mode, src, chunk, n, digests, dst, stride, lens);
        if (rc != CW_ERR_UNSUPPORTED) return rc;
    }
    return host_pass(ctx, true, true, halg, calg, mode, src, chunk, n, digests, dst, stride, lens);
}

int cw_hash_compress_packed(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src, size_t chunk,
                            size_t n, void *digests, void *packed, size_t packed_cap, uint64_t *offsets, uint32_t *lens)
{
    int rc = check_common(ctx, chunk, n, src);
    if (rc != CW_OK) return rc;
    if ((int)halg < 0 || (int)halg > 2 || (int)calg < 0 || (int)calg > 1 || (int)mode < 0 || (int)mode > 1) return CW_ERR_ARG;
    if (!offsets || (n && (!packed || !lens || !digests))) return CW_ERR_ARG;
    if (n == 0) {
        offsets[0] = 0;
        return CW_OK;
    }
    return host_pass_packed(ctx, halg, calg, mode, src, chunk, n, digests, packed, packed_cap, offsets, lens);
}

int cw_submit(cw_ctx *ctx, cw_hash_alg halg, cw_comp_alg calg, cw_mode mode, const void *src, size_t chunk, size_t n,
              void *digests, void *dst, size_t stride, uint32_t *lens, cw_complete_fn cb, void *user, cw_ticket *ticket)
{
    int rc = check_common(ctx, chunk, n, src);
    if (rc != CW_OK) return rc;
    if (!ticket) return CW_ERR_ARG;
    if ((int)halg < 0 || (int)halg > 2 || (int)calg < 0 || (int)calg > 1 || (int)mode < 0 || (int)mode > 1) return CW_ERR_ARG;
    if (n && (!dst || !lens || !digests)) return CW_ERR_ARG;
    if (stride < cw_compress_bound(calg, chunk)) return CW_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->q_mu);
    if (!ctx->worker.joinable()) ctx->worker = std::thread(worker_loop, ctx);
    Job j{ctx->next_ticket++, halg, calg, mode, src, chunk, n, digests, dst, stride, lens, cb, user};
    *ticket = j.ticket;
    ctx->pending[j.ticket] = true;
    if (batchable(chunk, n)) {
        lk.unlock();
        if (batch_submit(ctx, j)) return CW_OK;
        lk.lock();
    }
    ctx->queue.push_back(j);
    ctx->q_cv.notify_one();
    return CW_OK;
}

int cw_wait(cw_ctx *ctx, cw_ticket ticket)
{
    if (!ctx) return CW_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->q_mu);
    if (ticket == 0 || ticket >= ctx->next_ticket) return CW_ERR_ARG;
    ctx->done_cv.wait(lk, [&] { return ctx->pending.find(ticket) == ctx->pending.end(); });
    auto it = ctx->finished.find(ticket);
    if (it == ctx->finished.end()) return CW_ERR_ARG;  // already waited for, or its status went to the completion callback
    const int rc = it->second;
    ctx->finished.erase(it);
    return rc;
}

int cw_decompress_dev(cw_ctx *ctx, cw_comp_alg alg, const void *d_blocks, size_t stride, const uint32_t *d_lens,
                      size_t chunk, size_t n, void *d_out, uint32_t *d_out_lens, void *stream)
{
    int rc = check_common(ctx, chunk, n, d_blocks);
    if (rc != CW_OK) return rc;
    if ((int)alg < 0 || (int)alg > 1 || (n && (!d_lens || !d_out || !d_out_lens))) return CW_ERR_ARG;
    if (alg == CW_COMP_LZF && stride < chunk) return CW_ERR_ARG;  // raw-stored chunks need chunk bytes per slot
    if (n == 0) return CW_OK;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const WorkSlot w = next_work(ctx);
    CW_CUDA(ctx, cudaMemsetAsync(w.p, 0, sizeof *w.p, st));
    ProfScope ps(ctx, st);
    CW_CUDA(ctx, cw_launch_decode(ctx->dev, (int)alg, (const uint8_t *)d_blocks, stride, d_lens, (uint32_t)chunk, n,
                                  (uint8_t *)d_out, d_out_lens, w.p, st));
    work_done(ctx, w, st);
    ctx->launches++;
    return CW_OK;
}

int cw_decompress(cw_ctx *ctx, cw_comp_alg alg, const void *blocks, size_t stride, const uint32_t *lens, size_t chunk, size_t n,
                  void *out, uint32_t *out_lens)
{
    int rc = check_common(ctx, chunk, n, blocks);
    if (rc != CW_OK) return rc;
    if ((int)alg < 0 || (int)alg > 1 || (n && (!lens || !out || !out_lens))) return CW_ERR_ARG;
    if (n == 0) return CW_OK;
    std::lock_guard<std::mutex> hg(ctx->host_mu);
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t *d_b = nullptr, *d_o = nullptr;
    uint32_t *d_l = nullptr, *d_ol = nullptr;
    cudaError_t e = cudaMalloc(&d_b, n * stride);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, n * chunk);
    if (e == cudaSuccess) e = cudaMalloc(&d_l, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_ol, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, blocks, n * stride, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_l, lens, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        rc = cw_decompress_dev(ctx, alg, d_b, stride, d_l, chunk, n, d_o, d_ol, ctx->stream);
        if (rc == CW_OK) {
            e = cudaMemcpyAsync(out, d_o, n * chunk, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_lens, d_ol, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        }
    }
    cudaFree(d_b);
    cudaFree(d_o);
    cudaFree(d_l);
    cudaFree(d_ol);
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? CW_ERR_NOMEM : fail(ctx, e, "cw_decompress");
    return rc;
}

int cw_pack_dev(cw_ctx *ctx, const void *d_blocks, size_t stride, const uint32_t *d_lens, const void *d_src, size_t chunk,
                size_t n, void *d_packed, size_t packed_cap, uint64_t *d_offsets, void *stream)
{
    int rc = check_common(ctx, chunk, n, d_blocks);
    if (rc != CW_OK) return rc;
    if (n && (!d_lens || !d_offsets || (!d_packed && packed_cap))) return CW_ERR_ARG;
    if (n == 0) return CW_OK;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    // the tile totals live in a per-call allocation ordered on the caller's stream (stream-ordered allocator), so calls
    // on different streams or from different threads never share scratch
    const size_t need = ((n + 1023) / 1024 + 1) * sizeof(unsigned long long);
    unsigned long long *scratch = nullptr;
    if (cudaMallocAsync(&scratch, need, st) != cudaSuccess) {
        cudaGetLastError();
        return CW_ERR_NOMEM;
    }
    const cudaError_t le = cw_launch_pack((const uint8_t *)d_blocks, stride, d_lens, (const uint8_t *)d_src, (uint32_t)chunk, n,
                                          (uint8_t *)d_packed, packed_cap, (unsigned long long *)d_offsets, scratch, st);
    cudaFreeAsync(scratch, st);
    if (le != cudaSuccess) return fail(ctx, le, "cw_launch_pack");
    ctx->launches += 3;
    return CW_OK;
}

int cw_digest_hist_dev(cw_ctx *ctx, const void *d_digests, size_t D, size_t n, uint64_t *d_hist, void *stream)
{
    if (!ctx || D == 0 || D > 1024 || !d_hist || (n && !d_digests)) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    CW_CUDA(ctx, cw_launch_digest_hist(ctx->dev, (const uint8_t *)d_digests, (uint32_t)D, n, (unsigned long long *)d_hist,
                                       stream ? (cudaStream_t)stream : ctx->stream));
    if (n) ctx->launches++;
    return CW_OK;
}

int cw_digest_byte_stats(const uint64_t *hist, size_t n_bytes, double *avg, double *stddev, double *avg_stddev)
{
    if (!hist || n_bytes == 0) return CW_ERR_ARG;
    double sum_sd = 0;
    for (size_t b = 0; b < n_bytes; b++) {
        const uint64_t *h = hist + b * 256;
        // the reference's map holds the values 0..254 from the start and 255 only once it has been seen
        const int present = h[255] ? 256 : 255;
        uint64_t total = 0;
        for (int v = 0; v < present; v++) total += h[v];
        const double a = (double)total / 255.0;
        double sd = 0;
        for (int v = 0; v < present; v++) {
            const double d = (double)h[v] - a;
            sd += d * d;
        }
        sd = std::sqrt(sd / 254.0);
        if (avg) avg[b] = a;
        if (stddev) stddev[b] = sd;
        sum_sd += sd;
    }
    if (avg_stddev) *avg_stddev = sum_sd / (double)n_bytes;
    return CW_OK;
}

int cw_digest_stats(cw_ctx *ctx, const void *digests, size_t D, size_t n, size_t n_bytes, double *avg, double *stddev,
                    double *avg_stddev)
{
    if (!ctx || D == 0 || D > 1024 || n_bytes == 0 || n_bytes > D || (n && !digests)) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> hg(ctx->host_mu);
    uint8_t *d_dig = nullptr;
    unsigned long long *d_hist = nullptr;
    std::vector<uint64_t> hist(D * 256);
    cudaError_t e = cudaMalloc(&d_hist, hist.size() * sizeof(uint64_t));
    if (e == cudaSuccess && n) e = cudaMalloc(&d_dig, n * D);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(d_dig, digests, n * D, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cw_launch_digest_hist(ctx->dev, d_dig, (uint32_t)D, n, d_hist, ctx->stream);
    if (e == cudaSuccess && n) ctx->launches++;
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(hist.data(), d_hist, hist.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_dig);
    cudaFree(d_hist);
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? CW_ERR_NOMEM : fail(ctx, e, "cw_digest_stats");
    return cw_digest_byte_stats(hist.data(), n_bytes, avg, stddev, avg_stddev);
}

int cw_dedup_dev(cw_ctx *ctx, const void *d_digests, size_t D, size_t n, uint32_t *d_first, uint64_t *d_n_unique, void *stream)
{
    if (!ctx || D == 0 || (D & 3u) || D > 1024 || !d_n_unique || (n && (!d_digests || !d_first))) return CW_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_digests) & 3u) || n >= (1ull << 31)) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    uint32_t *table = nullptr;  // per-call slot table, ordered on the caller's stream (see cw_pack_dev)
    if (cudaMallocAsync(&table, slots * sizeof(uint32_t), st) != cudaSuccess) {
        cudaGetLastError();
        return CW_ERR_NOMEM;
    }
    const cudaError_t le = cw_launch_dedup((const uint8_t *)d_digests, (uint32_t)D, n, table, (uint32_t)slots, d_first,
                                           (unsigned long long *)d_n_unique, st);
    cudaFreeAsync(table, st);
    if (le != cudaSuccess) return fail(ctx, le, "cw_launch_dedup");
    if (n) ctx->launches += 2;
    return CW_OK;
}

int cw_synth_dev(cw_ctx *ctx, void *d_dst, size_t chunk, uint64_t first_chunk, size_t n, uint64_t seed, void *stream)
{
    if (!ctx || !d_dst || chunk == 0 || (chunk & 7u) || (reinterpret_cast<uintptr_t>(d_dst) & 7u)) return CW_ERR_ARG;
    if (chunk > CW_MAX_CHUNK) return CW_ERR_UNSUPPORTED;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    CW_CUDA(ctx, cw_launch_synth((uint8_t *)d_dst, (uint32_t)chunk, first_chunk, n, seed,
                                 stream ? (cudaStream_t)stream : ctx->stream));
    ctx->launches++;
    return CW_OK;
}

uint64_t cw_launch_count(const cw_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int cw_prof_enable(cw_ctx *ctx, int on)
{
    if (!ctx) return CW_ERR_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->prof = on != 0;
    return CW_OK;
}

int cw_prof_read(cw_ctx *ctx, double *kernel_ms, uint64_t *launches)
{
    if (!ctx) return CW_ERR_ARG;
    CW_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        evs.swap(ctx->prof_ev);
    }
    for (auto &p : evs) {
        CW_CUDA(ctx, cudaEventSynchronize(p.second));
        float ms = 0;
        CW_CUDA(ctx, cudaEventElapsedTime(&ms, p.first, p.second));
        ctx->prof_ms += ms;
        ctx->prof_launches++;
    }
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        for (auto &p : evs) {
            ctx->ev_pool.push_back(p.first);
            ctx->ev_pool.push_back(p.second);
        }
    }
    if (kernel_ms) *kernel_ms = ctx->prof_ms;
    if (launches) *launches = ctx->prof_launches;
    ctx->prof_ms = 0;
    ctx->prof_launches = 0;
    return CW_OK;
}

}  // extern "C"
