// dfl_api.cu -- C ABI of libdeflate_b200.so (see include/deflate_b200.h).
//
// Host side of the boundary only: argument marshalling, device buffer ownership, container
// framing decisions, the streaming handle.  All arithmetic on the data happens in the kernels of
// dfl_kernels.cu; there is no CPU implementation to fall back to.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>
#include <vector>

#include "../../include/deflate_b200.h"
#include "dfl_internal.h"

using namespace dfl;

namespace {

thread_local std::string t_cuda_err;
thread_local int t_profiling = 0;
thread_local uint32_t t_peers = 1;   // pipelines issued concurrently by the calling thread
thread_local std::vector<std::pair<const char*, float>> t_stage_ms;
thread_local uint64_t t_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};

int cuda_fail(cudaError_t e, const char* where) {
    t_cuda_err = std::string(where) + ": " + cudaGetErrorString(e);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return DFL_E_NODEVICE;
    if (e == cudaErrorMemoryAllocation) return DFL_E_NOMEM;
    return DFL_E_CUDA;
}
#define CK(expr)                                           \
    do {                                                   \
        cudaError_t e_ = (expr);                           \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr); \
    } while (0)

int device_count_cached() {
    static int count = -1;
    static std::once_flag once;
    std::call_once(once, [] {
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess) { c = 0; (void)cudaGetLastError(); }
        count = c;
    });
    return count;
}

// Streams, scratch and kernel attributes belong to one CUDA device: everything that caches them is keyed by it.
int current_device() {
    int dev = 0;
    if (device_count_cached() <= 0 || cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return dev;
}
// Makes `dev` current for the lifetime of the guard (an encoder handle stays on the device it was created on,
// whatever device the calling thread has selected since).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (dev < 0 || device_count_cached() <= 0) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess) prev = cur;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
int dev_alloc(T*& p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    p = reinterpret_cast<T*>(q);
    return DFL_OK;
}
template <typename T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// Input that is still on its way to the device: slice k = [lo[k], lo[k + 1]) is complete once
// ev[k] has fired (recorded on the copy stream by feed(k)).
// The pipeline launches its first two stages once per slice, and every launch ends in a partly idle last wave
// (one k_match CTA runs 1.7 ms at Default), so slices should be few; the first must be short so that the kernels
// start early.  At Default the link is ~7x faster than the kernels (19 ms against 134 ms per GiB): slices may grow
// geometrically without the kernels ever waiting.  With a short chain budget copy and kernels run at about the
// same pace and the CTAs are short: uniform slices.
constexpr size_t kCopySlice = 128u << 20;
constexpr size_t kCopyFirst = 32u << 20;
inline std::vector<size_t> plan_slices(size_t n, bool long_kernels) {
    std::vector<size_t> lo{0};
    if (n == 0) return lo;
    size_t at = kCopyFirst;
    while (at < n) {
        lo.push_back(at);
        at = long_kernels ? (at == kCopyFirst ? (size_t)256u << 20 : at * 4) : at + kCopySlice;
    }
    lo.push_back(n);
    return lo;
}

// ---- pageable host memory <-> device, staged by this library -------------------------------------------------
// cudaMemcpyAsync on pageable memory is a single-threaded staged copy (measured on the B200 hosts: 6-10 GB/s in,
// 4-5 GB/s out, against 55 GB/s for pinned memory) and would bound every host-facing call.  Large pageable
// transfers go through a ring of pinned slots instead: a few worker threads move the bytes between the caller's
// memory and the slots in parallel, the DMA engine moves the slots.  The workers only memcpy -- every CUDA call
// stays on the calling thread.
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool* p = new CopyPool();   // lives as long as the process: its threads never outlive it
        return *p;
    }
    void submit(void* dst, const void* src, size_t n, std::atomic<int>* done) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(Job{dst, src, n, done});
        }
        cv_job_.notify_one();
    }
    void wait(const std::atomic<int>* done) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return done->load(std::memory_order_acquire) != 0; });
    }

private:
    struct Job { void* dst; const void* src; size_t n; std::atomic<int>* done; };
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw >= 12 ? 4 : (hw >= 6 ? 3 : (hw >= 3 ? 2 : 1));
        if (const char* e = getenv("DFL_COPY_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 32) n = (unsigned)v; }
        for (unsigned i = 0; i < n; i++) std::thread([this] { run(); }).detach();
    }
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [&] { return !q_.empty(); });
                j = q_.front();
                q_.pop_front();
            }
            memcpy(j.dst, j.src, j.n);
            {
                std::lock_guard<std::mutex> lk(mu_);
                j.done->store(1, std::memory_order_release);
            }
            cv_done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_;
    std::deque<Job> q_;
};

constexpr size_t kStageSlot = 1u << 20;     // bytes per pinned slot
constexpr size_t kStageSlots = 32;          // slots per direction
constexpr size_t kStageAhead = 16;          // device-to-host: DMAs kept in flight in front of the memcpy jobs
constexpr size_t kStageMin = 2u << 20;      // smaller pageable transfers are left to the driver

inline bool host_pointer_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Copy jobs hold pointers into the caller's buffer, the slot rings and the completion flags of the call that
// submitted them: whichever way that call returns -- also with a CUDA error half way -- it first waits for every
// job it has submitted.
struct JobDrain {
    CopyPool& pool;
    std::atomic<int>* done;
    size_t submitted = 0;
    ~JobDrain() { for (size_t i = 0; i < submitted; i++) pool.wait(&done[i]); }
};

struct Stager {   // one per Context, allocated on first use
    uint8_t* ring[2] = {nullptr, nullptr};   // [0] host-to-device, [1] device-to-host
    cudaEvent_t ev[2][kStageSlots] = {};
    uint64_t seq = 0;                         // host-to-device slots are handed out round robin across calls
    bool ok[2] = {false, false};
    int init(int d) {   // each direction's ring is pinned when it is first needed
        if (ok[d]) return DFL_OK;
        CK(cudaMallocHost(reinterpret_cast<void**>(&ring[d]), kStageSlot * kStageSlots));
        for (auto& e : ev[d]) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ok[d] = true;
        return DFL_OK;
    }
    ~Stager() {
        for (int d = 0; d < 2; d++) {
            if (ring[d]) cudaFreeHost(ring[d]);
            for (auto e : ev[d]) if (e) cudaEventDestroy(e);
        }
    }
    // Returns once h_src has been read completely; the last DMAs may still be running on `st`.
    int h2d(uint8_t* d_dst, const uint8_t* h_src, size_t len, cudaStream_t st) {
        int rc = init(0);
        if (rc) return rc;
        CopyPool& pool = CopyPool::get();
        const size_t n_chunks = (len + kStageSlot - 1) / kStageSlot;
        std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[n_chunks]);
        for (size_t i = 0; i < n_chunks; i++) done[i].store(0, std::memory_order_relaxed);
        JobDrain drain{pool, done.get()};
        size_t enq = 0;   // chunks whose DMA has been queued (in order)
        auto chunk_len = [&](size_t i) { return (i + 1 == n_chunks) ? len - i * kStageSlot : kStageSlot; };
        auto queue_dma = [&](size_t i) -> int {
            const size_t slot = (size_t)((seq + i) % kStageSlots);
            CK(cudaMemcpyAsync(d_dst + i * kStageSlot, ring[0] + slot * kStageSlot, chunk_len(i), cudaMemcpyHostToDevice, st));
            CK(cudaEventRecord(ev[0][slot], st));
            return DFL_OK;
        };
        for (size_t i = 0; i < n_chunks; i++) {
            while (enq < i && done[enq].load(std::memory_order_acquire)) { if ((rc = queue_dma(enq))) return rc; enq++; }
            if (i >= kStageSlots)   // the slot's previous chunk of this call must be on its way before its event means anything
                while (enq <= i - kStageSlots) { pool.wait(&done[enq]); if ((rc = queue_dma(enq))) return rc; enq++; }
            const size_t slot = (size_t)((seq + i) % kStageSlots);
            CK(cudaEventSynchronize(ev[0][slot]));   // the DMA that last read this slot (this call or an earlier one)
            pool.submit(ring[0] + slot * kStageSlot, h_src + i * kStageSlot, chunk_len(i), &done[i]);
            drain.submitted = i + 1;
        }
        while (enq < n_chunks) { pool.wait(&done[enq]); if ((rc = queue_dma(enq))) return rc; enq++; }
        seq += n_chunks;
        return DFL_OK;
    }
    // Small writes: the caller gathers bytes in the next host-to-device slot itself (one pass over the bytes,
    // no driver staging) and sends it off when it is full.  No h2d() between gather_begin and gather_commit.
    int gather_begin(uint8_t** slot_ptr) {
        int rc = init(0);
        if (rc) return rc;
        const size_t slot = (size_t)(seq % kStageSlots);
        CK(cudaEventSynchronize(ev[0][slot]));
        *slot_ptr = ring[0] + slot * kStageSlot;
        return DFL_OK;
    }
    int gather_commit(uint8_t* d_dst, size_t fill, cudaStream_t st) {
        const size_t slot = (size_t)(seq % kStageSlots);
        CK(cudaMemcpyAsync(d_dst, ring[0] + slot * kStageSlot, fill, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ev[0][slot], st));
        seq++;
        return DFL_OK;
    }
    // Blocking: h_dst is complete on return.  d_src must be ready in stream order on `st`.
    int d2h(uint8_t* h_dst, const uint8_t* d_src, size_t len, cudaStream_t st) {
        int rc = init(1);
        if (rc) return rc;
        CopyPool& pool = CopyPool::get();
        const size_t n_chunks = (len + kStageSlot - 1) / kStageSlot;
        std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[n_chunks]);
        for (size_t i = 0; i < n_chunks; i++) done[i].store(0, std::memory_order_relaxed);
        JobDrain drain{pool, done.get()};
        auto chunk_len = [&](size_t i) { return (i + 1 == n_chunks) ? len - i * kStageSlot : kStageSlot; };
        for (size_t i = 0; i < n_chunks + kStageAhead; i++) {
            if (i < n_chunks) {
                if (i >= kStageSlots) pool.wait(&done[i - kStageSlots]);   // the slot has been emptied
                const size_t slot = i % kStageSlots;
                CK(cudaMemcpyAsync(ring[1] + slot * kStageSlot, d_src + i * kStageSlot, chunk_len(i), cudaMemcpyDeviceToHost, st));
                CK(cudaEventRecord(ev[1][slot], st));
            }
            if (i >= kStageAhead && i - kStageAhead < n_chunks) {
                const size_t j = i - kStageAhead, slot = j % kStageSlots;
                CK(cudaEventSynchronize(ev[1][slot]));
                pool.submit(h_dst + j * kStageSlot, ring[1] + slot * kStageSlot, chunk_len(j), &done[j]);
                drain.submitted = j + 1;
            }
        }
        for (size_t i = 0; i < n_chunks; i++) pool.wait(&done[i]);
        return DFL_OK;
    }
};

struct Context {
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copy_ev;
    Buffers buf;
    uint8_t* d_in = nullptr;    // staging for host-buffer calls
    size_t d_in_cap = 0;
    uint8_t* d_out = nullptr;
    size_t d_out_cap = 0;
    uint8_t* d_out2 = nullptr;      // second output scratch (pieces alternate, so one can be copied out while the other fills)
    size_t d_out2_cap = 0;
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t d2h_ev[2] = {nullptr, nullptr};
    uint32_t* d_tok_in = nullptr;   // staging for dfl_encode_tokens
    size_t d_tok_cap = 0;
    DevMeta* h_meta = nullptr;  // pinned
    std::unique_ptr<Stager> stager;
    bool ok = false;

    // Host memory -> device on `st`; returns once `src` may be reused (pinned memory: only if wait_pinned).
    int copy_in(uint8_t* d_dst, const uint8_t* src, size_t len, cudaStream_t st, bool wait_pinned) {
        if (!len) return DFL_OK;
        const bool pageable = host_pointer_is_pageable(src);
        if (pageable && len >= kStageMin) {
            if (!stager) stager.reset(new Stager());
            return stager->h2d(d_dst, src, len, st);
        }
        CK(cudaMemcpyAsync(d_dst, src, len, cudaMemcpyHostToDevice, st));   // pageable: staged by the driver before it returns
        if (!pageable && wait_pinned) CK(cudaStreamSynchronize(st));
        return DFL_OK;
    }
    // Device -> host memory in stream order on `st`; blocking.
    int copy_out(uint8_t* dst, const uint8_t* d_src, size_t len, cudaStream_t st) {
        if (!len) return DFL_OK;
        if (len >= kStageMin && host_pointer_is_pageable(dst)) {
            if (!stager) stager.reset(new Stager());
            return stager->d2h(dst, d_src, len, st);
        }
        CK(cudaMemcpyAsync(dst, d_src, len, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return DFL_OK;
    }

    int init() {
        if (ok) return DFL_OK;
        if (device_count_cached() <= 0) {
            t_cuda_err = "no CUDA device";
            return DFL_E_NODEVICE;
        }
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&d2h_stream, cudaStreamNonBlocking));
        for (auto& e : d2h_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaMallocHost(reinterpret_cast<void**>(&h_meta), sizeof(DevMeta)));
        int rc = dev_alloc(buf.meta, 1);
        if (rc) return rc;
        ok = true;
        return DFL_OK;
    }
    void free_buffers() {
        Buffers& b = buf;
        dev_free(b.K); dev_free(b.K2); dev_free(b.off); dev_free(b.seq_tab); dev_free(b.Mf); dev_free(b.Mq); dev_free(b.segtok);
        dev_free(b.seg_e_pos); dev_free(b.seg_e_key); dev_free(b.seg_e_tok);
        dev_free(b.seg_x_pos); dev_free(b.seg_x_key); dev_free(b.seg_x_tok);
        dev_free(b.seg_start_pos); dev_free(b.seg_start_key); dev_free(b.seg_bad);
        dev_free(b.seg_cnt); dev_free(b.seg_off); dev_free(b.hist); dev_free(b.cost); dev_free(b.tables);
        dev_free(b.blk_type); dev_free(b.blk_bit); dev_free(b.blk_in); dev_free(b.adler_part);
        b.tok = nullptr;
        b.cap_n = 0;
        b.cap_quarter = false;
    }
    ~Context() {
        if (!ok) return;
        free_buffers();
        dev_free(buf.meta);
        dev_free(d_in); dev_free(d_out); dev_free(d_tok_in);
        if (h_meta) cudaFreeHost(h_meta);
        for (auto e : copy_ev) cudaEventDestroy(e);
        for (auto e : d2h_ev) if (e) cudaEventDestroy(e);
        if (d2h_stream) cudaStreamDestroy(d2h_stream);
        dev_free(d_out2);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }

    // Scratch for inputs of up to n bytes (history + payload).
    int ensure(size_t n, bool quarter) {
        Buffers& b = buf;
        if (b.cap_n >= n && b.cap_n > 0 && (b.cap_quarter || !quarter)) return DFL_OK;
        size_t cap = n + n / 8 + 65536;
        if (cap < b.cap_n) cap = b.cap_n;      // growing one kind of scratch never shrinks another
        if (cap > 0xfffffff0ull) cap = 0xfffffff0ull;
        if (cap < n) return DFL_E_ARG;
        cudaStreamSynchronize(stream);
        quarter = quarter || b.cap_quarter;
        free_buffers();
        size_t n_win = (cap + kWindow - 1) / kWindow + 1;
        size_t n_seg = parse_max_segments(cap) + 1;
        size_t n_blk = max_blocks_for((uint32_t)cap) + 1;
        size_t n_chunk = (cap + kAdlerChunk - 1) / kAdlerChunk + 1;
        int rc = 0;
        if ((rc = dev_alloc(b.K, n_win * kWindow))) return rc;
        if ((rc = dev_alloc(b.K2, n_win * kWindow))) return rc;
        if ((rc = dev_alloc(b.off, n_win * kWindow))) return rc;
        if ((rc = dev_alloc(b.Mf, cap))) return rc;
        if (quarter && (rc = dev_alloc(b.Mq, cap))) return rc;
        if ((rc = dev_alloc(b.segtok, parse_buffer_words(cap)))) return rc;
        if ((rc = dev_alloc(b.seg_e_pos, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_e_key, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_e_tok, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_x_pos, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_x_key, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_x_tok, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_start_pos, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_start_key, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_bad, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_cnt, n_seg))) return rc;
        if ((rc = dev_alloc(b.seg_off, n_seg))) return rc;
        if ((rc = dev_alloc(b.hist, n_blk * 320))) return rc;
        if ((rc = dev_alloc(b.cost, n_blk))) return rc;
        if ((rc = dev_alloc(b.tables, n_blk))) return rc;
        if ((rc = dev_alloc(b.blk_type, n_blk))) return rc;
        if ((rc = dev_alloc(b.blk_bit, n_blk + 1))) return rc;
        if ((rc = dev_alloc(b.blk_in, n_blk + 1))) return rc;
        if ((rc = dev_alloc(b.adler_part, 2 * n_chunk))) return rc;
        b.tok = reinterpret_cast<uint32_t*>(b.K);   // the candidate lists are dead once k_match has run; the token stream reuses them
        b.cap_n = cap;
        b.cap_quarter = quarter;
        return DFL_OK;
    }
    int ensure_stage(uint8_t*& p, size_t& cap, size_t need) {
        if (cap >= need) return DFL_OK;
        cudaStreamSynchronize(stream);
        dev_free(p);
        cap = 0;
        size_t c = need + need / 8 + 4096;
        int rc = dev_alloc(p, c);
        if (rc) return rc;
        cap = c;
        return DFL_OK;
    }
};

struct InputArrival {
    std::vector<cudaEvent_t>* ev;
    std::vector<size_t> lo;    // slice boundaries, n_slices + 1 of them
    size_t n_slices;
    const uint8_t* h_src;      // host source; slice k is copied by feed(k) right before it is waited for, so
    uint8_t* d_dst;            // that with pageable memory (a blocking, staged copy) the kernels of slice k
    size_t n;                  // run while slice k + 1 is being staged
    cudaStream_t copy_stream;
    Context* ctx;
    int feed(size_t k) const {
        int rc = ctx->copy_in(d_dst + lo[k], h_src + lo[k], lo[k + 1] - lo[k], copy_stream, false);
        if (rc) return rc;
        CK(cudaEventRecord((*ev)[k], copy_stream));
        return DFL_OK;
    }
    size_t slices_covering(size_t upto) const {   // how many leading slices hold [0, upto)
        size_t m = 0;
        while (m < n_slices && lo[m] < upto) m++;
        return m;
    }
};

// Scratch that outlives a call, per calling thread and device: the context of the one-shot calls and the two pools
// of 16 the batch calls rotate over.  dfl_trim() gives all of it back.
std::map<int, std::unique_ptr<Context>>& tls_contexts() {
    thread_local std::map<int, std::unique_ptr<Context>> ctxs;
    return ctxs;
}
std::map<int, std::vector<std::unique_ptr<Context>>>& tls_batch_pools(int which) {
    thread_local std::map<int, std::vector<std::unique_ptr<Context>>> pools[2];
    return pools[which];
}
Context& tls_context() {   // one per calling thread and device
    std::unique_ptr<Context>& ctx = tls_contexts()[current_device()];
    if (!ctx) ctx.reset(new Context());
    return *ctx;
}

struct StageTimer {
    cudaStream_t st;
    bool on;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    StageTimer(cudaStream_t s, bool enabled) : st(s), on(enabled) {
        if (on) mark("start");
    }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        names.push_back(name);
    }
    void finish() {
        if (!on) return;
        t_stage_ms.clear();
        for (size_t i = 1; i < ev.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            t_stage_ms.push_back({names[i], ms});
        }
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
    }
};

uint32_t wrap_header_bytes(int wrap, size_t gz_hdr_len) {
    if (wrap == DFL_ZLIB) return 2;
    if (wrap == DFL_GZIP) return gz_hdr_len ? (uint32_t)gz_hdr_len : 10;
    return 0;
}

// Runs the kernel pipeline for one piece of a stream.  d_in holds `n` bytes of which the first
// `begin` are dictionary only.  On success *out_bytes is the number of bytes produced in d_out
// (container header included when hdr_bytes > 0, trailer included when final && wrap == zlib).
// RFC 1952 member header the reference gets from gzip-header 1.0 `GzBuilder::new().into_header()`
// (lib.rs:251, writer.rs:341-357): no flags, MTIME 0, XFL 0, OS 255 (unknown).
const uint8_t kGzipDefaultHeader[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};

// Continuation of a stream in pieces without flushes (streaming handle): what the previous piece handed over.
struct PieceIn {
    uint32_t init_key = 0;                 // parser state at `begin`
    uint32_t parse_end = 0;                // 0 = to the end of the data
    int open_piece = 0;                    // more input follows without a flush: complete blocks only
    const uint32_t* h_carry_tok = nullptr; // tokens not coded yet (host memory)
    uint32_t n_carry_tok = 0;
    uint32_t carry_in_pos = 0;             // offset in d_in of their first input byte
    uint32_t bits_n = 0, bits_v = 0;       // the incomplete last byte of the output so far
};
thread_local const PieceIn* t_piece_in = nullptr;   // set by the streaming handle around its pipeline call

// issue_pipeline queues every stage and the read-back of the bookkeeping on `st` without waiting;
// finish_pipeline waits for it and turns the bookkeeping into the call's result.  dfl_compress_device_batch
// keeps several of these in flight on different streams; everything else runs them back to back.
int issue_pipeline(Context& c, cudaStream_t st, StageTimer& tm, const uint8_t* d_in, size_t n, size_t begin,
                   const dfl_options* opt, int wrap, uint32_t hdr_bytes, int final_block, int sync_marker, uint8_t* d_out,
                   size_t out_cap, const uint32_t* d_tokens_override = nullptr, uint64_t n_tokens_override = 0,
                   int stop_after_tokens = 0, const uint8_t* gz_hdr = nullptr, const InputArrival* arrival = nullptr) {
    if (n >= 0xfffffff0ull) return DFL_E_UNSUPPORTED;   // 32-bit positions; see DESIGN.md "limits"
    if (opt->special != 0) return DFL_E_UNSUPPORTED;    // compression_options.rs:52-59: placeholders
    EncodeJob j;
    j.d_in = d_in;
    j.n = (uint32_t)n;
    j.begin = (uint32_t)begin;
    j.prm = make_params(opt->max_hash_checks, opt->lazy_if_less_than, opt->matching_type);
    j.final_block = final_block;
    j.sync_marker = sync_marker;
    j.d_out = d_out;
    j.out_cap = out_cap & ~(size_t)15;
    j.hdr_bytes = hdr_bytes;
    j.isize = (uint32_t)((n - begin) & 0xffffffffull);
    j.d_tokens_override = d_tokens_override;
    j.n_tokens_override = n_tokens_override;
    j.stop_after_tokens = stop_after_tokens;
    j.peers = t_peers;
    const PieceIn none;
    const PieceIn& pin = t_piece_in ? *t_piece_in : none;
    j.init_key = pin.init_key;
    j.parse_end = pin.parse_end ? pin.parse_end : (uint32_t)n;
    j.open_piece = pin.open_piece;
    j.n_carry_tok = pin.n_carry_tok;
    j.carry_in_pos = pin.carry_in_pos;
    j.carry_bits_n = pin.bits_n;
    j.carry_bits_v = pin.bits_v;
    if ((reinterpret_cast<uintptr_t>(d_out) & 15u) != 0) return DFL_E_ARG;

    int rc = c.ensure(n, j.prm.need_quarter != 0);
    if (rc) return rc;
    Buffers& b = c.buf;
    CK(cudaMemsetAsync(b.meta, 0, sizeof(DevMeta), st));
    const bool seq_lz = (d_tokens_override == nullptr) && use_seq_lz77(j.prm, j.begin, j.open_piece, j.init_key, j.n_carry_tok) &&
                        j.parse_end == j.n;
    if (seq_lz) {
        if (!b.seq_tab) { int arc = dev_alloc(b.seq_tab, 3 * kWindow); if (arc) return arc; }
        if (arrival)
            for (size_t k = 0; k < arrival->n_slices; k++) { { int frc = arrival->feed(k); if (frc) return frc; } CK(cudaStreamWaitEvent(st, (*arrival->ev)[k], 0)); }
        CK(launch_lz77_seq(j, b, st));
        tm.mark("lz77_sequential");
    }
    const bool need_lz = (d_tokens_override == nullptr) && !seq_lz;
    const bool need_match = need_lz && j.prm.mode != kRle && j.prm.checks > 0;
    if (need_match) {
        const uint32_t w_end = n_windows(j), w_sort0 = first_sort_window(j), w_match0 = first_match_window(j);
        if (arrival) {
            // the input is still being copied: sort and match every window as soon as its bytes
            // (and the 272 bytes of look-ahead behind it) are on the device
            uint32_t w_sorted = w_sort0, w_matched = w_match0;
            for (size_t k = 0; k < arrival->n_slices; k++) {
                { int frc = arrival->feed(k); if (frc) return frc; }
                CK(cudaStreamWaitEvent(st, (*arrival->ev)[k], 0));
                const size_t have = (k + 1 == arrival->n_slices) ? n : arrival->lo[k + 1];
                uint32_t w_ok = (have >= n) ? w_end : (uint32_t)((have - 272) / kWindow);
                if (w_ok > w_end) w_ok = w_end;
                CK(launch_window_sort(j, b, st, w_sorted, w_ok));
                if (w_ok > w_sorted) w_sorted = w_ok;
                CK(launch_match(j, b, st, w_matched, w_sorted));
                if (w_sorted > w_matched) w_matched = w_sorted;
            }
            tm.mark("window_sort+match");
        } else {
            if (arrival)
                for (size_t k = 0; k < arrival->n_slices; k++) { { int frc = arrival->feed(k); if (frc) return frc; } CK(cudaStreamWaitEvent(st, (*arrival->ev)[k], 0)); }
            CK(launch_window_sort(j, b, st, w_sort0, w_end));
            tm.mark("window_sort");
            CK(launch_match(j, b, st, w_match0, w_end));
            tm.mark("match");
        }
    } else if (arrival && !seq_lz) {
        for (size_t k = 0; k < arrival->n_slices; k++) { { int frc = arrival->feed(k); if (frc) return frc; } CK(cudaStreamWaitEvent(st, (*arrival->ev)[k], 0)); }
    }
    if (wrap == DFL_ZLIB && final_block && !stop_after_tokens) {
        CK(launch_adler32(d_in + begin, n - begin, b, st));
        tm.mark("adler32");
    }
    if (wrap == DFL_GZIP && final_block && !stop_after_tokens) {
        CK(launch_crc32(d_in + begin, n - begin, b, st));
        tm.mark("crc32");
    }
    if (need_lz) {
        CK(launch_parse(j, b, st));
        tm.mark("parse");
        if (j.n_carry_tok)   // the sorted lists the token array shares its memory with are dead only now (the parser resolves long matches on them)
            CK(cudaMemcpyAsync(b.tok, pin.h_carry_tok, (size_t)j.n_carry_tok * 4, cudaMemcpyHostToDevice, st));
        CK(launch_token_layout(j, b, st));
        tm.mark("token_layout");
    }
    if (!stop_after_tokens) {
        CK(launch_block_stats(j, b, st));
        tm.mark("block_stats");
        CK(launch_block_codes(j, b, st));
        tm.mark("block_codes");
        CK(launch_block_scan(j, b, st));
        tm.mark("block_scan");
        CK(launch_pack(j, b, st));
        tm.mark("pack");
        CK(launch_finalize(j, b, wrap, st));
        if (wrap == DFL_GZIP && hdr_bytes > 0 && hdr_bytes + 16 <= j.out_cap)   // the header bytes are the caller's (or the default)
            CK(cudaMemcpyAsync(d_out, gz_hdr ? gz_hdr : kGzipDefaultHeader, hdr_bytes, cudaMemcpyHostToDevice, st));
        tm.mark("finalize");
    }
    CK(cudaMemcpyAsync(c.h_meta, b.meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, st));
    return DFL_OK;
}

int finish_pipeline(Context& c, cudaStream_t st, size_t n, size_t begin, size_t* out_bytes) {
    CK(cudaStreamSynchronize(st));
    const DevMeta& m = *c.h_meta;
    t_counters[0] = m.n_tokens;
    t_counters[1] = m.n_blocks;
    t_counters[2] = parse_n_seg(n - begin, parse_geom(n - begin, kLazy));   // (greedy runs on large inputs use half as many)
    t_counters[3] = m.n_repaired_par;
    t_counters[4] = m.n_repaired_seq;
    t_counters[5] = (uint64_t)g_launch_count;
    t_counters[6] = m.n_stored;
    t_counters[7] = m.n_fixed;
    if (out_bytes) *out_bytes = (size_t)m.out_bytes;
    if (m.err == 100) return DFL_E_OVERFLOW;
    if (m.err != 0) {
        t_cuda_err = "device invariant violated, code " + std::to_string(m.err);
        return DFL_E_INTERNAL;
    }
    return DFL_OK;
}

int run_pipeline(Context& c, cudaStream_t st, const uint8_t* d_in, size_t n, size_t begin, const dfl_options* opt, int wrap,
                 uint32_t hdr_bytes, int final_block, int sync_marker, uint8_t* d_out, size_t out_cap, size_t* out_bytes,
                 const uint32_t* d_tokens_override = nullptr, uint64_t n_tokens_override = 0, int stop_after_tokens = 0,
                 const uint8_t* gz_hdr = nullptr, const InputArrival* arrival = nullptr) {
    g_launch_count = 0;
    StageTimer tm(st, t_profiling != 0);
    int rc = issue_pipeline(c, st, tm, d_in, n, begin, opt, wrap, hdr_bytes, final_block, sync_marker, d_out, out_cap,
                            d_tokens_override, n_tokens_override, stop_after_tokens, gz_hdr, arrival);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    rc = finish_pipeline(c, st, n, begin, out_bytes);
    tm.finish();
    return rc;
}

bool valid_wrap(int wrap) { return wrap == DFL_RAW || wrap == DFL_ZLIB || wrap == DFL_GZIP; }

}  // namespace

// ============================================================================ misc
extern "C" const char* dfl_strerror(int status) {
    switch (status) {
        case DFL_OK: return "ok";
        case DFL_AGAIN: return "internal buffer full, call again";
        case DFL_E_ARG: return "invalid argument";
        case DFL_E_NOMEM: return "out of (device) memory";
        case DFL_E_CUDA: return "CUDA error";
        case DFL_E_NODEVICE: return "no CUDA device (this library has no CPU fallback)";
        case DFL_E_OVERFLOW: return "output buffer too small";
        case DFL_E_STATE: return "encoder is in the wrong state for this call";
        case DFL_E_UNSUPPORTED: return "not supported";
        case DFL_E_INTERNAL: return "internal error";
        case DFL_E_NCCL: return "NCCL unavailable or a collective failed";
        default: return "unknown status";
    }
}
extern "C" const char* dfl_last_cuda_error(void) { return t_cuda_err.c_str(); }
extern "C" int dfl_version(void) { return DFL_VERSION; }
extern "C" int dfl_device_count(void) { return device_count_cached(); }

extern "C" int dfl_options_preset(int preset, dfl_options* out) {
    if (!out) return DFL_E_ARG;
    switch (preset) {   // compression_options.rs:14-20,126-178
        case DFL_PRESET_FAST: *out = {1, 0, 0, 0}; return DFL_OK;
        case DFL_PRESET_DEFAULT: *out = {128, 32, 1, 0}; return DFL_OK;
        case DFL_PRESET_BEST: *out = {1768, 128, 1, 0}; return DFL_OK;
        case DFL_PRESET_HUFFMAN_ONLY: *out = {0, 0, 0, 0}; return DFL_OK;
        case DFL_PRESET_RLE: *out = {0, 0, 1, 0}; return DFL_OK;
        default: return DFL_E_ARG;
    }
}

extern "C" size_t dfl_bound(size_t n, int wrap) {
    // Every block costs at most its stored form (+1 bit of slack), 5 bytes per 32767-byte chunk,
    // at most n/31744 + 1 blocks, plus the largest container (gzip: 10 + 8) and write slack.
    return n + 5 * (n / 32767 + 1) + 6 * (n / 31744 + 2) + 64 + (wrap == DFL_GZIP ? 320 : 0);
}

extern "C" int dfl_set_profiling(int enabled) {
    int old = t_profiling;
    t_profiling = enabled;
    return old;
}
extern "C" int dfl_last_stage_times(const char** names, float* ms, int cap) {
    int n = (int)t_stage_ms.size();
    for (int i = 0; i < n && i < cap; i++) {
        if (names) names[i] = t_stage_ms[i].first;
        if (ms) ms[i] = t_stage_ms[i].second;
    }
    return n;
}
extern "C" int dfl_last_counters(uint64_t* out, int cap) {
    for (int i = 0; i < 8 && i < cap; i++) out[i] = t_counters[i];
    return 8;
}

// ============================================================================ one-shot
// Inputs too long for one pipeline run (32-bit positions) take the streaming handle's route: the same bytes, in
// bounded pieces.  DFL_ONESHOT_PIECE_LIMIT lowers the switch-over point (test hook).
static size_t oneshot_piece_limit() {
    static const size_t v = [] {
        const char* e = getenv("DFL_ONESHOT_PIECE_LIMIT");
        size_t x = e ? (size_t)strtoull(e, nullptr, 10) : 0;
        return x ? x : ((size_t)3 << 30);
    }();
    return v;
}


// One stream encoded as open pieces straight from a device buffer (the streaming handle's scheme without the
// host copy of the input): every piece sees the 32 KiB in front of it and the input of the tokens it inherits,
// and writes into one of two scratch outputs; its complete bytes then go to the caller's buffer -- device to
// device, or to host memory on a separate stream while the next piece is already running.  Used for device
// buffers too long for one pipeline run and for large host calls (where it overlaps both copy directions with
// the kernels: the input keeps arriving in slices on the copy stream).
struct PieceIO {
    size_t piece = (size_t)1 << 30;         // bytes parsed per piece
    const InputArrival* arrival = nullptr;  // the device buffer is still being filled from the host
    uint8_t* d_out = nullptr;               // device sink ...
    uint8_t* h_out = nullptr;               // ... or host sink
    size_t out_cap = 0;
};

static int compress_pieces(Context& c, cudaStream_t st, const uint8_t* d_in, size_t n, const dfl_options* opt, int wrap,
                           const uint8_t* gz_hdr, size_t gz_hdr_len, const PieceIO& io, size_t* out_len) {
    size_t parse_pos = 0, carry_in = 0, out_off = 0, sum_upto = 0, fed = 0;
    uint32_t parse_key = 0, bits_n = 0, bits_v = 0, adler = 1, crc = 0;
    std::vector<uint32_t> carry_tok;
    bool header_written = false;
    bool d2h_pending[2] = {false, false};
    int rc = DFL_OK;
    const int saved_prof = t_profiling;
    auto feed_to = [&](size_t upto) -> int {   // issue the host-to-device slices covering [0, upto)
        if (!io.arrival) return DFL_OK;
        const size_t want = io.arrival->slices_covering(upto);
        for (; fed < want && fed < io.arrival->n_slices; fed++) { int frc = io.arrival->feed(fed); if (frc) return frc; }
        return DFL_OK;
    };
    for (uint32_t k = 0;; k++) {
        const size_t piece_end = (n - parse_pos) > io.piece ? parse_pos + io.piece : n;
        const bool last = piece_end == n;
        size_t base = parse_pos > kWindow ? parse_pos - kWindow : 0;
        if (!carry_tok.empty() && carry_in < base) base = carry_in;
        base &= ~(size_t)15;
        const size_t n_piece = piece_end - base;
        const uint32_t hdr = header_written ? 0u : wrap_header_bytes(wrap, gz_hdr_len);
        const size_t bound = dfl_bound(n_piece, wrap) + gz_hdr_len + 64;
        uint8_t*& scratch = (k & 1u) ? c.d_out2 : c.d_out;
        size_t& scratch_cap = (k & 1u) ? c.d_out2_cap : c.d_out_cap;
        if (d2h_pending[k & 1u]) {              // the bytes of piece k - 2 must have left this scratch buffer
            CK(cudaEventSynchronize(c.d2h_ev[k & 1u]));
            d2h_pending[k & 1u] = false;
        }
        if ((rc = c.ensure_stage(scratch, scratch_cap, bound))) return rc;
        if ((rc = feed_to(piece_end))) return rc;
        if (io.arrival)
            for (size_t q = 0; q < fed; q++) CK(cudaStreamWaitEvent(st, (*io.arrival->ev)[q], 0));
        PieceIn pin;
        pin.init_key = parse_key;
        pin.open_piece = last ? 0 : 1;
        pin.parse_end = last ? 0u : (uint32_t)(n_piece - (kMaxMatch + 1));
        pin.h_carry_tok = carry_tok.data();
        pin.n_carry_tok = (uint32_t)carry_tok.size();
        pin.carry_in_pos = (uint32_t)(carry_in - base);
        pin.bits_n = bits_n;
        pin.bits_v = bits_v;
        t_piece_in = &pin;
        t_profiling = 0;
        g_launch_count = 0;
        StageTimer tm(st, false);
        rc = issue_pipeline(c, st, tm, d_in + base, n_piece, parse_pos - base, opt, hdr ? wrap : DFL_RAW, hdr, last ? 1 : 0, 0,
                            scratch, scratch_cap, nullptr, 0, 0, gz_hdr);
        t_piece_in = nullptr;
        t_profiling = saved_prof;
        if (rc) { cudaStreamSynchronize(st); return rc; }
        // while this piece runs, the next one's input is put on its way (a blocking, staged copy for pageable memory)
        if (!last && (rc = feed_to(piece_end + io.piece < n ? piece_end + io.piece : n))) return rc;
        size_t produced = 0;
        if ((rc = finish_pipeline(c, st, n_piece, parse_pos - base, &produced))) return rc;
        const DevMeta m = *c.h_meta;
        // the container checksum is folded piece by piece (the scratch is sized for one piece) and written below
        if (wrap != DFL_RAW && piece_end > sum_upto) {
            const size_t len = piece_end - sum_upto;
            if (wrap == DFL_ZLIB) CK(launch_adler32(d_in + sum_upto, len, c.buf, st));
            else CK(launch_crc32(d_in + sum_upto, len, c.buf, st));
            CK(cudaMemcpyAsync(c.h_meta, c.buf.meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (wrap == DFL_ZLIB) adler = adler32_combine(adler, c.h_meta->adler, len);
            else crc = crc32_combine(crc, c.h_meta->crc, len);
            sum_upto = piece_end;
        }
        const size_t full = last ? (size_t)m.stream_bytes : (size_t)(m.stream_bits >> 3);
        const uint32_t left_bits = last ? 0u : (uint32_t)(m.stream_bits & 7ull);
        if (out_off + hdr + full + 16 > io.out_cap) { *out_len = out_off + hdr + full + 16; return DFL_E_OVERFLOW; }
        if (hdr + full) {
            if (io.h_out) {   // the pipeline has finished (its bookkeeping was read): copy out beside the next piece
                CK(cudaMemcpyAsync(io.h_out + out_off, scratch, hdr + full, cudaMemcpyDeviceToHost, c.d2h_stream));
                CK(cudaEventRecord(c.d2h_ev[k & 1u], c.d2h_stream));
                d2h_pending[k & 1u] = true;
            } else {
                CK(cudaMemcpyAsync(io.d_out + out_off, scratch, hdr + full, cudaMemcpyDeviceToDevice, st));
            }
        }
        uint8_t last_byte = 0;
        if (left_bits) CK(cudaMemcpyAsync(&last_byte, scratch + hdr + full, 1, cudaMemcpyDeviceToHost, st));
        const size_t coded = (size_t)m.n_blocks * kBlockTokens;
        const size_t rem = !last && m.n_tokens > coded ? (size_t)(m.n_tokens - coded) : 0;
        std::vector<uint32_t> next_carry(rem);
        if (rem) CK(cudaMemcpyAsync(next_carry.data(), c.buf.tok + coded, rem * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        out_off += hdr + full;
        header_written = true;
        bits_n = left_bits;
        bits_v = left_bits ? last_byte : 0u;
        carry_tok.swap(next_carry);
        if (last) break;
        carry_in = base + (size_t)m.in_coded_end;
        parse_pos = base + m.end_pos;
        parse_key = m.end_key;
    }
    if (io.h_out) CK(cudaStreamSynchronize(c.d2h_stream));
    uint8_t tr[8];
    size_t tr_len = 0;
    if (wrap == DFL_ZLIB) {            // lib.rs:192-196
        for (int k = 0; k < 4; k++) tr[k] = (uint8_t)(adler >> (24 - 8 * k));
        tr_len = 4;
    } else if (wrap == DFL_GZIP) {     // lib.rs:260-265
        const uint32_t isize = (uint32_t)(n & 0xffffffffull);
        for (int k = 0; k < 4; k++) { tr[k] = (uint8_t)(crc >> (8 * k)); tr[4 + k] = (uint8_t)(isize >> (8 * k)); }
        tr_len = 8;
    }
    if (tr_len) {
        if (out_off + tr_len > io.out_cap) { *out_len = out_off + tr_len; return DFL_E_OVERFLOW; }
        if (io.h_out) memcpy(io.h_out + out_off, tr, tr_len);
        else {
            CK(cudaMemcpyAsync(io.d_out + out_off, tr, tr_len, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
        }
        out_off += tr_len;
    }
    *out_len = out_off;
    return DFL_OK;
}

extern "C" int dfl_compress_device(const void* d_in, size_t n, const dfl_options* opt, int wrap, const uint8_t* gz_hdr,
                                   size_t gz_hdr_len, void* d_out, size_t out_cap, size_t* out_len, void* stream) {
    if (!opt || !d_out || !out_len || (!d_in && n) || !valid_wrap(wrap)) return DFL_E_ARG;
    if (wrap != DFL_GZIP || !gz_hdr || gz_hdr_len == 0) { gz_hdr = nullptr; gz_hdr_len = 0; }
    if (gz_hdr_len > 0xffffu) return DFL_E_ARG;
    Context& c = tls_context();
    int rc = c.init();
    if (rc) return rc;
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : c.stream;
    if (n >= oneshot_piece_limit()) {
        if (opt->special != 0 || (reinterpret_cast<uintptr_t>(d_in) & 15u) != 0) return opt->special ? DFL_E_UNSUPPORTED : DFL_E_ARG;
        PieceIO io;
        io.piece = oneshot_piece_limit() < ((size_t)1 << 30) ? oneshot_piece_limit() : ((size_t)1 << 30);
        io.d_out = reinterpret_cast<uint8_t*>(d_out);
        io.out_cap = out_cap;
        return compress_pieces(c, st, reinterpret_cast<const uint8_t*>(d_in), n, opt, wrap, gz_hdr, gz_hdr_len, io, out_len);
    }
    return run_pipeline(c, st, reinterpret_cast<const uint8_t*>(d_in), n, 0, opt, wrap, wrap_header_bytes(wrap, gz_hdr_len), 1,
                        0, reinterpret_cast<uint8_t*>(d_out), out_cap, out_len, nullptr, 0, 0, gz_hdr);
}

extern "C" int dfl_compress(const uint8_t* in, size_t n, const dfl_options* opt, int wrap, const uint8_t* gz_hdr,
                            size_t gz_hdr_len, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!opt || !out || !out_len || (!in && n) || !valid_wrap(wrap)) return DFL_E_ARG;
    if (wrap != DFL_GZIP || !gz_hdr || gz_hdr_len == 0) { gz_hdr = nullptr; gz_hdr_len = 0; }
    if (gz_hdr_len > 0xffffu) return DFL_E_ARG;
    if (n >= oneshot_piece_limit()) {
        dfl_encoder* e = dfl_encoder_new(opt, wrap, gz_hdr, gz_hdr_len);
        if (!e) return DFL_E_NOMEM;
        const size_t lim = oneshot_piece_limit();
        if (lim < ((size_t)1 << 28)) dfl_encoder_set_piece_bytes(e, lim < 4096 ? 4096 : lim);
        int rc = dfl_encoder_write(e, in, n, nullptr);
        if (rc == DFL_OK) rc = dfl_encoder_flush(e, DFL_FLUSH_FINISH);
        if (rc == DFL_OK) {
            const uint8_t* p = nullptr;
            size_t len = 0;
            dfl_encoder_take_output(e, &p, &len);
            *out_len = len;
            if (len > out_cap) rc = DFL_E_OVERFLOW;
            else memcpy(out, p, len);
        }
        dfl_encoder_free(e);
        return rc;
    }
    Context& c = tls_context();
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.ensure_stage(c.d_in, c.d_in_cap, n + 64))) return rc;
    size_t bound = dfl_bound(n, wrap) + gz_hdr_len;
    if ((rc = c.ensure_stage(c.d_out, c.d_out_cap, bound + 64))) return rc;
    // Optional (DFL_HOST_PIECE_MIB): run large calls as open pieces, so that the copy of a piece's output overlaps
    // the next piece's kernels as well.  Measured on B200 it does not pay -- 1 GiB at Default 6162 vs 6161 MiB/s,
    // at Fast 20998 vs 24088 MiB/s with 256 MiB pieces, worse with smaller ones (the per-piece synchronisation
    // costs what the overlap saves) -- so it is off unless asked for.
    static const size_t host_piece = [] {
        const char* e = getenv("DFL_HOST_PIECE_MIB");
        size_t x = e ? (size_t)strtoull(e, nullptr, 10) : 0;
        return x << 20;
    }();
    // host -> device in slices on a second stream; the pipeline's first two stages start on a
    // slice as soon as it has landed (writer.rs callers pay PCIe: SURVEY 8(f) rank 1)
    InputArrival arrival{&c.copy_ev, {}, 0, in, c.d_in, n, c.copy_stream, &c};
    arrival.lo = plan_slices(n, opt->max_hash_checks >= 16 && !(host_piece && n >= 2 * host_piece));
    arrival.n_slices = arrival.lo.size() - 1;
    while (c.copy_ev.size() < arrival.n_slices) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c.copy_ev.push_back(e);
    }
    if (host_piece && n >= 2 * host_piece && opt->special == 0) {
        PieceIO io;
        io.piece = host_piece;
        io.arrival = &arrival;
        io.h_out = out;
        io.out_cap = out_cap;
        return compress_pieces(c, c.stream, c.d_in, n, opt, wrap, gz_hdr, gz_hdr_len, io, out_len);
    }
    size_t produced = 0;
    rc = run_pipeline(c, c.stream, c.d_in, n, 0, opt, wrap, wrap_header_bytes(wrap, gz_hdr_len), 1, 0, c.d_out, c.d_out_cap,
                      &produced, nullptr, 0, 0, gz_hdr, arrival.n_slices ? &arrival : nullptr);
    if (rc) return rc;
    *out_len = produced;
    if (produced > out_cap) return DFL_E_OVERFLOW;
    return c.copy_out(out, c.d_out, produced, c.stream);
}

extern "C" int dfl_compress_device_piece(const void* d_in, size_t n_total, size_t dict_len, const dfl_options* opt, int flush_mode,
                                         void* d_out, size_t out_cap, size_t* out_len, void* stream) {
    if (!opt || !d_out || !out_len || (!d_in && n_total) || dict_len > n_total) return DFL_E_ARG;
    if (flush_mode != DFL_FLUSH_SYNC && flush_mode != DFL_FLUSH_FINISH) return DFL_E_ARG;
    Context& c = tls_context();
    int rc = c.init();
    if (rc) return rc;
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : c.stream;
    // only the last 32 KiB in front of the piece can be referenced (matching.rs:102-106)
    const uint8_t* p = reinterpret_cast<const uint8_t*>(d_in);
    size_t skip = dict_len > kWindow ? dict_len - kWindow : 0;
    skip &= ~(size_t)15;   // keep the 16-byte alignment of the staging loads
    return run_pipeline(c, st, p + skip, n_total - skip, dict_len - skip, opt, DFL_RAW, 0, flush_mode == DFL_FLUSH_FINISH ? 1 : 0,
                        flush_mode == DFL_FLUSH_SYNC ? 1 : 0, reinterpret_cast<uint8_t*>(d_out), out_cap, out_len);
}

// ---- many independent streams (PNG IDAT-like chunks): a pool of contexts, each with
// its own stream and scratch, keeps kBatchLanes pipelines in flight so that small inputs fill the GPU.
extern "C" int dfl_compress_device_batch(size_t count, const void* const* d_in, const size_t* n, const dfl_options* opt,
                                         int wrap, void* const* d_out, const size_t* out_cap, size_t* out_len,
                                         int* status) {
    if (!opt || !valid_wrap(wrap) || (count && (!d_in || !n || !d_out || !out_cap || !out_len))) return DFL_E_ARG;
    constexpr size_t kBatchLanes = 16;
    std::vector<std::unique_ptr<Context>>& pool = tls_batch_pools(0)[current_device()];
    const size_t lanes = count < kBatchLanes ? count : kBatchLanes;
    while (pool.size() < lanes) pool.emplace_back(new Context());
    int first_err = DFL_OK;
    const int saved_prof = t_profiling;
    t_profiling = 0;
    t_peers = (uint32_t)lanes;
    g_launch_count = 0;
    for (size_t base = 0; base < count; base += lanes) {
        const size_t m = (count - base) < lanes ? (count - base) : lanes;
        std::vector<int> rc(m, DFL_OK);
        for (size_t k = 0; k < m; k++) {
            const size_t i = base + k;
            Context& c = *pool[k];
            rc[k] = (!d_out[i] || (!d_in[i] && n[i])) ? DFL_E_ARG : c.init();
            if (rc[k]) continue;
            StageTimer tm(c.stream, false);
            rc[k] = issue_pipeline(c, c.stream, tm, reinterpret_cast<const uint8_t*>(d_in[i]), n[i], 0, opt, wrap,
                                   wrap_header_bytes(wrap, 0), 1, 0, reinterpret_cast<uint8_t*>(d_out[i]), out_cap[i]);
        }
        for (size_t k = 0; k < m; k++) {
            const size_t i = base + k;
            Context& c = *pool[k];
            if (rc[k] == DFL_OK) rc[k] = finish_pipeline(c, c.stream, n[i], 0, &out_len[i]);
            else if (c.ok) cudaStreamSynchronize(c.stream);
            if (status) status[i] = rc[k];
            if (rc[k] != DFL_OK && first_err == DFL_OK) first_err = rc[k];
        }
    }
    t_profiling = saved_prof;
    t_peers = 1;
    return first_err;
}

// Host-buffer batch: the members rotate over a pool of pipelines; while the host waits for one member's size
// (needed for its device-to-host copy) the other lanes keep the GPU and both copy directions busy.
extern "C" int dfl_compress_batch(size_t count, const uint8_t* const* in, const size_t* n, const dfl_options* opt, int wrap,
                                  uint8_t* const* out, const size_t* out_cap, size_t* out_len, int* status) {
    if (!opt || !valid_wrap(wrap) || (count && (!in || !n || !out || !out_cap || !out_len))) return DFL_E_ARG;
    constexpr size_t kBatchLanes = 16;
    std::vector<std::unique_ptr<Context>>& pool = tls_batch_pools(1)[current_device()];
    const size_t lanes = count < kBatchLanes ? count : kBatchLanes;
    while (pool.size() < lanes) pool.emplace_back(new Context());
    std::vector<size_t> member(lanes, SIZE_MAX);   // the member in flight on each lane
    std::vector<int> lane_rc(lanes, DFL_OK);
    int first_err = DFL_OK;
    const int saved_prof = t_profiling;
    t_profiling = 0;
    t_peers = (uint32_t)lanes;
    g_launch_count = 0;
    auto settle = [&](size_t k) {   // member on lane k: wait for its size, put its bytes on their way to the host
        const size_t i = member[k];
        if (i == SIZE_MAX) return;
        Context& c = *pool[k];
        int rc = lane_rc[k];
        if (rc == DFL_OK) rc = finish_pipeline(c, c.stream, n[i], 0, &out_len[i]);
        else if (c.ok) cudaStreamSynchronize(c.stream);
        if (rc == DFL_OK && out_len[i] > out_cap[i]) rc = DFL_E_OVERFLOW;
        if (rc == DFL_OK && cudaMemcpyAsync(out[i], c.d_out, out_len[i], cudaMemcpyDeviceToHost, c.stream) != cudaSuccess)
            rc = DFL_E_CUDA;
        if (status) status[i] = rc;
        if (rc != DFL_OK && first_err == DFL_OK) first_err = rc;
        member[k] = SIZE_MAX;
    };
    for (size_t i = 0; i < count; i++) {
        const size_t k = i % lanes;
        settle(k);
        Context& c = *pool[k];
        int rc = (!out[i] || (!in[i] && n[i])) ? DFL_E_ARG : (n[i] >= ((size_t)1 << 32) - 64 ? DFL_E_UNSUPPORTED : c.init());
        if (rc == DFL_OK) rc = c.ensure_stage(c.d_in, c.d_in_cap, n[i] + 64);
        if (rc == DFL_OK) rc = c.ensure_stage(c.d_out, c.d_out_cap, dfl_bound(n[i], wrap) + 64);
        // same stream as the previous member's copy out of d_out: ordered behind it
        if (rc == DFL_OK && n[i] && cudaMemcpyAsync(c.d_in, in[i], n[i], cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
            rc = DFL_E_CUDA;
        if (rc == DFL_OK) {
            StageTimer tm(c.stream, false);
            rc = issue_pipeline(c, c.stream, tm, c.d_in, n[i], 0, opt, wrap, wrap_header_bytes(wrap, 0), 1, 0, c.d_out, c.d_out_cap);
        }
        member[k] = i;
        lane_rc[k] = rc;
    }
    for (size_t k = 0; k < lanes; k++) settle(k);
    for (size_t k = 0; k < lanes; k++)
        if (pool[k]->ok && cudaStreamSynchronize(pool[k]->stream) != cudaSuccess && first_err == DFL_OK) first_err = DFL_E_CUDA;
    t_profiling = saved_prof;
    t_peers = 1;
    return first_err;
}

extern "C" int dfl_crc32_device(const void* d_in, size_t n, uint32_t* crc, void* stream) {
    if (!crc || (!d_in && n)) return DFL_E_ARG;
    Context& c = tls_context();
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.ensure(n ? n : 1, false))) return rc;
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : c.stream;
    CK(launch_crc32(reinterpret_cast<const uint8_t*>(d_in), n, c.buf, st));
    CK(cudaMemcpyAsync(c.h_meta, c.buf.meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *crc = c.h_meta->crc;
    return DFL_OK;
}

extern "C" int dfl_adler32_device(const void* d_in, size_t n, uint32_t* adler, void* stream) {
    if (!adler || (!d_in && n)) return DFL_E_ARG;
    Context& c = tls_context();
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.ensure(n ? n : 1, false))) return rc;
    cudaStream_t st = stream ? reinterpret_cast<cudaStream_t>(stream) : c.stream;
    CK(launch_adler32(reinterpret_cast<const uint8_t*>(d_in), n, c.buf, st));
    CK(cudaMemcpyAsync(c.h_meta, c.buf.meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *adler = c.h_meta->adler;
    return DFL_OK;
}

extern "C" int dfl_encode_tokens(const uint8_t* in, size_t n, const uint32_t* tokens, size_t n_tokens, uint8_t* out,
                                 size_t out_cap, size_t* out_len) {
    if (!out || !out_len || (!in && n) || (!tokens && n_tokens) || n_tokens > n) return DFL_E_ARG;
    Context& c = tls_context();
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.ensure_stage(c.d_in, c.d_in_cap, n + 64))) return rc;
    if ((rc = c.ensure_stage(c.d_out, c.d_out_cap, dfl_bound(n, DFL_RAW) + 64))) return rc;
    if (c.d_tok_cap < n_tokens + 16) {
        dev_free(c.d_tok_in);
        c.d_tok_cap = 0;
        if ((rc = dev_alloc(c.d_tok_in, n_tokens + 4096))) return rc;
        c.d_tok_cap = n_tokens + 4096;
    }
    if (n) CK(cudaMemcpyAsync(c.d_in, in, n, cudaMemcpyHostToDevice, c.stream));
    if (n_tokens) CK(cudaMemcpyAsync(c.d_tok_in, tokens, n_tokens * 4, cudaMemcpyHostToDevice, c.stream));
    dfl_options opt = {0, 0, 0, 0};
    size_t produced = 0;
    rc = run_pipeline(c, c.stream, c.d_in, n, 0, &opt, DFL_RAW, 0, 1, 0, c.d_out, c.d_out_cap, &produced, c.d_tok_in, n_tokens);
    if (rc) return rc;
    *out_len = produced;
    if (produced > out_cap) return DFL_E_OVERFLOW;
    CK(cudaMemcpyAsync(out, c.d_out, produced, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    return DFL_OK;
}

extern "C" int dfl_lz77_tokens(const uint8_t* in, size_t n, const dfl_options* opt, uint32_t* tokens, size_t tokens_cap,
                               size_t* n_tokens) {
    if (!opt || !n_tokens || (!in && n) || (!tokens && tokens_cap)) return DFL_E_ARG;
    Context& c = tls_context();
    int rc = c.init();
    if (rc) return rc;
    if ((rc = c.ensure_stage(c.d_in, c.d_in_cap, n + 64))) return rc;
    if ((rc = c.ensure_stage(c.d_out, c.d_out_cap, 4096))) return rc;
    if (n) CK(cudaMemcpyAsync(c.d_in, in, n, cudaMemcpyHostToDevice, c.stream));
    size_t produced = 0;
    rc = run_pipeline(c, c.stream, c.d_in, n, 0, opt, DFL_RAW, 0, 1, 0, c.d_out, c.d_out_cap, &produced, nullptr, 0, 1);
    if (rc) return rc;
    size_t T = (size_t)c.h_meta->n_tokens;
    *n_tokens = T;
    if (T > tokens_cap) return DFL_E_OVERFLOW;
    if (T) CK(cudaMemcpyAsync(tokens, c.buf.tok, T * 4, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    return DFL_OK;
}

// ============================================================================ streaming
// write() sends input straight to the device; the bytes there are encoded
//   * when flush(Sync) / flush(Finish) is called -- a *closed* piece: everything is parsed and coded, and
//   * on their own once `piece_bytes` have accumulated -- an *open* piece: the reference's stream has no
//     seam there (lib.rs:408-433: the bytes do not depend on how write() chops the input), so the piece
//     stops parsing 259 bytes before the end of what it has (every decision before that point has seen its
//     full 258-byte look-ahead), codes only complete 31744-token blocks, and hands the parser state, the
//     uncoded tokens and the incomplete last output byte to the next piece.
// Either way the last 32 KiB stay as dictionary, so memory is bounded and a stream may be longer than 4 GiB.
//
// Two device buffers alternate.  An open piece is only *issued* by write(); while its kernels run, further
// writes land in the other buffer behind a reserved prefix, and when the piece is settled (at the next
// piece, flush, or checksum) its tail -- dictionary, unparsed bytes, the input of carried tokens -- is
// copied device-to-device into that prefix.  So the caller's copy-in overlaps the kernels, every input byte
// crosses the link once, and the host keeps no copy of the input.
namespace {

struct ByteBuf {   // growable host bytes without value-initialisation
    uint8_t* p = nullptr;
    size_t n = 0, cap = 0;
    ~ByteBuf() { free(p); }
    bool reserve(size_t need) {
        if (need <= cap) return true;
        size_t c = cap ? cap : 4096;
        while (c < need) c += c / 2 + 4096;
        uint8_t* q = static_cast<uint8_t*>(realloc(p, c));
        if (!q) return false;
        p = q;
        cap = c;
        return true;
    }
};

// history an open piece can hand over: window + look-ahead tail + the input of < 31744 carried tokens
constexpr size_t kStreamReserve = ((size_t)kBlockTokens * kMaxMatch + kWindow + 4 * (kMaxMatch + 1) + 4096 + 15) & ~(size_t)15;
constexpr size_t kStreamSlack = 256;          // bytes behind the data that kernels may touch
constexpr size_t kPendBytes = 1u << 20;       // small writes are gathered on the host first
constexpr size_t kPendDirect = 256u << 10;    // writes of at least this size go to the device directly

struct StreamBuf {
    uint8_t* d = nullptr;
    size_t cap = 0;
    size_t org = kStreamReserve;   // position of stream offset `off_org`: appended bytes start here
    uint64_t off_org = 0;
    size_t lo = kStreamReserve;    // first valid byte (installed history), 16-byte aligned
    size_t new_n = 0;              // bytes appended at org
    size_t end() const { return org + new_n; }
    size_t pos(uint64_t off) const { return (size_t)((int64_t)org + (int64_t)(off - off_org)); }
};

}  // namespace

namespace {
// Device resources of handles that were freed: streams, the pipeline's scratch, the two stream buffers.  A new
// handle on the same device takes them over instead of allocating (cudaMalloc/cudaFree of a few GiB cost more than
// encoding a piece; image encoders create one handle per picture).
struct HandleRes {
    std::unique_ptr<Context> ctx;
    uint8_t* d[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    int device = -1;
};
// DFL_HANDLE_POOL=n keeps up to n parked (0 = free everything with the handle); a parked handle that has run
// 256 MiB pieces holds about 5 GiB of device memory.
inline size_t handle_pool_max() {
    static const size_t v = [] {
        const char* e = getenv("DFL_HANDLE_POOL");
        return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)4;
    }();
    return v;
}
std::mutex g_handle_pool_mu;
std::vector<HandleRes> g_handle_pool;
}  // namespace

struct dfl_encoder {
    dfl_options opt;
    int wrap;
    std::unique_ptr<Context> ctx;   // streams and scratch, taken from / returned to a pool (encoder_init, ~dfl_encoder)
    int device = -1;
    StreamBuf sb[2];
    int f = 0;                   // sb[f] receives writes
    bool busy = false;           // an open piece is running on sb[1 - f]
    size_t busy_n = 0, busy_begin = 0;
    uint32_t busy_hdr = 0;
    uint64_t n_issued = 0;       // pieces issued so far; piece k writes into output scratch k & 1
    const uint8_t* bulk_src = nullptr;   // settled piece whose bytes are still in its scratch buffer
    size_t bulk_n = 0;
    cudaEvent_t in_ev = nullptr;     // input copies issued so far are on the device (copy stream)
    cudaEvent_t hist_ev = nullptr;   // the latest history hand-over has left its source buffer (compute stream)
    std::vector<uint8_t> pend;   // small writes not yet sent to the device (only while no device has been seen)
    uint8_t* gather = nullptr;   // small writes are gathered in a pinned slot of the context's stager ...
    size_t gather_n = 0;         // ... this many bytes so far
    size_t pending() const { return pend.size() + gather_n; }
    uint64_t parse_off = 0;      // stream offset of the first byte not parsed yet
    uint32_t parse_key = 0;      // parser state there (parse_state_key)
    std::vector<uint32_t> carry_tok;   // parsed but not yet coded (fewer than 31744)
    uint64_t carry_off = 0;      // stream offset of the first carried token's input byte
    uint32_t bits_n = 0, bits_v = 0;   // incomplete last byte of the stream so far
    ByteBuf out;                 // produced bytes not yet handed to the caller
    size_t out_pos = 0;
    bool header_written = false;
    bool finished = false;
    uint32_t adler = 1;          // Adler-32 of stream bytes [0, sum_off) (device-computed, folded on the host)
    uint32_t crc = 0;            // CRC-32 likewise (gzip)
    uint64_t sum_off = 0;
    uint64_t total_in = 0;       // bytes accepted by write()
    size_t piece_bytes = 256u << 20;   // appended bytes that trigger an open piece
    std::vector<uint8_t> gz_hdr; // gzip member header to emit (GzBuilder::into_header(), writer.rs:341-357)
    double t_append = 0, t_room = 0, t_wait = 0, t_small = 0, t_issue = 0, t_bulk = 0;   // host seconds (DFL_STREAM_TRACE)
    ~dfl_encoder() {
        if (getenv("DFL_STREAM_TRACE"))
            fprintf(stderr, "[dfl stream] pieces %llu: append %.1f ms (grow %.1f), wait for kernels %.1f, state %.1f, issue %.1f, bytes out %.1f\n",
                    (unsigned long long)n_issued, 1e3 * t_append, 1e3 * t_room, 1e3 * t_wait, 1e3 * t_small, 1e3 * t_issue, 1e3 * t_bulk);
        if (ctx && ctx->ok) {
            cudaStreamSynchronize(ctx->stream);
            cudaStreamSynchronize(ctx->copy_stream);
            cudaStreamSynchronize(ctx->d2h_stream);
            std::lock_guard<std::mutex> lk(g_handle_pool_mu);
            if (g_handle_pool.size() < handle_pool_max()) {
                HandleRes r;
                r.ctx = std::move(ctx);
                for (int i = 0; i < 2; i++) { r.d[i] = sb[i].d; r.cap[i] = sb[i].cap; sb[i].d = nullptr; }
                r.device = device;
                g_handle_pool.push_back(std::move(r));
            }
        }
        for (StreamBuf& b : sb) dev_free(b.d);
        if (in_ev) cudaEventDestroy(in_ev);
        if (hist_ev) cudaEventDestroy(hist_ev);
    }
};

namespace {

struct HostTimer {   // adds the lifetime of the object to *acc
    double* acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit HostTimer(double* a) : acc(a) {}
    ~HostTimer() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

constexpr size_t kOpenTail = kMaxMatch + 1;   // bytes an open piece leaves unparsed
enum { kPieceOpen = 0 };   // besides DFL_FLUSH_SYNC / DFL_FLUSH_FINISH

int encoder_init(dfl_encoder* e) {
    if (!e->ctx) {
        int dev = -1;
        if (device_count_cached() > 0) cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(g_handle_pool_mu);
            for (size_t i = 0; i < g_handle_pool.size(); i++) {
                if (g_handle_pool[i].device != dev) continue;
                HandleRes r = std::move(g_handle_pool[i]);
                g_handle_pool.erase(g_handle_pool.begin() + i);
                e->ctx = std::move(r.ctx);
                for (int k = 0; k < 2; k++) { e->sb[k].d = r.d[k]; e->sb[k].cap = r.cap[k]; }
                break;
            }
        }
        if (!e->ctx) e->ctx.reset(new Context());
        e->device = dev;
    }
    int rc = e->ctx->init();
    if (rc) return rc;
    if (!e->in_ev) CK(cudaEventCreateWithFlags(&e->in_ev, cudaEventDisableTiming));
    if (!e->hist_ev) CK(cudaEventCreateWithFlags(&e->hist_ev, cudaEventDisableTiming));
    return DFL_OK;
}

// Room for `more` further appended bytes in the fill buffer (grown geometrically up to one piece).
int encoder_room(dfl_encoder* e, size_t more) {
    StreamBuf& b = e->sb[e->f];
    const size_t need = b.end() + more + kStreamSlack;
    if (b.cap >= need) return DFL_OK;
    HostTimer ht(&e->t_room);
    size_t full = kStreamReserve + 16 + e->piece_bytes + kStreamSlack;
    size_t want = b.cap ? full : kStreamReserve + (1u << 20);   // small streams stay small; the second request gets it all
    if (want < need) want = need;
    Context& c = *e->ctx;
    CK(cudaStreamSynchronize(c.copy_stream));
    CK(cudaStreamSynchronize(c.stream));
    uint8_t* q = nullptr;
    int rc = dev_alloc(q, want);
    if (rc) return rc;
    if (b.d) {
        const size_t from = e->busy ? b.org : b.lo;   // the history arrives later while a piece is running
        if (b.end() > from) CK(cudaMemcpy(q + from, b.d + from, b.end() - from, cudaMemcpyDeviceToDevice));
        dev_free(b.d);
    }
    b.d = q;
    b.cap = want;
    return DFL_OK;
}

// Append host bytes to the fill buffer.  Returns once `src` may be reused.
int encoder_append(dfl_encoder* e, const uint8_t* src, size_t len) {
    if (!len) return DFL_OK;
    HostTimer ht(&e->t_append);
    int rc = encoder_init(e);
    if (rc) return rc;
    if ((rc = encoder_room(e, len))) return rc;
    StreamBuf& b = e->sb[e->f];
    Context& c = *e->ctx;
    // pageable sources are staged (by this library's copy threads or, if small, by the driver) before this returns;
    // a pinned one is read by the DMA itself, so wait for that
    if ((rc = c.copy_in(b.d + b.end(), src, len, c.copy_stream, true))) return rc;
    b.new_n += len;
    return DFL_OK;
}

int encoder_push_pending(dfl_encoder* e) {
    if (e->gather_n) {
        HostTimer ht(&e->t_append);
        int rc = encoder_room(e, e->gather_n);
        if (rc) return rc;
        StreamBuf& b = e->sb[e->f];
        Context& c = *e->ctx;
        if ((rc = c.stager->gather_commit(b.d + b.end(), e->gather_n, c.copy_stream))) return rc;
        b.new_n += e->gather_n;
        e->gather = nullptr;
        e->gather_n = 0;
    }
    if (e->pend.empty()) return DFL_OK;
    int rc = encoder_append(e, e->pend.data(), e->pend.size());
    e->pend.clear();
    return rc;
}

// Adler-32 / CRC-32 of stream bytes [sum_off, to) -- all inside buffer b -- folded into e->adler / e->crc.
// Runs on the compute stream and waits for it.
int encoder_fold_checksum(dfl_encoder* e, const StreamBuf& b, uint64_t to) {
    if (e->wrap == DFL_RAW || to <= e->sum_off) { if (to > e->sum_off) e->sum_off = to; return DFL_OK; }
    Context& c = *e->ctx;
    const size_t len = (size_t)(to - e->sum_off);
    int rc = c.ensure(len, false);
    if (rc) return rc;
    CK(cudaEventRecord(e->in_ev, c.copy_stream));
    CK(cudaStreamWaitEvent(c.stream, e->in_ev, 0));
    const uint8_t* src = b.d + b.pos(e->sum_off);
    if (e->wrap == DFL_ZLIB) CK(launch_adler32(src, len, c.buf, c.stream));
    else CK(launch_crc32(src, len, c.buf, c.stream));
    CK(cudaMemcpyAsync(c.h_meta, c.buf.meta, sizeof(DevMeta), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    // arithmetic on two device results
    if (e->wrap == DFL_ZLIB) e->adler = adler32_combine(e->adler, c.h_meta->adler, len);
    else e->crc = crc32_combine(e->crc, c.h_meta->crc, len);
    e->sum_off = to;
    return DFL_OK;
}

// Queue the pipeline for everything in the fill buffer; writes continue in the other buffer.
int encoder_issue(dfl_encoder* e, int mode) {
    HostTimer ht(&e->t_issue);
    Context& c = *e->ctx;
    StreamBuf& x = e->sb[e->f];
    const bool open = (mode == kPieceOpen);
    const size_t n = x.end() - x.lo;
    const size_t begin = x.pos(e->parse_off) - x.lo;
    const size_t coded_from = (!e->carry_tok.empty() && e->carry_off < e->parse_off) ? x.pos(e->carry_off) - x.lo : begin;
    int rc;
    if (!x.d && (rc = encoder_room(e, 0))) return rc;   // an empty stream still needs a valid pointer
    const size_t bound = dfl_bound(n - coded_from, e->wrap) + e->gz_hdr.size() + 64;
    // pieces alternate between two output buffers: the previous one's bytes leave theirs while this one runs
    uint8_t*& scratch = (e->n_issued & 1u) ? c.d_out2 : c.d_out;
    size_t& scratch_cap = (e->n_issued & 1u) ? c.d_out2_cap : c.d_out_cap;
    if ((rc = c.ensure_stage(scratch, scratch_cap, bound))) return rc;
    CK(cudaEventRecord(e->in_ev, c.copy_stream));
    CK(cudaStreamWaitEvent(c.stream, e->in_ev, 0));
    const uint32_t hdr = e->header_written ? 0u : wrap_header_bytes(e->wrap, e->gz_hdr.size());
    PieceIn pin;
    pin.init_key = e->parse_key;
    pin.parse_end = open ? (uint32_t)(n - kOpenTail) : 0u;
    pin.open_piece = open ? 1 : 0;
    pin.h_carry_tok = e->carry_tok.data();
    pin.n_carry_tok = (uint32_t)e->carry_tok.size();
    pin.carry_in_pos = (uint32_t)(e->carry_tok.empty() ? begin : x.pos(e->carry_off) - x.lo);
    pin.bits_n = e->bits_n;
    pin.bits_v = e->bits_v;
    // The trailer is appended on the host from the running checksum, so the kernels see wrap == RAW
    // unless the header still has to be written.
    t_piece_in = &pin;
    g_launch_count = 0;
    StageTimer tm(c.stream, false);
    rc = issue_pipeline(c, c.stream, tm, x.d + x.lo, n, begin, &e->opt, (hdr ? e->wrap : DFL_RAW), hdr,
                        mode == DFL_FLUSH_FINISH ? 1 : 0, mode == DFL_FLUSH_SYNC ? 1 : 0, scratch, scratch_cap, nullptr, 0, 0,
                        e->gz_hdr.empty() ? nullptr : e->gz_hdr.data());
    t_piece_in = nullptr;
    if (rc) { cudaStreamSynchronize(c.stream); return rc; }
    e->n_issued++;
    e->busy = true;
    e->busy_n = n;
    e->busy_begin = begin;
    e->busy_hdr = hdr;
    // the other buffer takes over; its origin is congruent to this one's end, which keeps the handed-over
    // history 16-byte aligned whatever its length (see encoder_settle)
    StreamBuf& y = e->sb[1 - e->f];
    y.org = kStreamReserve + (x.end() & 15u);
    y.off_org = x.off_org + x.new_n;
    y.lo = y.org;
    y.new_n = 0;
    e->f = 1 - e->f;
    return DFL_OK;
}

// Wait for the piece in flight, collect its bytes and state, hand its tail to the fill buffer.
int encoder_settle(dfl_encoder* e, int mode) {
    if (!e->busy) return DFL_OK;
    Context& c = *e->ctx;
    StreamBuf& x = e->sb[1 - e->f];
    StreamBuf& y = e->sb[e->f];
    const bool open = (mode == kPieceOpen);
    e->busy = false;
    size_t produced = 0;
    int rc;
    {
        HostTimer ht(&e->t_wait);
        rc = finish_pipeline(c, c.stream, e->busy_n, e->busy_begin, &produced);
    }
    if (rc) return rc;
    HostTimer ht(&e->t_small);
    const DevMeta m = *c.h_meta;
    const uint32_t hdr = e->busy_hdr;
    // bytes that are complete: all of them for a closed piece, all but the last partial one for an open piece
    const size_t full = open ? (size_t)(m.stream_bits >> 3) : (size_t)m.stream_bytes;
    const uint32_t left_bits = open ? (uint32_t)(m.stream_bits & 7ull) : 0u;
    const uint8_t* scratch = ((e->n_issued - 1) & 1u) ? c.d_out2 : c.d_out;
    uint8_t last_byte = 0;
    if (left_bits) CK(cudaMemcpyAsync(&last_byte, scratch + hdr + full, 1, cudaMemcpyDeviceToHost, c.stream));
    // tokens behind the last complete block wait for the next piece
    const size_t coded = (size_t)m.n_blocks * kBlockTokens;
    const size_t rem = open && m.n_tokens > coded ? (size_t)(m.n_tokens - coded) : 0;
    std::vector<uint32_t> next_carry(rem);
    if (rem) CK(cudaMemcpyAsync(next_carry.data(), c.buf.tok + coded, rem * 4, cudaMemcpyDeviceToHost, c.stream));
    const uint64_t x_lo_off = x.off_org - (uint64_t)(x.org - x.lo);   // stream offset of x.d[x.lo]
    const uint64_t x_end_off = x.off_org + x.new_n;
    if ((rc = encoder_fold_checksum(e, x, x_end_off))) return rc;   // ends with a wait for the stream
    CK(cudaStreamSynchronize(c.stream));
    e->bulk_src = scratch;       // fetched by encoder_fetch_bulk, for an open piece after the next one is under way
    e->bulk_n = hdr + full;
    e->bits_v = left_bits ? last_byte : 0u;
    e->bits_n = left_bits;
    e->carry_tok.swap(next_carry);
    e->header_written = true;
    if (open) {
        e->parse_off = x_lo_off + m.end_pos;
        e->parse_key = m.end_key;
        e->carry_off = x_lo_off + m.in_coded_end;
    } else {
        e->parse_off = x_end_off;
        e->parse_key = 0;
        e->carry_off = x_end_off;
    }
    // keep the dictionary of the next position to parse and the input of the carried tokens: they move in
    // front of the bytes written since.  y.org was chosen congruent to x.end(), the kept range starts on a
    // 16-byte boundary of x (x.lo is one), so it lands on one in y.
    uint64_t keep_off = e->parse_off - x_lo_off > kWindow ? e->parse_off - kWindow : x_lo_off;
    if (!e->carry_tok.empty() && e->carry_off < keep_off) keep_off = e->carry_off;
    const size_t keep_pos = x.pos(keep_off) & ~(size_t)15;
    const size_t hist = x.end() - keep_pos;
    if (hist > y.org || keep_pos < x.lo) {
        t_cuda_err = "stream history exceeds its reserve";
        return DFL_E_INTERNAL;
    }
    y.lo = y.org - hist;
    if (hist) {
        if (!y.d && (rc = encoder_room(e, 0))) return rc;
        CK(cudaMemcpyAsync(y.d + y.lo, x.d + keep_pos, hist, cudaMemcpyDeviceToDevice, c.stream));
        // x is the next buffer to be written to: not before this copy has read it
        CK(cudaEventRecord(e->hist_ev, c.stream));
        CK(cudaStreamWaitEvent(c.copy_stream, e->hist_ev, 0));
    }
    return DFL_OK;
}

// Bytes of the latest settled piece: device scratch -> e->out, on a stream of their own (the compute stream may
// already be running the next piece).  A pageable destination makes this a blocking copy.
int encoder_fetch_bulk(dfl_encoder* e) {
    if (!e->bulk_n) { e->bulk_src = nullptr; return DFL_OK; }
    HostTimer ht(&e->t_bulk);
    Context& c = *e->ctx;
    if (!e->out.reserve(e->out.n + e->bulk_n + 16)) return DFL_E_NOMEM;
    int rc = c.copy_out(e->out.p + e->out.n, e->bulk_src, e->bulk_n, c.d2h_stream);
    if (rc) return rc;
    e->out.n += e->bulk_n;
    e->bulk_n = 0;
    e->bulk_src = nullptr;
    return DFL_OK;
}

int encoder_emit(dfl_encoder* e, int mode) {
    int rc = encoder_init(e);
    if (rc) return rc;
    if ((rc = encoder_settle(e, kPieceOpen))) return rc;
    if ((rc = encoder_push_pending(e))) return rc;
    const StreamBuf& x = e->sb[e->f];
    const bool open = (mode == kPieceOpen);
    if (open && x.end() < x.pos(e->parse_off) + kOpenTail + 1) return encoder_fetch_bulk(e);   // not enough look-ahead to decide anything yet
    if ((rc = encoder_issue(e, mode))) return rc;
    if ((rc = encoder_fetch_bulk(e))) return rc;   // the previous piece's bytes, while this one runs
    if (open) return DFL_OK;   // settled by whoever comes next
    if ((rc = encoder_settle(e, mode))) return rc;
    if ((rc = encoder_fetch_bulk(e))) return rc;
    if (mode == DFL_FLUSH_FINISH) {
        if (!e->out.reserve(e->out.n + 8)) return DFL_E_NOMEM;
        if (e->wrap == DFL_ZLIB) {   // lib.rs:192-196 / writer.rs:235-245: Adler-32, big endian
            const uint32_t a = e->adler;
            const uint8_t t[4] = {(uint8_t)(a >> 24), (uint8_t)(a >> 16), (uint8_t)(a >> 8), (uint8_t)a};
            memcpy(e->out.p + e->out.n, t, 4);
            e->out.n += 4;
        } else if (e->wrap == DFL_GZIP) {   // writer.rs:408-426: CRC-32 and the input size, little endian
            const uint32_t v[2] = {e->crc, (uint32_t)(e->total_in & 0xffffffffull)};
            for (uint32_t x32 : v)
                for (int k = 0; k < 4; k++) e->out.p[e->out.n++] = (uint8_t)(x32 >> (8 * k));
        }
        e->finished = true;
    }
    return DFL_OK;
}

}  // namespace

extern "C" dfl_encoder* dfl_encoder_new(const dfl_options* opt, int wrap, const uint8_t* gz_hdr, size_t gz_hdr_len) {
    if (!opt || !valid_wrap(wrap) || gz_hdr_len > 0xffffu) return nullptr;
    dfl_encoder* e = new (std::nothrow) dfl_encoder();
    if (!e) return nullptr;
    e->opt = *opt;
    e->wrap = wrap;
    if (wrap == DFL_GZIP && gz_hdr && gz_hdr_len) e->gz_hdr.assign(gz_hdr, gz_hdr + gz_hdr_len);
    return e;
}

// A small write: copied once, into pinned memory that the DMA engine reads.
static int encoder_gather(dfl_encoder* e, const uint8_t* src, size_t len) {
    HostTimer ht(&e->t_append);
    int rc = encoder_init(e);
    if (rc) return rc;
    Context& c = *e->ctx;
    if (!c.stager) c.stager.reset(new Stager());
    while (len) {
        if (!e->gather && (rc = c.stager->gather_begin(&e->gather))) return rc;
        const size_t m = len < kStageSlot - e->gather_n ? len : kStageSlot - e->gather_n;
        memcpy(e->gather + e->gather_n, src, m);
        e->gather_n += m;
        src += m;
        len -= m;
        if (e->gather_n == kStageSlot && (rc = encoder_push_pending(e))) return rc;
    }
    return DFL_OK;
}

// Appended bytes that trigger the next open piece: the first pieces of a stream are short, so that the kernels
// start early (nothing runs while the first piece is being copied in), then piece_bytes.
static size_t piece_target(const dfl_encoder* e) {
    const uint64_t k = e->n_issued;
    const size_t ramp = k == 0 ? ((size_t)32 << 20) : (k == 1 ? ((size_t)128 << 20) : e->piece_bytes);
    return ramp < e->piece_bytes ? ramp : e->piece_bytes;
}

extern "C" int dfl_encoder_write(dfl_encoder* e, const uint8_t* buf, size_t n, size_t* consumed) {
    DeviceGuard on_device(e ? e->device : -1);   // the handle's streams and buffers live on the device it was created on
    if (!e || (!buf && n)) return DFL_E_ARG;
    if (e->finished) return DFL_E_STATE;
    int rc = DFL_OK;
    size_t done = 0;
    while (done < n) {
        // take at most one piece at a time, so that neither buffer nor a device call grows without bound
        const size_t have = e->sb[e->f].new_n + e->pending();
        const size_t target = piece_target(e);
        const size_t room = target > have ? target - have : 0;
        const size_t take = (n - done) < room ? (n - done) : room;
        if (take) {
            if (take < kPendDirect && device_count_cached() > 0) {
                rc = encoder_gather(e, buf + done, take);
            } else if (take < kPendDirect) {   // no device: keep the bytes, flush() will report it
                try {
                    e->pend.insert(e->pend.end(), buf + done, buf + done + take);
                } catch (const std::bad_alloc&) {
                    if (consumed) *consumed = done;
                    return done ? DFL_OK : DFL_E_NOMEM;
                }
                rc = e->pend.size() >= kPendBytes ? encoder_push_pending(e) : DFL_OK;
            } else {
                rc = encoder_push_pending(e);
                if (rc == DFL_OK) rc = encoder_append(e, buf + done, take);
            }
            if (rc) { if (consumed) *consumed = done; return rc; }
            e->total_in += take;
            done += take;
        }
        if (e->sb[e->f].new_n + e->pending() >= target) {
            rc = encoder_emit(e, kPieceOpen);
            if (rc) { if (consumed) *consumed = done; return rc; }
        }
    }
    if (consumed) *consumed = n;
    return DFL_OK;
}

extern "C" int dfl_encoder_flush(dfl_encoder* e, int mode) {
    DeviceGuard on_device(e ? e->device : -1);   // the handle's streams and buffers live on the device it was created on
    if (!e || (mode != DFL_FLUSH_SYNC && mode != DFL_FLUSH_FINISH)) return DFL_E_ARG;
    if (e->finished) return mode == DFL_FLUSH_FINISH ? DFL_OK : DFL_E_STATE;
    return encoder_emit(e, mode);
}

extern "C" int dfl_encoder_take_output(dfl_encoder* e, const uint8_t** p, size_t* len) {
    DeviceGuard on_device(e ? e->device : -1);   // the handle's streams and buffers live on the device it was created on
    if (!e || !p || !len) return DFL_E_ARG;
    // bytes of a piece still running become visible once it is settled; do that here if it costs no wait
    if (e->busy && cudaStreamQuery(e->ctx->stream) == cudaSuccess) {
        int rc = encoder_settle(e, kPieceOpen);
        if (rc == DFL_OK) rc = encoder_fetch_bulk(e);
        if (rc) return rc;
    }
    *p = e->out.p + e->out_pos;
    *len = e->out.n - e->out_pos;
    return DFL_OK;
}

extern "C" void dfl_encoder_advance_output(dfl_encoder* e, size_t n) {
    if (!e) return;
    e->out_pos += n;
    if (e->out_pos >= e->out.n) {
        e->out.n = 0;
        e->out_pos = 0;
    }
}

extern "C" int dfl_encoder_set_piece_bytes(dfl_encoder* e, size_t bytes) {
    if (!e || bytes < 4096 || bytes > 0x80000000ull) return DFL_E_ARG;
    e->piece_bytes = bytes;
    return DFL_OK;
}

extern "C" uint32_t dfl_encoder_checksum(dfl_encoder* e) {
    DeviceGuard on_device(e ? e->device : -1);   // the handle's streams and buffers live on the device it was created on
    if (!e || e->wrap == DFL_RAW) return 1;   // NoChecksum::current_hash (checksum.rs:26-28)
    if (encoder_init(e) != DFL_OK || encoder_settle(e, kPieceOpen) != DFL_OK || encoder_fetch_bulk(e) != DFL_OK ||
        encoder_push_pending(e) != DFL_OK)
        return 0;
    const StreamBuf& b = e->sb[e->f];
    if (encoder_fold_checksum(e, b, b.off_org + b.new_n) != DFL_OK) return 0;
    return e->wrap == DFL_ZLIB ? e->adler : e->crc;
}

extern "C" int dfl_encoder_reset(dfl_encoder* e, const uint8_t* gz_hdr, size_t gz_hdr_len) {
    DeviceGuard on_device(e ? e->device : -1);   // the handle's streams and buffers live on the device it was created on
    if (!e || gz_hdr_len > 0xffffu) return DFL_E_ARG;
    if (!e->finished) {
        int rc = encoder_emit(e, DFL_FLUSH_FINISH);   // output_all() (writer.rs:112-115,218-223)
        if (rc) return rc;
    }
    for (StreamBuf& b : e->sb) {
        b.org = b.lo = kStreamReserve;
        b.off_org = 0;
        b.new_n = 0;
    }
    e->busy = false;
    e->n_issued = 0;
    e->bulk_n = 0;
    e->bulk_src = nullptr;
    e->pend.clear();
    e->gather = nullptr;
    e->gather_n = 0;
    e->parse_off = 0;
    e->parse_key = 0;
    e->carry_tok.clear();
    e->carry_off = 0;
    e->bits_n = e->bits_v = 0;
    e->header_written = false;
    e->finished = false;
    e->adler = 1;
    e->crc = 0;
    e->sum_off = 0;
    e->total_in = 0;
    // reset() installs the default header, reset_with_builder() the caller's (writer.rs:394-406)
    e->gz_hdr.clear();
    if (e->wrap == DFL_GZIP && gz_hdr && gz_hdr_len) e->gz_hdr.assign(gz_hdr, gz_hdr + gz_hdr_len);
    return DFL_OK;
}

extern "C" void dfl_encoder_free(dfl_encoder* e) {
    if (!e) return;
    DeviceGuard on_device(e->device);
    delete e;
}

// ============================================================================ scratch
// Everything the library keeps between calls on behalf of the calling thread -- the one-shot context (29 B of
// device scratch per input byte of the largest call so far, staging buffers, streams), the batch pools -- and the
// process-wide pool of parked handle resources is released; the next call builds what it needs again (cudaMalloc
// of a GiB-sized context costs 0.1 - 2 s on the B200 hosts, which is why nothing is freed unasked).  Live
// dfl_encoder handles are not touched.
extern "C" int dfl_trim(void) {
    int saved = -1;
    const bool have_dev = device_count_cached() > 0 && cudaGetDevice(&saved) == cudaSuccess;
    auto on_device = [&](int dev) { if (have_dev && dev >= 0) cudaSetDevice(dev); };
    for (auto& kv : tls_contexts()) { on_device(kv.first); kv.second.reset(); }
    tls_contexts().clear();
    for (int which = 0; which < 2; which++) {
        for (auto& kv : tls_batch_pools(which)) { on_device(kv.first); kv.second.clear(); }
        tls_batch_pools(which).clear();
    }
    {
        std::lock_guard<std::mutex> lk(g_handle_pool_mu);
        for (HandleRes& r : g_handle_pool) {
            on_device(r.device);
            r.ctx.reset();
            for (int i = 0; i < 2; i++) { dev_free(r.d[i]); r.cap[i] = 0; }
        }
        g_handle_pool.clear();
    }
    if (have_dev) cudaSetDevice(saved);
    return DFL_OK;
}
