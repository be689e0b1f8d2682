// vmp_stage.hpp — host side of the scan upload: copy of the caller's (pageable) cloud into the handle's pinned staging
// area by a few helper threads, optionally checking the time order of a raw scan in the same pass.
//
// Why: a 200 000-point scan is 2.4 MB (3.2 MB raw); one host core moves it at ~10 GB/s = 0.2-0.3 ms, which is as long as
// the whole device side of the scan.  The copy is embarrassingly parallel, so the handle keeps a small pool of helpers
// (default: up to 11 + the calling thread, two cores left alone, parts of at least 32 KB; VMP_COPY_THREADS=n sets the number of helpers, 0
// disables them.  Measured on the C2 scan, 2.4 MB, 16-core host: vmp_scan end to end 2446 scans/s with 3 helpers, 2588 with 7, 2681 with 11).  Helpers spin briefly after a job (back-to-back scans
// find them awake) and then sleep on a condition variable (at sensor rate they cost nothing).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace vmp {

class StagePool {
public:
    StagePool() = default;
    ~StagePool() { shutdown(); }
    StagePool(const StagePool&) = delete;
    StagePool& operator=(const StagePool&) = delete;

    // dst <- src (bytes).  check_stride > 0: src is an array of records of check_stride floats whose last float is a time
    // offset; returns false if those are not non-decreasing (lio_builder.cpp:75 sorts in that case).
    bool copy(void* dst, const void* src, size_t bytes, int check_stride = 0) {
        begin(dst, src, bytes, check_stride);
        return end();
    }

    // The same copy in two halves: begin() hands the helpers their parts and returns at once, end() does the caller's own part and
    // waits for the helpers.  In between the caller is free (vmp_scan launches the scan's graph there).  A copy too small for
    // the helpers is done in end().
    void begin(void* dst, const void* src, size_t bytes, int check_stride = 0) {
        constexpr size_t MIN_PAR = 512 * 1024;
        drain();
        job_.chunked = false; job_.gather = 0;
        job_.dst = (char*)dst; job_.src = (const char*)src; job_.bytes = bytes; job_.check_stride = check_stride;
        inline_ = bytes < MIN_PAR || !ensure_started();
        if (inline_) return;
        int parts = (int)workers_.size() + 1;
        const int by_size = (int)(bytes / (32 * 1024));                   // no more parts than 32 KB pieces
        if (parts > by_size) parts = by_size < 2 ? 2 : by_size;
        // whole records per part, cache-line friendly
        const size_t rec = check_stride > 0 ? (size_t)check_stride * sizeof(float) : 64;
        job_.rec = rec; job_.nrec = bytes / rec; job_.parts = parts;
        sorted_.store(true, std::memory_order_relaxed);
        remaining_.store(parts - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);        // generation bump under the lock: a helper about to sleep cannot miss it
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    }
    bool end() {
        if (inline_) return slice(job_.dst, job_.src, job_.bytes, 0, job_.bytes, job_.check_stride);
        bool ok = run_part(0);
        while (remaining_.load(std::memory_order_acquire) != 0) cpu_relax();
        return ok && sorted_.load(std::memory_order_relaxed);
    }

    // dst[3 i .. 3 i + 2] <- src[stride i .. stride i + 2] for n records (x y z out of the caller's point type, lio_builder.cpp:224-229), by
    // the helpers and the caller together
    void gather_xyz(float* dst, const float* src, int stride, size_t n) {
        drain();
        job_.chunked = false;
        job_.dst = (char*)dst; job_.src = (const char*)src; job_.nrec = n; job_.rec = (size_t)stride * sizeof(float); job_.bytes = n * job_.rec;
        job_.check_stride = 0; job_.gather = stride;
        if (n * 12 < 256 * 1024 || !ensure_started()) { gather_slice(job_, 0, n); job_.gather = 0; return; }
        int parts = (int)workers_.size() + 1;
        const int by_size = (int)(n * 12 / (32 * 1024));
        if (parts > by_size) parts = by_size < 2 ? 2 : by_size;
        job_.parts = parts;
        remaining_.store(parts - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        run_part(0);
        drain();
        job_.gather = 0;
    }

    // A copy the helpers run through WITHOUT the caller: chunk after chunk, every helper its piece of each, one arrival counter per chunk.
    // The caller only waits for chunk c (wait_chunk) and ships it while the helpers are already on the next ones - no barrier and no
    // wake-up between chunks (vmp_scan: the DMA copies of a streamed upload).  Without helpers wait_chunk copies the chunk itself.
    static constexpr int MAX_CHUNKS = 16;
    int helpers() { ensure_started(); return (int)workers_.size(); }
    void begin_chunks(void* dst, const void* src, size_t bytes, size_t chunk_bytes) {
        drain();
        job_.gather = 0;
        job_.dst = (char*)dst; job_.src = (const char*)src; job_.bytes = bytes; job_.check_stride = 0;
        job_.chunk = chunk_bytes; job_.nchunks = (int)((bytes + chunk_bytes - 1) / chunk_bytes);
        inline_ = bytes < 256 * 1024 || job_.nchunks > MAX_CHUNKS || !ensure_started();
        if (inline_) return;
        job_.chunked = true;
        job_.parts = (int)workers_.size();
        for (int c = 0; c < job_.nchunks; c++) done_[c].store(0, std::memory_order_relaxed);
        remaining_.store(job_.parts, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    }
    void wait_chunk(int c) {
        if (inline_) {
            const size_t b0 = (size_t)c * job_.chunk, b1 = std::min(job_.bytes, b0 + job_.chunk);
            if (b1 > b0) std::memcpy(job_.dst + b0, job_.src + b0, b1 - b0);
            return;
        }
        while (done_[c].load(std::memory_order_acquire) != job_.parts) cpu_relax();
    }
    void end_chunks() { drain(); job_.chunked = false; }

    void shutdown() {
        if (workers_.empty()) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
    }

private:
    struct Job { char* dst; const char* src; size_t bytes, rec, nrec; int parts, check_stride; size_t chunk; int nchunks; bool chunked; int gather; };

    static void gather_slice(const Job& j, size_t r0, size_t r1) {
        float* d = reinterpret_cast<float*>(j.dst);
        const float* s = reinterpret_cast<const float*>(j.src);
        const size_t st = (size_t)j.gather;
        for (size_t i = r0; i < r1; i++) { d[3 * i] = s[st * i]; d[3 * i + 1] = s[st * i + 1]; d[3 * i + 2] = s[st * i + 2]; }
    }

    void drain() { while (remaining_.load(std::memory_order_acquire) != 0) cpu_relax(); }

    // helper `part` of `parts`: its 64-byte-aligned piece of every chunk, in chunk order
    void run_chunks(int part) {
        const Job& j = job_;
        for (int c = 0; c < j.nchunks; c++) {
            const size_t c0 = (size_t)c * j.chunk, c1 = std::min(j.bytes, c0 + j.chunk), len = c1 - c0;
            const size_t per = ((len + j.parts - 1) / j.parts + 63) / 64 * 64;
            const size_t b0 = std::min(len, per * (size_t)part), b1 = std::min(len, per * (size_t)(part + 1));
            if (b1 > b0) std::memcpy(j.dst + c0 + b0, j.src + c0 + b0, b1 - b0);
            done_[c].fetch_add(1, std::memory_order_release);
        }
    }

    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }

    bool ensure_started() {
        if (started_) return !workers_.empty();
        started_ = true;
        const unsigned hc = std::thread::hardware_concurrency();
        int n = hc >= 4 ? (int)std::min(11u, hc - 2) : 1;
        if (const char* e = std::getenv("VMP_COPY_THREADS")) n = std::atoi(e);
        if (hc > 0 && (unsigned)n + 1 > hc) n = (int)hc - 1;
        if (n <= 0) return false;
        try {
            for (int i = 0; i < n; i++) workers_.emplace_back([this, i] { worker(i + 1); });
        } catch (...) {
            shutdown();
            return false;
        }
        return true;
    }

    // copy [b0, b1) of the job; with check_stride also verify the order inside the slice and across its left edge
    static bool slice(void* dst, const void* src, size_t bytes, size_t b0, size_t b1, int check_stride) {
        (void)bytes;
        bool ok = true;
        if (check_stride <= 0) {
            std::memcpy((char*)dst + b0, (const char*)src + b0, b1 - b0);
            return true;
        }
        // block-wise so that the order check reads what the copy just pulled into the cache
        const size_t rec = (size_t)check_stride * sizeof(float);
        constexpr size_t BLK = 64 * 1024;
        for (size_t p = b0; p < b1; p += BLK - BLK % rec) {
            const size_t e = std::min(b1, p + (BLK - BLK % rec));
            std::memcpy((char*)dst + p, (const char*)src + p, e - p);
            const float* f = reinterpret_cast<const float*>((const char*)dst + p);
            const size_t n = (e - p) / rec;
            float prev = p > 0 ? *reinterpret_cast<const float*>((const char*)src + p - sizeof(float)) : f[check_stride - 1];
            bool o = true;
            for (size_t i = 0; i < n; i++) { const float t = f[i * check_stride + check_stride - 1]; o &= !(t < prev); prev = t; }
            ok &= o;
        }
        return ok;
    }

    bool run_part(int id) {
        const Job& j = job_;
        const size_t per = (j.nrec + j.parts - 1) / j.parts;
        size_t r0 = std::min(j.nrec, per * (size_t)id), r1 = std::min(j.nrec, per * (size_t)(id + 1));
        if (j.gather) { gather_slice(j, r0, r1); return true; }
        size_t b0 = r0 * j.rec, b1 = (id == j.parts - 1) ? j.bytes : r1 * j.rec;     // the last part takes the tail bytes
        if (b1 <= b0) return true;
        return slice(j.dst, j.src, j.bytes, b0, b1, j.check_stride);
    }

    void worker(int id) {
        unsigned long long seen = 0;
        for (;;) {
            // spin for a while, then sleep
            unsigned long long g = gen_.load(std::memory_order_acquire);
            for (int spin = 0; g == seen && spin < 40000; spin++) { cpu_relax(); g = gen_.load(std::memory_order_acquire); }
            if (g == seen) {
                std::unique_lock<std::mutex> lk(mu_);
                sleepers_.fetch_add(1, std::memory_order_release);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_release);
                g = gen_.load(std::memory_order_acquire);
            }
            seen = g;
            if (stop_) return;
            if (job_.chunked) {
                if (id <= job_.parts) { run_chunks(id - 1); remaining_.fetch_sub(1, std::memory_order_release); }
            } else if (id < job_.parts) {
                if (!run_part(id)) sorted_.store(false, std::memory_order_relaxed);
                remaining_.fetch_sub(1, std::memory_order_release);
            }
        }
    }

    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::atomic<unsigned long long> gen_{0};
    std::atomic<int> remaining_{0}, sleepers_{0};
    std::atomic<int> done_[MAX_CHUNKS] = {};
    std::atomic<bool> sorted_{true};
    Job job_{};
    bool started_ = false, stop_ = false, inline_ = true;
};

}  // namespace vmp
