// Host side of the reference-signature call: a list of pageable float32 clips -> one pinned staging buffer.
//
// R/processor.py:88-105 and R/training/trainer_unfreeze.py:856-860 hand the extractor ordinary (pageable) NumPy arrays, so
// the drop-in has to move 123 MB per cfg2 batch into pinned memory before the H2D copy can run at PCIe speed.  That copy is
// pure memory traffic on the host; it is done here by a small persistent pool of native threads (no GIL, no per-clip Python
// task) with NON-TEMPORAL stores: the destination is written once and read next by the DMA engine, so pulling it through the
// cache (and paying a read-for-ownership of every line) only adds a third more DRAM traffic.  A job is cut into slices of
// bytes, not clips, so ragged batches balance as well as uniform ones.
//
// stx_host_pack is synchronous (returns when the bytes are in place).  stx_host_pack_begin / _wait / _end run a job of several
// chunks on a native driver thread: feature_extraction.py waits for chunk c right before it enqueues that chunk's H2D copy, so
// packing chunk c + 1 overlaps the copy of chunk c, with no interpreter thread (and no GIL hand-off) in the loop.
#include "stx_common.h"
#include <immintrin.h>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <sched.h>
#include <thread>

namespace stx {
namespace {

struct Segment { const unsigned char* src; unsigned char* dst; size_t bytes; };

__attribute__((target("avx2"))) void copy_stream_avx2(unsigned char* dst, const unsigned char* src, size_t n) {
    // head: up to the first 32-byte boundary of the destination
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head > n) head = n;
    std::memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    std::memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}

void copy_bytes(unsigned char* dst, const unsigned char* src, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) copy_stream_avx2(dst, src, n);
    else std::memcpy(dst, src, n);
}

// A fixed pool of worker threads; one job at a time (callers serialise on job_mutex).
class PackPool {
public:
    explicit PackPool(int n) : stop_(false), generation_(0), pending_(0) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { run(i); });
    }
    ~PackPool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++generation_; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    // copies all segments, split into `parts` byte ranges handled by the workers (and the calling thread)
    void run_job(const std::vector<Segment>& segs, int parts) {
        std::lock_guard<std::mutex> job(job_mutex_);
        size_t total = 0;
        for (const auto& s : segs) total += s.bytes;
        if (parts < 1) parts = 1;
        if (parts > size() + 1) parts = size() + 1;
        if (total < (size_t)1 << 20 || parts == 1) {       // small jobs: not worth waking anybody
            for (const auto& s : segs) copy_bytes(s.dst, s.src, s.bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> l(m_);
            segs_ = &segs; total_ = total; parts_ = parts; next_part_ = 0; pending_ = parts;
            ++generation_;
        }
        cv_.notify_all();
        work();                                             // the caller takes parts as well
        std::unique_lock<std::mutex> l(m_);
        done_cv_.wait(l, [this] { return pending_ == 0; });
        segs_ = nullptr;
    }

private:
    void copy_part(int part) {
        // byte range [lo, hi) of the concatenation of all segments, cut on 64-byte boundaries
        const size_t per = ((total_ + parts_ - 1) / parts_ + 63) & ~size_t(63);
        const size_t lo = std::min(total_, per * part), hi = std::min(total_, per * (part + 1));
        size_t pos = 0;
        for (const auto& s : *segs_) {
            const size_t a = std::max(lo, pos), b = std::min(hi, pos + s.bytes);
            if (a < b) copy_bytes(s.dst + (a - pos), s.src + (a - pos), b - a);
            pos += s.bytes;
            if (pos >= hi) break;
        }
    }
    void work() {
        for (;;) {
            int part;
            {
                std::lock_guard<std::mutex> l(m_);
                if (!segs_ || next_part_ >= parts_) return;
                part = next_part_++;
            }
            copy_part(part);
            bool last;
            { std::lock_guard<std::mutex> l(m_); last = (--pending_ == 0); }
            if (last) done_cv_.notify_all();
        }
    }
    void run(int) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, job_mutex_;
    std::condition_variable cv_, done_cv_;
    bool stop_;
    unsigned long long generation_;
    const std::vector<Segment>* segs_ = nullptr;
    size_t total_ = 0;
    int parts_ = 0, next_part_ = 0, pending_;
};

// one pool per process, sized once from the cores this process may run on (never resized: callers may be concurrent)
PackPool* pool() {
    static PackPool* p = [] {
        int cores = (int)std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        return new PackPool(std::max(0, std::min(cores, 32) - 1));      // never destroyed: workers sleep on a condition variable
    }();
    return p;
}

// A multi-chunk packing job driven by its own native thread: chunk c is packed (by the pool) after chunk c - 1, and
// `done` counts the chunks in place.  The caller waits per chunk without holding any interpreter lock.
struct PackJob {
    std::vector<std::vector<Segment>> chunks;
    int threads = 1;
    int done = 0;
    std::mutex m;
    std::condition_variable cv;
    std::thread driver;
};

}  // namespace
}  // namespace stx

extern "C" {

void* stx_host_pack_begin(const void* const* h_src, const int64_t* n_bytes, void* h_dst_base, const int64_t* dst_byte_offsets,
                          int count, const int32_t* chunk_starts, int n_chunks, int threads) {
    using namespace stx;
    if (count < 0 || n_chunks < 0 || threads < 1 || (count > 0 && (!h_src || !n_bytes || !h_dst_base || !dst_byte_offsets)) ||
        (n_chunks > 0 && !chunk_starts)) {
        set_error("stx_host_pack_begin: bad argument");
        return nullptr;
    }
    PackJob* job = new PackJob;
    job->threads = threads;
    job->chunks.resize(n_chunks);
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = chunk_starts[c], b1 = chunk_starts[c + 1];
        if (b0 < 0 || b1 < b0 || b1 > count) { set_error("stx_host_pack_begin: bad chunk %d", c); delete job; return nullptr; }
        for (int i = b0; i < b1; ++i) {
            if (n_bytes[i] < 0 || dst_byte_offsets[i] < 0 || (n_bytes[i] > 0 && !h_src[i])) { set_error("stx_host_pack_begin: bad segment %d", i); delete job; return nullptr; }
            if (n_bytes[i] > 0)
                job->chunks[c].push_back({static_cast<const unsigned char*>(h_src[i]),
                                          static_cast<unsigned char*>(h_dst_base) + dst_byte_offsets[i], (size_t)n_bytes[i]});
        }
    }
    job->driver = std::thread([job] {
        for (size_t c = 0; c < job->chunks.size(); ++c) {
            if (!job->chunks[c].empty()) pool()->run_job(job->chunks[c], job->threads);
            { std::lock_guard<std::mutex> l(job->m); job->done = (int)c + 1; }
            job->cv.notify_all();
        }
    });
    return job;
}

int stx_host_pack_wait(void* h_job, int chunk) {
    using namespace stx;
    PackJob* job = static_cast<PackJob*>(h_job);
    if (!job) { set_error("stx_host_pack_wait: null job"); return STX_EINVAL; }
    const int want = (chunk < 0 || chunk >= (int)job->chunks.size()) ? (int)job->chunks.size() : chunk + 1;
    std::unique_lock<std::mutex> l(job->m);
    job->cv.wait(l, [&] { return job->done >= want; });
    return 0;
}

int stx_host_pack_end(void* h_job) {
    using namespace stx;
    PackJob* job = static_cast<PackJob*>(h_job);
    if (!job) return 0;
    if (job->driver.joinable()) job->driver.join();
    delete job;
    return 0;
}

int stx_host_pack(const void* const* h_src, const int64_t* n_bytes, void* h_dst_base, const int64_t* dst_byte_offsets,
                  int count, int threads) {
    using namespace stx;
    if (count < 0 || threads < 1 || (count > 0 && (!h_src || !n_bytes || !h_dst_base || !dst_byte_offsets))) {
        set_error("stx_host_pack: bad argument");
        return STX_EINVAL;
    }
    std::vector<Segment> segs;
    segs.reserve(count);
    for (int i = 0; i < count; ++i) {
        if (n_bytes[i] < 0 || dst_byte_offsets[i] < 0 || (n_bytes[i] > 0 && !h_src[i])) { set_error("stx_host_pack: bad segment %d", i); return STX_EINVAL; }
        if (n_bytes[i] > 0)
            segs.push_back({static_cast<const unsigned char*>(h_src[i]),
                            static_cast<unsigned char*>(h_dst_base) + dst_byte_offsets[i], (size_t)n_bytes[i]});
    }
    if (segs.empty()) return 0;
    pool()->run_job(segs, threads);
    return 0;
}

}  // extern "C"
