// see host_copy.h
#include "host_copy.h"

#include <emmintrin.h>

#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace xdtts {
namespace {

struct Chunk {
    void* dst;
    const void* src;
    size_t bytes;
    std::atomic<int>* pending;
    bool stream;   // destination is a DMA staging buffer: write it with non-temporal stores
};

// memcpy whose stores bypass the caches.  A staging buffer filled with ordinary stores sits, dirty, in the private caches
// of the cores that wrote it; the DMA engine then has to snoop every line out of them -- measured on the B200 host:
// 10 MB staged by 8 threads left the host -> device copy at 8 GB/s instead of 46 GB/s.  SSE2 is baseline x86-64.
void copy_chunk(const Chunk& c) {
    if (!c.stream || ((uintptr_t)c.dst & 15) || (c.bytes & 15)) {
        memcpy(c.dst, c.src, c.bytes);
        return;
    }
    __m128i* d = (__m128i*)c.dst;
    const __m128i* s = (const __m128i*)c.src;
    const size_t n = c.bytes / 16;
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        const __m128i a = _mm_loadu_si128(s + i), b = _mm_loadu_si128(s + i + 1), e = _mm_loadu_si128(s + i + 2), f = _mm_loadu_si128(s + i + 3);
        _mm_stream_si128(d + i, a);
        _mm_stream_si128(d + i + 1, b);
        _mm_stream_si128(d + i + 2, e);
        _mm_stream_si128(d + i + 3, f);
    }
    for (; i < n; i++) _mm_stream_si128(d + i, _mm_loadu_si128(s + i));
    _mm_sfence();
}

class Copier {
  public:
    Copier() {
        // enough threads to keep up with the DMA engine (~25 GB/s per direction; one thread copies ~5 GB/s): half the
        // cores, at most 8.  XDTTS_COPY_THREADS overrides (0: the calling thread copies alone).
        const int hw = (int)std::thread::hardware_concurrency();
        int n = hw > 0 ? (hw / 2 < 8 ? (hw / 2 < 1 ? 1 : hw / 2) : 8) : 4;
        if (const char* e = getenv("XDTTS_COPY_THREADS")) n = atoi(e);
        if (hw > 0 && n > hw) n = hw;
        if (n < 0) n = 0;
        for (int i = 0; i < n; i++) threads_.emplace_back([this] { loop(); });
    }
    ~Copier() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void push(const Chunk& c) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(c);
        }
        cv_.notify_one();
    }
    bool run_one() {   // the waiting caller copies too
        Chunk c;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (q_.empty()) return false;
            c = q_.front();
            q_.pop_front();
        }
        copy_chunk(c);
        c.pending->fetch_sub(1, std::memory_order_release);
        return true;
    }
    int threads() const { return (int)threads_.size(); }

  private:
    void loop() {
        for (;;) {
            Chunk c;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                c = q_.front();
                q_.pop_front();
            }
            copy_chunk(c);
            c.pending->fetch_sub(1, std::memory_order_release);
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Chunk> q_;
    std::vector<std::thread> threads_;
    bool stop_ = false;
};

Copier& copier() {
    static Copier* c = new Copier();   // never destroyed: helper threads must not be joined from a static destructor
    return *c;
}

}  // namespace

void host_copy_async(void* dst, const void* src, size_t bytes, std::atomic<int>* pending, bool to_staging) {
    const size_t chunk = 512 * 1024;
    Copier& c = copier();
    for (size_t off = 0; off < bytes; off += chunk) {
        const size_t n = bytes - off < chunk ? bytes - off : chunk;
        pending->fetch_add(1, std::memory_order_relaxed);
        c.push(Chunk{(char*)dst + off, (const char*)src + off, n, pending, to_staging});
    }
}

void host_copy_wait(std::atomic<int>* pending) {
    Copier& c = copier();
    while (pending->load(std::memory_order_acquire) > 0)
        if (!c.run_one()) std::this_thread::yield();
}

int host_copy_threads() { return copier().threads(); }

}  // namespace xdtts
