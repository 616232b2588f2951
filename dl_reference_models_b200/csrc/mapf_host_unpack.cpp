// Host side of mapf_step_host's packed device->host transfer: a small spin-then-sleep thread pool that expands
// the bit-packed agent records (mapf_pack_kernel.cuh) into the caller's plain arrays while the next slice's PCIe
// copy is in flight.  The expansion is exact: cell codes and mask bits are copied bit for bit, the goal delta is
// looked up in the same float table the kernels use, the reward is 0.5f * an exact integer.
#include "mapf_host_unpack.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define MAPF_X86 1
#else
#define MAPF_X86 0
#endif

namespace mapf {

namespace {

inline void cpu_relax() {
#if MAPF_X86
    _mm_pause();
#else
    std::this_thread::yield();
#endif
}

inline uint64_t spread3_generic(uint32_t w) {  // 8 three-bit fields -> 8 bytes
    uint64_t x = 0;
    for (int i = 0; i < 8; ++i) x |= (uint64_t)((w >> (3 * i)) & 7u) << (8 * i);
    return x;
}

inline uint64_t spread1_generic(uint32_t m) {  // 5 bits -> 5 bytes of 0/1
    uint64_t x = 0;
    for (int i = 0; i < 5; ++i) x |= (uint64_t)((m >> i) & 1u) << (8 * i);
    return x;
}

// One body, two instantiations: the BMI2 one (pdep does an 8-cell expansion in one instruction) is compiled with
// the target attribute and only called when the CPU reports the extension.
#define MAPF_UNPACK_BODY(SPREAD3, SPREAD1)                                                        \
    const int V2 = j.V2, RS = j.RS, nfull = V2 >> 3, rem = V2 & 7;                                \
    const int tail_bytes = (rem * 3 + 5 + 7) / 8;                                                 \
    for (int64_t a = b0; a < b1; ++a) {                                                           \
        const uint8_t *r = j.packed + (a - j.a0) * RS;                                            \
        uint8_t *o = j.obs + a * V2;                                                              \
        for (int c = 0; c < nfull; ++c) {                                                         \
            uint32_t w;                                                                           \
            memcpy(&w, r, 4); /* 3 payload bytes; the 4th belongs to the record's tail */         \
            const uint64_t x = SPREAD3(w & 0xFFFFFFu);                                            \
            memcpy(o + 8 * c, &x, 8);                                                             \
            r += 3;                                                                               \
        }                                                                                         \
        uint32_t w;                                                                               \
        memcpy(&w, r, 4);                                                                         \
        for (int i = 0; i < rem; ++i) o[nfull * 8 + i] = (uint8_t)((w >> (3 * i)) & 7u);          \
        const uint64_t mx = SPREAD1((w >> (3 * rem)) & 31u);                                      \
        memcpy(j.mask + a * 5, &mx, 5);                                                           \
        r += tail_bytes;                                                                          \
        j.goal_delta[2 * a] = j.gdt_row[(int)(int8_t)r[0] + 128];                                 \
        j.goal_delta[2 * a + 1] = j.gdt_col[(int)(int8_t)r[1] + 128];                             \
        j.reward[a] = 0.5f * (float)(int8_t)r[2];                                                 \
        if (j.blocking_prev) j.blocking_prev[a] = r[3];                                           \
    }

void unpack_range_generic(const UnpackJob &j, int64_t b0, int64_t b1) {
    MAPF_UNPACK_BODY(spread3_generic, spread1_generic)
}

#if MAPF_X86
#define MAPF_PDEP3(w) _pdep_u64((w), 0x0707070707070707ULL)
#define MAPF_PDEP1(m) _pdep_u64((m), 0x0101010101ULL)
__attribute__((target("bmi2"))) void unpack_range_bmi2(const UnpackJob &j, int64_t b0, int64_t b1) {
    MAPF_UNPACK_BODY(MAPF_PDEP3, MAPF_PDEP1)
}
#endif

}  // namespace

struct HostPool {
    int nthreads = 1, nchunks = 1;
    bool bmi2 = false;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> generation{0};
    std::atomic<int> next{0}, done{0};
    std::atomic<bool> stop{false};
    UnpackJob job{};

    void run_chunks() {
        for (;;) {
            const int c = next.fetch_add(1, std::memory_order_acq_rel);
            if (c >= nchunks) return;
            const int64_t n = job.a1 - job.a0;
            const int64_t b0 = job.a0 + n * c / nchunks, b1 = job.a0 + n * (c + 1) / nchunks;
#if MAPF_X86
            if (bmi2) unpack_range_bmi2(job, b0, b1); else
#endif
            unpack_range_generic(job, b0, b1);
            done.fetch_add(1, std::memory_order_acq_rel);
        }
    }

    void worker() {
        uint64_t seen = 0;
        for (;;) {
            // spin for a while (the slices of one step, and back-to-back steps, arrive within ~100 us), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            int polls = 0;
            while (generation.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_acquire)) {
                cpu_relax();
                if ((++polls & 255) == 0 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] {
                        return generation.load(std::memory_order_acquire) != seen || stop.load(std::memory_order_acquire);
                    });
                    break;
                }
            }
            if (stop.load(std::memory_order_acquire)) return;
            seen = generation.load(std::memory_order_acquire);
            run_chunks();
        }
    }
};

HostPool *host_pool_create(int threads) {
    HostPool *p = new HostPool();
    p->nthreads = threads < 1 ? 1 : threads;
    p->nchunks = p->nthreads * 4;  // fixed for the pool's lifetime (late wakers can never claim a stale chunk)
#if MAPF_X86
    p->bmi2 = __builtin_cpu_supports("bmi2");
#endif
    p->next.store(p->nchunks);
    for (int i = 1; i < p->nthreads; ++i) p->workers.emplace_back([p] { p->worker(); });
    return p;
}

void host_pool_destroy(HostPool *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop.store(true, std::memory_order_release);
    }
    p->cv.notify_all();
    for (auto &t : p->workers) t.join();
    delete p;
}

int host_pool_threads(const HostPool *p) { return p ? p->nthreads : 0; }

void host_pool_unpack(HostPool *p, const UnpackJob &job) {
    if (job.a1 <= job.a0) return;
    p->job = job;
    p->done.store(0, std::memory_order_relaxed);
    p->next.store(0, std::memory_order_release);
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->generation.fetch_add(1, std::memory_order_acq_rel);
    }
    p->cv.notify_all();
    p->run_chunks();
    while (p->done.load(std::memory_order_acquire) < p->nchunks) cpu_relax();
}

}  // namespace mapf
