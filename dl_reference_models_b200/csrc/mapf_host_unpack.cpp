// Host side of mapf_step_host's packed device->host transfer: a small spin-then-sleep thread pool that expands
// the bit-packed agent records (mapf_pack_kernel.cuh) into the caller's plain arrays while the next slice's PCIe
// copy is in flight.  The expansion is exact: cell codes and mask bits are copied bit for bit, the goal delta is
// looked up in the same float table the kernels use, the reward is 0.5f * an exact integer.
#include "mapf_host_unpack.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

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

// the three streams of a slice's packed block (layout: mapf_host_unpack.h) and the output arrays, all at agent b0
struct Cursor {
    const uint8_t *bits;  // window + mask bits, pack_obs_bytes(V2) per agent
    const int8_t *diff;   // (d_row, d_col) per agent
    const int8_t *rew2;   // 2 * reward per agent
    uint8_t *obs;
    int8_t *mask;
    float *gd, *rw;
};

// Scalar expansion of n agents.  One body, two instantiations: the BMI2 one (pdep expands 8 cells in one
// instruction) carries the target attribute and is only called when the CPU reports the extension.  Window sizes
// are 9 / 25 / 49 (sensor range 1..3): always whole 8-cell chunks plus one cell that shares a byte with the mask.
#define MAPF_UNPACK_SCALAR(NAME, ATTR, SPREAD3, SPREAD1)                                          \
    ATTR void NAME(int V2, Cursor c, int64_t n, const float *gr, const float *gc) {               \
        const int nfull = V2 >> 3, PB = nfull * 3 + 1;                                            \
        for (int64_t a = 0; a < n; ++a) {                                                         \
            for (int k = 0; k < nfull; ++k) {                                                     \
                uint32_t w = (uint32_t)c.bits[3 * k] | ((uint32_t)c.bits[3 * k + 1] << 8) |       \
                             ((uint32_t)c.bits[3 * k + 2] << 16);                                 \
                const uint64_t x = SPREAD3(w);                                                    \
                memcpy(c.obs + 8 * k, &x, 8);                                                     \
            }                                                                                     \
            const uint32_t t = c.bits[3 * nfull];                                                 \
            c.obs[8 * nfull] = (uint8_t)(t & 7u);                                                 \
            const uint64_t mx = SPREAD1(t >> 3);                                                  \
            memcpy(c.mask, &mx, 5);                                                               \
            c.gd[0] = gr[c.diff[0]];                                                              \
            c.gd[1] = gc[c.diff[1]];                                                              \
            c.rw[0] = 0.5f * (float)c.rew2[0];                                                    \
            c.bits += PB; c.diff += 2; c.rew2 += 1; c.obs += V2; c.mask += 5; c.gd += 2; c.rw += 1; \
        }                                                                                         \
    }

MAPF_UNPACK_SCALAR(unpack_scalar_generic, , spread3_generic, spread1_generic)

#if MAPF_X86
#define MAPF_PDEP3(w) _pdep_u64((w), 0x0707070707070707ULL)
#define MAPF_PDEP1(m) _pdep_u64((m), 0x0101010101ULL)
MAPF_UNPACK_SCALAR(unpack_scalar_bmi2, __attribute__((target("bmi2"))), MAPF_PDEP3, MAPF_PDEP1)

// AVX-512 VBMI expansion.  An agent's bits are laid out as nfull + 1 quadwords (one per 3-byte chunk, one for the
// tail byte) by a byte permute, vpmultishiftqb pulls the 3-bit / 1-bit fields to byte positions, and a second
// permute compacts windows and masks of 8 / (nfull + 1) agents for two masked stores -- 7 instructions per group
// instead of ~25 per agent.  Goal differences and rewards go 16 agents at a time (convert, IEEE divide / multiply:
// the same float values as the kernels' table).
struct VbmiPlan {
    alignas(64) uint8_t gather[64], shift[64], keep[64], obs_idx[64], mask_idx[64];
    int agents;            // agents per 64-byte group
    uint64_t load_mask, obs_mask, mask_mask;
};

VbmiPlan make_vbmi_plan(int V2) {
    VbmiPlan p;
    memset(&p, 0, sizeof(p));
    const int nfull = V2 >> 3, QA = nfull + 1, PB = nfull * 3 + 1;
    p.agents = 8 / QA;
    for (int k = 0; k < p.agents; ++k) {
        for (int q = 0; q < QA; ++q)
            for (int j = 0; j < 8; ++j) {
                const int at = (k * QA + q) * 8 + j;
                if (q < nfull) {
                    p.gather[at] = (uint8_t)(k * PB + q * 3 + (j < 3 ? j : 2));
                    p.shift[at] = (uint8_t)(3 * j);
                    p.keep[at] = 7;
                } else {  // tail byte: cell V2-1 in bits 0..2, mask entries in bits 3..7
                    p.gather[at] = (uint8_t)(k * PB + 3 * nfull);
                    p.shift[at] = (uint8_t)(j == 0 ? 0 : (j <= 5 ? 2 + j : 0));
                    p.keep[at] = (uint8_t)(j == 0 ? 7 : (j <= 5 ? 1 : 0));
                }
            }
        for (int i = 0; i < V2; ++i) p.obs_idx[k * V2 + i] = (uint8_t)(k * QA * 8 + i);
        for (int m = 0; m < 5; ++m) p.mask_idx[k * 5 + m] = (uint8_t)(k * QA * 8 + nfull * 8 + 1 + m);
    }
    auto low = [](int n) { return n >= 64 ? ~0ULL : ((1ULL << n) - 1); };
    p.load_mask = low(p.agents * PB);
    p.obs_mask = low(p.agents * V2);
    p.mask_mask = low(p.agents * 5);
    return p;
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi,bmi2")))
void unpack_vbmi(int V2, Cursor c, int64_t n, const float *gr, const float *gc, float den_row, float den_col) {
    const VbmiPlan p = make_vbmi_plan(V2);
    const int PB = (V2 >> 3) * 3 + 1, G = p.agents;
    const __m512i gather = _mm512_load_si512(p.gather), shift = _mm512_load_si512(p.shift),
                  keep = _mm512_load_si512(p.keep), obs_idx = _mm512_load_si512(p.obs_idx),
                  mask_idx = _mm512_load_si512(p.mask_idx);
    const int64_t groups = n / G;
    for (int64_t g = 0; g < groups; ++g) {
        const __m512i raw = _mm512_maskz_loadu_epi8(p.load_mask, c.bits + g * G * PB);
        const __m512i quads = _mm512_permutexvar_epi8(gather, raw);
        const __m512i cells = _mm512_and_si512(_mm512_multishift_epi64_epi8(shift, quads), keep);
        _mm512_mask_storeu_epi8(c.obs + g * G * V2, p.obs_mask, _mm512_permutexvar_epi8(obs_idx, cells));
        _mm512_mask_storeu_epi8(c.mask + g * G * 5, p.mask_mask, _mm512_permutexvar_epi8(mask_idx, cells));
    }
    // den_row > 0: normalised goal delta = difference / denominator (IEEE division, the kernels' table values)
    alignas(64) float dv[16];
    for (int i = 0; i < 16; ++i) dv[i] = (i & 1) ? den_col : den_row;
    const __m512 denv = _mm512_load_ps(dv);
    const __m512 half = _mm512_set1_ps(0.5f);
    const int64_t n16 = n / 16;
    for (int64_t b = 0; b < n16; ++b) {
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(c.diff + b * 32));
        const __m512 lo = _mm512_cvtepi32_ps(_mm512_cvtepi8_epi32(_mm256_castsi256_si128(d)));
        const __m512 hi = _mm512_cvtepi32_ps(_mm512_cvtepi8_epi32(_mm256_extracti128_si256(d, 1)));
        _mm512_storeu_ps(c.gd + b * 32, _mm512_div_ps(lo, denv));
        _mm512_storeu_ps(c.gd + b * 32 + 16, _mm512_div_ps(hi, denv));
        const __m128i r = _mm_loadu_si128(reinterpret_cast<const __m128i *>(c.rew2 + b * 16));
        _mm512_storeu_ps(c.rw + b * 16, _mm512_mul_ps(_mm512_cvtepi32_ps(_mm512_cvtepi8_epi32(r)), half));
    }
    // left-overs: windows / masks of the last n % G agents, goal deltas / rewards of the last n % 16
    {
        Cursor t = c;
        const int64_t done = groups * G;
        t.bits += done * PB; t.obs += done * V2; t.mask += done * 5;
        alignas(8) float sink_gd[2 * 8], sink_rw[8];
        alignas(8) int8_t zero[2 * 8] = {0};
        for (int64_t a = done; a < n; ++a) {
            Cursor one = t;
            one.diff = zero; one.rew2 = zero; one.gd = sink_gd; one.rw = sink_rw;
            unpack_scalar_bmi2(V2, one, 1, gr, gc);
            t.bits += PB; t.obs += V2; t.mask += 5;
        }
        for (int64_t a = n16 * 16; a < n; ++a) {
            c.gd[2 * a] = gr[c.diff[2 * a]];
            c.gd[2 * a + 1] = gc[c.diff[2 * a + 1]];
            c.rw[a] = 0.5f * (float)c.rew2[a];
        }
    }
}
#endif

constexpr int kBlock = 64;  // chunk boundaries fall on whole 64-agent blocks

enum { ISA_GENERIC = 0, ISA_BMI2 = 1, ISA_VBMI = 2 };

#if MAPF_X86
// bytes [src, src + n) -> dst with non-temporal 64-byte stores (unaligned head / tail: plain copies)
__attribute__((target("avx512f")))
inline void stream_out(void *dst_, const void *src_, size_t n) {
    uint8_t *dst = static_cast<uint8_t *>(dst_);
    const uint8_t *src = static_cast<const uint8_t *>(src_);
    size_t head = (size_t)(-(intptr_t)reinterpret_cast<uintptr_t>(dst)) & 63u;
    if (head > n) head = n;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    for (; n >= 64; n -= 64, dst += 64, src += 64)
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst), _mm512_loadu_si512(src));
    if (n) memcpy(dst, src, n);
}

void unpack_vbmi(int V2, Cursor c, int64_t n, const float *gr, const float *gc, float den_row, float den_col);

// MAPF_HOST_NT=1: expand kNtBlock agents into a cache-resident block, then stream whole lines out with non-temporal
// stores.  Plain stores read every line of the output arrays before they overwrite it (read for ownership); where the
// host's memory system is what bounds mapf_step_host -- many ranks on one node -- that read is a third of the traffic.
// (On a single-GPU host the call is PCIe-paced and the staged copy only costs: 1.77e9 vs 2.08e9 agent-steps/s.)
constexpr int kNtBlock = 512;
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi,bmi2")))
void unpack_agents_nt(const UnpackJob &j, int64_t b0, int64_t b1) {
    alignas(64) static thread_local uint8_t s_obs[kNtBlock * 49];
    alignas(64) static thread_local int8_t s_mask[kNtBlock * 5];
    alignas(64) static thread_local float s_gd[kNtBlock * 2], s_rw[kNtBlock];
    const int V2 = j.V2, PB = pack_obs_bytes(V2);
    const int64_t na = j.a1 - j.a0;
    const float *gr = j.gdt_row + 128, *gc = j.gdt_col + 128;
    if (j.bytes_src) memcpy(j.bytes_dst + b0, j.bytes_src + (b0 - j.a0), (size_t)(b1 - b0));
    for (int64_t blk = b0; blk < b1; blk += kNtBlock) {
        const int64_t n = (b1 - blk < kNtBlock) ? b1 - blk : kNtBlock, off = blk - j.a0;
        Cursor c;
        c.bits = j.packed + off * PB;
        c.diff = reinterpret_cast<const int8_t *>(j.packed + na * PB) + off * 2;
        c.rew2 = reinterpret_cast<const int8_t *>(j.packed + na * (PB + 2)) + off;
        c.obs = s_obs; c.mask = s_mask; c.gd = s_gd; c.rw = s_rw;
        unpack_vbmi(V2, c, n, gr, gc, j.den_row, j.den_col);
        stream_out(j.obs + blk * V2, s_obs, (size_t)n * V2);
        stream_out(j.mask + blk * 5, s_mask, (size_t)n * 5);
        stream_out(j.goal_delta + blk * 2, s_gd, (size_t)n * 8);
        stream_out(j.reward + blk, s_rw, (size_t)n * 4);
    }
    _mm_sfence();
}
#endif

// Agents [b0, b1) of the job.
void unpack_agents(const UnpackJob &j, int64_t b0, int64_t b1, int isa, bool nt = false) {
#if MAPF_X86
    if (nt && isa == ISA_VBMI) return unpack_agents_nt(j, b0, b1);
#endif
    const int V2 = j.V2, PB = pack_obs_bytes(V2);
    const int64_t na = j.a1 - j.a0, off = b0 - j.a0;
    Cursor c;
    c.bits = j.packed + off * PB;
    c.diff = reinterpret_cast<const int8_t *>(j.packed + na * PB) + off * 2;
    c.rew2 = reinterpret_cast<const int8_t *>(j.packed + na * (PB + 2)) + off;
    c.obs = j.obs + b0 * V2; c.mask = j.mask + b0 * 5; c.gd = j.goal_delta + b0 * 2; c.rw = j.reward + b0;
    const float *gr = j.gdt_row + 128, *gc = j.gdt_col + 128;
    if (j.bytes_src) memcpy(j.bytes_dst + b0, j.bytes_src + off, (size_t)(b1 - b0));
#if MAPF_X86
    if (isa == ISA_VBMI) return unpack_vbmi(V2, c, b1 - b0, gr, gc, j.den_row, j.den_col);
    if (isa == ISA_BMI2) return unpack_scalar_bmi2(V2, c, b1 - b0, gr, gc);
#endif
    unpack_scalar_generic(V2, c, b1 - b0, gr, gc);
}

}  // namespace

// A step is a short queue of jobs (one per slice).  The caller's thread submits job c right after it has enqueued
// slice c's GPU work and goes on enqueueing; the workers take the jobs in order, each job only once its slice has
// arrived -- signalled by a 32-bit ticket the GPU writes into pinned host memory behind the slice's copy (no CUDA
// call on a worker thread).  Workers spin for a while after a step (the slices of one step, and back-to-back steps,
// arrive within ~100 us), then sleep on a condition variable.
constexpr int kMaxJobs = 32;

struct QueuedJob {
    UnpackJob job{};
    std::atomic<const volatile uint32_t *> ticket{nullptr};
    std::atomic<uint32_t> ticket_value{0};
    // chunk claims carry the step number in the upper half: a worker that is late by a whole step can never claim
    // a chunk of a job whose ticket it has not waited for
    std::atomic<uint64_t> next{0};
    std::atomic<int> done{0};
    std::atomic<int64_t> t_first{0}, t_last{0};  // trace: first chunk claimed / last chunk done (steady_clock ns)
};

struct HostPool {
    int nthreads = 1, nchunks = 1;
    int isa = ISA_GENERIC;
    std::atomic<bool> nt{false};   // staged expansion with non-temporal stores (hosts bound by their memory system)
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> epoch{0};       // one per step
    std::atomic<int> published{0};        // jobs of this step submitted so far
    std::atomic<bool> closed{true};       // no more jobs will be submitted this step
    std::atomic<int> sleepers{0};
    std::atomic<bool> stop{false}, failed{false};
    QueuedJob jobs[kMaxJobs];

    void run_job(QueuedJob &q) {
        const uint64_t tag = q.next.load(std::memory_order_acquire) >> 32;
        const volatile uint32_t *ticket = q.ticket.load(std::memory_order_relaxed);
        const uint32_t want = q.ticket_value.load(std::memory_order_relaxed);
        if (ticket) {
            const auto t0 = std::chrono::steady_clock::now();
            uint32_t polls = 0;
            while (*ticket != want) {
                if (stop.load(std::memory_order_relaxed) || failed.load(std::memory_order_relaxed) ||
                    (q.next.load(std::memory_order_relaxed) >> 32) != tag)
                    return;
                // a ticket that never comes means the GPU work in front of it failed: give up instead of hanging
                if ((++polls & 0xFFFFu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                    failed.store(true, std::memory_order_release);
                    return;
                }
                cpu_relax();
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        for (;;) {
            uint64_t v = q.next.load(std::memory_order_acquire);
            int c = -1;
            while ((v >> 32) == tag && (int)(v & 0xFFFFFFFFu) < nchunks) {
                if (q.next.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) { c = (int)(v & 0xFFFFFFFFu); break; }
            }
            if (c < 0) return;
            if (c == 0) q.t_first.store(std::chrono::steady_clock::now().time_since_epoch().count(), std::memory_order_relaxed);
            const UnpackJob &job = q.job;
            const int64_t n = job.a1 - job.a0, nblk = (n + kBlock - 1) / kBlock;
            int64_t b0 = job.a0 + kBlock * (nblk * c / nchunks), b1 = job.a0 + kBlock * (nblk * (c + 1) / nchunks);
            if (b1 > job.a1) b1 = job.a1;
            if (b0 < b1) unpack_agents(job, b0, b1, isa, nt.load(std::memory_order_relaxed));
            if (q.done.fetch_add(1, std::memory_order_acq_rel) + 1 == nchunks)
                q.t_last.store(std::chrono::steady_clock::now().time_since_epoch().count(), std::memory_order_relaxed);
        }
    }

    // take this step's jobs in order until the step is closed and every published job has been visited
    void run_step() {
        for (int q = 0; q < kMaxJobs; ++q) {
            while (published.load(std::memory_order_acquire) <= q) {
                if (closed.load(std::memory_order_acquire) && published.load(std::memory_order_acquire) <= q) return;
                if (stop.load(std::memory_order_relaxed)) return;
                cpu_relax();
            }
            run_job(jobs[q]);
        }
    }

    void worker() {
        uint64_t seen = 0;
        for (;;) {
            const auto t0 = std::chrono::steady_clock::now();
            int polls = 0;
            while (epoch.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_acquire)) {
                cpu_relax();
                if ((++polls & 255) == 0 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) {
                    std::unique_lock<std::mutex> lk(mu);
                    sleepers.fetch_add(1, std::memory_order_acq_rel);
                    cv.wait(lk, [&] {
                        return epoch.load(std::memory_order_acquire) != seen || stop.load(std::memory_order_acquire);
                    });
                    sleepers.fetch_sub(1, std::memory_order_acq_rel);
                    break;
                }
            }
            if (stop.load(std::memory_order_acquire)) return;
            seen = epoch.load(std::memory_order_acquire);
            run_step();
        }
    }
};

HostPool *host_pool_create(int threads, const int *cpus, int ncpus) {
    HostPool *p = new HostPool();
    p->nthreads = threads < 1 ? 1 : threads;
    p->nchunks = p->nthreads * 4;  // fixed for the pool's lifetime
#if MAPF_X86
    if (__builtin_cpu_supports("bmi2")) p->isa = ISA_BMI2;
    if (p->isa == ISA_BMI2 && __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512bw") &&
        __builtin_cpu_supports("avx512vl"))
        p->isa = ISA_VBMI;
    if (const char *ov = getenv("MAPF_HOST_NT")) p->nt.store(atoi(ov) != 0);
    if (const char *ov = getenv("MAPF_HOST_ISA")) {  // generic | bmi2 | vbmi: cap the instruction set (tests)
        const int cap = ov[0] == 'g' ? ISA_GENERIC : ov[0] == 'b' ? ISA_BMI2 : ISA_VBMI;
        if (cap < p->isa) p->isa = cap;
    }
#endif
    for (int i = 1; i < p->nthreads; ++i) {
        p->workers.emplace_back([p] { p->worker(); });
#if defined(__linux__)
        if (cpus && ncpus > 0) {   // one core per worker, inside this rank's share of the node (no migration, no sharing)
            cpu_set_t set;
            CPU_ZERO(&set);
            CPU_SET(cpus[(i - 1) % ncpus], &set);
            pthread_setaffinity_np(p->workers.back().native_handle(), sizeof(set), &set);
        }
#endif
    }
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

void host_pool_set_nt(HostPool *p, bool on) { if (p) p->nt.store(on, std::memory_order_release); }

int host_pool_threads(const HostPool *p) { return p ? p->nthreads : 0; }

void host_pool_begin(HostPool *p) {
    p->failed.store(false, std::memory_order_relaxed);
    p->published.store(0, std::memory_order_relaxed);
    p->closed.store(false, std::memory_order_release);
    p->epoch.fetch_add(1, std::memory_order_acq_rel);
    if (p->sleepers.load(std::memory_order_acquire) > 0) {
        { std::lock_guard<std::mutex> lk(p->mu); }
        p->cv.notify_all();
    }
}

bool host_pool_submit(HostPool *p, const UnpackJob &job, const volatile uint32_t *ticket, uint32_t ticket_value) {
    const int q = p->published.load(std::memory_order_relaxed);
    if (q >= kMaxJobs) return false;
    if (job.a1 <= job.a0) return true;
    QueuedJob &slot = p->jobs[q];
    slot.job = job;
    slot.ticket.store(ticket, std::memory_order_relaxed);
    slot.ticket_value.store(ticket_value, std::memory_order_relaxed);
    slot.done.store(0, std::memory_order_relaxed);
    slot.next.store((p->epoch.load(std::memory_order_relaxed) & 0xFFFFFFFFu) << 32, std::memory_order_release);
    p->published.store(q + 1, std::memory_order_release);
    return true;
}

bool host_pool_finish(HostPool *p) {
    p->closed.store(true, std::memory_order_release);
    p->run_step();
    const int n = p->published.load(std::memory_order_acquire);
    for (int q = 0; q < n; ++q)
        while (p->jobs[q].done.load(std::memory_order_acquire) < p->nchunks) {
            if (p->failed.load(std::memory_order_acquire)) return false;
            cpu_relax();
        }
    return !p->failed.load(std::memory_order_acquire);
}

void host_pool_job_times(const HostPool *p, int q, int64_t *first_ns, int64_t *last_ns) {
    *first_ns = p->jobs[q].t_first.load(std::memory_order_relaxed);
    *last_ns = p->jobs[q].t_last.load(std::memory_order_relaxed);
}

// What the host side of mapf_step_host can hope for on THIS machine: streaming fill (write-only) and copy rates of
// `threads` threads over private chunks of a buffer far larger than the caches (best of three passes, GB/s of bytes
// written).  The expansion writes every delivered byte once, so delivered bytes / fill rate is its floor.
void host_memory_probe(int threads, int64_t bytes, const int *cpus, int ncpus, double *fill_gbs, double *copy_gbs) {
    if (threads < 1) threads = 1;
    const int64_t chunk = (bytes / threads) & ~int64_t(63);
    std::vector<uint8_t> a((size_t)(chunk * threads) + 64), b((size_t)(chunk * threads) + 64);
    memset(a.data(), 1, a.size());
    memset(b.data(), 2, b.size());
    double best[2] = {0.0, 0.0};
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 4; ++rep) {
            std::atomic<int> ready{0};
            std::atomic<bool> go{false};
            std::vector<std::thread> ts;
            for (int t = 0; t < threads; ++t)
                ts.emplace_back([&, t] {
                    ready.fetch_add(1);
                    while (!go.load(std::memory_order_acquire)) cpu_relax();
                    uint8_t *dst = b.data() + (size_t)t * chunk;
                    if (mode == 0) memset(dst, rep, (size_t)chunk);
                    else memcpy(dst, a.data() + (size_t)t * chunk, (size_t)chunk);
                });
#if defined(__linux__)
            if (cpus && ncpus > 0)
                for (int t = 0; t < threads; ++t) {
                    cpu_set_t set;
                    CPU_ZERO(&set);
                    CPU_SET(cpus[t % ncpus], &set);
                    pthread_setaffinity_np(ts[t].native_handle(), sizeof(set), &set);
                }
#endif
            while (ready.load() < threads) cpu_relax();
            const auto t0 = std::chrono::steady_clock::now();
            go.store(true, std::memory_order_release);
            for (auto &t : ts) t.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const double gbs = (double)(chunk * threads) / dt / 1e9;
            if (rep > 0 && gbs > best[mode]) best[mode] = gbs;
        }
    if (fill_gbs) *fill_gbs = best[0];
    if (copy_gbs) *copy_gbs = best[1];
}

void host_pool_unpack(HostPool *p, const UnpackJob &job) {
    host_pool_begin(p);
    host_pool_submit(p, job, nullptr, 0);
    host_pool_finish(p);
}

}  // namespace mapf
