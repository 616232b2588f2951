// Host side of the packed device->host transfer of mapf_step_host (see mapf_pack_kernel.cuh for the record
// format).  Internal to libmapf_b200.so: not part of the C ABI.
#pragma once
#include <cstdint>

namespace mapf {

// bytes of the bit-packed observation window + action mask of one agent: 8 cells -> 3 bytes, the remaining
// (V2 & 7) cells share the tail bytes with the 5 mask bits
inline int pack_obs_bytes(int V2) {
    const int rem = V2 & 7;
    return (V2 >> 3) * 3 + (rem * 3 + 5 + 7) / 8;
}
// A packed block of n agents is three streams, one after the other:
//   [n x pack_obs_bytes: window + mask bits][n x (int8 d_row, int8 d_col)][n x int8 2*reward]
inline int pack_record_bytes(int V2) { return pack_obs_bytes(V2) + 3; }

struct UnpackJob {
    const uint8_t *packed;  // packed block of agents [a0, a1)
    int64_t a0, a1;         // global agent indices (env * N + agent)
    int V2;                 // 9, 25 or 49 (sensor range 1..3)
    uint8_t *obs;           // [BN, V2]   full host arrays (global indexing); may be null
    int8_t *mask;           // [BN, 5]
    float *goal_delta;      // [BN, 2]
    float *reward;          // [BN]
    const float *gdt_row, *gdt_col;  // 256-entry tables indexed by (int8 delta + 128)
    float den_row, den_col;          // the tables' rule: value = delta / den (1 when not normalised)
    const uint8_t *bytes_src;        // optional per-agent byte channel carried as is (blocking_prev): agent a0's byte
    uint8_t *bytes_dst;              // ... and its full host array (global indexing)
};

struct HostPool;
// threads >= 1 (the caller's thread counts as one); cpus / ncpus: optional cores to pin the workers to, one each
HostPool *host_pool_create(int threads, const int *cpus = nullptr, int ncpus = 0);
void host_pool_destroy(HostPool *p);
int host_pool_threads(const HostPool *p);
// expansion through cache-resident blocks and non-temporal whole-line stores (AVX-512 hosts; a no-op elsewhere); set
// between steps.  Default: MAPF_HOST_NT at pool creation, else off.
void host_pool_set_nt(HostPool *p, bool on);
// One step = begin, submit (non-blocking, at most 32 jobs; a job starts once *ticket == ticket_value, or at once
// if ticket is null), finish (the caller's thread joins in and returns when every job is done).  Single caller.
void host_pool_begin(HostPool *p);
bool host_pool_submit(HostPool *p, const UnpackJob &job, const volatile uint32_t *ticket, uint32_t ticket_value);
bool host_pool_finish(HostPool *p);  // false: a ticket did not arrive within 20 s (the GPU work before it failed)
// trace: steady_clock nanoseconds at which job q of the last step was started / completed
void host_pool_job_times(const HostPool *p, int q, int64_t *first_ns, int64_t *last_ns);
// streaming fill / copy rate of `threads` host threads (GB/s of bytes written), optionally pinned like the pool
void host_memory_probe(int threads, int64_t bytes, const int *cpus, int ncpus, double *fill_gbs, double *copy_gbs);
// begin + submit + finish of one job
void host_pool_unpack(HostPool *p, const UnpackJob &job);

}  // namespace mapf
