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
// record: [obs+mask bits][int8 d_row][int8 d_col][int8 2*reward][u8 blocking_prev]
inline int pack_record_bytes(int V2) { return pack_obs_bytes(V2) + 4; }

struct UnpackJob {
    const uint8_t *packed;  // records of agents [a0, a1), record of agent a at packed + (a - a0) * RS
    int64_t a0, a1;         // global agent indices (env * N + agent)
    int V2, RS;
    uint8_t *obs;           // [BN, V2]   full host arrays (global indexing); may be null
    int8_t *mask;           // [BN, 5]
    float *goal_delta;      // [BN, 2]
    float *reward;          // [BN]
    uint8_t *blocking_prev; // [BN]
    const float *gdt_row, *gdt_col;  // 256-entry tables indexed by (int8 delta + 128)
};

struct HostPool;
HostPool *host_pool_create(int threads);  // threads >= 1 (the caller's thread counts as one)
void host_pool_destroy(HostPool *p);
int host_pool_threads(const HostPool *p);
// blocking: the agents of the job are split over the pool's threads and the calling thread
void host_pool_unpack(HostPool *p, const UnpackJob &job);

}  // namespace mapf
