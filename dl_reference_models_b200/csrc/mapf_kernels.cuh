// mapf_kernels.cuh -- sm_100a kernels of the batched MAPF environment transition.
//
// Reference semantics: src/environments/reference_model_multi_agent.py ("ENV:line").
// Mapping: one env per G-lane sub-warp group (G = 4/8/16/32 >= num_agents), one agent per lane.
//  * move resolution (ENV:502-526) is the reference's sequential agent-index order, executed as
//    an N-iteration shuffle/ballot chain inside the group;
//  * staggered observations (ENV:528-536) use the snapshot identity: agent i sees
//    agent a at new[a] if a <= i else old[a]; a step with a lifelong goal reassignment shows
//    everybody the final state instead (ENV:565-575);
//  * the lock heuristic (ENV:389-438) runs on three per-agent shift registers and ballots;
//  * lifelong goal resampling (ENV:284-304) picks the k-th candidate of a shared-memory
//    bitmap (free & ~occupied & ~goals), k from Philox4x32-10 or from a replay hook;
//  * the wait-for graph (no reference counterpart) is resolved by pointer jumping.
// No tensor cores: there is no dense contraction on this path; the kernels are integer/byte
// work bounded by HBM traffic and instruction issue.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mapf_b200.h"

namespace mapf {

constexpr int PAD = MAPF_MAX_SENSOR_RANGE;  // obstacle padding around the map in map_rows
constexpr uint32_t NOCELL = 0x7FFF7FFFu;    // packed (row, col) that never matches a real cell
constexpr uint32_t LFAR = 0x40000000u;      // linear code of an absent agent (far from any window)

// ------------------------------------------------------------------ Philox4x32-10
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(unsigned long long seed, long long stream) {
        unsigned long long k = seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(stream + 1));
        k0 = (uint32_t)k;
        k1 = (uint32_t)(k >> 32);
    }
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a0 = k0, a1 = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ a0, n1 = lo1, n2 = hi0 ^ c3 ^ a1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a0 += 0x9E3779B9u; a1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// One action draw: Philox keyed by (seed, global env id), counter = (call counter, agent quad).
__device__ __forceinline__ int sample_action(unsigned long long seed, long long env_global, int agent,
                                             unsigned long long counter, int mask_bits /* <0: unmasked */) {
    Philox ph(seed ^ 0xA511E9B3ull, env_global);
    // one Philox block serves the 4 agents of a quad: agent a takes word a & 3 of block a >> 2
    const uint4 x4 = ph((uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)(agent >> 2), 0x41435421u);
    const int w = agent & 3;
    const uint32_t x = w == 0 ? x4.x : w == 1 ? x4.y : w == 2 ? x4.z : x4.w;
    if (mask_bits < 0) return (int)__umulhi(x, 5u);    // scripts/benchmark_multi_agent_env.py:38-39
    const int n = __popc(mask_bits);                   // scripts/benchmark_multi_agent_env.py:42-57
    return n ? (int)__fns((unsigned)mask_bits, 0, (int)__umulhi(x, (uint32_t)n) + 1) : 0;
}

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ int prow(uint32_t p) { return (int)(short)(p & 0xFFFFu); }
__device__ __forceinline__ int pcol(uint32_t p) { return (int)(short)(p >> 16); }
__device__ __forceinline__ uint32_t pack_rc(int r, int c) {
    return ((uint32_t)r & 0xFFFFu) | ((uint32_t)c << 16);
}
// linear code with a 1024-wide row pitch: windows become 1-D range tests
__device__ __forceinline__ uint32_t lin(uint32_t p) { return (uint32_t)(prow(p) * 1024 + pcol(p)); }

template <int V> struct WinBits { using type = uint32_t; };
template <> struct WinBits<7> { using type = unsigned long long; };

// obstacle/OOB bit of cell (r, c), valid for -PAD <= r < R+PAD, -PAD <= c < C+PAD
__device__ __forceinline__ bool map_blocked(const uint32_t *rows, int wpr, int r, int c) {
    int bit = c + PAD;
    return (rows[(r + PAD) * wpr + (bit >> 5)] >> (bit & 31)) & 1u;
}

// k-th (0-based) set bit of a bitmap in shared memory; -1 if k >= popcount
__device__ __forceinline__ int select_kth(const uint32_t *bm, int words, int k) {
    for (int w = 0; w < words; ++w) {
        uint32_t x = bm[w];
        int c = __popc(x);
        if (k < c) return w * 32 + (int)__fns(x, 0, k + 1);
        k -= c;
    }
    return -1;
}

// Shared memory carve-up (32-bit words).  Per CTA: [shared map rows][shared free bitmap],
// then per group: arrays for the pair loop, scratch bitmap, (per-env map copy), and per warp
// the staging buffers for the byte outputs.
struct SmemLayout {
    int map_rows_off, free_off;    // CTA-wide (shared map)
    int gdt_off, gdt_words;        // CTA-wide goal-delta table: (2R-1) row quotients, then (2C-1) col quotients
    int grp_off, grp_words;        // per group block
    int g_new, g_snap, g_goal, g_int, g_delta, g_scratch, g_map, g_free;  // offsets inside a group block
    int g_rowm, g_colm, g_growm, g_gcolm;   // final-position / goal bucket masks (padded by KB on both sides)
    int g_orow, g_ocol, g_trow, g_tcol;     // old-position / move-target bucket masks (move resolution)
    int g_mask_words;
    int stage_off, stage_words;    // per warp staging (obs + mask), 16-byte aligned
    int total_words;
};

struct KParams {
    int B, N, R, C;
    int steps_per_episode, lifelong, lock_enabled, dw, lw, nearby, min_nb, eps_floor;
    int normalize, deterministic, per_env_maps, auto_reset;
    int observe_only;  // reset kernel: rebuild the observation channels from the current state, write nothing else
    long long env_id_base;
    unsigned long long seed;
    float den0, den1;  // ENV:152-155
    // map tables (device).  Shared map: one copy; per-env maps: B copies, strides below.
    const uint32_t *map_rows;   // [(R+2*PAD) * wpr] bit (c+PAD) of padded row (r+PAD) = obstacle/OOB
    const uint32_t *free_bits;  // [fw] bit (r*C+c) = free cell
    const int32_t *num_free;    // [1] or [B]
    const uint32_t *env_tables; // CTA-wide table image of the env-per-thread kernel (EnvLayout, shared map only)
    int wpr, map_words, fw;
    // state
    uint32_t *positions, *goals, *starts;  // int16 pairs viewed as u32: row | col << 16
    uint8_t *agent_flags;
    uint32_t *lock_gp, *lock_mv, *lock_fm;
    int16_t *lock_dist;
    int4 *env_words;  // [B, 4] int4 = 16 words
    double *env_metrics;
    // step / reset inputs
    const int8_t *actions;
    const uint32_t *goal_override;  // int16 pairs
    const int32_t *goal_rank;
    const uint8_t *reset_mask;
    const uint32_t *starts_override, *goals_override;
    // outputs
    uint8_t *o_local_obs;
    int8_t *o_action_mask;
    float2 *o_goal_delta;
    uint8_t *o_blocking_prev;
    float *o_reward;
    uint8_t *o_terminated, *o_truncated, *o_step_flags, *o_agent_step_flags;
    int4 *o_info;  // [B, 4] int4 = MAPF_INFO_WORDS int32
    int8_t *o_next_actions;  // fused benchmark sampler (scripts/benchmark_multi_agent_env.py:38-57)
    int sample_mode;         // 0 off, 1 uniform over the new action mask, 2 uniform over 0..4
    int env_prefetch;        // env-per-thread step kernel: L2 prefetch of the first tile's state ahead of griddepcontrol.wait
    int inner_steps;         // lane-per-agent step kernel: env steps per launch (mapf_step_many), >= 1
    long long out_step_stride;   // ... and how many envs further each of them writes its outputs
    unsigned long long sample_counter;
    uint32_t *err_bits;
    SmemLayout L;  // computed once on the host (mapf_create)
};

// bucket-mask padding: interactions reach SR + 1 cells (snapshot positions differ from final ones by <= 1)
__host__ __device__ constexpr int bucket_pad(int sr) { return sr + 1; }

__host__ __device__ inline SmemLayout make_layout(int G, int V2, int N, int wpr, int R, int C, int SR, int fw,
                                                  int per_env_maps, int threads) {
    SmemLayout L;
    int map_words = (R + 2 * PAD) * wpr;
    int o = 0;
    L.map_rows_off = o; o += per_env_maps ? 0 : map_words;
    L.free_off = o; o += per_env_maps ? 0 : fw;
    L.gdt_off = o; L.gdt_words = (2 * R - 1) + (2 * C - 1); o += L.gdt_words;
    o = (o + 3) & ~3;
    int g = 0;
    int GA = G;  // arrays padded to G entries (G is a multiple of 4); the kernels rely on these five coming first
    L.g_new = g; g += GA;
    L.g_snap = g; g += GA;
    L.g_goal = g; g += GA;
    L.g_int = g; g += GA;
    L.g_delta = g; g += GA;
    const int KB = bucket_pad(SR);
    // maps up to 32 x 32 get the same mask size whatever their shape: the kernels' compile-time modes address the
    // eight masks (and the scratch bitmap behind them) with immediates
    const int small_map = R <= 32 && C <= 32;
    const int mrw = (small_map ? 32 : R) + 2 * KB, mcw = (small_map ? 32 : C) + 2 * KB;
    L.g_rowm = g; g += mrw;
    L.g_colm = g; g += mcw;
    L.g_growm = g; g += mrw;
    L.g_gcolm = g; g += mcw;
    L.g_orow = g; g += mrw;
    L.g_ocol = g; g += mcw;
    L.g_trow = g; g += mrw;
    L.g_tcol = g; g += mcw;
    g = (g + 3) & ~3;
    L.g_mask_words = g - L.g_rowm;
    L.g_scratch = g; g += fw;
    L.g_map = g; g += per_env_maps ? map_words : 0;
    L.g_free = g; g += per_env_maps ? fw : 0;
    g = (g + 3) & ~3;
    L.grp_words = g;
    L.grp_off = o; o += g * (threads / G);
    int envs_per_warp = 32 / G;
    int obs_bytes = envs_per_warp * N * V2 + 32;   // +32: alignment slack (16 head + 16 tail)
    int mask_bytes = envs_per_warp * N * 5 + 32;
    int sw = ((obs_bytes + 15) / 16) * 4 + ((mask_bytes + 15) / 16) * 4;
    L.stage_words = sw;
    L.stage_off = o; o += sw * (threads / 32);
    L.total_words = o;
    return L;
}

// Copy `nbytes` staged bytes to global memory with 16-byte stores where possible.  The staged
// bytes start at smem + (dst & 15) so that source and destination share their 16-byte phase.
__device__ __forceinline__ void warp_copy_out(uint8_t *dst, const uint8_t *stage, int nbytes, int lane) {
    uintptr_t d = (uintptr_t)dst;
    int phase = (int)(d & 15);
    if (phase == 0 && (nbytes & 15) == 0) {  // the common case: whole 16-byte chunks, aligned
        const uint4 *s4 = reinterpret_cast<const uint4 *>(stage) + lane;
        uint4 *d4 = reinterpret_cast<uint4 *>(dst) + lane;
        const int n16 = nbytes >> 4;
        if (lane < n16) d4[0] = s4[0];            // up to 1 KB without a loop (C3: 800 B of windows per warp)
        if (lane + 32 < n16) d4[32] = s4[32];
        for (int i = lane + 64; i < n16; i += 32) d4[i - lane] = s4[i - lane];
        return;
    }
    const uint8_t *src = stage + phase;  // src[i] <-> dst[i]
    int head = phase ? 16 - phase : 0;
    if (head > nbytes) head = nbytes;
    if (lane < head) dst[lane] = src[lane];
    int body = (nbytes - head) & ~15;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    for (int i = lane; i < (body >> 4); i += 32) d4[i] = s4[i];
    int tail = nbytes - head - body;
    if (lane < tail) dst[head + body + lane] = src[head + body + lane];
}

// ------------------------------------------------------------------ the pair loop
// Per lane (= agent i of one env) scan all agents a of the env from shared memory and build
//   * the egocentric window (ENV:707-747) and the action mask (ENV:749-773),
//   * (FULL) lock neighbours (ENV:389-398), intent blocking (ENV:609-623), co-location count
//     (ENV:659-666) and the wait-for pointer.
template <int G, int SR, bool FULL>
struct PairOut {
    typename WinBits<2 * SR + 1>::type occ, xgoal;
    uint32_t nbmask;
    int nbcount, red, colocated, wf_next;
    bool intent_hit;
};

template <int G, int SR, bool FULL>
__device__ __forceinline__ void pair_loop(const uint32_t *s_new, const uint32_t *s_snap,
                                          const uint32_t *s_goal, const uint32_t *s_int,
                                          const uint32_t *s_delta, int gl, uint32_t my_new_lin,
                                          uint32_t my_new_packed, uint32_t my_intended, bool failed,
                                          int nearby, PairOut<G, SR, FULL> &o) {
    constexpr int V = 2 * SR + 1;
    using WB = typename WinBits<V>::type;
    o.occ = 0; o.xgoal = 0; o.nbmask = 0; o.nbcount = 0; o.red = 0; o.colocated = 0;
    o.wf_next = -1; o.intent_hit = false;
    const uint32_t origin = my_new_lin - (uint32_t)(SR * 1024 + SR);
    const uint32_t nb_origin = my_new_lin - (uint32_t)(nearby * 1024 + nearby);
    const uint32_t nb_side = 2u * (uint32_t)nearby;
#pragma unroll
    for (int a4 = 0; a4 < G; a4 += 4) {
        const uint4 vn = *reinterpret_cast<const uint4 *>(s_new + a4);
        const uint4 vs = *reinterpret_cast<const uint4 *>(s_snap + a4);
        const uint4 vg = *reinterpret_cast<const uint4 *>(s_goal + a4);
        uint4 vi = make_uint4(0, 0, 0, 0), vd = make_uint4(0, 0, 0, 0);
        if (FULL) {
            vi = *reinterpret_cast<const uint4 *>(s_int + a4);
            vd = *reinterpret_cast<const uint4 *>(s_delta + a4);
        }
        const uint32_t an[4] = {vn.x, vn.y, vn.z, vn.w};
        const uint32_t as[4] = {vs.x, vs.y, vs.z, vs.w};
        const uint32_t ag[4] = {vg.x, vg.y, vg.z, vg.w};
        const uint32_t ai[4] = {vi.x, vi.y, vi.z, vi.w};
        const uint32_t ad[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int a = a4 + k;
            const bool other = (a != gl);
            // snapshot position of agent a as seen by me (F3): new if a <= me, else old/snap
            const uint32_t ps = (a <= gl) ? an[k] : as[k];
            uint32_t d = ps - origin;
            uint32_t t1 = d & 1023u, t2 = d >> 10;
            // a lower-index agent on MY cell (injected states only, ENV:658-666) does not own it (ENV:200-205): skip it
            if (other && t1 < (uint32_t)V && t2 < (uint32_t)V && !(a < gl && ps == my_new_lin)) o.occ |= (WB)1 << (t2 * V + t1);
            d = ag[k] - origin;
            t1 = d & 1023u; t2 = d >> 10;
            if (other && t1 < (uint32_t)V && t2 < (uint32_t)V) o.xgoal |= (WB)1 << (t2 * V + t1);
            if (FULL) {
                d = an[k] - nb_origin;
                t1 = d & 1023u; t2 = d >> 10;
                if (other && t1 <= nb_side && t2 <= nb_side) {
                    int m = abs((int)t1 - nearby) + abs((int)t2 - nearby);
                    if (m <= nearby) {
                        if (m > 0) { o.nbmask |= 1u << a; o.nbcount++; o.red += (int)ad[k]; }
                        else o.colocated++;
                    }
                }
                if (other && ai[k] == my_new_lin) o.intent_hit = true;
                if (other && failed && an[k] == my_intended) o.wf_next = a;
            }
        }
    }
    (void)my_new_packed;
}

// Bucket masks: rowm[r + KB] / colm[c + KB] hold one bit per agent whose (final) row / column is r / c.
// OR-ing 2K+1 consecutive entries of each and AND-ing the two gives every agent inside the
// (2K+1)^2 box around a cell -- the only agents a lane can interact with -- in O(K) instead of O(N).
template <int K, int KB>
__device__ __forceinline__ uint32_t gather_box(const uint32_t *rowm, const uint32_t *colm, int r, int c) {
    uint32_t mr = 0, mc = 0;
#pragma unroll
    for (int d = 0; d <= 2 * K; ++d) {
        mr |= rowm[r + (KB - K) + d];
        mc |= colm[c + (KB - K) + d];
    }
    return mr & mc;
}

// Same outputs as pair_loop<G, SR, true>, visiting only the candidate agents.
template <int G, int SR>
__device__ __forceinline__ void scan_candidates(const uint32_t *s_new, const uint32_t *s_snap,
                                                const uint32_t *s_goal, const uint32_t *s_int,
                                                const uint32_t *s_delta, int gl, uint32_t cand, uint32_t gcand,
                                                uint32_t my_new_lin, uint32_t my_intended, bool failed,
                                                int nearby, PairOut<G, SR, true> &o) {
    constexpr int V = 2 * SR + 1;
    using WB = typename WinBits<V>::type;
    o.occ = 0; o.xgoal = 0; o.nbmask = 0; o.nbcount = 0; o.red = 0; o.colocated = 0;
    o.wf_next = -1; o.intent_hit = false;
    const uint32_t origin = my_new_lin - (uint32_t)(SR * 1024 + SR);
    const uint32_t nb_origin = my_new_lin - (uint32_t)(nearby * 1024 + nearby);
    const uint32_t nb_side = 2u * (uint32_t)nearby;
    while (cand) {
        const int a = __ffs(cand) - 1;
        cand &= cand - 1;
        const uint32_t an = s_new[a];
        const uint32_t ps = (a <= gl) ? an : s_snap[a];  // snapshot position of agent a as seen by me (F3)
        uint32_t d = ps - origin;
        uint32_t t1 = d & 1023u, t2 = d >> 10;
        // a lower-index agent on MY cell (injected states only, ENV:658-666) does not own it (ENV:200-205): skip it
        if (t1 < (uint32_t)V && t2 < (uint32_t)V && !(a < gl && ps == my_new_lin)) o.occ |= (WB)1 << (t2 * V + t1);
        d = an - nb_origin;
        t1 = d & 1023u; t2 = d >> 10;
        if (t1 <= nb_side && t2 <= nb_side) {
            const int m = abs((int)t1 - nearby) + abs((int)t2 - nearby);
            if (m <= nearby) {
                if (m > 0) { o.nbmask |= 1u << a; o.nbcount++; o.red += (int)s_delta[a]; }
                else o.colocated++;
            }
        }
        if (s_int[a] == my_new_lin) o.intent_hit = true;
        if (failed && an == my_intended) o.wf_next = a;
    }
    while (gcand) {
        const int a = __ffs(gcand) - 1;
        gcand &= gcand - 1;
        const uint32_t d = s_goal[a] - origin;
        const uint32_t t1 = d & 1023u, t2 = d >> 10;
        if (t1 < (uint32_t)V && t2 < (uint32_t)V) o.xgoal |= (WB)1 << (t2 * V + t1);
    }
}

// Window value bytes (ENV:730-745 priority) + action mask (ENV:761-771) into the staging area.
template <int SR>
__device__ __forceinline__ uint32_t emit_window(const uint32_t *rows, int wpr, int r, int c,
                                                typename WinBits<2 * SR + 1>::type occ,
                                                typename WinBits<2 * SR + 1>::type xgoal,
                                                uint32_t goal_lin, uint32_t my_lin,
                                                uint8_t *stage_obs /* may be null */) {
    constexpr int V = 2 * SR + 1;
    using WB = typename WinBits<V>::type;
    WB obst = 0;
    const int sb = c - SR + PAD, word = sb >> 5, sh = sb & 31;
#pragma unroll
    for (int wr = 0; wr < V; ++wr) {
        const uint32_t *rw = rows + (r - SR + wr + PAD) * wpr + word;
        uint32_t bits = __funnelshift_r(rw[0], rw[1], sh) & ((1u << V) - 1u);
        obst |= (WB)bits << (wr * V);
    }
    WB own = 0;
    {
        uint32_t d = goal_lin - (my_lin - (uint32_t)(SR * 1024 + SR));
        uint32_t t1 = d & 1023u, t2 = d >> 10;
        if (t1 < (uint32_t)V && t2 < (uint32_t)V) own = (WB)1 << (t2 * V + t1);
    }
    const WB agent = occ & ~obst;
    const WB blocked = obst | occ;
    const WB g3 = own & ~blocked;
    const WB g4 = xgoal & ~blocked & ~own;
    if (stage_obs) {
        // cell code = b0 + 2*b1 + 4*b2 with b0 = obstacle|own goal, b1 = agent|own goal, b2 = other goal
        // (1 = 001, 2 = 010, 3 = 011, 4 = 100).  Four cells at a time: a nibble times 0x00204081 puts
        // bit j of the nibble at bit 8*j, so three multiplies build four output bytes.
        const WB b0 = obst | g3, b1 = agent | g3, b2 = g4;
#pragma unroll
        for (int k = 0; k < V * V; k += 4) {
            const uint32_t n0 = (uint32_t)(b0 >> k) & 15u, n1 = (uint32_t)(b1 >> k) & 15u,
                           n2 = (uint32_t)(b2 >> k) & 15u;
            const uint32_t w = ((n0 * 0x00204081u) & 0x01010101u) | (((n1 * 0x00204081u) & 0x01010101u) << 1) |
                               (((n2 * 0x00204081u) & 0x01010101u) << 2);
            stage_obs[k] = (uint8_t)w;
            if (k + 1 < V * V) stage_obs[k + 1] = (uint8_t)(w >> 8);
            if (k + 2 < V * V) stage_obs[k + 2] = (uint8_t)(w >> 16);
            if (k + 3 < V * V) stage_obs[k + 3] = (uint8_t)(w >> 24);
        }
    }
    constexpr int ctr = SR * V + SR;
    uint32_t m = 1u;
    m |= (uint32_t)(((~blocked) >> (ctr - V)) & 1) << 1;  // up
    m |= (uint32_t)(((~blocked) >> (ctr + 1)) & 1) << 2;  // right
    m |= (uint32_t)(((~blocked) >> (ctr + V)) & 1) << 3;  // down
    m |= (uint32_t)(((~blocked) >> (ctr - 1)) & 1) << 4;  // left
    return m;
}

// ENV:330-335.  The quotients (goal - pos) / denominator take only 2R-1 / 2C-1 distinct values; the
// CTA tabulates them once with IEEE division (__fdiv_rn), lanes look them up.
__device__ __forceinline__ void fill_goal_delta_table(float *gdt, int R, int C, int normalize, float den0,
                                                      float den1, int tid, int nthreads) {
    const int nr = 2 * R - 1, n = nr + 2 * C - 1;
    for (int i = tid; i < n; i += nthreads) {
        const bool row = i < nr;
        const float d = (float)(row ? i - (R - 1) : i - nr - (C - 1));
        gdt[i] = normalize ? __fdiv_rn(d, row ? den0 : den1) : d;
    }
}
__device__ __forceinline__ float2 goal_delta(const float *gdt, int R, int C, uint32_t goal, uint32_t pos) {
    const int i0 = prow(goal) - prow(pos) + (R - 1), i1 = pcol(goal) - pcol(pos) + (C - 1) + 2 * R - 1;
    return make_float2(gdt[i0], gdt[i1]);
}

// Draw 2N distinct free cells (ENV:267-282) by symmetric rejection: every slot draws uniformly;
// a slot that equals a lower-numbered slot redraws.  The rule is invariant under relabelling of
// cells, hence the result is uniform over ordered tuples of distinct cells.
// Slot order: starts 0..N-1, goals N..2N-1.  Returns packed start / goal for this lane.
template <int G>
__device__ __forceinline__ void draw_layout(const Philox &ph, uint32_t &rng_counter, int F,
                                            const uint32_t *free_bm, int fw, int C, int N, int gl,
                                            unsigned gbase, unsigned gmask, bool need,
                                            uint32_t &start, uint32_t &goal) {
    const unsigned full = 0xFFFFFFFFu;
    int cs = -1, cg = -1;       // drawn cell ids
    bool rs = need, rg = need;  // must (re)draw
    uint32_t rounds = 0;        // rounds consumed by THIS env (group-uniform => shard invariant)
    for (;;) {
        const unsigned needmask = __ballot_sync(full, rs || rg);
        if (!needmask) break;
        const bool grp_need = (needmask & gmask) != 0;
        if (rs || rg) {
            uint4 x = ph(rng_counter + rounds, (uint32_t)gl, 0x52455345u /* "RESE" */, 0);
            if (rs) cs = select_kth(free_bm, fw, (int)__umulhi(x.x, (uint32_t)F));
            if (rg) cg = select_kth(free_bm, fw, (int)__umulhi(x.y, (uint32_t)F));
        }
        rs = false; rg = false;
#pragma unroll
        for (int a = 0; a < G; ++a) {
            int os = __shfl_sync(full, cs, gbase + a);
            int og = __shfl_sync(full, cg, gbase + a);
            if (need && a < N) {
                if (a < gl && os == cs) rs = true;   // lower start slot
                if (os == cg) rg = true;             // every start slot is lower than a goal slot
                if (a < gl && og == cg) rg = true;   // lower goal slot
            }
        }
        if (grp_need) rounds++;
    }
    rng_counter += rounds;
    if (need) {
        start = pack_rc(cs / C, cs % C);
        goal = pack_rc(cg / C, cg % C);
    }
}

// ============================================================================ step kernel
#ifndef MAPF_MOVE_CHAIN
#define MAPF_MOVE_CHAIN 0
#endif
#ifndef MAPF_STEP_MIN_CTAS
#define MAPF_STEP_MIN_CTAS 4
#endif
// MODE: 0 = lifelong / lock-metric switches read from the parameters; 1 = lifelong + lock metrics, 2 = episodic +
// lock metrics as compile-time constants (the untaken branches and their uniform tests drop out; same results).
template <int G, int SR, int MODE = 0>
__global__ void __launch_bounds__(256, MAPF_STEP_MIN_CTAS) mapf_step_kernel(const KParams p) {
    const bool kLifelong = MODE == 1 ? true : MODE == 2 ? false : p.lifelong;
    const bool kLock = MODE != 0 ? true : p.lock_enabled;
    constexpr int V = 2 * SR + 1, V2 = V * V;
    using WB = typename WinBits<V>::type;
    extern __shared__ __align__(16) uint32_t smem[];
    const unsigned full = 0xFFFFFFFFu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = tid % G, grp = tid / G;
    const int groups = MODE ? 256 / G : blockDim.x / G;   // MODE != 0 is only chosen for 256-thread CTAs
    const unsigned gbase = (unsigned)(lane / G) * G;
    const unsigned gmask = (G == 32) ? full : (((1u << G) - 1u) << gbase);
    const int N = p.N;
    const uint32_t mybit = 1u << gl, lower = mybit - 1u;

    constexpr int KB = bucket_pad(SR);
    const SmemLayout &L = p.L;
    uint32_t *gsm = smem + L.grp_off + grp * L.grp_words;
    // the five per-agent arrays open the group block at fixed offsets (make_layout: G words each) -> immediates
    uint32_t *s_new = gsm, *s_snap = gsm + G, *s_goal = gsm + 2 * G;
    // MODE != 0 is only chosen for maps up to 32 x 32: fixed mask size MS (make_layout), everything an immediate
    constexpr int MS = 32 + 2 * KB, MASK_WORDS = (8 * MS + 3) & ~3;
    uint32_t *s_int = gsm + 3 * G, *s_delta = gsm + 4 * G;
    uint32_t *s_scratch = MODE ? gsm + 5 * G + MASK_WORDS : gsm + L.g_scratch;
    uint32_t *s_rowm = MODE ? gsm + 5 * G : gsm + L.g_rowm, *s_colm = MODE ? gsm + 5 * G + MS : gsm + L.g_colm;
    uint32_t *s_growm = MODE ? gsm + 5 * G + 2 * MS : gsm + L.g_growm, *s_gcolm = MODE ? gsm + 5 * G + 3 * MS : gsm + L.g_gcolm;
    uint32_t *s_orow = MODE ? gsm + 5 * G + 4 * MS : gsm + L.g_orow, *s_ocol = MODE ? gsm + 5 * G + 5 * MS : gsm + L.g_ocol;
    uint32_t *s_trow = MODE ? gsm + 5 * G + 6 * MS : gsm + L.g_trow, *s_tcol = MODE ? gsm + 5 * G + 7 * MS : gsm + L.g_tcol;
    const int mask_quads = MODE ? (MASK_WORDS >> 2) : (L.g_mask_words >> 2);
    const float *gdt = reinterpret_cast<const float *>(smem + L.gdt_off);
    // CTA-wide tables, loaded once; the CTA then walks over tiles of `groups` envs (persistent grid)
    fill_goal_delta_table(reinterpret_cast<float *>(smem + L.gdt_off), p.R, p.C, p.normalize, p.den0, p.den1,
                          tid, blockDim.x);
    const uint32_t *rows = gsm + L.g_map, *freebm = gsm + L.g_free;
    if (!p.per_env_maps) {
        uint32_t *mr = smem + L.map_rows_off, *fb = smem + L.free_off;
        for (int i = tid; i < p.map_words; i += blockDim.x) mr[i] = p.map_rows[i];
        for (int i = tid; i < p.fw; i += blockDim.x) fb[i] = p.free_bits[i];
        rows = mr; freebm = fb;
    }
    // Programmatic dependent launch (the step is launched with stream serialization relaxed): the tables above are
    // immutable between launches, so this much may run while the previous launch's last CTAs finish; env state is
    // touched only after the previous launch has completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    uint8_t *stage = reinterpret_cast<uint8_t *>(smem + L.stage_off + warp * L.stage_words);
    const int envs_per_warp = 32 / G;
    const int obs_stage_bytes = ((envs_per_warp * N * V2 + 32 + 15) / 16) * 16;
    uint8_t *stage_obs = stage;
    uint8_t *stage_mask = stage + obs_stage_bytes;
    uint32_t errs = 0;
    const int ntiles = (p.B + groups - 1) / groups;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int env = tile * groups + grp;
    const bool env_ok = env < p.B;
    const bool act = env_ok && gl < N;
    // mapf_step_many: `inner_steps` consecutive env steps in ONE launch (small batches are launch-rate bound).  Step
    // it > 0 takes the actions the fused sampler drew at the end of step it - 1 (kept in a register) and writes its
    // outputs `out_step_stride` envs further into the caller's [K, B, ...] buffers (0: every step overwrites).
    int carried_action = 0;
    for (int it = 0; it < p.inner_steps; ++it) {
    const size_t so = (size_t)it * (size_t)p.out_step_stride, aoff = so * (size_t)N;
    {   // clear the group's bucket masks (contiguous, 16-byte aligned, multiple of 4 words)
        uint4 *mz = reinterpret_cast<uint4 *>(s_rowm);
        for (int i = gl; i < mask_quads; i += G) mz[i] = make_uint4(0, 0, 0, 0);
    }
    if (p.per_env_maps && env_ok) {
        uint32_t *mr = gsm + L.g_map, *fb = gsm + L.g_free;
        const uint32_t *src = p.map_rows + (size_t)env * p.map_words;
        for (int i = gl; i < p.map_words; i += G) mr[i] = src[i];
        const uint32_t *fsrc = p.free_bits + (size_t)env * p.fw;
        for (int i = gl; i < p.fw; i += G) fb[i] = fsrc[i];
    }
    __syncwarp();

    // ---------------------------------------------------------------- load
    const size_t ai = (size_t)(env_ok ? env : 0) * N + (gl < N ? gl : 0);
    uint32_t pos = act ? p.positions[ai] : NOCELL;
    uint32_t goal = act ? p.goals[ai] : NOCELL - 1;
    uint32_t aflags = act ? p.agent_flags[ai] : 0u;
    int action = it > 0 ? carried_action : ((act && p.actions) ? (int)p.actions[ai] : 0);
    if (action < 0 || action > 4) { errs |= MAPF_DEV_ERR_INVALID_ACTION; action = 0; }
    int4 w0 = make_int4(0, 0, 0, 0), w1 = w0, w2 = w0, w3 = w0;
    if (env_ok) {
        const int4 *ew = p.env_words + (size_t)env * 4;
        w0 = ew[0]; w1 = ew[1]; w2 = ew[2]; w3 = ew[3];
    }
    int step_count = w0.x + 1;  // ENV:475
    int lock_count = w0.y, lock_prev = w0.z, goals_total = w0.w;
    int blocking_total = w1.x, dl_events = w1.y, ll_events = w1.z, dl_steps = w1.w;
    int ll_steps = w2.x;
    uint32_t rng_counter = (uint32_t)w2.y;
    int ep_return_x2 = w2.z, wfg_steps = w2.w;
    int episodes = w3.x;
    int lock_head = w3.y;  // ring slot the next distance goes to (MAPF_W_LOCK_HEAD)
    // Lock history is fetched here, ahead of the warp barriers of the move resolution, so that its
    // DRAM latency overlaps with them (the compiler cannot hoist loads across __syncwarp).
    uint32_t gpr = 0, mvr = 0, fmr = 0;
    int ring_old = 0;
    const int count_after = lock_count + 1;
    if (lock_head < 0 || lock_head >= p.lw) lock_head = 0;
    const int slot_new = lock_head;
    const int slot_next = (lock_head + 1 == p.lw) ? 0 : lock_head + 1;
    if (kLock && act) {
        gpr = p.lock_gp[ai]; mvr = p.lock_mv[ai]; fmr = p.lock_fm[ai];
        if (count_after >= p.lw && p.lw > 1)  // slot_next still holds the oldest row of the window, ENV:432
            ring_old = (int)p.lock_dist[((size_t)env * p.lw + slot_next) * N + gl];
    }

    // ---------------------------------------------------------------- move resolution, ENV:502-526
    const int r0 = prow(pos), c0 = pcol(pos);
    const int nr = r0 + (action == 3) - (action == 1);
    const int nc = c0 + (action == 2) - (action == 4);
    const uint32_t intended = pack_rc(nr, nc);  // ENV:514-515 (kept even when invalid)
    bool wants = act && action != 0 && !map_blocked(rows, p.wpr, nr, nc);
#if MAPF_MOVE_CHAIN
    // The reference's sequential agent loop as an N-step shuffle/ballot chain: at agent i's turn its
    // target is held iff some other lane's *current* cell equals it.
    const uint32_t target = wants ? intended : NOCELL;
    uint32_t cur = pos;
#pragma unroll
    for (int i = 0; i < G; ++i) {
        if (i >= N) break;
        const uint32_t t = __shfl_sync(full, target, gbase + i);
        const bool hit = (gl != i) && (cur == t);  // occupied by somebody else at agent i's turn
        const unsigned blocked = __ballot_sync(full, hit) & gmask;
        if (gl == i && t != NOCELL && !blocked) cur = t;
    }
    const uint32_t newpos = cur;
#else
    // Sequential semantics without the N-step chain: agent i is blocked iff its target cell is held
    // at its turn, i.e. by a higher-index agent that has not moved yet, by a lower-index agent that
    // stayed, or by a lower-index agent that moved in first (same target).  Occupants and co-claimants
    // of a cell come from row/column bucket masks (their AND is the exact cell); dependencies point to
    // lower indices only, so the fix-point below settles in (longest follow-chain) rounds.
    if (act) {
        atomicOr(&s_orow[r0 + KB], mybit);
        atomicOr(&s_ocol[c0 + KB], mybit);
    }
    if (wants) {
        atomicOr(&s_trow[nr + KB], mybit);
        atomicOr(&s_tcol[nc + KB], mybit);
    }
    __syncwarp();
    uint32_t occ_lo = 0, dup_lo = 0;
    bool blocked_hi = false;
    if (wants) {
        const uint32_t occm = s_orow[nr + KB] & s_ocol[nc + KB] & ~mybit;
        blocked_hi = (occm & ~lower) != 0;
        occ_lo = occm & lower;
        dup_lo = s_trow[nr + KB] & s_tcol[nc + KB] & lower;
    }
    const uint32_t deps = occ_lo | dup_lo;
    bool known = !wants || blocked_hi || deps == 0;
    bool moves = wants && !blocked_hi && deps == 0;
    for (;;) {
        const unsigned kb = __ballot_sync(full, known);
        if (kb == full) break;
        const unsigned Kn = (kb & gmask) >> gbase;
        const unsigned Mv = (__ballot_sync(full, moves) & gmask) >> gbase;
        if (!known) {
            const bool blk = ((occ_lo & Kn & ~Mv) | (dup_lo & Mv)) != 0;  // a settled dependency already blocks
            if (blk || (deps & ~Kn) == 0) { known = true; moves = !blk; }
        }
    }
    const uint32_t newpos = moves ? intended : pos;
#endif
    const bool moved = act && (newpos != pos);
    const bool failed = act && action != 0 && !moved;  // ENV:583

    // ---------------------------------------------------------------- goals, ENV:538-563
    bool on_goal = act && newpos == goal;
    int reward_x2 = 0;
    bool gstep = false;
    uint32_t goal_new = goal;
    unsigned arrivals = 0;
    if (!kLifelong) {
        if (on_goal && !(aflags & MAPF_AF_REACHED)) {
            aflags |= MAPF_AF_REACHED | MAPF_AF_COMPLETED_ONCE;
            reward_x2 += 1; gstep = true;
        }
    } else {
        if (on_goal) {
            reward_x2 += 1; gstep = true;
            aflags |= MAPF_AF_COMPLETED_ONCE;
            aflags &= ~MAPF_AF_REACHED;
        }
        arrivals = __ballot_sync(full, on_goal) & gmask;
        unsigned pending = arrivals;
        const Philox ph(p.seed, p.env_id_base + env);
        while (__any_sync(full, pending != 0)) {  // ENV:284-304, in agent-index order
            const int i = pending ? (__ffs(pending) - 1 - (int)gbase) : -1;
            if (pending) for (int w = gl; w < p.fw; w += G) s_scratch[w] = freebm[w];
            __syncwarp();
            if (pending && act) {
                const uint32_t oc = (gl <= i) ? newpos : pos;  // occupancy snapshot of agent i
                int cell = prow(oc) * p.C + pcol(oc);
                atomicAnd(&s_scratch[cell >> 5], ~(1u << (cell & 31)));
                cell = prow(goal_new) * p.C + pcol(goal_new);
                atomicAnd(&s_scratch[cell >> 5], ~(1u << (cell & 31)));
            }
            __syncwarp();
            if (pending) {
                uint32_t ng = NOCELL;
                const size_t oi = (size_t)env * N + i;
                if (p.goal_override) {
                    uint32_t ov = p.goal_override[oi];
                    if (prow(ov) >= 0) ng = ov;
                }
                if (ng == NOCELL) {
                    int n = 0;
                    for (int w = 0; w < p.fw; ++w) n += __popc(s_scratch[w]);
                    int k = -1;
                    if (p.goal_rank) k = p.goal_rank[oi];
                    if (k < 0 && n > 0) {
                        uint4 x = ph(rng_counter, (uint32_t)i, 0x474F414Cu /* "GOAL" */, 0);
                        k = (int)__umulhi(x.x, (uint32_t)n);
                        rng_counter++;
                    }
                    int cell = (n > 0 && k < n) ? select_kth(s_scratch, p.fw, k) : -1;
                    if (cell >= 0) ng = pack_rc(cell / p.C, cell % p.C);
                    else errs |= MAPF_DEV_ERR_NO_GOAL_CELL;
                }
                if (gl == i && ng != NOCELL) goal_new = ng;
            }
            pending &= pending - 1;
            __syncwarp();
        }
        goals_total += __popc(arrivals);
    }
    if (!kLifelong) goals_total += __popc(__ballot_sync(full, gstep) & gmask);
    const bool reassigned = arrivals != 0;
    // ENV:555: an arrived lifelong agent is no longer "on goal"; others keep their test
    bool cur_on_goal = act && newpos == goal_new;
    const bool reached_goal_scratch = kLifelong ? false : cur_on_goal;

    // ---------------------------------------------------------------- lock bookkeeping, ENV:581-594
    int dist_now = abs(prow(goal_new) - prow(newpos)) + abs(pcol(goal_new) - pcol(newpos));
    int delta = 0;
    if (kLock) {
        const bool prev_on_goal = kLifelong ? false : (pos == goal_new);
        const bool gp = kLifelong ? gstep : (!prev_on_goal && cur_on_goal);
        gpr = (gpr << 1) | (gp ? 1u : 0u);
        mvr = (mvr << 1) | (moved ? 1u : 0u);
        fmr = (fmr << 1) | (failed ? 1u : 0u);
        if (act) {
            if (count_after >= p.lw && p.lw > 1) delta = ring_old - dist_now;
            p.lock_dist[((size_t)env * p.lw + slot_new) * N + gl] = (int16_t)dist_now;
        }
        lock_head = slot_next;
    }

    // ---------------------------------------------------------------- interaction scan
    const uint32_t my_lin = act ? lin(newpos) : LFAR;
    s_new[gl] = my_lin;
    s_snap[gl] = act ? lin(reassigned ? newpos : pos) : LFAR;
    const uint32_t goal_for_obs = reassigned ? goal_new : goal;  // ENV:565-575 vs in-loop obs
    s_goal[gl] = act ? lin(goal_for_obs) : LFAR;
    // intent of agents that have not (sticky-)reached, ENV:619-621
    s_int[gl] = (act && !(aflags & MAPF_AF_REACHED)) ? lin(intended) + 0u : LFAR + 1u;
    s_delta[gl] = (uint32_t)delta;
    if (act) {
        atomicOr(&s_rowm[prow(newpos) + KB], 1u << gl);
        atomicOr(&s_colm[pcol(newpos) + KB], 1u << gl);
        atomicOr(&s_growm[prow(goal_for_obs) + KB], 1u << gl);
        atomicOr(&s_gcolm[pcol(goal_for_obs) + KB], 1u << gl);
    }
    __syncwarp();
    PairOut<G, SR, true> po;
    {
        uint32_t cand = 0, gcand = 0;
        if (act) {
            const uint32_t others = ((N >= 32) ? 0xFFFFFFFFu : ((1u << N) - 1u)) & ~(1u << gl);
            // every interaction lies within SR+1 cells of my final position unless the lock
            // neighbourhood is wider than that: then fall back to scanning all agents
            cand = (p.nearby <= KB) ? (gather_box<KB, KB>(s_rowm, s_colm, prow(newpos), pcol(newpos)) & others) : others;
            gcand = gather_box<SR, KB>(s_growm, s_gcolm, prow(newpos), pcol(newpos)) & others;
        }
        scan_candidates<G, SR>(s_new, s_snap, s_goal, s_int, s_delta, gl, cand, gcand, my_lin,
                               failed ? lin(intended) : LFAR + 2u, failed, p.nearby, po);
    }

    // ---------------------------------------------------------------- observation channels
    const int my_env_in_warp = lane / G;
    uint8_t *my_stage_obs = nullptr;
    const size_t warp_env0 = (size_t)tile * groups + (size_t)(warp * envs_per_warp);
    const size_t warp_env0_out = warp_env0 + so;   // where this step's rows go in the output buffers
    if (p.o_local_obs) {
        uint8_t *dst0 = p.o_local_obs + warp_env0_out * N * V2;
        my_stage_obs = stage_obs + ((uintptr_t)dst0 & 15) + (size_t)(my_env_in_warp * N + gl) * V2;
    }
    uint32_t amask = 1u;
    if (act)
        amask = emit_window<SR>(rows, p.wpr, prow(newpos), pcol(newpos), po.occ, po.xgoal,
                                lin(goal_for_obs), my_lin, my_stage_obs);
    float2 gd = act ? goal_delta(gdt, p.R, p.C, goal_for_obs, newpos) : make_float2(0.f, 0.f);
    const uint32_t bp_prev_out = (aflags >> 2) & 1u;  // value of the previous step, ENV:322 (F5)

    // ---------------------------------------------------------------- lock detection, ENV:400-438,595-606
    bool dl_step = false, ll_step = false, dl_event = false, ll_event = false;
    if (kLock) {
        const uint32_t mdw = p.dw >= 32 ? full : ((1u << p.dw) - 1u);
        const uint32_t mlw = p.lw >= 32 ? full : ((1u << p.lw) - 1u);
        const unsigned Gd = (__ballot_sync(full, (gpr & mdw) != 0) & gmask) >> gbase;
        const unsigned Md = (__ballot_sync(full, (mvr & mdw) != 0) & gmask) >> gbase;
        const unsigned Fd = (__ballot_sync(full, (fmr & mdw) != 0) & gmask) >> gbase;
        const unsigned Gl = (__ballot_sync(full, (gpr & mlw) != 0) & gmask) >> gbase;
        const unsigned Ml = (__ballot_sync(full, (mvr & mlw) != 0) & gmask) >> gbase;
        const bool focal = act && !cur_on_goal && po.nbcount >= p.min_nb;  // ENV:391-396
        const unsigned P = po.nbmask | (1u << gl);
        const bool dl_f = focal && !(P & Gd) && !(P & Md) && (P & Fd);
        const bool ll_f = focal && !(P & Gl) && (P & Ml) && (po.red + delta <= p.eps_floor);
        const bool dl_any = (__ballot_sync(full, dl_f) & gmask) != 0;
        const bool ll_any = (__ballot_sync(full, ll_f) & gmask) != 0;
        dl_step = count_after >= p.dw && dl_any;
        ll_step = !dl_step && count_after >= p.lw && ll_any;
        dl_event = dl_step && !(lock_prev & 1);
        ll_event = ll_step && !(lock_prev & 2);
        lock_prev = (dl_step ? 1 : 0) | (ll_step ? 2 : 0);
        dl_steps += dl_step; ll_steps += ll_step; dl_events += dl_event; ll_events += ll_event;
        lock_count = count_after;
    }

    // ---------------------------------------------------------------- blocking, ENV:608-625
    const bool blocking = act && (aflags & MAPF_AF_REACHED) && !moved && po.intent_hit;
    aflags = (aflags & ~MAPF_AF_BLOCKING_PREV) | (blocking ? MAPF_AF_BLOCKING_PREV : 0u);
    blocking_total += __popc(__ballot_sync(full, blocking) & gmask);

    // ---------------------------------------------------------------- wait-for graph (pointer jumping)
    bool wf_cycle = false;
    if (__any_sync(full, po.wf_next >= 0)) {   // no edge in the whole warp (the common case with masked actions): no cycle
        int ptr = po.wf_next;  // -1: no outgoing edge
        unsigned reach = ptr >= 0 ? (1u << ptr) : 0u;
#pragma unroll
        for (int s = 1; s < G; s <<= 1) {
            const int src = gbase + (ptr >= 0 ? ptr : gl);
            const unsigned r2 = __shfl_sync(full, reach, src);
            const int p2 = __shfl_sync(full, ptr, src);
            if (ptr >= 0) { reach |= r2; ptr = p2; }
        }
        wf_cycle = act && ((reach >> gl) & 1u);
    }
    const bool wf_any = (__ballot_sync(full, wf_cycle) & gmask) != 0;
    wfg_steps += wf_any;

    // ---------------------------------------------------------------- rewards & termination, ENV:658-690
    reward_x2 -= 2 * po.colocated;
    const int n_on = __popc(__ballot_sync(full, act && reached_goal_scratch) & gmask);
    bool terminated = false, truncated = false;
    if (!kLifelong && n_on == N) {
        reward_x2 += 2; terminated = true;
    } else if (step_count >= p.steps_per_episode) {
        if (!kLifelong && !reached_goal_scratch) reward_x2 -= 2;
        terminated = true; truncated = true;  // F6
    }
    if (!act) reward_x2 = 0;
    int rsum = reward_x2;
#pragma unroll
    for (int s = G / 2; s > 0; s >>= 1) rsum += __shfl_xor_sync(full, rsum, s);
    ep_return_x2 += rsum;
    const bool done = env_ok && (terminated || truncated);

    // ---------------------------------------------------------------- per-step outputs
    if (act) {
        if (p.o_goal_delta) p.o_goal_delta[ai + aoff] = gd;
        if (p.o_blocking_prev) p.o_blocking_prev[ai + aoff] = (uint8_t)bp_prev_out;
        if (p.o_reward) p.o_reward[ai + aoff] = 0.5f * (float)reward_x2;
        if (p.o_agent_step_flags)
            p.o_agent_step_flags[ai + aoff] = (uint8_t)((moved ? MAPF_ASF_MOVED : 0) | (failed ? MAPF_ASF_FAILED_MOVE : 0) |
                                                 (gstep ? MAPF_ASF_GOAL_REACHED : 0) | (blocking ? MAPF_ASF_BLOCKING : 0) |
                                                 (wf_cycle ? MAPF_ASF_WFG_CYCLE : 0) | (cur_on_goal ? MAPF_ASF_ON_GOAL : 0));
    }
    const unsigned comp = __ballot_sync(full, act && (aflags & MAPF_AF_COMPLETED_ONCE)) & gmask;
    const unsigned reach = __ballot_sync(full, act && (aflags & MAPF_AF_REACHED)) & gmask;
    const int goals_step = __popc(__ballot_sync(full, gstep) & gmask);
    const int blocking_step = __popc(__ballot_sync(full, blocking) & gmask);
    if (env_ok && gl == 0) {
        if (p.o_info) {  // integer sources of info["__all__"], ENV:639-656
            int4 *io = p.o_info + ((size_t)env + so) * 4;
            io[0] = make_int4(goals_step, kLifelong ? goals_total : __popc(reach), blocking_step, blocking_total);
            io[1] = make_int4(dl_step, ll_step, dl_event, ll_event);
            io[2] = make_int4(dl_events, ll_events, dl_steps, ll_steps);
            io[3] = make_int4(__popc(comp), step_count, __popc(reach), wfg_steps);
        }
        if (p.o_terminated) p.o_terminated[env + so] = terminated;
        if (p.o_truncated) p.o_truncated[env + so] = truncated;
        if (p.o_step_flags)
            p.o_step_flags[env + so] = (uint8_t)((terminated ? MAPF_SF_TERMINATED : 0) | (truncated ? MAPF_SF_TRUNCATED : 0) |
                                            (dl_step ? MAPF_SF_DEADLOCK_STEP : 0) | (ll_step ? MAPF_SF_LIVELOCK_STEP : 0) |
                                            (dl_event ? MAPF_SF_DEADLOCK_EVENT : 0) | (ll_event ? MAPF_SF_LIVELOCK_EVENT : 0) |
                                            (reassigned ? MAPF_SF_GOAL_REASSIGNED : 0) | (wf_any ? MAPF_SF_WFG_CYCLE : 0));
    }

    // ---------------------------------------------------------------- episode end: metrics, auto-reset
    uint32_t out_pos = newpos, out_goal = goal_new, out_start = 0;
    bool write_start = false;
    // episode-end metric sums, src/trainers/callbacks.py:152,173,335-345
    {
        if (done && gl == 0) {
            double *m = p.env_metrics + (size_t)env * MAPF_METRIC_COUNT;
            const double gt = kLifelong ? (double)goals_total : (double)__popc(reach);  // ENV:630-633
            m[MAPF_M_EPISODES] += 1.0;
            m[MAPF_M_RETURN_SUM] += 0.5 * (double)ep_return_x2;
            m[MAPF_M_LENGTH_SUM] += (double)step_count;
            m[MAPF_M_SUCCESS_SUM] += (terminated && !truncated) ? 1.0 : 0.0;
            m[MAPF_M_GOALS_REACHED_SUM] += gt;
            m[MAPF_M_BLOCKING_COUNT_SUM] += (double)blocking_total;
            m[MAPF_M_DEADLOCK_COUNT_SUM] += (double)dl_events;
            m[MAPF_M_LIVELOCK_COUNT_SUM] += (double)ll_events;
            m[MAPF_M_DEADLOCK_STEPS_SUM] += (double)dl_steps;
            m[MAPF_M_LIVELOCK_STEPS_SUM] += (double)ll_steps;
            m[MAPF_M_THROUGHPUT_SUM] += gt / (double)(step_count > 1 ? step_count : 1);  // ENV:655
            m[MAPF_M_COMPLETION_RATIO_SUM] += (double)__popc(comp) / (double)N;           // ENV:638
            m[MAPF_M_WFG_CYCLE_STEPS_SUM] += (double)wfg_steps;
        }
    }
    if (done) episodes += 1;

    const bool do_reset = done && p.auto_reset;
    __syncwarp();
    if (__any_sync(full, do_reset)) {  // ENV:440-472 inside the launch (benchmark loop semantics)
        uint32_t st = 0, gg = 0;
        bool sample = do_reset && !p.deterministic;
        const int F = p.num_free[p.per_env_maps ? (env_ok ? env : 0) : 0];
        if (sample && F < 2 * N) { errs |= MAPF_DEV_ERR_TOO_FEW_CELLS; sample = false; }
        const Philox ph(p.seed, p.env_id_base + env);
        draw_layout<G>(ph, rng_counter, F, freebm, p.fw, p.C, N, gl, gbase, gmask, sample && act, st, gg);
        if (do_reset) {
            if (p.deterministic) { st = act ? p.starts[ai] : NOCELL; gg = goal_new; }  // F7
            else if (sample) { write_start = true; out_start = st; }
            else { st = newpos; gg = goal_new; }
            out_pos = st; out_goal = gg;
            aflags = 0; gpr = mvr = fmr = 0;
            step_count = 0; lock_count = 0; lock_head = 0; lock_prev = 0; goals_total = 0; blocking_total = 0;
            dl_events = ll_events = dl_steps = ll_steps = 0; ep_return_x2 = 0; wfg_steps = 0;
            // first observation of the next episode (final-state, no staggering at reset)
            const uint32_t l = act ? lin(st) : LFAR;
            s_new[gl] = l; s_snap[gl] = l; s_goal[gl] = act ? lin(gg) : LFAR;
        }
        __syncwarp();
        if (do_reset) {
            PairOut<G, SR, false> pr;
            pair_loop<G, SR, false>(s_new, s_snap, s_goal, s_int, s_delta, gl, act ? lin(st) : LFAR, st,
                                    0u, false, p.nearby, pr);
            if (act) {
                amask = emit_window<SR>(rows, p.wpr, prow(st), pcol(st), pr.occ, pr.xgoal, lin(gg),
                                        lin(st), my_stage_obs);
                if (p.o_goal_delta) p.o_goal_delta[ai + aoff] = goal_delta(gdt, p.R, p.C, gg, st);
                if (p.o_blocking_prev) p.o_blocking_prev[ai + aoff] = 0;
            }
        }
    }

    // ---------------------------------------------------------------- fused action sampler for the next step
    if (p.sample_mode && act) {
        carried_action = sample_action(p.seed, p.env_id_base + env, gl, p.sample_counter + (unsigned long long)it,
                                       p.sample_mode == 1 ? (int)amask : -1);
        p.o_next_actions[ai] = (int8_t)carried_action;
    }

    // ---------------------------------------------------------------- byte outputs through staging
    if (p.o_action_mask) {
        uint8_t *dst0 = reinterpret_cast<uint8_t *>(p.o_action_mask) + warp_env0_out * N * 5;
        uint8_t *ms = stage_mask + ((uintptr_t)dst0 & 15) + (size_t)(my_env_in_warp * N + gl) * 5;
        if (act) {
#pragma unroll
            for (int k = 0; k < 5; ++k) ms[k] = (uint8_t)((amask >> k) & 1u);
        }
    }
    __syncwarp();
    {
        long long envs_left = (long long)p.B - (long long)warp_env0;
        int nenv = envs_left < 0 ? 0 : (envs_left < envs_per_warp ? (int)envs_left : envs_per_warp);
        if (nenv > 0) {
            if (p.o_local_obs)
                warp_copy_out(p.o_local_obs + warp_env0_out * N * V2, stage_obs, nenv * N * V2, lane);
            if (p.o_action_mask)
                warp_copy_out(reinterpret_cast<uint8_t *>(p.o_action_mask) + warp_env0_out * N * 5,
                              stage_mask, nenv * N * 5, lane);
        }
    }

    // ---------------------------------------------------------------- state write-back
    if (act) {
        p.positions[ai] = out_pos;
        if (out_goal != goal) p.goals[ai] = out_goal;
        if (write_start) p.starts[ai] = out_start;
        p.agent_flags[ai] = (uint8_t)aflags;
        if (kLock) { p.lock_gp[ai] = gpr; p.lock_mv[ai] = mvr; p.lock_fm[ai] = fmr; }
    }
    if (env_ok && gl == 0) {
        int4 *ew = p.env_words + (size_t)env * 4;
        ew[0] = make_int4(step_count, lock_count, lock_prev, goals_total);
        ew[1] = make_int4(blocking_total, dl_events, ll_events, dl_steps);
        ew[2] = make_int4(ll_steps, (int)rng_counter, ep_return_x2, wfg_steps);
        ew[3] = make_int4(episodes, lock_head, w3.z, w3.w);
    }
    __syncwarp();
    }  // inner steps
    }  // tile loop
    errs = __reduce_or_sync(full, errs);
    if (errs && lane == 0) atomicOr(p.err_bits, errs);
}

// ============================================================================ reset kernel
template <int G, int SR>
__global__ void __launch_bounds__(256) mapf_reset_kernel(const KParams p) {
    constexpr int V = 2 * SR + 1, V2 = V * V;
    extern __shared__ __align__(16) uint32_t smem[];
    const unsigned full = 0xFFFFFFFFu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gl = tid % G, grp = tid / G;
    const int groups = blockDim.x / G;
    const int env = blockIdx.x * groups + grp;
    const unsigned gbase = (unsigned)(lane / G) * G;
    const unsigned gmask = (G == 32) ? full : (((1u << G) - 1u) << gbase);
    const int N = p.N;
    const bool env_ok = env < p.B;
    const bool sel = env_ok && (!p.reset_mask || p.reset_mask[env] != 0);
    const bool act = sel && gl < N;

    const SmemLayout &L = p.L;
    uint32_t *gsm = smem + L.grp_off + grp * L.grp_words;
    uint32_t *s_new = gsm, *s_snap = gsm + G, *s_goal = gsm + 2 * G;
    uint32_t *s_int = gsm + 3 * G, *s_delta = gsm + 4 * G;
    const float *gdt = reinterpret_cast<const float *>(smem + L.gdt_off);
    fill_goal_delta_table(reinterpret_cast<float *>(smem + L.gdt_off), p.R, p.C, p.normalize, p.den0, p.den1,
                          tid, blockDim.x);
    const uint32_t *rows, *freebm;
    if (p.per_env_maps) {
        uint32_t *mr = gsm + L.g_map, *fb = gsm + L.g_free;
        if (env_ok) {
            const uint32_t *src = p.map_rows + (size_t)env * p.map_words;
            for (int i = gl; i < p.map_words; i += G) mr[i] = src[i];
            const uint32_t *fsrc = p.free_bits + (size_t)env * p.fw;
            for (int i = gl; i < p.fw; i += G) fb[i] = fsrc[i];
        }
        rows = mr; freebm = fb;
    } else {
        uint32_t *mr = smem + L.map_rows_off, *fb = smem + L.free_off;
        for (int i = tid; i < p.map_words; i += blockDim.x) mr[i] = p.map_rows[i];
        for (int i = tid; i < p.fw; i += blockDim.x) fb[i] = p.free_bits[i];
        rows = mr; freebm = fb;
    }
    __syncthreads();
    uint8_t *stage = reinterpret_cast<uint8_t *>(smem + L.stage_off + warp * L.stage_words);
    const int envs_per_warp = 32 / G;
    const int obs_stage_bytes = ((envs_per_warp * N * V2 + 32 + 15) / 16) * 16;
    uint8_t *stage_obs = stage, *stage_mask = stage + obs_stage_bytes;

    const size_t ai = (size_t)(env_ok ? env : 0) * N + (gl < N ? gl : 0);
    int4 w2 = make_int4(0, 0, 0, 0), w3 = w2;
    if (env_ok) { w2 = p.env_words[(size_t)env * 4 + 2]; w3 = p.env_words[(size_t)env * 4 + 3]; }
    uint32_t rng_counter = (uint32_t)w2.y;
    uint32_t errs = 0;

    // layout: overrides > deterministic table (ENV:452-455) > Philox draw (ENV:267-282)
    uint32_t st = NOCELL, gg = NOCELL - 1;
    const bool have_override = p.starts_override && p.goals_override;
    bool sample = sel && !have_override && !p.deterministic && !p.observe_only;
    const int F = p.num_free[p.per_env_maps ? (env_ok ? env : 0) : 0];
    if (sample && F < 2 * N) { errs |= MAPF_DEV_ERR_TOO_FEW_CELLS; sample = false; }
    const Philox ph(p.seed, p.env_id_base + env);
    draw_layout<G>(ph, rng_counter, F, freebm, p.fw, p.C, N, gl, gbase, gmask, sample && act, st, gg);
    bool write_layout = false;
    if (act) {
        if (p.observe_only) { st = p.positions[ai]; gg = p.goals[ai]; }
        else if (have_override) { st = p.starts_override[ai]; gg = p.goals_override[ai]; write_layout = true; }
        else if (p.deterministic) { st = p.starts[ai]; gg = p.goals[ai]; }
        else if (sample) write_layout = true;
        else { st = p.positions[ai]; gg = p.goals[ai]; }
    }

    if (have_override && !p.observe_only) {
        // A layout is 2N distinct cells in the reference (rng.choice(..., replace=False), ENV:277; the tables of
        // get_grid.py).  Two agents with one goal would need the goal-owner grid's last-index-wins rule (ENV:207-212)
        // that the kernels do not keep, two with one start are the injected states of ENV:658-666 (set_state is the
        // door for those): an override that repeats a cell is refused, loudly.
        const unsigned ms = __match_any_sync(full, act ? st : (0x80000000u | (uint32_t)lane));
        const unsigned mg = __match_any_sync(full, act ? gg : (0x80000000u | (uint32_t)lane));
        if (act && (((ms | mg) & gmask & ~(1u << lane)) != 0)) errs |= MAPF_DEV_ERR_DUPLICATE_LAYOUT;
    }
    const uint32_t l = act ? lin(st) : LFAR;
    s_new[gl] = l; s_snap[gl] = l; s_goal[gl] = act ? lin(gg) : LFAR;
    __syncwarp();
    PairOut<G, SR, false> pr;
    pair_loop<G, SR, false>(s_new, s_snap, s_goal, s_int, s_delta, gl, l, st, 0u, false, p.nearby, pr);

    const int my_env_in_warp = lane / G;
    const size_t warp_env0 = (size_t)blockIdx.x * groups + (size_t)(warp * envs_per_warp);
    // outputs of unselected envs must stay untouched: selected groups write their channels directly
    if (act) {
        uint8_t tmp[V2];
        uint32_t amask = emit_window<SR>(rows, p.wpr, prow(st), pcol(st), pr.occ, pr.xgoal, lin(gg), l, tmp);
        if (p.o_local_obs) {
            uint8_t *dst = p.o_local_obs + ai * V2;
#pragma unroll
            for (int k = 0; k < V2; ++k) dst[k] = tmp[k];
        }
        if (p.o_action_mask) {
            int8_t *dst = p.o_action_mask + ai * 5;
#pragma unroll
            for (int k = 0; k < 5; ++k) dst[k] = (int8_t)((amask >> k) & 1u);
        }
        if (p.o_goal_delta) p.o_goal_delta[ai] = goal_delta(gdt, p.R, p.C, gg, st);
        if (p.observe_only) {
            if (p.o_blocking_prev) p.o_blocking_prev[ai] = (p.agent_flags[ai] >> 2) & 1u;
        } else {
        if (p.o_blocking_prev) p.o_blocking_prev[ai] = 0;
        if (p.o_reward) p.o_reward[ai] = 0.f;
        if (p.o_agent_step_flags) p.o_agent_step_flags[ai] = (st == gg) ? MAPF_ASF_ON_GOAL : 0;
        // state, ENV:441-450
        p.positions[ai] = st;
        if (write_layout) { p.starts[ai] = st; p.goals[ai] = gg; }
        p.agent_flags[ai] = 0;
        if (p.lock_gp) { p.lock_gp[ai] = 0; p.lock_mv[ai] = 0; p.lock_fm[ai] = 0; }
        if (p.lock_dist)
            for (int s = 0; s < p.lw; ++s) p.lock_dist[((size_t)env * p.lw + s) * N + gl] = 0;
        }
    }
    if (sel && gl == 0 && !p.observe_only) {
        if (p.o_terminated) p.o_terminated[env] = 0;
        if (p.o_truncated) p.o_truncated[env] = 0;
        if (p.o_step_flags) p.o_step_flags[env] = 0;
        if (p.o_info) {
            int4 *io = p.o_info + (size_t)env * 4;
            io[0] = io[1] = io[2] = io[3] = make_int4(0, 0, 0, 0);
        }
        int4 *ew = p.env_words + (size_t)env * 4;
        ew[0] = make_int4(0, 0, 0, 0);
        ew[1] = make_int4(0, 0, 0, 0);
        ew[2] = make_int4(0, (int)rng_counter, 0, 0);
        ew[3] = make_int4(w3.x, 0, w3.z, w3.w);
    }
    (void)stage_obs; (void)stage_mask; (void)my_env_in_warp; (void)warp_env0;
    errs = __reduce_or_sync(full, errs);
    if (errs && lane == 0) atomicOr(p.err_bits, errs);
}

// ============================================================================ small kernels
// ENV:306-328 flat float32 observation, one thread per output element.
__global__ void mapf_pack_flat_kernel(const uint8_t *local_obs, const float2 *goal_delta,
                                      const uint8_t *bp, const int8_t *mask, float *flat, long long BN,
                                      int V2, int gdist, int use_bp, int use_mask, int D) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= BN * D) return;
    long long a = idx / D;
    int k = (int)(idx - a * D);
    float v;
    if (k < V2) v = (float)local_obs[a * V2 + k];
    else {
        k -= V2;
        float2 g = goal_delta[a];
        if (k == 0) v = g.x;
        else if (k == 1) v = g.y;
        else {
            k -= 2;
            if (gdist && k == 0) v = __fadd_rn(fabsf(g.x), fabsf(g.y));  // ENV:320
            else {
                k -= gdist;
                if (use_bp && k == 0) v = (float)bp[a];
                else { k -= use_bp; v = (float)mask[a * 5 + k]; }
            }
        }
    }
    flat[idx] = v;
    (void)use_mask;
}

__global__ void mapf_sample_actions_kernel(const int8_t *mask, int8_t *actions, long long BN, int N,
                                           unsigned long long seed, long long env_id_base,
                                           unsigned long long counter) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= BN) return;
    long long env = idx / N;
    int a = (int)(idx - env * N);
    int bits = -1;
    if (mask) {
        const int8_t *m = mask + idx * 5;
        bits = 0;
        for (int k = 0; k < 5; ++k) bits |= (m[k] != 0) << k;
    }
    actions[idx] = (int8_t)sample_action(seed, env_id_base + env, a, counter, bits);
}

// Occupancy heat-map of the reference's evaluator (main.py:153-155, 265-267): counts[r, c] += number of agents of
// the selected envs standing on (r, c).  Integer counts, so the order of the atomics does not matter; a CTA first
// accumulates in a shared-memory histogram when the map fits.
__global__ void mapf_occupancy_kernel(const uint32_t *positions, const uint8_t *active, long long BN, int N, int R, int C,
                                      unsigned long long *counts, int use_smem) {
    extern __shared__ unsigned int hist[];
    const int cells = R * C;
    if (use_smem) {
        for (int i = threadIdx.x; i < cells; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
    }
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < BN; idx += (long long)gridDim.x * blockDim.x) {
        if (active && !active[idx / N]) continue;
        const uint32_t p = positions[idx];
        const int r = prow(p), c = pcol(p);
        if (r < 0 || r >= R || c < 0 || c >= C) continue;   // main.py:266
        if (use_smem) atomicAdd(&hist[r * C + c], 1u);
        else atomicAdd(&counts[r * C + c], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += blockDim.x)
            if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
    }
}

// Single-agent shortest paths on the shared map (the obstacle-aware distance field behind the
// reference's classical planners, scripts/a-star.py:123-126 uses the Manhattan heuristic of the same 4-neighbour
// moves).  One warp per source cell, one lane per map row (R, C <= 32): the BFS frontier is a bitboard, one wavefront
// per iteration: next = (left | right | up | down) & free & ~visited, up / down by warp shuffles.
// table[src * R*C + cell] = number of moves from src to cell, 255 = unreachable (or > 254).
__global__ void mapf_distance_table_kernel(const uint32_t *free_bits, int R, int C, uint8_t *table) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int src = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int cells = R * C;
    if (src >= cells) return;
    // free cells of my row as a C-bit word (the free bitmap is cell-linear)
    uint32_t freerow = 0;
    if (lane < R) {
        const int b0 = lane * C;
        const uint64_t w = (uint64_t)free_bits[b0 >> 5] | ((uint64_t)(((b0 >> 5) + 1) * 32 < cells + 32 ? free_bits[(b0 >> 5) + 1] : 0u) << 32);
        freerow = (uint32_t)(w >> (b0 & 31)) & (C >= 32 ? full : ((1u << C) - 1u));
    }
    uint8_t *out = table + (size_t)src * cells;
    for (int c = 0; c < C; ++c) if (lane < R) out[lane * C + c] = 255;
    const int sr = src / C, sc = src % C;
    uint32_t frontier = (lane == sr) ? ((1u << sc) & freerow) : 0u, visited = frontier;
    for (int d = 0; d < 255; ++d) {
        if (!__any_sync(full, frontier != 0)) break;
        for (uint32_t f = frontier; f; f &= f - 1) out[lane * C + (__ffs(f) - 1)] = (uint8_t)d;
        const uint32_t up = __shfl_up_sync(full, frontier, 1), dn = __shfl_down_sync(full, frontier, 1);
        uint32_t nf = (frontier << 1) | (frontier >> 1) | (lane > 0 ? up : 0u) | (lane < 31 ? dn : 0u);
        nf &= freerow & ~visited;
        visited |= nf;
        frontier = nf;
    }
}

// out[b, n] = table[goal][position] of every agent (int16; -1 = unreachable)
__global__ void mapf_goal_path_lengths_kernel(const uint32_t *positions, const uint32_t *goals, const uint8_t *table,
                                              long long BN, int R, int C, int16_t *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BN) return;
    const uint32_t p = positions[i], g = goals[i];
    const int pc = prow(p) * C + pcol(p), gc = prow(g) * C + pcol(g);
    const uint8_t d = table[(size_t)gc * (R * C) + pc];
    out[i] = d == 255 ? (int16_t)-1 : (int16_t)d;
}

// Deterministic reduction of env_metrics[B,K]: CTA k reduces metric k in a fixed order
// (strided partial sums, then a shared-memory tree), so the result does not depend on timing.
__global__ void mapf_metrics_reduce_kernel(const double *env_metrics, int B, double *out) {
    __shared__ double sm[256];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int e = threadIdx.x; e < B; e += blockDim.x) s += env_metrics[(size_t)e * MAPF_METRIC_COUNT + k];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int st = blockDim.x / 2; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = sm[0];
}

}  // namespace mapf
