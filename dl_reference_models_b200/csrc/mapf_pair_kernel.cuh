// mapf_pair_kernel.cuh -- env-per-lane-pair step kernel (sm_100a) of the batched MAPF transition.
//
// Reference semantics: src/environments/reference_model_multi_agent.py ("ENV:line").
//
// Same formulation as the env-per-thread kernel (mapf_env_kernel.cuh: owner grids as bitboards in shared memory, the
// agents walked in index order like ENV:502-563, staggered observations read off the evolving board, rare per-env
// events served by the whole warp), with TWO lanes per env.  Why: the env-per-thread launch is ONE wave of 14 warps
// per SM (a GPU-filling batch offers 13.8 tiles of 32 envs per SM and each needs 15 KB of shared memory and 128
// registers per thread); every warp is a chain of ~11 k mostly dependent instructions at ~8 cycles each and 3.5 warps
// per scheduler cannot cover that (issue slots 0.41 used, no pipe above 50 %).  Splitting an env over two lanes keeps
// the bytes of shared memory per env and doubles the warps per SM (28 x 16 envs, <= 72 registers per thread):
//   * lane = e + 16 h: env slot e of the tile, half h.  In every quad of agents (4 q .. 4 q + 3) half h owns agents
//     4 q + 2 h and 4 q + 2 h + 1: state moves in 64-bit accesses, two lanes cover one 128-bit quad;
//   * only the part of the walk that must see the board in agent order is serialised: per quad the h = 0 lanes move
//     their two agents and read the raw window rows (snapshot k, ENV:528-536), then the h = 1 lanes do (about 15
//     instructions per agent); everything else -- table look-ups, byte rows, lock history, goal deltas, the sampler --
//     runs on both halves at once;
//   * the two owner boards of an env sit side by side in one 64-bit word per map row, [row][env slot]: one LDS.64
//     of a half-warp fetches the occupancy and the goal row of 16 envs conflict-free (a wavefront per 16 agents and
//     two planes, the same as two 32-bit loads per 32 agents of the one-lane kernel);
//   * the byte rows of a quad still meet in ONE stage row per env (no extra shared memory): each half builds the 2 V^2
//     bytes of its two agents as words and stores them through a funnel shift by 16 h bits -- the second half starts
//     2 V^2 = 2 (mod 4) bytes into the row -- plus one 16-bit store each for the word the halves share.
// Supported: N a multiple of 4, shared map up to 64 x 32 (the env-per-thread kernel's limits).  Results are identical
// to the other two step kernels (tests/test_gpu_kernel_equivalence.py).
#pragma once

#include "mapf_env_kernel.cuh"

namespace mapf {

constexpr int PAIR_EPW = 16;   // envs per warp

// stage row of one env and quad: V2 observation words + 5 action-mask words; a stride = 2 (mod 4) with an odd half keeps
// the 16 rows of a half on distinct even bank offsets and the other half (an odd number of words further) off them
__host__ __device__ constexpr int pair_stage_stride(int V2) {
    int s = V2 + 5;
    while ((s & 3) != 2) ++s;
    return s;
}

// EnvLayout of the pair kernel: same tables, per-warp block = [stage][boards (uint2 [row][16])][agent records]
__host__ __device__ inline EnvLayout make_pair_layout(int N, int R, int C, int SR, int fw, int warps) {
    const int V = 2 * SR + 1, V2 = V * V;
    EnvLayout E = make_env_layout(N, R, C, SR, fw, 1);
    E.board_rows = (E.board_rows + 3) & ~3;                      // cleared 4 rows (512 B) per warp instruction
    int w = PAIR_EPW * pair_stage_stride(V2) * 4;
    w = (w + 15) & ~15;
    w += E.board_rows * PAIR_EPW * 8;
    w += 2 * E.nq * 32 * 4;                                       // records: u32 [agent of the half][lane]
    w = (w + 15) & ~15;
    E.warp_bytes = w;
    E.total_bytes = E.tables_bytes + w * warps + 1024;            // window rows beyond the last board stay inside
    return E;
}

// record slot of agent a of env slot e: [(quad, k)][lane = e + 16 h], a = 4 quad + 2 h + k
__device__ __forceinline__ int pair_rec_idx(int a, int e) {
    return ((((a >> 2) << 1) | (a & 1)) << 5) + e + (((a >> 1) & 1) << 4);
}
// bits 0..1 of m -> bytes 0/1 (two agents)
__device__ __forceinline__ uint32_t spread2(uint32_t m) { return ((m & 3u) * 0x00000081u) & 0x00000101u; }

template <int SR, bool FAST = false>
__global__ void __launch_bounds__(896, 1) mapf_step_pair_kernel(const KParams p, const EnvLayout E) {
    const bool kLifelong = FAST ? true : p.lifelong;
    const bool kLock = FAST ? true : p.lock_enabled;
    constexpr int V = 2 * SR + 1, V2 = V * V;
    constexpr uint32_t VM = (1u << V) - 1u;
    constexpr uint32_t M4 = VM << 2;
    constexpr int OBS_W = V2;                        // observation words per quad and env
    constexpr int STRIDE = pair_stage_stride(V2);
    constexpr int STAGE_BYTES = (PAIR_EPW * STRIDE * 4 + 15) & ~15;
    constexpr int CTR = SR * V + SR;
    constexpr int PADR = ENV_ROW_PAD;
    constexpr int EPW = PAIR_EPW;
    constexpr int NWH = (V2 - 1) / 2;                // full words of a half's 2 * V2 bytes (one half-word follows)
    using WB = typename WinBits<V>::type;
    extern __shared__ __align__(16) unsigned char esm[];
    const unsigned full = 0xFFFFFFFFu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = blockDim.x >> 5;
    const int e = lane & 15, h = lane >> 4;
    const int N = p.N, R = p.R, C = p.C, NQ = E.nq;

    // ------------------------------------------------------------------ CTA-wide tables (built on the host)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.env_tables);
        uint4 *dst = reinterpret_cast<uint4 *>(esm);
        for (int i = tid; i < (E.tables_bytes >> 4); i += blockDim.x) dst[i] = src[i];
    }
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    const char *t1b = reinterpret_cast<const char *>(esm + ENV_T1_OFF);
    const uint8_t *kth = esm + ENV_KTH_OFF;
    const WB *lut = reinterpret_cast<const WB *>(esm + ENV_LUT_OFF);
    const uint32_t *freerow = reinterpret_cast<const uint32_t *>(esm + ENV_FREEROW_OFF);
    const uint32_t *freebits = reinterpret_cast<const uint32_t *>(esm + ENV_FREEBITS_OFF);
    const float *gdt = reinterpret_cast<const float *>(esm + ENV_GDT_OFF);
    const uint32_t esm_s = (uint32_t)__cvta_generic_to_shared(esm);
    const uint32_t t1_s = esm_s + ENV_T1_OFF, kth_s = esm_s + ENV_KTH_OFF, lut_s = esm_s + ENV_LUT_OFF;
    const uint32_t *freepre = reinterpret_cast<const uint32_t *>(esm + E.pre_off);
    auto kth_free = [&](int k) -> int {
        int lo = 0;
#pragma unroll
        for (int stp = 32; stp >= 1; stp >>= 1)
            if (lo + stp < p.fw && (int)freepre[lo + stp] <= k) lo += stp;
        return lo * 32 + (int)__fns(freebits[lo], 0, k - (int)freepre[lo] + 1);
    };

    unsigned char *wsm = esm + E.tables_bytes + warp * E.warp_bytes;
    uint32_t *stage_w = reinterpret_cast<uint32_t *>(wsm);                       // [env slot][STRIDE] words
    const uint32_t my_stage_s = (uint32_t)__cvta_generic_to_shared(stage_w + e * STRIDE);
    uint2 *brd_w = reinterpret_cast<uint2 *>(wsm + STAGE_BYTES);                 // boards [row][env slot]: .x occupancy, .y goals
    uint32_t *rec_w = reinterpret_cast<uint32_t *>(brd_w + E.board_rows * EPW);  // agent records [agent of the half][lane]
    uint2 *brd = brd_w + e;                                                      // my env: row r at brd[r * EPW]
    uint32_t *rec = rec_w + lane;                                                // my agents: (quad q, k) at rec[(2 q + k) * 32]
    const int ntiles = (p.B + EPW - 1) / EPW;
    uint32_t errs = 0;
    const uint32_t mdw = p.dw >= 32 ? full : ((1u << p.dw) - 1u);
    const uint32_t mlw = p.lw >= 32 ? full : ((1u << p.lw) - 1u);
    const uint32_t allN = (N >= 32) ? full : ((1u << N) - 1u);

    // Final-state observation of agent `lane` of ONE env, the whole warp on that env (ENV:565-575 after a goal
    // reassignment, ENV:459-468 after a reset) -- straight to global memory.  brd_e: the env's boards (row r at [r * EPW]).
    auto emit_final = [&](size_t abe, long long eg, const uint2 *brd_e, uint32_t code, uint32_t gcode, uint32_t bp_bit,
                          bool ctr2) {
        if (lane >= N) return;
        const int r = (int)(code >> 5), c = (int)(code & 31u);
        const WB obst = lut[code];
        const int sa = c > SR ? c - SR : 0, sb2 = (c < SR ? SR - c : 0) + 2;
        const uint2 *brow = brd_e + (r - SR) * EPW;
        uint32_t acc[(4 * V2 + 31) / 32 + 1];
#pragma unroll
        for (int j = 0; j < (int)(sizeof(acc) / sizeof(acc[0])); ++j) acc[j] = 0;
        uint32_t blk_up = 0, blk_mid = 0, blk_dn = 0;
#pragma unroll
        for (int wr = 0; wr < V; ++wr) {
            const uint2 b = brow[wr * EPW];
            const uint32_t o4 = (wr * V >= 2 ? (uint32_t)(obst >> (wr * V - 2)) : (uint32_t)(obst << 2)) & M4;
            const uint32_t MC = (wr == SR) ? (M4 & ~(4u << SR)) : M4;
            const uint32_t occ4 = ((b.x >> sa) << sb2) & MC;
            const uint32_t agent4 = occ4 & ~o4, blk4 = occ4 | o4;
            const uint32_t g4 = ((b.y >> sa) << sb2) & M4 & ~blk4;
            if (wr == SR - 1) blk_up = blk4;
            if (wr == SR) blk_mid = blk4;
            if (wr == SR + 1) blk_dn = blk4;
            const uint32_t t = *reinterpret_cast<const uint32_t *>(t1b + o4) +
                               (*reinterpret_cast<const uint32_t *>(t1b + agent4) << 1) +
                               (*reinterpret_cast<const uint32_t *>(t1b + g4) << 2);
            const int bitpos = 4 * V * wr, wi = bitpos >> 5, sh = bitpos & 31;
            acc[wi] |= t << sh;
            if (sh + 4 * V > 32) acc[wi + 1] |= t >> (32 - sh);
        }
        const uint32_t am = 1u | ((~blk_up >> (2 + SR)) & 1u) << 1 | ((~blk_mid >> (2 + SR + 1)) & 1u) << 2 |
                            ((~blk_dn >> (2 + SR)) & 1u) << 3 | ((~blk_mid >> (2 + SR - 1)) & 1u) << 4;
        if (p.o_local_obs) {
            uint8_t *ob = p.o_local_obs + (abe + lane) * V2;
#pragma unroll
            for (int n = 0; n < V2; ++n) ob[n] = (uint8_t)((acc[n >> 3] >> (4 * (n & 7))) & 0xFu);
            const int dr = (int)(gcode >> 5) - r + SR, dc = (int)(gcode & 31u) - c + SR;
            if ((unsigned)dr < (unsigned)V && (unsigned)dc < (unsigned)V) {   // own goal: code 3
                const int ci = dr * V + dc;
                const bool occ_other = (ci != CTR) && ((brd_e[(gcode >> 5) * EPW].x >> (gcode & 31u)) & 1u);
                if (!((obst >> ci) & 1) && !occ_other) ob[ci] = 3;
            }
            if (ctr2) ob[CTR] = 2;   // injected co-location (ENV:737-739)
        }
        if (p.o_action_mask) {
#pragma unroll
            for (int k = 0; k < 5; ++k) p.o_action_mask[(abe + lane) * 5 + k] = (int8_t)((am >> k) & 1u);
        }
        if (p.o_goal_delta) {
            const int gi0 = (int)(gcode >> 5) - r + (R - 1), gi1 = (int)(gcode & 31u) - c + (C - 1) + 2 * R - 1;
            p.o_goal_delta[abe + lane] = make_float2(gdt[gi0], gdt[gi1]);
        }
        if (p.o_blocking_prev) p.o_blocking_prev[abe + lane] = (uint8_t)bp_bit;
        if (p.sample_mode) {
            const uint4 rnd = sample_quad(p.seed, eg, lane >> 2, p.sample_counter);
            const uint32_t x = qget(rnd, lane & 3);
            const uint32_t na = p.sample_mode == 1 ? kth[am * 8 + __umulhi(x, (uint32_t)__popc(am))] : __umulhi(x, 5u);
            p.o_next_actions[abe + lane] = (int8_t)na;
        }
    };
    // OR / AND / sum over the two lanes of an env
    auto por = [&](uint32_t x) { return x | __shfl_xor_sync(full, x, 16); };

    for (int tile = blockIdx.x * warps + warp; tile < ntiles; tile += gridDim.x * warps) {
    const int env = tile * EPW + e;
    const bool ok = env < p.B;
    const size_t ab = (size_t)(ok ? env : 0) * N;
    const long long env_global = p.env_id_base + env;
    const size_t env0 = (size_t)tile * EPW;

    int lock_count = 0, lock_head = 0;
    if (ok) {
        const int *ew1 = reinterpret_cast<const int *>(p.env_words + (size_t)env * 4);
        lock_count = ew1[MAPF_W_LOCK_COUNT];
        lock_head = ew1[MAPF_W_LOCK_HEAD];
    }
    const int count_after = lock_count + 1;
    if (lock_head < 0 || lock_head >= p.lw) lock_head = 0;
    const int slot_new = lock_head;
    const int slot_next = (lock_head + 1 == p.lw) ? 0 : lock_head + 1;
    const bool use_ring = kLock && count_after >= p.lw && p.lw > 1;

    // masks over agent indices; until the walk is over every lane holds the bits of ITS agents only
    uint32_t reached_m = 0, completed_m = 0, bprev_m = 0;
    bool degen = false;
    uint32_t notown_m = 0;   // MAPF_AF_NOT_OWNER bits of my agents (injected co-location carried over)
    // ---------------------------------------------------------------- pre-pass: agent records and owner boards of the state before the step
    {
        uint4 *z = reinterpret_cast<uint4 *>(brd_w);
        for (int i = lane; i < E.board_rows * (EPW * 8 / 16); i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    uint32_t dup = 0;
    for (int q = 0; q < NQ; ++q) {
        const int i0 = 4 * q + 2 * h;
        uint2 pq = make_uint2(0u, 0u), gq = pq;
        uint32_t act2 = 0, fl2 = 0;
        if (ok) {
            pq = *reinterpret_cast<const uint2 *>(p.positions + ab + i0);
            gq = *reinterpret_cast<const uint2 *>(p.goals + ab + i0);
            if (p.actions) act2 = *reinterpret_cast<const uint16_t *>(p.actions + ab + i0);
            fl2 = *reinterpret_cast<const uint16_t *>(p.agent_flags + ab + i0);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t fl = (fl2 >> (8 * k)) & 0xFFu;
            reached_m |= (fl & 1u) << (i0 + k);
            completed_m |= ((fl >> 1) & 1u) << (i0 + k);
            bprev_m |= ((fl >> 2) & 1u) << (i0 + k);
            notown_m |= ((fl >> 3) & 1u) << (i0 + k);
            const uint32_t code = code_of(k ? pq.y : pq.x), gcode = code_of(k ? gq.y : gq.x);
            int a = (int)(int8_t)(act2 >> (8 * k));
            if (a < 0 || a > 4) { errs |= MAPF_DEV_ERR_INVALID_ACTION; a = 0; }
            const uint32_t nbi = __byte_perm((uint32_t)CTR | ((uint32_t)(CTR - V) << 8) | ((uint32_t)(CTR + 1) << 16) |
                                             ((uint32_t)(CTR + V) << 24), (uint32_t)(CTR - 1), (uint32_t)a) & 0xFFu;
            const uint32_t tblocked = ((uint32_t)lut[code] >> nbi) & 1u;
            rec[(2 * q + k) * 32] = code | ((uint32_t)a << 11) | (tblocked << 14) | (gcode << 16);
            if (ok) {
                if (!((fl >> 3) & 1u)) {   // only the owner of a cell marks it (always, in legal states)
                    const uint32_t cb0 = 1u << (code & 31u);
                    dup |= atomicOr(&brd[(code >> 5) * EPW].x, cb0) & cb0;   // two agents on one cell: an injected state
                } else degen = true;
                atomicOr(&brd[(gcode >> 5) * EPW].y, 1u << (gcode & 31u));
            }
        }
    }
    degen |= dup != 0;
    degen = __shfl_xor_sync(full, (int)degen, 16) != 0 || degen;
    __syncwarp();   // boards and records complete (both halves wrote them)
    uint32_t solo_m = 0xFFFFFFFFu;   // bit a = agent a owns the cell it stands on (ENV:200-205); env-wide, degenerate states only
    const uint32_t notown_env = por(notown_m);   // (a warp-wide shuffle: outside the per-env branch)
    if (degen) {   // rare: among agents that claim the same cell the highest index owns it
        solo_m = ~notown_env;
        const uint32_t claim = solo_m;
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j)
                if (((claim >> j) & 1u) && ((rec_w[pair_rec_idx(i, e)] ^ rec_w[pair_rec_idx(j, e)]) & REC_CODE) == 0u) {
                    solo_m &= ~(1u << i);
                    break;
                }
    }

    uint32_t moved_m = 0, failed_m = 0, gstep_m = 0, ongoal_m = 0;
    uint32_t Gd = 0, Md = 0, Fd = 0, Gl = 0, Ml = 0;
    const unsigned act16 = __ballot_sync(full, ok) & 0xFFFFu;
    const bool masked_sampler = p.sample_mode == 1;

    // ---------------------------------------------------------------- the agent walk (ENV:502-563), a quad at a time
    int4 w0 = make_int4(0, 0, 0, 0), w1 = w0, w2 = w0, w3 = w0;
    for (int q = 0; q < NQ; ++q) {
        const int i0 = 4 * q + 2 * h;
        const uint32_t rv0 = rec[(2 * q) * 32], rv1 = rec[(2 * q + 1) * 32];
        uint2 gpq = make_uint2(0u, 0u), mvq = gpq, fmq = gpq;
        uint32_t ringq = 0;
        if (kLock && ok) {
            gpq = *reinterpret_cast<const uint2 *>(p.lock_gp + ab + i0);
            mvq = *reinterpret_cast<const uint2 *>(p.lock_mv + ab + i0);
            fmq = *reinterpret_cast<const uint2 *>(p.lock_fm + ab + i0);
            if (use_ring) ringq = *reinterpret_cast<const uint32_t *>(p.lock_dist + ((size_t)env * p.lw + slot_next) * N + i0);
        }
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (p.sample_mode) rnd = sample_quad(p.seed, env_global, q, p.sample_counter);
        // ---- what the serial turn needs, prepared by both halves at once
        uint32_t code[2] = {rv0 & REC_CODE, rv1 & REC_CODE};
        const uint32_t gcode[2] = {rv0 >> 16, rv1 >> 16};
        const uint32_t act[2] = {(rv0 >> 11) & 7u, (rv1 >> 11) & 7u};
        const uint32_t tcode[2] = {code[0] + (uint32_t)action_delta(act[0]), code[1] + (uint32_t)action_delta(act[1])};
        const bool wants[2] = {ok && act[0] != 0 && !(rv0 & 0x4000u), ok && act[1] != 0 && !(rv1 & 0x4000u)};
        uint2 raw[2][V];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int wr = 0; wr < V; ++wr) raw[k][wr] = make_uint2(0u, 0u);
        uint32_t mvb = 0, occo = 0;
        // ---- the serial turn of one half: moves in agent order and the raw window rows right after each own move,
        // when the occupancy board IS snapshot k (ENV:528-536).  Goals do not change inside the walk.
        auto turn = [&]() {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int trow = (int)tcode[k] >> 5;   // a blocked target may lie one row outside the board: inside the allocation, ignored
                const uint32_t tb = 1u << (tcode[k] & 31u);
                const uint32_t tv = brd[trow * EPW].x;
                const bool moves = wants[k] & ((tv & tb) == 0u);
                atomicAnd(&brd[(int)(code[k] >> 5) * EPW].x, moves ? ~(1u << (code[k] & 31u)) : 0xFFFFFFFFu);
                atomicOr(&brd[trow * EPW].x, moves ? tb : 0u);
                code[k] = moves ? tcode[k] : code[k];
                mvb |= moves ? (1u << k) : 0u;
                const uint2 *brow = &brd[((int)(code[k] >> 5) - SR) * EPW];
#pragma unroll
                for (int wr = 0; wr < V; ++wr) raw[k][wr] = brow[wr * EPW];
                occo |= ((brd[(int)(gcode[k] >> 5) * EPW].x >> (gcode[k] & 31u)) & 1u) << k;
            }
        };
        if (h == 0) turn();
        __syncwarp();
        if (h == 1) turn();
        __syncwarp();

        // ---- everything else of the quad, both halves at once (my agents: i0, i0 + 1)
        uint32_t ds[2] = {0, 0};
        uint32_t masks2 = 0, next2 = 0;
        uint32_t win[2][V];
        int patch[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t rvk = k ? rv1 : rv0;
            const uint32_t bit = 1u << (i0 + k);
            const bool moves = (mvb >> k) & 1u;
            const bool failed = ok && act[k] != 0 && !moves;  // ENV:583
            moved_m |= moves ? bit : 0u;
            failed_m |= failed ? bit : 0u;
            const uint32_t cd = code[k], gc = gcode[k];
            const uint32_t ocode = moves ? cd - (uint32_t)action_delta(act[k]) : cd;
            // ENV:538-563.  A lifelong arrival gets its new goal after the walk, in agent order.
            const bool on_goal = ok && cd == gc;
            bool gstep = false, cur_on_goal = on_goal;
            if (!kLifelong) {
                if (on_goal && !(reached_m & bit)) { reached_m |= bit; completed_m |= bit; gstep = true; }
            } else if (on_goal) {
                gstep = true;
                completed_m |= bit; reached_m &= ~bit;
                cur_on_goal = false;   // ENV:555
            }
            gstep_m |= gstep ? bit : 0u;
            ongoal_m |= cur_on_goal ? bit : 0u;
            // ENV:581-594 lock history
            uint32_t delta16 = 0;
            if (kLock) {
                const bool prev_on_goal = kLifelong ? false : (ocode == gc);
                const bool gp = kLifelong ? gstep : (!prev_on_goal && cur_on_goal);
                const uint32_t g2 = ((k ? gpq.y : gpq.x) << 1) | (gp ? 1u : 0u);
                const uint32_t m2 = ((k ? mvq.y : mvq.x) << 1) | (moves ? 1u : 0u);
                const uint32_t f2 = ((k ? fmq.y : fmq.x) << 1) | (failed ? 1u : 0u);
                if (k) { gpq.y = g2; mvq.y = m2; fmq.y = f2; } else { gpq.x = g2; mvq.x = m2; fmq.x = f2; }
                const uint32_t okbit = ok ? bit : 0u;
                Gd |= (g2 & mdw) ? okbit : 0u;
                Md |= (m2 & mdw) ? okbit : 0u;
                Fd |= (f2 & mdw) ? okbit : 0u;
                Gl |= (g2 & mlw) ? okbit : 0u;
                Ml |= (m2 & mlw) ? okbit : 0u;
                const int dist = abs((int)(gc >> 5) - (int)(cd >> 5)) + abs((int)(gc & 31u) - (int)(cd & 31u));
                ds[k] = (uint32_t)dist;
                const int ring_old = (int)(int16_t)(k ? (ringq >> 16) : (ringq & 0xFFFFu));
                delta16 = (use_ring && ok) ? (uint32_t)(ring_old - dist) << 16 : 0u;
            }
            rec[(2 * q + k) * 32] = cd | (rvk & 0x7800u) | delta16;
            // ---------------------------------------------------- window rows of the agent from the raw rows of its snapshot
            const int r = (int)(cd >> 5), c = (int)(cd & 31u);
            const WB obst = lut_ld<WB>(lut_s + cd * (uint32_t)sizeof(WB));
            const int sa = c > SR ? c - SR : 0, sb2 = (c < SR ? SR - c : 0) + 2;   // window columns start at c - SR
            uint32_t blk_up = 0, blk_mid = 0, blk_dn = 0;
#pragma unroll
            for (int wr = 0; wr < V; ++wr) {
                // rows outside the map read neighbouring shared memory: masked by the obstacle plane
                const uint32_t bx = raw[k][wr].x, by = raw[k][wr].y;
                const uint32_t o4 = (wr * V >= 2 ? (uint32_t)(obst >> (wr * V - 2)) : (uint32_t)(obst << 2)) & M4;
                const uint32_t MC = (wr == SR) ? (M4 & ~(4u << SR)) : M4;   // my own cell is not "another agent" (ENV:737)
                const uint32_t occ4 = ((bx >> sa) << sb2) & MC;
                const uint32_t agent4 = occ4 & ~o4;
                const uint32_t blk4 = occ4 | o4;
                const uint32_t g4 = ((by >> sa) << sb2) & M4 & ~blk4;
                if (wr == SR - 1) blk_up = blk4;
                if (wr == SR) blk_mid = blk4;
                if (wr == SR + 1) blk_dn = blk4;
                // plane 0 = obstacle | other's goal, plane 1 = other agent | other's goal: nibble codes 1 / 2 / 3 (3 -> OTHER_GOAL
                // by the PRMT below); t1 has one word per bank, lanes that share a bank share the address
                win[k][wr] = lds_pure(t1_s + (o4 | g4)) + (lds_pure(t1_s + (agent4 | g4)) << 1);
            }
            // ENV:761-771: a direction is valid iff its neighbour is neither obstacle nor agent
            const uint32_t am = 1u | ((~blk_up >> (2 + SR)) & 1u) << 1 | ((~blk_mid >> (2 + SR + 1)) & 1u) << 2 |
                                ((~blk_dn >> (2 + SR)) & 1u) << 3 | ((~blk_mid >> (2 + SR - 1)) & 1u) << 4;
            masks2 |= am << (8 * k);
            {   // own goal (code 3): the goal plane wrote 4 there; patched after the quad's words are stored
                const int dr = (int)(gc >> 5) - r + SR, dc = (int)(gc & 31u) - c + SR;
                const bool in = (unsigned)dr < (unsigned)V && (unsigned)dc < (unsigned)V;
                const int ci = in ? dr * V + dc : 0;
                const bool occ_other = (ci != CTR) && ((occo >> k) & 1u);
                patch[k] = (in && !((obst >> ci) & 1) && !occ_other) ? V2 * k + ci : -1;
            }
            {   // both samplers, selected afterwards (stored only when the sampler is fused in)
                const uint32_t x = h ? (k ? rnd.w : rnd.z) : (k ? rnd.y : rnd.x);
                const uint32_t na_masked = lds_pure_u8(kth_s + am * 8u + __umulhi(x, (uint32_t)__popc(am)));
                const uint32_t na = masked_sampler ? na_masked : __umulhi(x, 5u);
                next2 |= na << (8 * k);
            }
        }
        if (ok) {
            *reinterpret_cast<uint2 *>(p.positions + ab + i0) = make_uint2(packed_of(code[0]), packed_of(code[1]));
            if (kLock) {
                *reinterpret_cast<uint2 *>(p.lock_gp + ab + i0) = gpq;
                *reinterpret_cast<uint2 *>(p.lock_mv + ab + i0) = mvq;
                *reinterpret_cast<uint2 *>(p.lock_fm + ab + i0) = fmq;
                *reinterpret_cast<uint32_t *>(p.lock_dist + ((size_t)env * p.lw + slot_new) * N + i0) = ds[0] | (ds[1] << 16);
            }
        }
        if (q + 1 == NQ && ok) {
            const int4 *ew4 = p.env_words + (size_t)env * 4;
            w0 = ew4[0]; w1 = ew4[1]; w2 = ew4[2]; w3 = ew4[3];
        }
        // ---- byte rows of my two agents: 2 * V2 bytes as NWH words and a half-word, then into the env's stage row
        // behind a funnel shift by 16 h bits (the second half starts 2 * V2 = 2 (mod 4) bytes into the row)
        uint32_t sw[NWH + 1];
        {
            uint32_t carry = 0;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                uint32_t acc[(4 * (V2 + 3) + 31) / 32 + 1];
#pragma unroll
                for (int j = 0; j < (int)(sizeof(acc) / sizeof(acc[0])); ++j) acc[j] = 0;
#pragma unroll
                for (int wr = 0; wr < V; ++wr) {
                    const uint32_t t = win[k][wr];
                    const int bitpos = 4 * (V * wr + k);  // agent 1's bytes start 1 byte into its first word (V2 = 1 mod 4)
                    const int wi = bitpos >> 5, sh = bitpos & 31;
                    acc[wi] |= t << sh;
                    if (sh + 4 * V > 32) acc[wi + 1] |= t >> (32 - sh);
                }
                constexpr int NWMAX = (V2 + 1 + 3) / 4;
                const int j0 = (V2 * k) >> 2;
                const int nw = (k + V2 + 3) >> 2;
#pragma unroll
                for (int m = 0; m < NWMAX; ++m) {
                    if (m >= nw) continue;
                    uint32_t w = nibbles_to_bytes_0124((m & 1) ? (acc[m >> 1] >> 16) : acc[m >> 1]);
                    if (m == 0 && k != 0) w |= carry;                    // leading partial word shared with agent 0
                    if (m == nw - 1 && k == 0) carry = w;                // trailing partial word of agent 0: finished by agent 1
                    else sw[j0 + m] = w;
                }
            }
        }
        const uint32_t fsh = 16u * (uint32_t)h;
        const uint32_t obs_s = my_stage_s + (uint32_t)((NWH + 1) * 4 * h);
#pragma unroll
        for (int m = 0; m < NWH; ++m) sts_stage(obs_s + 4u * m, __funnelshift_r(sw[m], sw[m + 1], fsh));
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(my_stage_s + (uint32_t)(4 * NWH) + 2u * h),
                     "h"((unsigned short)((h ? sw[0] : sw[NWH]) & 0xFFFFu)));
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (patch[k] >= 0) sts_stage_u8(my_stage_s + (uint32_t)(2 * V2 * h + patch[k]), 3u);
        {   // action masks of my two agents: 10 bytes = 2 words and a half-word, same scheme behind the observation words
            const uint32_t m0 = masks2 & 0x1Fu, m1 = (masks2 >> 8) & 0x1Fu;
            const uint32_t lo0 = spread4(m0), lo1 = spread4(m1);
            const uint32_t ms0 = lo0, ms1 = (m0 >> 4) | (lo1 << 8), ms2 = (lo1 >> 24) | ((m1 >> 4) << 8);
            const uint32_t msk_s = my_stage_s + (uint32_t)(4 * OBS_W) + 12u * h;
            sts_stage(msk_s, __funnelshift_r(ms0, ms1, fsh));
            sts_stage(msk_s + 4u, __funnelshift_r(ms1, ms2, fsh));
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(my_stage_s + (uint32_t)(4 * OBS_W + 8) + 2u * h),
                         "h"((unsigned short)((h ? ms0 : ms2) & 0xFFFFu)));
        }
        if (ok) {
            if (p.o_goal_delta) {   // ENV:330-335
                float2 gd[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int gi0 = (int)(gcode[k] >> 5) - (int)(code[k] >> 5) + (R - 1);
                    const int gi1 = (int)(gcode[k] & 31u) - (int)(code[k] & 31u) + (C - 1) + 2 * R - 1;
                    gd[k] = make_float2(gdt[gi0], gdt[gi1]);
                }
                *reinterpret_cast<float4 *>(p.o_goal_delta + ab + i0) = make_float4(gd[0].x, gd[0].y, gd[1].x, gd[1].y);
            }
            if (p.o_blocking_prev)
                *reinterpret_cast<uint16_t *>(p.o_blocking_prev + ab + i0) = (uint16_t)spread2(bprev_m >> i0);
            if (p.sample_mode) *reinterpret_cast<uint16_t *>(p.o_next_actions + ab + i0) = (uint16_t)next2;
        }
        // ------------------------------------------------ coalesced flush of the stage rows (one per env)
        // row `s` holds quad q of env env0 + s: agent index (env0 + s) * N + 4 * q
        __syncwarp();
        {
            const size_t agent0 = env0 * N + (size_t)(4 * q);
            for (int w = lane; w < OBS_W + 5; w += 32) {
                const bool is_obs = w < OBS_W;
                unsigned char *gp = is_obs ? (p.o_local_obs ? p.o_local_obs + agent0 * V2 + 4 * w : nullptr)
                                           : (p.o_action_mask ? reinterpret_cast<unsigned char *>(p.o_action_mask) +
                                                                    agent0 * 5 + 4 * (w - OBS_W) : nullptr);
                const uint32_t gstride = (uint32_t)N * (uint32_t)(is_obs ? V2 : 5);
                const uint32_t *src = stage_w + w;
                if (gp) {
                    if (act16 == 0xFFFFu) {
#pragma unroll
                        for (int s = 0; s < EPW; ++s)
                            *reinterpret_cast<uint32_t *>(gp + (size_t)s * gstride) = src[s * STRIDE];
                    } else {
#pragma unroll 1
                        for (int s = 0; s < EPW; ++s)
                            if ((act16 >> s) & 1u)
                                *reinterpret_cast<uint32_t *>(gp + (size_t)s * gstride) = src[s * STRIDE];
                    }
                }
            }
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------- both halves' masks become env-wide
    moved_m = por(moved_m); failed_m = por(failed_m); gstep_m = por(gstep_m); ongoal_m = por(ongoal_m);
    reached_m = por(reached_m); completed_m = por(completed_m); bprev_m = por(bprev_m);
    if (kLock) { Gd = por(Gd); Md = por(Md); Fd = por(Fd); Gl = por(Gl); Ml = por(Ml); }

    // ---------------------------------------------------------------- the rest of the env words
    int step_count = w0.x + 1;  // ENV:475
    int lock_prev = w0.z, goals_total = w0.w;
    int blocking_total = w1.x, dl_events = w1.y, ll_events = w1.z, dl_steps = w1.w;
    int ll_steps = w2.x;
    uint32_t rng_counter = (uint32_t)w2.y;
    int ep_return_x2 = w2.z, wfg_steps = w2.w;
    int episodes = w3.x;

    // ---------------------------------------------------------------- injected co-location (ENV:658-666), kept off the walk
    // (both lanes of the env replay the same thing; the centre patches they write are the same bytes)
    if (__any_sync(full, degen)) {
        __syncwarp();   // the flush above wrote those windows
        if (degen && ok) {
            for (int i = 0; i < N; ++i) {
                const uint32_t rvi = rec_w[pair_rec_idx(i, e)], newc = rvi & REC_CODE, biti = 1u << i;
                const bool mvd = (moved_m >> i) & 1u;
                const uint32_t oldc = mvd ? newc - (uint32_t)action_delta((rvi >> 11) & 7u) : newc;
                bool owned_by_other = false;
                for (int a2 = 0; a2 < N; ++a2) {
                    if (a2 == i) continue;
                    const uint32_t rva = rec_w[pair_rec_idx(a2, e)];
                    uint32_t ca = rva & REC_CODE;   // where a2 stands at agent i's turn: moved already only if a2 < i
                    if (a2 > i && ((moved_m >> a2) & 1u)) ca -= (uint32_t)action_delta((rva >> 11) & 7u);
                    if (ca == oldc) {
                        if (mvd) solo_m &= ~(1u << a2);                       // ENV:523
                        else if ((solo_m >> a2) & 1u) owned_by_other = true;  // the cell has an owner, and it is not me
                    }
                }
                if (mvd) solo_m |= biti;                                       // ENV:525
                else if (!(solo_m & biti) && owned_by_other && p.o_local_obs)
                    p.o_local_obs[(ab + i) * V2 + CTR] = 2;
            }
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------- lifelong goal reassignment (ENV:284-304, 547-556)
    if (kLock) lock_head = slot_next;
    const int arrivals = __popc(gstep_m);
    goals_total += arrivals;  // lifelong: every arrival; else first arrivals (ENV:545,562)
    const uint32_t pend = kLifelong ? gstep_m : 0u;
    const bool reassigned = pend != 0;
    {
        // the WARP serves each of its reassigned envs together: lane = agent for the roll-back / re-emission,
        // lane = map row for the candidate scan (see mapf_env_kernel.cuh)
        unsigned rw = __ballot_sync(full, pend != 0) & 0xFFFFu;
        while (rw) {
            const int es = __ffs(rw) - 1;
            rw &= rw - 1;
            uint32_t pe = __shfl_sync(full, pend, es);
            const uint32_t mv_e = __shfl_sync(full, moved_m, es), bp_e = __shfl_sync(full, bprev_m, es);
            const uint32_t solo_e = __shfl_sync(full, degen ? solo_m : 0xFFFFFFFFu, es);
            const uint32_t rc_e = __shfl_sync(full, rng_counter, es);
            const int slot_new_e = __shfl_sync(full, slot_new, es), slot_next_e = __shfl_sync(full, slot_next, es);
            const bool use_ring_e = __shfl_sync(full, (int)use_ring, es) != 0;
            const long long eg = p.env_id_base + (long long)(env0 + es);
            const size_t abe = (env0 + es) * (size_t)N, enve = env0 + es;
            uint2 *brd_e = brd_w + es;
            const int my_rec = pair_rec_idx(lane < N ? lane : 0, es);
            uint32_t rng_inc = 0, err_e = 0, ongoal_fix = 0;
            while (pe) {   // arrivals in agent order (ENV:284-304)
                const int i = __ffs(pe) - 1;
                pe &= pe - 1;
                // roll the env's occupancy board back to snapshot i: later movers leave their new cell, then re-take the old one
                const uint32_t later = mv_e & ~((2u << i) - 1u);
                const bool und = lane < N && ((later >> lane) & 1u);
                uint32_t nc = 0, oc = 0;
                if (und) {
                    const uint32_t rv = rec_w[my_rec];
                    nc = rv & REC_CODE;
                    oc = nc - (uint32_t)action_delta((rv >> 11) & 7u);
                    atomicAnd(&brd_e[(nc >> 5) * EPW].x, ~(1u << (nc & 31u)));
                }
                const uint32_t gold = code_of(p.goals[abe + i]);
                if (lane == 0) brd_e[(gold >> 5) * EPW].y &= ~(1u << (gold & 31u));   // ENV:288
                __syncwarp();
                if (und) atomicOr(&brd_e[(oc >> 5) * EPW].x, 1u << (oc & 31u));
                __syncwarp();
                uint32_t ng = 0xFFFFFFFFu;
                if (p.goal_override) {
                    const uint32_t ov = p.goal_override[abe + i];
                    if (prow(ov) >= 0) ng = code_of(ov);
                }
                if (ng == 0xFFFFFFFFu) {   // candidates = free, unoccupied, nobody's goal; rows lane and lane + 32
                    uint32_t c0 = 0, c1 = 0;
                    if (lane < R) { const uint2 b = brd_e[lane * EPW]; c0 = freerow[lane] & ~b.x & ~b.y; }
                    if (lane + 32 < R) { const uint2 b = brd_e[(lane + 32) * EPW]; c1 = freerow[lane + 32] & ~b.x & ~b.y; }
                    const int n0 = __popc(c0), n1 = __popc(c1);
                    int pre0 = n0, pre1 = n1;   // inclusive prefix sums over the lanes
#pragma unroll
                    for (int sft = 1; sft < 32; sft <<= 1) {
                        const int v0 = __shfl_up_sync(full, pre0, sft), v1 = __shfl_up_sync(full, pre1, sft);
                        if (lane >= sft) { pre0 += v0; pre1 += v1; }
                    }
                    const int tot0 = __shfl_sync(full, pre0, 31), n = tot0 + __shfl_sync(full, pre1, 31);
                    int kk = -1;
                    if (p.goal_rank) kk = p.goal_rank[abe + i];
                    if (kk < 0 && n > 0) {
                        const Philox ph(p.seed, eg);
                        const uint4 x = ph(rc_e + rng_inc, (uint32_t)i, 0x474F414Cu /* "GOAL" */, 0);
                        kk = (int)__umulhi(x.x, (uint32_t)n);
                        rng_inc++;
                    }
                    if (n > 0 && kk < n) {   // row-major order: rows 0..31, then 32..63
                        const bool hit0 = kk >= pre0 - n0 && kk < pre0;
                        const bool hit1 = kk >= tot0 + pre1 - n1 && kk < tot0 + pre1;
                        uint32_t mine = 0;
                        if (hit0) mine = (uint32_t)(lane * 32) + __fns(c0, 0, kk - (pre0 - n0) + 1);
                        if (hit1) mine = (uint32_t)((lane + 32) * 32) + __fns(c1, 0, kk - tot0 - (pre1 - n1) + 1);
                        const unsigned hb = __ballot_sync(full, hit0 || hit1);
                        ng = __shfl_sync(full, mine, __ffs(hb) - 1);
                    } else {
                        err_e |= MAPF_DEV_ERR_NO_GOAL_CELL;
                    }
                }
                if (ng == 0xFFFFFFFFu) { ng = gold; ongoal_fix |= 1u << i; }   // no cell: the old goal stays, the agent is on it
                if (lane == 0) {
                    brd_e[(ng >> 5) * EPW].y |= 1u << (ng & 31u);
                    p.goals[abe + i] = packed_of(ng);
                    if (kLock) {   // ENV:591: distance to the NEW goal
                        const int ri = pair_rec_idx(i, es);
                        const uint32_t rv = rec_w[ri], cde = rv & REC_CODE;
                        const int dist = abs((int)(ng >> 5) - (int)(cde >> 5)) + abs((int)(ng & 31u) - (int)(cde & 31u));
                        p.lock_dist[((size_t)enve * p.lw + slot_new_e) * N + i] = (int16_t)dist;
                        if (use_ring_e) {
                            const int ring_old = (int)p.lock_dist[((size_t)enve * p.lw + slot_next_e) * N + i];
                            rec_w[ri] = (rv & 0xFFFFu) | ((uint32_t)(ring_old - dist) << 16);
                        }
                    }
                }
                // roll forward again
                if (und) atomicAnd(&brd_e[(oc >> 5) * EPW].x, ~(1u << (oc & 31u)));
                __syncwarp();
                if (und) atomicOr(&brd_e[(nc >> 5) * EPW].x, 1u << (nc & 31u));
                __syncwarp();
            }
            // ENV:565-575: everybody of this env shows the final state; lane = agent
            {
                const uint32_t cde = lane < N ? (rec_w[my_rec] & REC_CODE) : 0u;
                const uint32_t gcd = lane < N ? code_of(p.goals[abe + lane]) : 0u;
                const bool ctr2 = !((solo_e >> lane) & 1u) && ((brd_e[(cde >> 5) * EPW].x >> (cde & 31u)) & 1u);
                emit_final(abe, eg, brd_e, cde, gcd, (bp_e >> lane) & 1u, ctr2);
            }
            if (e == es) { rng_counter += rng_inc; errs |= err_e; ongoal_m |= ongoal_fix; }
            __syncwarp();
        }
    }

    // ---------------------------------------------------------------- epilogue: owner masks, locks, blocking, wait-for graph
    // The boards are dead: .x of row r + PADR / .y of row c + PADR get bit a for agent a's final row / column.
    __syncwarp();
    {
        uint4 *z = reinterpret_cast<uint4 *>(brd_w);
        for (int i = lane; i < E.board_rows * (EPW * 8 / 16); i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    if (ok) {
#pragma unroll 1
        for (int al = 0; al < 2 * NQ; ++al) {
            const int i = ((al >> 1) << 2) + 2 * h + (al & 1);
            const uint32_t cde = rec[al * 32] & REC_CODE;
            atomicOr(&brd[((cde >> 5) + PADR) * EPW].x, 1u << i);
            atomicOr(&brd[((cde & 31u) + PADR) * EPW].y, 1u << i);
        }
    }
    __syncwarp();
    uint32_t coloc_any = 0, wf_alive = 0, flags_any = 0;   // flags_any: bit 0 deadlock, bit 1 livelock participant set found
    uint32_t wf_m = 0, blocking_m = 0;
    const uint32_t intent_m = allN & ~reached_m;  // ENV:619-621: only agents that have not (sticky-)reached press
#pragma unroll 1
    for (int al = 0; al < (ok ? 2 * NQ : 0); ++al) {
        const int i = ((al >> 1) << 2) + 2 * h + (al & 1);
        const uint32_t bit = 1u << i;
        const uint32_t rv = rec[al * 32];
        const uint32_t cde = rv & REC_CODE;
        const int r = (int)(cde >> 5), c = (int)(cde & 31u);
        const uint2 *prow_ = &brd[(r + PADR) * EPW], *pcol_ = &brd[(c + PADR) * EPW];
        const uint32_t here = prow_->x & pcol_->y;   // agents on my cell (me included)
        if (here & ~bit) coloc_any |= bit;
        // ENV:389-438 neighbours within Manhattan distance `nearby`, via the row / column masks
        if (kLock && !(ongoal_m & bit)) {
            uint32_t nb = 0;
            if (p.nearby == 2) {
                const uint32_t c0 = pcol_->y;
                const uint32_t c1 = c0 | pcol_[-EPW].y | pcol_[EPW].y;
                const uint32_t c2 = c1 | pcol_[-2 * EPW].y | pcol_[2 * EPW].y;
                nb = (prow_->x & c2) | ((prow_[-EPW].x | prow_[EPW].x) & c1) | ((prow_[-2 * EPW].x | prow_[2 * EPW].x) & c0);
            } else {
                uint32_t u = 0;
                for (int w = 0; w <= p.nearby; ++w) {
                    const int dd = p.nearby - w;
                    if (c - w >= 0) u |= pcol_[-w * EPW].y;
                    if (c + w < C) u |= pcol_[w * EPW].y;
                    uint32_t rm = 0;
                    if (r - dd >= 0) rm |= prow_[-dd * EPW].x;
                    if (r + dd < R) rm |= prow_[dd * EPW].x;
                    nb |= rm & u;
                }
            }
            nb &= ~here;
            if (__popc(nb) >= p.min_nb) {
                const uint32_t P = nb | bit;
                if (!(P & Gd) && !(P & Md) && (P & Fd)) flags_any |= 1u;
                if (!(P & Gl) && (P & Ml)) {
                    uint32_t rest = nb;
                    int red = (int)rv >> 16;
                    while (rest) {
                        const int a = __ffs(rest) - 1;
                        rest &= rest - 1;
                        red += (int)rec_w[pair_rec_idx(a, e)] >> 16;
                    }
                    if (red <= p.eps_floor) flags_any |= 2u;
                }
            }
        }
        // intended cell (kept even when invalid, ENV:514-515) -> who stands there
        uint32_t owner = here & ~bit;
        if (!(moved_m & bit)) {
            owner = 0;
            if (!(rv & 0x4000u) && (reached_m != 0u || (failed_m & bit))) {
                const uint32_t tc = cde + (uint32_t)action_delta((rv >> 11) & 7u);
                owner = brd[((tc >> 5) + PADR) * EPW].x & brd[((tc & 31u) + PADR) * EPW].y & ~bit;
            }
        }
        if (intent_m & bit) blocking_m |= owner;   // ENV:609-623 (filtered below)
        if ((failed_m & bit) && owner) {   // wait-for edge i -> owner (kept in the action bits of the record)
            rec[al * 32] = (rv & ~0xF800u) | ((uint32_t)(31 - __clz(owner)) << 11);
            wf_alive |= bit;
        }
    }
    coloc_any = por(coloc_any); wf_alive = por(wf_alive); flags_any = por(flags_any); blocking_m = por(blocking_m);
    const bool dl_any = flags_any & 1u, ll_any = (flags_any & 2u) != 0;
    blocking_m &= reached_m & ~moved_m;
    // wait-for cycles: strip agents whose target is gone or that nobody waits for, until stable
    if (__any_sync(full, wf_alive != 0)) {
        __syncwarp();   // the edges written by the other half
        uint32_t alive = wf_alive;
        for (;;) {
            uint32_t keep = 0, targets = 0, rest = alive;
            while (rest) {
                const int i = __ffs(rest) - 1;
                rest &= rest - 1;
                const uint32_t t = (rec_w[pair_rec_idx(i, e)] >> 11) & 31u;
                if ((alive >> t) & 1u) { keep |= 1u << i; targets |= 1u << t; }
            }
            keep &= targets;
            const bool changed = keep != alive;
            alive = keep;
            if (!__any_sync(full, changed)) break;
        }
        wf_m = alive;
    }
    const bool wf_any = wf_m != 0;
    wfg_steps += wf_any;
    const int blocking_step = __popc(blocking_m);
    blocking_total += blocking_step;

    // ---------------------------------------------------------------- lock detection result, ENV:595-606
    bool dl_step = false, ll_step = false, dl_event = false, ll_event = false;
    if (kLock) {
        dl_step = count_after >= p.dw && dl_any;
        ll_step = !dl_step && count_after >= p.lw && ll_any;
        dl_event = dl_step && !(lock_prev & 1);
        ll_event = ll_step && !(lock_prev & 2);
        lock_prev = (dl_step ? 1 : 0) | (ll_step ? 2 : 0);
        dl_steps += dl_step; ll_steps += ll_step; dl_events += dl_event; ll_events += ll_event;
        lock_count = count_after;
    }

    // ---------------------------------------------------------------- rewards & termination, ENV:658-690
    bool terminated = false, truncated = false;
    const uint32_t scratch_on = kLifelong ? 0u : ongoal_m;   // reached_goal scratch (ENV:555)
    uint32_t bonus_m = 0, penalty_m = 0;
    if (!kLifelong && __popc(scratch_on) == N) { terminated = true; bonus_m = allN; }
    else if (step_count >= p.steps_per_episode) {
        terminated = true; truncated = true;  // F6
        if (!kLifelong) penalty_m = allN & ~scratch_on;
    }
    const bool done = ok && (terminated || truncated);
    int rsum = __popc(gstep_m) + 2 * __popc(bonus_m) - 2 * __popc(penalty_m);
    uint32_t coloc_pairs2 = 0;   // 2 * (co-located others), summed over my agents
    for (int q = 0; q < NQ; ++q) {
        const int i0 = 4 * q + 2 * h;
        const uint32_t gs = spread2(gstep_m >> i0), bl = spread2(blocking_m >> i0);
        const uint32_t asf2 = spread2(moved_m >> i0) * MAPF_ASF_MOVED + spread2(failed_m >> i0) * MAPF_ASF_FAILED_MOVE +
                              gs * MAPF_ASF_GOAL_REACHED + bl * MAPF_ASF_BLOCKING +
                              spread2(wf_m >> i0) * MAPF_ASF_WFG_CYCLE + spread2(ongoal_m >> i0) * MAPF_ASF_ON_GOAL;
        uint32_t af2 = spread2(reached_m >> i0) * MAPF_AF_REACHED + spread2(completed_m >> i0) * MAPF_AF_COMPLETED_ONCE +
                       bl * MAPF_AF_BLOCKING_PREV;
        if (degen) af2 += spread2(~solo_m >> i0) * MAPF_AF_NOT_OWNER;   // the owner grid outlives the step (ENV:102)
        const uint32_t pos2 = gs + 2u * spread2(bonus_m >> i0), neg2 = 2u * spread2(penalty_m >> i0);
        float rw[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int rx2 = (int)((pos2 >> (8 * k)) & 0xFFu) - (int)((neg2 >> (8 * k)) & 0xFFu);
            if ((coloc_any >> (i0 + k)) & 1u) {   // ENV:658-666, -1 per co-located pair member (injected states only)
                const uint32_t cde = rec[(2 * q + k) * 32] & REC_CODE;
                const int others = __popc(brd[((cde >> 5) + PADR) * EPW].x & brd[((cde & 31u) + PADR) * EPW].y) - 1;
                rx2 -= 2 * others;
                coloc_pairs2 += 2u * (uint32_t)others;
            }
            rw[k] = 0.5f * (float)rx2;
        }
        if (ok) {
            if (p.o_reward) *reinterpret_cast<float2 *>(p.o_reward + ab + i0) = make_float2(rw[0], rw[1]);
            if (p.o_agent_step_flags) *reinterpret_cast<uint16_t *>(p.o_agent_step_flags + ab + i0) = (uint16_t)asf2;
            *reinterpret_cast<uint16_t *>(p.agent_flags + ab + i0) = (uint16_t)af2;
        }
    }
    coloc_pairs2 += __shfl_xor_sync(full, coloc_pairs2, 16);
    rsum -= (int)coloc_pairs2;
    ep_return_x2 += rsum;
    const int n_comp = __popc(completed_m), n_reach = __popc(reached_m);
    if (ok && h == 0) {
        if (p.o_info) {  // integer sources of info["__all__"], ENV:639-656
            int4 *io = p.o_info + (size_t)env * 4;
            io[0] = make_int4(arrivals, kLifelong ? goals_total : n_reach, blocking_step, blocking_total);
            io[1] = make_int4(dl_step, ll_step, dl_event, ll_event);
            io[2] = make_int4(dl_events, ll_events, dl_steps, ll_steps);
            io[3] = make_int4(n_comp, step_count, n_reach, wfg_steps);
        }
        if (p.o_terminated) p.o_terminated[env] = terminated;
        if (p.o_truncated) p.o_truncated[env] = truncated;
        if (p.o_step_flags)
            p.o_step_flags[env] = (uint8_t)((terminated ? MAPF_SF_TERMINATED : 0) | (truncated ? MAPF_SF_TRUNCATED : 0) |
                                            (dl_step ? MAPF_SF_DEADLOCK_STEP : 0) | (ll_step ? MAPF_SF_LIVELOCK_STEP : 0) |
                                            (dl_event ? MAPF_SF_DEADLOCK_EVENT : 0) | (ll_event ? MAPF_SF_LIVELOCK_EVENT : 0) |
                                            (reassigned ? MAPF_SF_GOAL_REASSIGNED : 0) | (wf_any ? MAPF_SF_WFG_CYCLE : 0));
    }

    // ---------------------------------------------------------------- episode end: metrics, auto-reset
    if (done) {
        if (h == 0) {   // episode-end metric sums, src/trainers/callbacks.py:152,173,335-345
            double *m = p.env_metrics + (size_t)env * MAPF_METRIC_COUNT;
            const double gt = kLifelong ? (double)goals_total : (double)n_reach;  // ENV:630-633
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
            m[MAPF_M_COMPLETION_RATIO_SUM] += (double)n_comp / (double)N;                 // ENV:638
            m[MAPF_M_WFG_CYCLE_STEPS_SUM] += (double)wfg_steps;
        }
        episodes += 1;
    }
    // ENV:440-472 inside the launch: the WARP resets each of its finished envs together, lane = agent
    unsigned rsw = __ballot_sync(full, done && p.auto_reset) & 0xFFFFu;
    if (rsw) {
        const int F = p.num_free[0];
        __syncwarp();
        while (rsw) {
            const int es = __ffs(rsw) - 1;
            rsw &= rsw - 1;
            const size_t enve = env0 + es, abe = enve * (size_t)N;
            const long long eg = p.env_id_base + (long long)enve;
            const uint32_t rc_e = __shfl_sync(full, rng_counter, es);
            uint2 *brd_e = brd_w + es;
            const bool mine = lane < N;
            uint32_t rounds = 0, err_e = 0;
            bool sample = !p.deterministic;
            if (sample && F < 2 * N) { err_e |= MAPF_DEV_ERR_TOO_FEW_CELLS; sample = false; }
            uint32_t st = 0, gg = 0;   // cell codes of agent `lane`
            if (sample) {
                const Philox ph(p.seed, eg);
                int cs = -1 - lane, cg = -33 - lane;   // distinct placeholders for the lanes beyond N
                bool rs = mine, rgn = mine;
                while (__any_sync(full, rs || rgn)) {
                    if (rs || rgn) {
                        const uint4 x = ph(rc_e + rounds, (uint32_t)lane, 0x52455345u /* "RESE" */, 0);
                        if (rs) cs = kth_free((int)__umulhi(x.x, (uint32_t)F));
                        if (rgn) cg = kth_free((int)__umulhi(x.y, (uint32_t)F));
                    }
                    rs = false; rgn = false;
                    if (N <= 16) {
                        const int gv = __shfl_sync(full, cg, (lane - 16) & 31);
                        const unsigned same = __match_any_sync(full, lane < 16 ? cs : gv);
                        const bool dupl = (same & ((1u << lane) - 1u)) != 0;
                        const bool dupg = __shfl_sync(full, (int)dupl, (lane + 16) & 31) != 0;
                        rs = mine && dupl;
                        rgn = mine && dupg;
                    } else {
                        for (int a = 0; a < N; ++a) {
                            const int os = __shfl_sync(full, cs, a), og = __shfl_sync(full, cg, a);
                            if (mine) {
                                if (a < lane && os == cs) rs = true;
                                if (os == cg) rgn = true;
                                if (a < lane && og == cg) rgn = true;
                            }
                        }
                    }
                    rounds++;
                }
                if (mine) {
                    st = (uint32_t)((cs / C) * 32 + cs % C);
                    gg = (uint32_t)((cg / C) * 32 + cg % C);
                }
            } else if (mine) {
                st = p.deterministic ? code_of(p.starts[abe + lane]) : (rec_w[pair_rec_idx(lane, es)] & REC_CODE);   // F7: goals stay
                gg = code_of(p.goals[abe + lane]);
            }
            if (mine) {
                p.positions[abe + lane] = packed_of(st);
                if (sample) { p.starts[abe + lane] = packed_of(st); p.goals[abe + lane] = packed_of(gg); }
                p.agent_flags[abe + lane] = 0;
                if (kLock) { p.lock_gp[abe + lane] = 0u; p.lock_mv[abe + lane] = 0u; p.lock_fm[abe + lane] = 0u; }
            }
            // boards of the new layout
            for (int r = lane; r < E.board_rows; r += 32) brd_e[r * EPW] = make_uint2(0u, 0u);
            __syncwarp();
            if (mine) {
                atomicOr(&brd_e[(st >> 5) * EPW].x, 1u << (st & 31u));
                atomicOr(&brd_e[(gg >> 5) * EPW].y, 1u << (gg & 31u));
            }
            __syncwarp();
            emit_final(abe, eg, brd_e, st, gg, 0u, false);
            if (e == es) {
                rng_counter += rounds; errs |= err_e;
                step_count = 0; lock_count = 0; lock_head = 0; lock_prev = 0; goals_total = 0; blocking_total = 0;
                dl_events = ll_events = dl_steps = ll_steps = 0; ep_return_x2 = 0; wfg_steps = 0;
            }
            __syncwarp();
        }
    }

    // ---------------------------------------------------------------- env words write-back
    if (ok && h == 0) {
        int4 *ew4 = p.env_words + (size_t)env * 4;
        ew4[0] = make_int4(step_count, lock_count, lock_prev, goals_total);
        ew4[1] = make_int4(blocking_total, dl_events, ll_events, dl_steps);
        ew4[2] = make_int4(ll_steps, (int)rng_counter, ep_return_x2, wfg_steps);
        ew4[3] = make_int4(episodes, lock_head, w3.z, w3.w);
    }
    __syncwarp();
    }  // tile loop
    errs = __reduce_or_sync(full, errs);
    if (errs && lane == 0) atomicOr(p.err_bits, errs);
}

}  // namespace mapf
