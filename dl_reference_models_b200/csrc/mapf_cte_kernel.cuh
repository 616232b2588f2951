// mapf_cte_kernel.cuh -- the reference's single-agent ("CTE": centralised training and execution) view of the
// same grid world, src/environments/reference_model_single_agent.py (cited "CTE:line"): one joint action for all
// agents, ONE observation (the full grid with agent / goal codes + the 5N action mask) and ONE scalar reward with
// blocking / move-after-goal penalties.  Same sequential move semantics as the multi-agent env (CTE:246-276).
//
// Mapping: one env per thread, agents walked in index order (the batch is the parallel axis; N <= 32 and the
// env has no lock metrics or lifelong mode, so the per-env work is small).  The [R,C] observation of a warp's 32
// envs is one contiguous block: all lanes copy the obstacle map into it with coalesced stores, then every thread
// patches its own env's 2N goal / agent cells (CTE:428-441).  The scalar reward is accumulated in double in
// exactly the reference's order of additions, so it is bit-identical to the Python float.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mapf_b200.h"

namespace mapf {

__global__ void __launch_bounds__(128) mapf_cte_kernel(const mapf_cte_args a, const int mode /* 0 step, 1 reset */) {
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long env0 = env - lane;
    const int N = a.num_agents, R = a.rows, C = a.cols, RC = R * C;
    const bool ok = env < a.num_envs;
    const bool sel = ok && (mode == 0 || !a.reset_mask || a.reset_mask[env] != 0);
    const unsigned sel_w = __ballot_sync(full, sel);

    int py[MAPF_MAX_AGENTS], px[MAPF_MAX_AGENTS], gy[MAPF_MAX_AGENTS], gx[MAPF_MAX_AGENTS];
    uint32_t once = 0;
    if (sel) {
        for (int i = 0; i < N; ++i) {
            const size_t k = ((size_t)env * N + i) * 2;
            py[i] = a.positions[k]; px[i] = a.positions[k + 1];
            gy[i] = a.goals[k]; gx[i] = a.goals[k + 1];
            if (mode == 0 && a.reached_once[(size_t)env * N + i]) once |= 1u << i;
        }
    }
    if (sel && mode == 1) {   // CTE:218-223 (the layout itself is installed by the caller)
        a.step_count[env] = 0;
        a.blocking_total[env] = 0.0;
        for (int i = 0; i < N; ++i) a.reached_once[(size_t)env * N + i] = 0;
    }
    if (sel && mode == 0) {
        const int step_count = a.step_count[env] + 1;   // CTE:238
        a.step_count[env] = step_count;
        double reward = 0.0;
        double blocking_step = 0.0, goals_step = 0.0;
        uint32_t moved = 0, on_goal = 0;
        int iy[MAPF_MAX_AGENTS], ix[MAPF_MAX_AGENTS];
        bool bad = false;
        for (int i = 0; i < N; ++i) {   // CTE:250-276, sequential in agent order
            int act = a.actions ? (int)a.actions[(size_t)env * N + i] : 0;
            if (act < 0 || act > 4) { bad = true; act = 0; }   // ValueError at CTE:375-377
            const int ny = py[i] + (act == 3) - (act == 1), nx = px[i] + (act == 2) - (act == 4);
            iy[i] = ny; ix[i] = nx;
            bool free_cell = ny >= 0 && ny < R && nx >= 0 && nx < C && a.grid[ny * C + nx] == 0;
            if (free_cell)
                for (int j = 0; j < N; ++j)
                    if (j != i && py[j] == ny && px[j] == nx) { free_cell = false; break; }
            if (free_cell) {
                if (ny != py[i] || nx != px[i]) moved |= 1u << i;
                py[i] = ny; px[i] = nx;
            }
            if (py[i] == gy[i] && px[i] == gx[i]) {
                on_goal |= 1u << i;
                if (!((once >> i) & 1u)) { once |= 1u << i; reward += 0.5; goals_step += 1.0; }
            }
        }
        if (bad) atomicOr(a.err_bits, MAPF_DEV_ERR_INVALID_ACTION);
        for (int i = 0; i < N; ++i)   // CTE:285-290
            for (int j = i + 1; j < N; ++j)
                if (py[i] == py[j] && px[i] == px[j]) reward -= 1.0;
        for (int b = 0; b < N; ++b) {   // CTE:292-307 intent-based local blocking penalty
            if (!((once >> b) & 1u) || ((moved >> b) & 1u)) continue;
            for (int o = 0; o < N; ++o) {
                if (o == b || ((once >> o) & 1u)) continue;
                if (iy[o] == py[b] && ix[o] == px[b]) { reward += a.blocking_penalty; blocking_step += 1.0; break; }
            }
        }
        const double blocking_total = a.blocking_total[env] + blocking_step;
        a.blocking_total[env] = blocking_total;
        for (int i = 0; i < N; ++i)   // CTE:309-315
            if (((once >> i) & 1u) && ((moved >> i) & 1u)) reward += a.move_after_goal_penalty;
        const uint32_t allN = N >= 32 ? full : ((1u << N) - 1u);
        bool term = false, trunc = false;
        if (on_goal == allN) { reward += (double)N; term = true; }                 // CTE:317-320
        else if (step_count >= a.steps_per_episode) {                                // CTE:329-337
            for (int i = 0; i < N; ++i) if (!((on_goal >> i) & 1u)) reward -= 1.0;
            term = true; trunc = true;
        }
        for (int i = 0; i < N; ++i) {
            const size_t k = ((size_t)env * N + i) * 2;
            a.positions[k] = (int16_t)py[i]; a.positions[k + 1] = (int16_t)px[i];
            a.reached_once[(size_t)env * N + i] = (once >> i) & 1u;
        }
        if (a.reward) a.reward[env] = reward;
        if (a.terminated) a.terminated[env] = term;
        if (a.truncated) a.truncated[env] = trunc;
        if (a.info) {
            double *io = a.info + (size_t)env * 4;
            io[0] = blocking_step; io[1] = goals_step; io[2] = (double)__popc(once); io[3] = blocking_total;
        }
    }

    // ------------------------------------------------------------------ observation, CTE:428-441
    const int D = RC + 5 * N;
    for (int e = 0; e < 32; ++e) {   // the map into the selected envs' blocks, coalesced
        if (!((sel_w >> e) & 1u)) continue;
        if (a.obs_grid) { uint8_t *dst = a.obs_grid + (size_t)(env0 + e) * RC; for (int c = lane; c < RC; c += 32) dst[c] = a.grid[c]; }
        if (a.flat_obs) { float *dst = a.flat_obs + (size_t)(env0 + e) * D; for (int c = lane; c < RC; c += 32) dst[c] = (float)a.grid[c]; }
    }
    __syncwarp();
    if (sel) {
        uint8_t *og = a.obs_grid ? a.obs_grid + (size_t)env * RC : nullptr;
        float *fo = a.flat_obs ? a.flat_obs + (size_t)env * D : nullptr;
        for (int i = 0; i < N; ++i) {   // goals first, agents overwrite goals
            const int c = gy[i] * C + gx[i];
            if (og) og[c] = (uint8_t)(2 * i + 3);
            if (fo) fo[c] = (float)(2 * i + 3);
        }
        for (int i = 0; i < N; ++i) {
            const int c = py[i] * C + px[i];
            if (og) og[c] = (uint8_t)(2 * i + 2);
            if (fo) fo[c] = (float)(2 * i + 2);
        }
        // CTE:466-489: allowed iff inside the map and the cell value is 0 or odd <=> no agent stands there
        // (goal codes and the obstacle code 1 are odd: the reference's rule, kept as is)
        for (int i = 0; i < N; ++i) {
            int m[5] = {1, py[i] > 0, px[i] < C - 1, py[i] < R - 1, px[i] > 0};
            for (int j = 0; j < N; ++j) {
                if (py[j] == py[i] - 1 && px[j] == px[i]) m[1] = 0;
                if (py[j] == py[i] && px[j] == px[i] + 1) m[2] = 0;
                if (py[j] == py[i] + 1 && px[j] == px[i]) m[3] = 0;
                if (py[j] == py[i] && px[j] == px[i] - 1) m[4] = 0;
            }
            for (int k = 0; k < 5; ++k) {
                if (a.action_mask) a.action_mask[((size_t)env * N + i) * 5 + k] = (int8_t)m[k];
                if (fo) fo[RC + 5 * i + k] = (float)m[k];
            }
        }
    }
}

}  // namespace mapf
