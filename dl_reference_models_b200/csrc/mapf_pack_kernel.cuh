// Bit-packing of the per-agent output channels for the host-buffer entry point (mapf_step_host).
//
// The host-buffer step is PCIe-bound: 43 B per agent leave the device (window 25 B, mask 5 B, goal delta 8 B,
// reward 4 B, blocking flag 1 B at sensor range 2).  The first four are low-entropy: window cells are codes 0..4,
// mask entries are 0/1, the goal delta is a table value indexed by an integer row/column difference, the reward is
// an integer number of halves.  This kernel rewrites those channels of a slice of n agents as three streams,
//     [n x (3 bits per window cell, 8 cells -> 3 bytes; last cell + 5 mask bits -> 1 byte)]
//     [n x (int8 d_row, int8 d_col)]  [n x int8 2*reward]                    (13 B per agent at sensor range 2),
// the block crosses PCIe, and host threads expand it into the caller's arrays (mapf_host_unpack.cpp) -- bit for
// bit the arrays the unpacked copy would have delivered.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mapf_host_unpack.h"

namespace mapf {

// One thread per agent of the range [0, n_agents) of the given (already advanced) channel pointers; `out` is the
// slice's packed block; out_bp / out_env*: where the requested byte channels (blocking_prev per agent; terminated,
// truncated, step_flags per env) go inside the block, null = not requested.  inv0 / inv1: factors that turn the goal-delta floats back into integer differences (the
// normalisation denominators, or 1).  The bit stream is staged in shared memory so its global stores are
// coalesced words.
__global__ void __launch_bounds__(256) mapf_pack_host_kernel(
    const uint8_t *__restrict__ obs, const int8_t *__restrict__ mask, const float2 *__restrict__ gd,
    const float *__restrict__ reward, uint8_t *__restrict__ out, long long n_agents, int V2, float inv0, float inv1,
    const uint8_t *__restrict__ bp, uint8_t *__restrict__ out_bp, const uint8_t *__restrict__ env0,
    uint8_t *__restrict__ out_env0, const uint8_t *__restrict__ env1, uint8_t *__restrict__ out_env1,
    const uint8_t *__restrict__ env2, uint8_t *__restrict__ out_env2, long long n_envs,
    uint8_t *__restrict__ diff_out = nullptr, uint8_t *__restrict__ rew_out = nullptr) {
    // diff_out / rew_out: where the slice's goal-difference and reward streams go when they do not follow the bit
    // stream directly (the batch-wide three-stream layout of mapf_step_host_records)
    extern __shared__ uint8_t pack_stage[];  // blockDim.x * PB bytes (+3 slack)
    const int nfull = V2 >> 3, PB = nfull * 3 + 1;
    const long long first = (long long)blockIdx.x * blockDim.x;
    const long long a = first + threadIdx.x;
    if (a < n_agents) {
        const uint8_t *o = obs + a * V2;
        uint8_t *dst = pack_stage + (size_t)threadIdx.x * PB;
        int k = 0;
        for (int c = 0; c < nfull; ++c) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) w |= (uint32_t)(o[c * 8 + i] & 7u) << (3 * i);
            dst[k++] = (uint8_t)w; dst[k++] = (uint8_t)(w >> 8); dst[k++] = (uint8_t)(w >> 16);
        }
        uint32_t w = o[nfull * 8] & 7u;
#pragma unroll
        for (int i = 0; i < 5; ++i) w |= (uint32_t)(mask[a * 5 + i] != 0) << (3 + i);
        dst[k] = (uint8_t)w;
        const float2 g = gd[a];
        uint8_t *diff = (diff_out ? diff_out : out + n_agents * PB) + a * 2;  // byte stores: the stream's offset may be odd
        diff[0] = (uint8_t)(int8_t)__float2int_rn(g.x * inv0);
        diff[1] = (uint8_t)(int8_t)__float2int_rn(g.y * inv1);
        (rew_out ? rew_out : out + n_agents * (PB + 2))[a] = (uint8_t)(int8_t)__float2int_rn(reward[a] * 2.0f);
        // byte channels that ride along unchanged (one DMA per slice instead of one per channel)
        if (out_bp) out_bp[a] = bp[a];
        if (a < n_envs) {
            if (out_env0) out_env0[a] = env0[a];
            if (out_env1) out_env1[a] = env1[a];
            if (out_env2) out_env2[a] = env2[a];
        }
    }
    __syncthreads();
    // coalesced copy of this block's share of the bit stream (first * PB is a multiple of 4: blockDim.x is)
    const long long left = n_agents - first;
    const int nrec = left < (long long)blockDim.x ? (int)left : (int)blockDim.x;
    const int nbytes = nrec * PB;
    uint8_t *gout = out + first * PB;
    const int nwords = nbytes >> 2;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(pack_stage);
    uint32_t *gw = reinterpret_cast<uint32_t *>(gout);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) gw[i] = sw[i];
    for (int i = (nwords << 2) + threadIdx.x; i < nbytes; i += blockDim.x) gout[i] = pack_stage[i];
}

// Stream-ordered "slice has arrived" signal: launched behind the slice's device-to-host copy, writes the step's
// ticket into pinned (device-mapped) host memory where the host threads poll it.
__global__ void mapf_ticket_kernel(volatile uint32_t *slot, uint32_t ticket) {
    *slot = ticket;
    __threadfence_system();
}

}  // namespace mapf
