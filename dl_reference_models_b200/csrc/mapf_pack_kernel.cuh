// Bit-packing of the per-agent output channels for the host-buffer entry point (mapf_step_host).
//
// The host-buffer step is PCIe-bound: 43 B per agent leave the device (window 25 B, mask 5 B, goal delta 8 B,
// reward 4 B, blocking flag 1 B at sensor range 2).  All of it is low-entropy: window cells are codes 0..4, mask
// entries are 0/1, the goal delta is a table value indexed by an integer row/column difference, the reward is an
// integer number of halves.  This kernel rewrites those channels as one record per agent,
//     [3 bits per window cell, 8 cells -> 3 bytes | (V2 & 7) cells + 5 mask bits][int8 d_row][int8 d_col]
//     [int8 2*reward][u8 blocking_prev]                                             (14 B at sensor range 2),
// the records cross PCIe, and host threads expand them into the caller's arrays (mapf_host_unpack.cpp) -- bit for
// bit the arrays the unpacked copy would have delivered.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mapf_host_unpack.h"

namespace mapf {

// One thread per agent of the range [0, n_agents) of the given (already advanced) channel pointers.
// inv0 / inv1: factors that turn the goal-delta floats back into integer differences (the normalisation
// denominators, or 1).  Records are staged in shared memory so the global stores are coalesced words.
__global__ void __launch_bounds__(256) mapf_pack_host_kernel(
    const uint8_t *__restrict__ obs, const int8_t *__restrict__ mask, const float2 *__restrict__ gd,
    const float *__restrict__ reward, const uint8_t *__restrict__ bp, uint8_t *__restrict__ out,
    long long n_agents, int V2, int RS, float inv0, float inv1) {
    extern __shared__ uint8_t pack_stage[];  // blockDim.x * RS bytes (+3 slack)
    const long long first = (long long)blockIdx.x * blockDim.x;
    const long long a = first + threadIdx.x;
    if (a < n_agents) {
        const uint8_t *o = obs + a * V2;
        uint8_t *dst = pack_stage + (size_t)threadIdx.x * RS;
        const int nfull = V2 >> 3, rem = V2 & 7;
        int k = 0;
        for (int c = 0; c < nfull; ++c) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) w |= (uint32_t)(o[c * 8 + i] & 7u) << (3 * i);
            dst[k++] = (uint8_t)w; dst[k++] = (uint8_t)(w >> 8); dst[k++] = (uint8_t)(w >> 16);
        }
        uint32_t w = 0;
        for (int i = 0; i < rem; ++i) w |= (uint32_t)(o[nfull * 8 + i] & 7u) << (3 * i);
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) m |= (uint32_t)(mask[a * 5 + i] != 0) << i;
        w |= m << (3 * rem);
        const int tail_bytes = (rem * 3 + 5 + 7) / 8;
        for (int i = 0; i < tail_bytes; ++i) dst[k++] = (uint8_t)(w >> (8 * i));
        const float2 g = gd[a];
        dst[k++] = (uint8_t)(int8_t)__float2int_rn(g.x * inv0);
        dst[k++] = (uint8_t)(int8_t)__float2int_rn(g.y * inv1);
        dst[k++] = (uint8_t)(int8_t)__float2int_rn(reward[a] * 2.0f);
        dst[k++] = bp ? bp[a] : (uint8_t)0;
    }
    __syncthreads();
    // coalesced copy of this block's records (first * RS is a multiple of 4: blockDim.x is)
    const long long left = n_agents - first;
    const int nrec = left < (long long)blockDim.x ? (int)left : (int)blockDim.x;
    const int nbytes = nrec * RS;
    uint8_t *gout = out + first * RS;
    const int nwords = nbytes >> 2;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(pack_stage);
    uint32_t *gw = reinterpret_cast<uint32_t *>(gout);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) gw[i] = sw[i];
    for (int i = (nwords << 2) + threadIdx.x; i < nbytes; i += blockDim.x) gout[i] = pack_stage[i];
}

}  // namespace mapf
