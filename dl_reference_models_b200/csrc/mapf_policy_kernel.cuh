// mapf_policy_kernel.cuh -- rollout-loop kernels around the env step (BASELINE config 5).
//
// mapf_policy_act_kernel: the reference's action-mask MLP (models/action_mask_model.py:8-67:
// Linear(F,64)-ReLU-Linear(64,64)-ReLU-{Linear(64,5), Linear(64,1)}, logits + log(mask + 1e-6)) evaluated
// straight from the env's output channels, plus the categorical draw, in ONE launch: the float feature block
// the PyTorch loop materialises (112 B per agent and step, written and re-read several times) never exists.
//   * a warp takes 32 agents: raw channels (u8 window, f32 goal delta, u8 pressure) -> bf16 feature tile in
//     shared memory -> three layers of m16n8k16 bf16 mma.sync with f32 accumulation; the accumulator layout of
//     one layer IS the A-operand layout of the next (two n8 tiles = one k16 chunk), so activations stay in
//     registers between layers;
//   * weights live in shared memory as bf16 (padded row strides: conflict-free fragment loads);
//   * the head's 5 logits + value are gathered inside each lane quad; lane 0 of the quad applies the mask,
//     does a stable softmax and draws by inverse CDF from one Philox word (keyed by the global env id).
// The op is tiny-tile GEMM work bounded by its ~50 B of HBM traffic per agent, not by tensor throughput, which
// is why the legacy mma.sync path is enough here (no TMA / TMEM pipeline to amortise at K <= 64).
//
// mapf_gae_kernel: generalised advantage estimation as a backwards scan, one thread per (env, agent) series.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mapf_kernels.cuh"

namespace mapf {

constexpr int POL_H = 64;          // hidden width of both trunk layers
constexpr int POL_W2_STRIDE = 72;  // bf16 row stride of W2 / head weights in shared memory (64 + 8: conflict-free)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
    return *reinterpret_cast<const uint32_t *>(&v);
}

// per-warp shared memory of mapf_policy_act_kernel (bytes)
__host__ __device__ constexpr int pol_raw_bytes(int v2) { return (((32 * v2 + 3) / 4 + 2) * 4 + 15) & ~15; }   // raw window bytes + slack, 16-byte multiple
__host__ __device__ constexpr int pol_warp_bytes(int kc1, int v2) {
    return 32 * (16 * kc1 + 8) * 2 /* bf16 feature tile */ + pol_raw_bytes(v2) + 160 /* masks */ + 32 * 8 * 4 /* head outputs */;
}

// KC1 = k16 chunks of the first layer (F <= 16 * KC1), W1 row stride = 16 * KC1 + 8
template <int KC1>
__global__ void __launch_bounds__(256) mapf_policy_act_kernel(const mapf_policy_args a) {
    constexpr int S1 = 16 * KC1 + 8;   // bf16 row stride of the feature tile and of W1
    extern __shared__ __align__(16) unsigned char psm[];
    __nv_bfloat16 *w1 = reinterpret_cast<__nv_bfloat16 *>(psm);                 // [64][S1]
    __nv_bfloat16 *w2 = w1 + POL_H * S1;                                        // [64][72]
    __nv_bfloat16 *w3 = w2 + POL_H * POL_W2_STRIDE;                             // [8][72]: 5 logits, value, 2 zero rows
    float *bias = reinterpret_cast<float *>(w3 + 8 * POL_W2_STRIDE);            // b1[64] b2[64] b3[8]
    unsigned char *wsm0 = reinterpret_cast<unsigned char *>(bias + 2 * POL_H + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = blockDim.x >> 5;
    {   // weights: already bf16 and padded on the host side (mapf_policy_pack_weights)
        const uint4 *src = reinterpret_cast<const uint4 *>(a.weights);
        uint4 *dst = reinterpret_cast<uint4 *>(psm);
        const int n16 = (int)((wsm0 - psm) >> 4);
        for (int i = tid; i < n16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int V2 = a.v2, F = a.feature_dim, N = a.num_agents;
    unsigned char *wsm = wsm0 + warp * pol_warp_bytes(KC1, V2);
    __nv_bfloat16 *xs = reinterpret_cast<__nv_bfloat16 *>(wsm);                 // [32][S1] bf16 features
    uint32_t *raw = reinterpret_cast<uint32_t *>(wsm + 32 * S1 * 2);            // the tile's window bytes, as loaded
    uint32_t *rawm = reinterpret_cast<uint32_t *>(wsm + 32 * S1 * 2 + pol_raw_bytes(V2));   // 32 x 5 mask bytes
    float *os = reinterpret_cast<float *>(wsm + 32 * S1 * 2 + pol_raw_bytes(V2) + 160);     // [32][8] head outputs
    const int g = lane >> 2, t = lane & 3;
    const long long BN = (long long)a.num_envs * N;
    const long long ntiles = (BN + 31) >> 5;
    const int NW = (V2 + 3) >> 2;   // words of one agent's window
    for (long long tile = (long long)blockIdx.x * warps + warp; tile < ntiles; tile += (long long)gridDim.x * warps) {
        const long long ag0 = tile << 5;
        const long long left = BN - ag0;
        const int na = left < 32 ? (int)left : 32;
        // ---------------------------------------------------------------- raw channels of the tile, coalesced
        {
            const uint32_t *ob32 = reinterpret_cast<const uint32_t *>(a.local_obs + ag0 * V2);   // 32 * V2 bytes: 4-byte aligned
            const int nwords = (na * V2 + 3) >> 2;
            for (int i = lane; i < (32 * V2 + 3) / 4 + 2; i += 32) raw[i] = i < nwords ? ob32[i] : 0u;
            if (a.action_mask) {
                const uint32_t *mk32 = reinterpret_cast<const uint32_t *>(a.action_mask + ag0 * 5);
                const int mwords = (na * 5 + 3) >> 2;
                for (int i = lane; i < 40; i += 32) rawm[i] = i < mwords ? mk32[i] : 0u;
                if (a.action_mask_out) {   // the rollout buffer's copy of the masks (saves a separate copy launch per step)
                    uint32_t *mo32 = reinterpret_cast<uint32_t *>(a.action_mask_out + ag0 * 5);
                    // word copies when the destination row is 4-byte aligned (a [T,B,N,5] row may start anywhere)
                    const int full_words = (reinterpret_cast<uintptr_t>(a.action_mask_out) & 3) == 0 ? (na * 5) >> 2 : 0;
                    for (int i = lane; i < full_words; i += 32) mo32[i] = mk32[i];
                    for (int i = full_words * 4 + lane; i < na * 5; i += 32) a.action_mask_out[ag0 * 5 + i] = a.action_mask[ag0 * 5 + i];
                }
            }
        }
        __syncwarp();
        // ---------------------------------------------------------------- feature row of agent `lane` (ENV:306-328 order: window, goal delta, pressure)
        {
            uint2 *row = reinterpret_cast<uint2 *>(xs + lane * S1);
            const int off = lane * V2;
            for (int j = 0; j < 4 * KC1; ++j) {   // 4 cells -> 4 bf16 per trip; columns beyond the window are zero
                uint32_t w = 0;
                if (j < NW) {
                    const int b = off + 4 * j;
                    w = __funnelshift_r(raw[b >> 2], raw[(b >> 2) + 1], (b & 3) * 8);
                    if (4 * j + 4 > V2) w &= (1u << (8 * (V2 - 4 * j))) - 1u;
                }
                const __nv_bfloat162 lo = __floats2bfloat162_rn((float)(w & 0xFFu), (float)((w >> 8) & 0xFFu));
                const __nv_bfloat162 hi = __floats2bfloat162_rn((float)((w >> 16) & 0xFFu), (float)(w >> 24));
                row[j] = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
            }
            if (lane < na) {
                const float2 gd = reinterpret_cast<const float2 *>(a.goal_delta)[ag0 + lane];
                xs[lane * S1 + V2] = __float2bfloat16(gd.x);
                xs[lane * S1 + V2 + 1] = __float2bfloat16(gd.y);
                if (a.blocking_prev) xs[lane * S1 + V2 + 2] = __float2bfloat16((float)a.blocking_prev[ag0 + lane]);
            }
            if (a.features_out) {   // float32 feature block for the learner (optional)
                float *fo = a.features_out + ag0 * F;
                const uint8_t *rb = reinterpret_cast<const uint8_t *>(raw);
                for (int i = lane; i < na * F; i += 32) {
                    const int r = i / F, c = i - r * F;
                    float v;
                    if (c < V2) v = (float)rb[r * V2 + c];
                    else if (c < V2 + 2) v = a.goal_delta[(ag0 + r) * 2 + (c - V2)];
                    else v = (float)a.blocking_prev[ag0 + r];
                    fo[i] = v;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int row0 = 16 * mt + g;   // this lane's rows: row0 and row0 + 8
            // ------------------------------------------------------------ layer 1
            uint32_t af[KC1][4];
#pragma unroll
            for (int kk = 0; kk < KC1; ++kk) {
                const uint32_t *p0 = reinterpret_cast<const uint32_t *>(xs + row0 * S1 + 16 * kk + 2 * t);
                const uint32_t *p1 = reinterpret_cast<const uint32_t *>(xs + (row0 + 8) * S1 + 16 * kk + 2 * t);
                af[kk][0] = p0[0]; af[kk][1] = p1[0]; af[kk][2] = p0[4]; af[kk][3] = p1[4];
            }
            uint32_t h1[POL_H / 16][4];   // A fragments of layer 2
#pragma unroll
            for (int j = 0; j < POL_H / 8; ++j) {
                const float2 bb = *reinterpret_cast<const float2 *>(bias + 8 * j + 2 * t);
                float d[4] = {bb.x, bb.y, bb.x, bb.y};
#pragma unroll
                for (int kk = 0; kk < KC1; ++kk) {
                    const uint32_t *wb = reinterpret_cast<const uint32_t *>(w1 + (8 * j + g) * S1 + 16 * kk + 2 * t);
                    mma_bf16_16816(d, af[kk], wb[0], wb[4]);
                }
                h1[j >> 1][(j & 1) * 2 + 0] = pack_relu_bf16(d[0], d[1]);   // row g,   cols 8j + 2t, +1
                h1[j >> 1][(j & 1) * 2 + 1] = pack_relu_bf16(d[2], d[3]);   // row g+8
            }
            // ------------------------------------------------------------ layer 2
            uint32_t h2[POL_H / 16][4];
#pragma unroll
            for (int j = 0; j < POL_H / 8; ++j) {
                const float2 bb = *reinterpret_cast<const float2 *>(bias + POL_H + 8 * j + 2 * t);
                float d[4] = {bb.x, bb.y, bb.x, bb.y};
#pragma unroll
                for (int kk = 0; kk < POL_H / 16; ++kk) {
                    const uint32_t *wb = reinterpret_cast<const uint32_t *>(w2 + (8 * j + g) * POL_W2_STRIDE + 16 * kk + 2 * t);
                    mma_bf16_16816(d, h1[kk], wb[0], wb[4]);
                }
                h2[j >> 1][(j & 1) * 2 + 0] = pack_relu_bf16(d[0], d[1]);
                h2[j >> 1][(j & 1) * 2 + 1] = pack_relu_bf16(d[2], d[3]);
            }
            // ------------------------------------------------------------ heads: columns 0..4 logits, 5 value
            const float2 bb = *reinterpret_cast<const float2 *>(bias + 2 * POL_H + 2 * t);
            float d[4] = {bb.x, bb.y, bb.x, bb.y};
#pragma unroll
            for (int kk = 0; kk < POL_H / 16; ++kk) {
                const uint32_t *wb = reinterpret_cast<const uint32_t *>(w3 + g * POL_W2_STRIDE + 16 * kk + 2 * t);
                mma_bf16_16816(d, h2[kk], wb[0], wb[4]);
            }
            *reinterpret_cast<float2 *>(os + row0 * 8 + 2 * t) = make_float2(d[0], d[1]);
            *reinterpret_cast<float2 *>(os + (row0 + 8) * 8 + 2 * t) = make_float2(d[2], d[3]);
        }
        __syncwarp();
        // ---------------------------------------------------------------- one agent per lane: mask, softmax, draw
        if (lane < na) {
            const long long ag = ag0 + lane;
            const float4 o03 = *reinterpret_cast<const float4 *>(os + lane * 8);
            const float2 o45 = *reinterpret_cast<const float2 *>(os + lane * 8 + 4);
            float l[5] = {o03.x, o03.y, o03.z, o03.w, o45.x};
            if (!a.no_masking && a.action_mask) {   // logits + clamp(log(mask + 1e-6), FLOAT_MIN), action_mask_model.py:57-61
                const uint8_t *mk = reinterpret_cast<const uint8_t *>(rawm) + lane * 5;
#pragma unroll
                for (int k = 0; k < 5; ++k) l[k] += mk[k] ? 9.99999e-07f : -13.815511f;
            }
            float mx = l[0];
#pragma unroll
            for (int k = 1; k < 5; ++k) mx = fmaxf(mx, l[k]);
            float pr[5], sum = 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) { pr[k] = __expf(l[k] - mx); sum += pr[k]; }
            const unsigned env = (unsigned)((unsigned long long)ag / (unsigned)N);   // B * N < 2^32 in any batch that fits a GPU
            const int agent = (int)(ag - (long long)env * N);
            Philox ph(a.seed ^ 0x504F4C49ull /* "POLI" */, a.env_id_base + env);
            const uint4 x = ph((uint32_t)a.counter, (uint32_t)(a.counter >> 32), (uint32_t)agent, 0x53414D50u /* "SAMP" */);
            const float u = ((float)(x.x >> 8) + 0.5f) * (1.0f / 16777216.0f) * sum;   // uniform in (0, sum)
            int act = 0;
            float cum = pr[0];
#pragma unroll
            for (int k = 1; k < 5; ++k) { if (u >= cum) act = k; cum += pr[k]; }
            if (a.actions) a.actions[ag] = (int8_t)act;
            if (a.actions64) a.actions64[ag] = (long long)act;
            if (a.logp) a.logp[ag] = l[act] - mx - __logf(sum);
            if (a.value) a.value[ag] = o45.y;
            if (a.logits_out) {
#pragma unroll
                for (int k = 0; k < 5; ++k) a.logits_out[ag * 5 + k] = l[k];
            }
        }
        __syncwarp();
    }
}

// Generalised advantage estimation (ppo.py:104-117 hyper-parameters are arguments): adv[t] = delta_t + gamma * lam *
// nd_t * adv[t+1], delta_t = r_t + gamma * V_{t+1} * nd_t - V_t, nd_t = 1 - done_t; one thread per (env, agent).
__global__ void mapf_gae_kernel(const float *rewards, const float *values, const uint8_t *dones, const float *last_value,
                                float *adv, float *ret, int T, long long BN, int N, float gamma, float lam) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BN) return;
    const long long env = i / N, B = BN / N;
    float running = 0.f, next_value = last_value[i];
    for (int t = T - 1; t >= 0; --t) {
        const float nd = dones[(long long)t * B + env] ? 0.f : 1.f;
        const float v = values[(long long)t * BN + i];
        const float delta = rewards[(long long)t * BN + i] + gamma * next_value * nd - v;
        running = delta + gamma * lam * nd * running;
        adv[(long long)t * BN + i] = running;
        ret[(long long)t * BN + i] = running + v;
        next_value = v;
    }
}

}  // namespace mapf
