// mlp_trunk.cu — the first (narrow-input) layer of the policy/value MLP and its weight gradient.
//
// Basic_MLP's only layer, Linear(obs_dim, H) + LeakyReLU (xuance/torch/representations/mlp.py:40-51), has a 3- or
// 4-wide reduction: it is a bandwidth-bound broadcast, not a GEMM, so it runs on the SIMT pipes:
//   forward   h1[b][n] = leaky(b0[n] + sum_i obs[b][i] W0[n][i])            writes B x H floats, reads B x obs_dim
//   backward  dW0[n][i] = sum_b dz1[b][n] obs[b][i],  db0[n] = sum_b dz1[b][n]   reads B x H floats once
// (dz1 comes from the dgrad tensor-core kernel; torch would run a [H x B] x [B x obs_dim] sgemm + a column reduction).
#include "common.cuh"

namespace xb {

constexpr int kMaxObs = 8;

// one thread = 4 consecutive output features of one row; a group of H/4 threads covers a row
template <int OBS>
__global__ void __launch_bounds__(256) trunk_fwd_kernel(const float* __restrict__ obs, int ld, const float* __restrict__ W0,
                                                        const float* __restrict__ b0, float slope, float4* __restrict__ h1,
                                                        int64_t B, int H, uint32_t* __restrict__ h1_signs) {
    const int tpr = H >> 2;                        // threads per row
    const int rows_per_block = blockDim.x / tpr;
    const int q = threadIdx.x % tpr, slot = threadIdx.x / tpr;
    if (slot >= rows_per_block) return;
    const int lane = threadIdx.x & 31;             // (h1_signs: tpr is a multiple of 32, so q % 8 == lane % 8 and a row's threads
    float w[4][OBS], bias[4];                      //  fill whole warps — the shuffles below stay inside one row)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        bias[e] = b0[4 * q + e];
#pragma unroll
        for (int i = 0; i < OBS; ++i) w[e][i] = W0[(4 * q + e) * OBS + i];
    }
    for (int64_t b = (int64_t)blockIdx.x * rows_per_block + slot; b < B; b += (int64_t)gridDim.x * rows_per_block) {
        float x[OBS];
#pragma unroll
        for (int i = 0; i < OBS; ++i) x[i] = __ldg(obs + b * ld + i);
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float a = bias[e];
#pragma unroll
            for (int i = 0; i < OBS; ++i) a += x[i] * w[e][i];
            o[e] = a > 0.f ? a : a * slope;
        }
        h1[b * tpr + q] = make_float4(o[0], o[1], o[2], o[3]);
        if (h1_signs) {         // bit l of word e of a 128-feature slice = (h1[128 slice + 4 l + e] > 0), as xb_gather_trunk_fwd
            const uint32_t b0_ = __ballot_sync(0xffffffffu, o[0] > 0.f), b1_ = __ballot_sync(0xffffffffu, o[1] > 0.f);
            const uint32_t b2_ = __ballot_sync(0xffffffffu, o[2] > 0.f), b3_ = __ballot_sync(0xffffffffu, o[3] > 0.f);
            if (lane < 4)
                h1_signs[b * (H >> 5) + (q >> 5) * 4 + lane] = lane == 0 ? b0_ : (lane == 1 ? b1_ : (lane == 2 ? b2_ : b3_));
        }
    }
}

// partial[blockIdx][n][OBS + 1] = sum over this block's rows; a second kernel adds the blocks in order
template <int OBS>
__global__ void __launch_bounds__(256) trunk_wgrad_kernel(const float4* __restrict__ dz1, const float* __restrict__ obs,
                                                          int ld, float* __restrict__ partial, int64_t B, int H) {
    extern __shared__ float red[];                 // [rows_per_block][H][OBS + 1]
    pdl_wait();                                    // (launched with the programmatic-serialization attribute, common.cuh)
    pdl_trigger();
    const int tpr = H >> 2;
    const int rows_per_block = blockDim.x / tpr;
    const int q = threadIdx.x % tpr, slot = threadIdx.x / tpr;
    float acc[4][OBS + 1];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int i = 0; i <= OBS; ++i) acc[e][i] = 0.f;
    if (slot < rows_per_block) {
        const int64_t step = (int64_t)gridDim.x * rows_per_block;
        constexpr int U = 8;                        // independent rows in flight per thread
        for (int64_t b0 = (int64_t)blockIdx.x * rows_per_block + slot; b0 < B; b0 += U * step) {
            float4 d[U];
            float x[U][OBS];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t b = b0 + u * step;
                const bool ok = b < B;
                d[u] = ok ? dz1[b * tpr + q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < OBS; ++i) x[u][i] = ok ? __ldg(obs + b * ld + i) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#pragma unroll
                    for (int i = 0; i < OBS; ++i) acc[e][i] += dv[e] * x[u][i];
                    acc[e][OBS] += dv[e];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int i = 0; i <= OBS; ++i) red[(slot * H + 4 * q + e) * (OBS + 1) + i] = acc[e][i];
    }
    __syncthreads();
    const int n_out = H * (OBS + 1);
    for (int k = threadIdx.x; k < n_out; k += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < rows_per_block; ++r) s += red[r * n_out + k];
        partial[(int64_t)blockIdx.x * n_out + k] = s;
    }
}

// one warp per output element: lanes stride over the per-block partials, then a fixed-order shuffle tree (deterministic)
__global__ void __launch_bounds__(256) trunk_wgrad_reduce_kernel(const float* __restrict__ partial, int n_blocks, int H,
                                                                 int OBS, float* __restrict__ dW0,
                                                                 float* __restrict__ db0) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n_out = H * (OBS + 1);
    if (k >= n_out) return;
    float s = 0.f;
    for (int p = lane; p < n_blocks; p += 32) s += partial[(int64_t)p * n_out + k];
    s = warp_sum(s);
    if (lane == 0) {
        const int n = k / (OBS + 1), i = k % (OBS + 1);
        if (i < OBS) dW0[n * OBS + i] = s;
        else db0[n] = s;
    }
}

constexpr int kTrunkBlocks = 3 * kNumSMs;

}  // namespace xb

using namespace xb;

#define XB_OBS_SWITCH(OBS, CALL)        \
    switch (OBS) {                      \
        case 1: { constexpr int O = 1; CALL; } break; \
        case 2: { constexpr int O = 2; CALL; } break; \
        case 3: { constexpr int O = 3; CALL; } break; \
        case 4: { constexpr int O = 4; CALL; } break; \
        case 5: { constexpr int O = 5; CALL; } break; \
        case 6: { constexpr int O = 6; CALL; } break; \
        case 7: { constexpr int O = 7; CALL; } break; \
        case 8: { constexpr int O = 8; CALL; } break; \
        default: return XB_E_UNSUPPORTED;             \
    }

static inline bool trunk_shape_ok(int obs_dim, int H) {
    return obs_dim >= 1 && obs_dim <= kMaxObs && H % 4 == 0 && H >= 4 && 256 % (H / 4) == 0;
}

extern "C" int xb_mlp_trunk_fwd(const float* obs, int ld, int obs_dim, const float* W0, const float* b0, float slope,
                                float* h1, int64_t B, int H, uint32_t* h1_signs, xb_stream_t stream) {
    if (!obs || !W0 || !b0 || !h1 || B <= 0 || ld < obs_dim) return XB_E_BADARG;
    if (!trunk_shape_ok(obs_dim, H) || ((uintptr_t)h1 & 15u)) return XB_E_UNSUPPORTED;
    if (h1_signs && H % 128 != 0) return XB_E_UNSUPPORTED;       // (a row's threads must fill whole warps)
    const int rows_per_block = 256 / (H / 4);
    const int grid = grid_for((B + rows_per_block - 1) / rows_per_block * 256, 256, 4);
    XB_OBS_SWITCH(obs_dim, (trunk_fwd_kernel<O><<<grid, 256, 0, (cudaStream_t)stream>>>(
                               obs, ld, W0, b0, slope, reinterpret_cast<float4*>(h1), B, H, h1_signs)));
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_mlp_trunk_wgrad_workspace_floats(int obs_dim, int H) { return kTrunkBlocks * H * (obs_dim + 1); }
extern "C" int xb_mlp_trunk_wgrad_parts(void) { return kTrunkBlocks; }

extern "C" int xb_mlp_trunk_wgrad(const float* dz1, const float* obs, int ld, int obs_dim, float* workspace, float* dW0,
                                  float* db0, int64_t B, int H, xb_stream_t stream) {
    if (!dz1 || !obs || !workspace || (dW0 && !db0) || B <= 0 || ld < obs_dim) return XB_E_BADARG;
    if (!trunk_shape_ok(obs_dim, H) || ((uintptr_t)dz1 & 15u)) return XB_E_UNSUPPORTED;
    const int rows_per_block = 256 / (H / 4);
    const int smem = rows_per_block * H * (obs_dim + 1) * (int)sizeof(float);
    if (smem > 48 * 1024) return XB_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaSuccess;
    XB_OBS_SWITCH(obs_dim, (err = launch_pdl(trunk_wgrad_kernel<O>, dim3(kTrunkBlocks), dim3(256), (size_t)smem, s, true,
                                             reinterpret_cast<const float4*>(dz1), obs, ld, workspace, B, H)));
    XB_CUDA(err);
    XB_LAUNCH_CHECK();
    if (!dW0) return 0;                      // partials only: the caller finishes with xb_mlp_backward_tail
    const int n_out = H * (obs_dim + 1);
    trunk_wgrad_reduce_kernel<<<(n_out + 7) / 8, 256, 0, s>>>(workspace, kTrunkBlocks, H, obs_dim, dW0, db0);
    XB_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------------ 3-logit heads
// A softmax over (z0, z1, z2) equals the softmax over (z0 - z2, z1 - z2, 0): log-probabilities, entropy and every gradient of a
// loss that depends on the logits through the softmax are unchanged, dL/dz2 = -(dL/dz0 + dL/dz1), and the hidden-layer
// gradient sum_j dL/dz_j w_j equals dL/dz0 (w0 - w2) + dL/dz1 (w1 - w2).  So a Discrete(3) actor head (MountainCar-v0,
// Acrobot-v1: xuance/torch/policies/categorical.py:26-32 is generic in action_dim) runs on the two-head dense kernels with
// FOLDED head parameters; its third logit is reported as 0 (logits are defined up to a per-row constant).
namespace xb {
__global__ void head3_fold_kernel(const float* __restrict__ w3, const float* __restrict__ b3, float* __restrict__ w2,
                                  float* __restrict__ b2, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H) {
        const float c = w3[2 * H + i];
        w2[i] = w3[i] - c;
        w2[H + i] = w3[H + i] - c;
    }
    if (i < 2) b2[i] = b3[i] - b3[2];
}
// rows 0, 1 of gw3 / entries 0, 1 of gb3 hold the two-head gradients; row 2 = -(row 0 + row 1)
__global__ void head3_unfold_grads_kernel(float* __restrict__ gw3, float* __restrict__ gb3, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H) gw3[2 * H + i] = -(gw3[i] + gw3[H + i]);
    if (i == 0) gb3[2] = -(gb3[0] + gb3[1]);
}
}  // namespace xb

extern "C" int xb_head3_fold(const float* w3, const float* b3, float* w2, float* b2, int H, xb_stream_t stream) {
    if (!w3 || !b3 || !w2 || !b2 || H < 2) return XB_E_BADARG;
    xb::head3_fold_kernel<<<(H + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w3, b3, w2, b2, H);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_head3_unfold_grads(float* gw3, float* gb3, int H, xb_stream_t stream) {
    if (!gw3 || !gb3 || H < 1) return XB_E_BADARG;
    xb::head3_unfold_grads_kernel<<<(H + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gw3, gb3, H);
    XB_LAUNCH_CHECK();
    return 0;
}
