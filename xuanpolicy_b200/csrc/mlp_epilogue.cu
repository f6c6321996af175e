// mlp_epilogue.cu — the non-GEMM half of the MLP backward: LeakyReLU' fused with the bias-gradient reduction.
//
// The policy/value MLP GEMMs stay in torch (cuBLAS).  What torch autograd wraps around them in the backward of
// every Linear+LeakyReLU block (xuance/torch/utils/layers.py:15-21, representations/mlp.py:40-47,
// policies/categorical.py:26-32) is two more passes over the [B, H] activation gradient:
//     leaky_relu_backward  (read dy, y; write dz)              ~13.7 us at B=65536, H=128
//     dz.sum(0)            (read dz; column reduction -> db)    ~82 us  (strided reduce: 400 GB/s)
// This kernel does both in ONE pass: dz = dy * (y > 0 ? 1 : slope), db[h] = sum_b dz[b,h].
// Traffic per element: 4 B dy + 4 B y read, 4 B dz written = 12 B (the bias gradient is free).
// Mapping: a CTA owns a contiguous slab of rows; thread (r, c) walks rows r, r+R, ... of float4 column group c,
// so every warp access is a fully coalesced 512 B row segment; column partials are reduced through shared memory,
// written per CTA, and the last CTA to finish sums the per-CTA partials in a fixed order (deterministic).
#include "common.cuh"

namespace xb {

constexpr int kEpiBlock = 256;
constexpr int kEpiUnroll = 4;

// y[b,h] = leaky_relu(y[b,h] + bias[h]) in place: the forward epilogue of Linear+LeakyReLU after a bias-free cuBLAS mm
// (torch's addmm runs a separate 36 us bias kernel plus a 10 us activation kernel at [65536,128]; this is one pass).
__global__ void __launch_bounds__(kEpiBlock)
    bias_act_fwd_kernel(float4* __restrict__ y, const float4* __restrict__ bias, float slope, int64_t n4, int H4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // blockDim.x is a multiple of H4 (checked by the host), so a thread always lands on the same column group
    const float4 b = bias[threadIdx.x % H4];
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * kEpiUnroll) {
        float4 v[kEpiUnroll];
#pragma unroll
        for (int u = 0; u < kEpiUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n4) v[u] = y[i];
        }
#pragma unroll
        for (int u = 0; u < kEpiUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n4) {
                float4 o;
                o.x = v[u].x + b.x; o.y = v[u].y + b.y; o.z = v[u].z + b.z; o.w = v[u].w + b.w;
                o.x = o.x > 0.f ? o.x : o.x * slope;
                o.y = o.y > 0.f ? o.y : o.y * slope;
                o.z = o.z > 0.f ? o.z : o.z * slope;
                o.w = o.w > 0.f ? o.w : o.w * slope;
                y[i] = o;
            }
        }
    }
}

__device__ __forceinline__ float4 lrelu_bwd4(const float4 g, const float4 o, float slope) {
    float4 z;
    z.x = o.x > 0.f ? g.x : g.x * slope;
    z.y = o.y > 0.f ? g.y : g.y * slope;
    z.z = o.z > 0.f ? g.z : g.z * slope;
    z.w = o.w > 0.f ? g.w : g.w * slope;
    return z;
}

__global__ void __launch_bounds__(kEpiBlock)
    act_bias_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, float slope,
                        float4* __restrict__ dz, float* __restrict__ dbias, float* __restrict__ partials,
                        unsigned int* __restrict__ ticket, int64_t B, int H4) {
    extern __shared__ float4 red[];  // [rows_per_pass][H4]
    __shared__ bool is_last;
    const int rows_per_pass = kEpiBlock / H4;
    const int c = threadIdx.x % H4, r = threadIdx.x / H4;
    const int64_t rows_per_cta = (B + gridDim.x - 1) / gridDim.x;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t row1 = row0 + rows_per_cta < B ? row0 + rows_per_cta : B;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_per_pass) {
        for (int64_t b0 = row0 + r; b0 < row1; b0 += (int64_t)rows_per_pass * kEpiUnroll) {
            float4 g[kEpiUnroll], o[kEpiUnroll];
#pragma unroll
            for (int u = 0; u < kEpiUnroll; ++u) {       // all loads first: 2*kEpiUnroll 16-byte requests in flight
                const int64_t b = b0 + (int64_t)u * rows_per_pass;
                if (b < row1) { g[u] = dy[b * H4 + c]; o[u] = y[b * H4 + c]; }
            }
#pragma unroll
            for (int u = 0; u < kEpiUnroll; ++u) {
                const int64_t b = b0 + (int64_t)u * rows_per_pass;
                if (b < row1) {
                    const float4 z = lrelu_bwd4(g[u], o[u], slope);
                    dz[b * H4 + c] = z;
                    acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
                }
            }
        }
        red[r * H4 + c] = acc;
    }
    __syncthreads();
    if (threadIdx.x < H4) {
        float4 s = red[threadIdx.x];
        for (int k = 1; k < rows_per_pass; ++k) {
            const float4 t = red[k * H4 + threadIdx.x];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        reinterpret_cast<float4*>(partials)[(int64_t)blockIdx.x * H4 + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {   // all threads of the last CTA reduce the per-CTA partials: thread (r, c) takes CTAs r, r+R, ...
        __threadfence();
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows_per_pass) {
            const float4* pp = reinterpret_cast<const float4*>(partials);
            int g = r;
            for (; g + 7 * rows_per_pass < (int)gridDim.x; g += 8 * rows_per_pass) {   // 8 independent loads in flight
                float4 t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = pp[(int64_t)(g + u * rows_per_pass) * H4 + c];
#pragma unroll
                for (int u = 0; u < 8; ++u) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
            }
            for (; g < (int)gridDim.x; g += rows_per_pass) {
                const float4 t = pp[(int64_t)g * H4 + c];
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            red[r * H4 + c] = s;
        }
        __syncthreads();
        if (threadIdx.x < H4) {
            float4 t = red[threadIdx.x];
            for (int k = 1; k < rows_per_pass; ++k) {
                const float4 u = red[k * H4 + threadIdx.x];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            reinterpret_cast<float4*>(dbias)[threadIdx.x] = t;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

}  // namespace xb

using namespace xb;

extern "C" int xb_bias_act_fwd(float* y, const float* bias, float slope, int64_t B, int H, xb_stream_t stream) {
    if (B <= 0 || !y || !bias) return XB_E_BADARG;
    if (H % 4 != 0 || H < 4 || kEpiBlock % (H / 4) != 0) return XB_E_UNSUPPORTED;
    if ((((uintptr_t)y) | ((uintptr_t)bias)) & 15u) return XB_E_BADARG;  // float4 accesses
    const int64_t n4 = B * (H / 4);
    bias_act_fwd_kernel<<<grid_for((n4 + kEpiUnroll - 1) / kEpiUnroll, kEpiBlock, 8), kEpiBlock, 0, (cudaStream_t)stream>>>(
        (float4*)y, (const float4*)bias, slope, n4, H / 4);
    XB_LAUNCH_CHECK();
    return 0;
}

// workspace: fp32 [4 + 592 * H]; word 0 is the ticket (zero-initialised by the caller)
extern "C" int xb_act_bias_bwd(const float* dy, const float* y, float slope, float* dz, float* dbias, float* workspace,
                               int64_t B, int H, xb_stream_t stream) {
    if (B <= 0 || !dy || !y || !dz || !dbias || !workspace) return XB_E_BADARG;
    if (H % 4 != 0 || H < 4 || H / 4 > kEpiBlock) return XB_E_UNSUPPORTED;
    if ((((uintptr_t)dy) | ((uintptr_t)y) | ((uintptr_t)dz) | ((uintptr_t)dbias) | ((uintptr_t)workspace)) & 15u)
        return XB_E_BADARG;  // float4 accesses
    const int H4 = H / 4;
    const int rows_per_pass = kEpiBlock / H4;
    int grid = kNumSMs * 2;
    const int64_t need = (B + rows_per_pass * kEpiUnroll * 2 - 1) / (rows_per_pass * kEpiUnroll * 2);
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    const size_t smem = (size_t)rows_per_pass * H4 * sizeof(float4);
    act_bias_bwd_kernel<<<grid, kEpiBlock, smem, (cudaStream_t)stream>>>(
        (const float4*)dy, (const float4*)y, slope, (float4*)dz, dbias, workspace + 4, (unsigned int*)workspace, B, H4);
    XB_LAUNCH_CHECK();
    return 0;
}
