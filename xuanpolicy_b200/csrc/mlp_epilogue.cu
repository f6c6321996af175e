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
        for (int64_t b = row0 + r; b < row1; b += rows_per_pass) {
            const float4 g = dy[b * H4 + c], o = y[b * H4 + c];
            float4 z;
            z.x = o.x > 0.f ? g.x : g.x * slope;
            z.y = o.y > 0.f ? g.y : g.y * slope;
            z.z = o.z > 0.f ? g.z : g.z * slope;
            z.w = o.w > 0.f ? g.w : g.w * slope;
            dz[b * H4 + c] = z;
            acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
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
    if (is_last) {
        __threadfence();
        for (int h = threadIdx.x; h < H4 * 4; h += blockDim.x) {
            float s = 0.f;
            for (int g = 0; g < (int)gridDim.x; ++g) s += partials[(int64_t)g * H4 * 4 + h];
            dbias[h] = s;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

}  // namespace xb

using namespace xb;

// workspace: fp32 [1 + grid_max * H] with grid_max = 592; word 0 is the ticket (zero-initialised by the caller)
extern "C" int xb_act_bias_bwd(const float* dy, const float* y, float slope, float* dz, float* dbias, float* workspace,
                               int64_t B, int H, xb_stream_t stream) {
    if (B <= 0 || !dy || !y || !dz || !dbias || !workspace) return XB_E_BADARG;
    if (H % 4 != 0 || H < 4 || H / 4 > kEpiBlock) return XB_E_UNSUPPORTED;
    const int H4 = H / 4;
    const int rows_per_pass = kEpiBlock / H4;
    int grid = kNumSMs * 4;
    const int64_t need = (B + rows_per_pass * 8 - 1) / (rows_per_pass * 8);  // at least ~8 passes per CTA
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    const size_t smem = (size_t)rows_per_pass * H4 * sizeof(float4);
    act_bias_bwd_kernel<<<grid, kEpiBlock, smem, (cudaStream_t)stream>>>(
        (const float4*)dy, (const float4*)y, slope, (float4*)dz, dbias, workspace + 4, (unsigned int*)workspace, B, H4);
    XB_LAUNCH_CHECK();
    return 0;
}
