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

// ---- narrow output heads (A <= 4 columns): bandwidth-bound matrix-vector work, not a GEMM ---------------------
// The last Linear of the actor (A = 1..4 outputs) and the critic (1 output) reads a [B, H] activation once and
// produces A numbers per row.  cuBLAS serves these shapes with gemv-style kernels at ~1.6 TB/s and torch autograd
// adds an outer-product kernel, a gemv, three reductions and two adds in the backward.  Here:
//   head_fwd:      out[b,a] = h[b,:] . W[a,:] + bias[a]                                  (one pass over h)
//   head_bwd_act:  dz[b,:]  = (sum_a dout[b,a] W[a,:]) * lrelu'(y[b,:])   (gradient w.r.t. the hidden pre-activation)
//                  db1[:]   = sum_b dz[b,:]      dW2[a,:] = sum_b dout[b,a] y[b,:]      db2[a] = sum_b dout[b,a]
//                  — the head's backward fused with the hidden layer's activation backward: y is read once,
//                  dz written once, the intermediate dL/dy is never materialised.
constexpr int kHeadMaxA = 4;
constexpr int kHeadMaxC = 4;   // float4 column groups per thread in head_fwd (H <= 512)
constexpr int kHeadBlock = 512;

constexpr int kHeadRows = 4;   // rows in flight per lane group

__global__ void __launch_bounds__(kEpiBlock)
    head_fwd_kernel(const float4* __restrict__ h, const float4* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ out, int64_t B, int H4, int A, int G) {
    // a group of G lanes (power of two <= 32, H4 % G == 0) owns kHeadRows rows per iteration; lane j covers column
    // groups j, j+G, ...; all row loads are issued before any arithmetic, then one shuffle tree per (row, output)
    const int lane = threadIdx.x & (G - 1);
    const int per = H4 / G;
    float4 w[kHeadMaxA][kHeadMaxC];
#pragma unroll
    for (int a = 0; a < kHeadMaxA; ++a)
#pragma unroll
        for (int k = 0; k < kHeadMaxC; ++k)
            w[a][k] = (a < A && k < per) ? W[a * H4 + lane + k * G] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int sub = (threadIdx.x & 31) / G;                 // which of the warp's 32/G lane groups
    // warp-uniform trip count (full-mask shuffles below): iterate on the warp's first lane group
    for (int64_t bw = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G; bw < B; bw += groups * kHeadRows) {
        float4 v[kHeadRows][kHeadMaxC];
#pragma unroll
        for (int u = 0; u < kHeadRows; ++u) {
            const int64_t b = bw + sub + (int64_t)u * groups;
            const int64_t row = b < B ? b : B - 1;
#pragma unroll
            for (int k = 0; k < kHeadMaxC; ++k)
                if (k < per) v[u][k] = h[row * H4 + lane + k * G];
        }
#pragma unroll
        for (int u = 0; u < kHeadRows; ++u) {
            const int64_t b = bw + sub + (int64_t)u * groups;
#pragma unroll
            for (int a = 0; a < kHeadMaxA; ++a) {
                if (a < A) {                                  // A is uniform: no shuffles for unused outputs
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < kHeadMaxC; ++k)
                        if (k < per) acc += v[u][k].x * w[a][k].x + v[u][k].y * w[a][k].y + v[u][k].z * w[a][k].z + v[u][k].w * w[a][k].w;
                    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0 && b < B) out[b * A + a] = acc + bias[a];
                }
            }
        }
    }
}

__device__ __forceinline__ void add4(float4& s, const float4 t) { s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
__device__ __forceinline__ void fma4(float4& s, float k, const float4 t) { s.x += k * t.x; s.y += k * t.y; s.z += k * t.z; s.w += k * t.w; }

// partial record per CTA (float4 slots): [0, H4) db1 | [H4, (1+A) H4) dW2 rows | slot (1+A) H4: db2 (padded to 4)
__global__ void __launch_bounds__(kHeadBlock)
    head_bwd_act_kernel(const float* __restrict__ dout, const float4* __restrict__ y, const float4* __restrict__ W2,
                        float slope, float4* __restrict__ dz, float* __restrict__ db1, float* __restrict__ dW2,
                        float* __restrict__ db2, float4* __restrict__ partials, unsigned int* __restrict__ ticket,
                        int64_t B, int H4, int A) {
    extern __shared__ float4 red[];  // kHeadBlock float4
    __shared__ bool is_last;
    const int R = kHeadBlock / H4;
    const int c = threadIdx.x % H4, r = threadIdx.x / H4;
    const int64_t rows_per_cta = (B + gridDim.x - 1) / gridDim.x;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t row1 = row0 + rows_per_cta < B ? row0 + rows_per_cta : B;
    const int n4 = (1 + A) * H4 + 1;             // float4 slots per partial record
    float4 w[kHeadMaxA];
#pragma unroll
    for (int a = 0; a < kHeadMaxA; ++a) w[a] = a < A ? W2[a * H4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc_w[kHeadMaxA];
#pragma unroll
    for (int a = 0; a < kHeadMaxA; ++a) acc_w[a] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc_d = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t b0 = row0 + r; b0 < row1; b0 += (int64_t)R * kEpiUnroll) {
        float4 o[kEpiUnroll];
        float g[kEpiUnroll][kHeadMaxA];
#pragma unroll
        for (int u = 0; u < kEpiUnroll; ++u) {
            const int64_t b = b0 + (int64_t)u * R;
            if (b < row1) {
                o[u] = y[b * H4 + c];
#pragma unroll
                for (int a = 0; a < kHeadMaxA; ++a) g[u][a] = a < A ? dout[b * A + a] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < kEpiUnroll; ++u) {
            const int64_t b = b0 + (int64_t)u * R;
            if (b < row1) {
                float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int a = 0; a < kHeadMaxA; ++a) {
                    fma4(dh, g[u][a], w[a]);
                    fma4(acc_w[a], g[u][a], o[u]);
                }
                if (c == 0) add4(acc_d, make_float4(g[u][0], g[u][1], g[u][2], g[u][3]));
                const float4 z = lrelu_bwd4(dh, o[u], slope);
                dz[b * H4 + c] = z;
                add4(acc_b, z);
            }
        }
    }
    // block reduction over the R row lanes, field by field; one partial record per CTA
    float4* my = partials + (int64_t)blockIdx.x * n4;
#pragma unroll
    for (int f = 0; f <= kHeadMaxA; ++f) {
        if (f <= A) {                                 // A is block-uniform, so the barriers are too
            red[r * H4 + c] = (f == 0) ? acc_b : acc_w[f == 0 ? 0 : f - 1];
            __syncthreads();
            if (threadIdx.x < H4) {
                float4 s = red[threadIdx.x];
                for (int k = 1; k < R; ++k) add4(s, red[k * H4 + threadIdx.x]);
                my[f * H4 + threadIdx.x] = s;
            }
            __syncthreads();
        }
    }
    if (c == 0) red[r] = acc_d;
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 s = red[0];
        for (int k = 1; k < R; ++k) add4(s, red[k]);
        my[(1 + A) * H4] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {   // deterministic final reduction: thread (slice, slot) sums CTAs slice, slice+S, ... of one slot
        __threadfence();
        for (int base = 0; base < n4; base += kHeadBlock) {
            const int width = n4 - base < kHeadBlock ? n4 - base : kHeadBlock;   // slots handled this round
            const int S = kHeadBlock / width;                                      // CTA slices per slot
            const int slot = base + (int)threadIdx.x % width, slice = (int)threadIdx.x / width;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (slice < S) {
                int gi = slice;
                for (; gi + 3 * S < (int)gridDim.x; gi += 4 * S) {
                    float4 t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) t[u] = partials[(int64_t)(gi + u * S) * n4 + slot];
#pragma unroll
                    for (int u = 0; u < 4; ++u) add4(s, t[u]);
                }
                for (; gi < (int)gridDim.x; gi += S) add4(s, partials[(int64_t)gi * n4 + slot]);
            }
            __syncthreads();
            if (slice < S) red[slice * width + (slot - base)] = s;
            __syncthreads();
            if ((int)threadIdx.x < width) {
                float4 t = red[threadIdx.x];
                for (int k = 1; k < S; ++k) add4(t, red[k * width + threadIdx.x]);
                const int sl = base + threadIdx.x;
                if (sl < H4) {
                    reinterpret_cast<float4*>(db1)[sl] = t;
                } else if (sl < (1 + A) * H4) {
                    reinterpret_cast<float4*>(dW2)[sl - H4] = t;
                } else {
                    const float v[4] = {t.x, t.y, t.z, t.w};
                    for (int a = 0; a < A; ++a) db2[a] = v[a];
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

}  // namespace xb

using namespace xb;

extern "C" int xb_head_fwd(const float* h, const float* W, const float* bias, float* out, int64_t B, int H, int A,
                           xb_stream_t stream) {
    if (B <= 0 || !h || !W || !bias || !out) return XB_E_BADARG;
    if (H % 4 != 0 || H < 4 || A < 1 || A > kHeadMaxA) return XB_E_UNSUPPORTED;
    if ((((uintptr_t)h) | ((uintptr_t)W)) & 15u) return XB_E_BADARG;
    const int H4 = H / 4;
    int G = 32;
    while (G > 1 && (H4 % G != 0)) G >>= 1;
    if (H4 / G > kHeadMaxC) return XB_E_UNSUPPORTED;
    const int64_t threads = B * G;
    head_fwd_kernel<<<grid_for(threads, kEpiBlock, 8), kEpiBlock, 0, (cudaStream_t)stream>>>(
        (const float4*)h, (const float4*)W, bias, out, B, H4, A, G);
    XB_LAUNCH_CHECK();
    return 0;
}

// workspace: fp32 [4 + 148 * ((1 + A) * H + 4)], word 0 = ticket, ZERO-INITIALISED once by the caller
extern "C" int xb_head_bwd_act(const float* dout, const float* y, const float* W2, float slope, float* dz, float* db1,
                               float* dW2, float* db2, float* workspace, int64_t B, int H, int A, xb_stream_t stream) {
    if (B <= 0 || !dout || !y || !W2 || !dz || !db1 || !dW2 || !db2 || !workspace) return XB_E_BADARG;
    if (H % 4 != 0 || H < 4 || kHeadBlock % (H / 4) != 0 || A < 1 || A > kHeadMaxA) return XB_E_UNSUPPORTED;
    if ((((uintptr_t)y) | ((uintptr_t)W2) | ((uintptr_t)dz) | ((uintptr_t)db1) | ((uintptr_t)dW2) | ((uintptr_t)workspace)) & 15u)
        return XB_E_BADARG;
    const int H4 = H / 4;
    const int R = kHeadBlock / H4;
    int grid = kNumSMs;                                  // one 512-thread CTA per SM: 148 partial records
    const int64_t need = (B + R * kEpiUnroll - 1) / (R * kEpiUnroll);
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    head_bwd_act_kernel<<<grid, kHeadBlock, kHeadBlock * sizeof(float4), (cudaStream_t)stream>>>(
        dout, (const float4*)y, (const float4*)W2, slope, (float4*)dz, db1, dW2, db2, (float4*)(workspace + 4),
        (unsigned int*)workspace, B, H4, A);
    XB_LAUNCH_CHECK();
    return 0;
}


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
