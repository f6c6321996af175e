// optim.cu — optimiser tail of PPOCLIP_Learner.update on ONE flat fp32 parameter/gradient buffer.
//
// Replaces (xuance/torch/learners/policy_gradient/ppoclip_learner.py:47-51)
//     torch.nn.utils.clip_grad_norm_(policy.parameters(), clip_grad_norm)
//     optimizer.step()      torch.optim.Adam(lr, eps=1e-5)            (runner_drl.py:71)
//     scheduler.step()      LinearLR(start 1.0 -> end 0.0, total_iters) (runner_drl.py:72-73)
// which in torch is ~20 small foreach launches plus host-side Python per update.  Here: two launches, no host
// arithmetic, the update counter / learning rate live on the device, so the whole update is graph-replayable.
//
// Kernel 1 reduces sum((grad*grad_scale)^2) (per-block partials, last block finishes deterministically) and
// derives the step scalars; kernel 2 applies clip + Adam element-wise.  Traffic: 4 B/param read in pass 1,
// 16 B read + 12 B written per param in pass 2 (params are 9-133 K floats here: launch-latency bound).
#include "optim.cuh"

namespace xb {

constexpr int kOptBlock = 256;

__global__ void __launch_bounds__(kOptBlock)
    grad_norm_kernel(const float* __restrict__ grad, int64_t n, int64_t* __restrict__ step_dev, AdamHyper h,
                     double* __restrict__ ws, float* __restrict__ lr_out, float* __restrict__ gnorm_out) {
    __shared__ double smem[32];
    __shared__ bool is_last;
    double sq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double g = (double)(grad[i] * h.grad_scale);
        sq += g * g;
    }
    grad_norm_finish(sq, step_dev, h, ws, lr_out, gnorm_out, smem, &is_last);
}

__global__ void __launch_bounds__(kOptBlock)
    adam_apply_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                      float* __restrict__ exp_avg_sq, int64_t n, AdamHyper h, const double* __restrict__ ws) {
    pdl_wait();
    pdl_trigger();
    const float gscale = h.grad_scale * (float)ws[1];
    const float step_size = (float)(ws[2] / ws[3]);
    const float bc2_sqrt = (float)ws[4];
    const float b1 = h.beta1, b2 = h.beta2, eps = h.eps;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i] * gscale;
        float m = exp_avg[i], v = exp_avg_sq[i];
        m = m + (g - m) * (1.0f - b1);           // exp_avg.lerp_(grad, 1 - beta1)
        v = v * b2 + (1.0f - b2) * g * g;        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        param[i] = param[i] - step_size * (m / denom);
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
    }
}

// adam_apply + the tf32 hi/lo operand split of the two H x H hidden-layer weights (dense_tc.cu split_weights_kernel) in one
// launch: the split of an updated weight is written right where the weight is updated, so the next forward / dgrad find
// their 3xTF32 operands ready and the per-update split launch disappears.
struct AdamSplit {
    int64_t off[2];   // element offset of W_s [N][K] inside the flat parameter buffer
    float* hi[2];
    float* lo[2];
    int toff[2];      // column offset inside the transposed, concatenated operand
    float* thi;
    float* tlo;
    int N, K, ldt;
};

__device__ __forceinline__ void split_tf32_opt(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

__global__ void __launch_bounds__(kOptBlock)
    adam_apply_split_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                            float* __restrict__ exp_avg_sq, int64_t n, AdamHyper h, const double* __restrict__ ws, AdamSplit sp) {
    pdl_wait();                                 // (launched with the programmatic-serialization attribute, common.cuh)
    pdl_trigger();
    const float gscale = h.grad_scale * (float)ws[1];
    const float step_size = (float)(ws[2] / ws[3]);
    const float bc2_sqrt = (float)ws[4];
    const float b1 = h.beta1, b2 = h.beta2, eps = h.eps;
    const int64_t nk = (int64_t)sp.N * sp.K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = grad[i] * gscale;
        float m = exp_avg[i], v = exp_avg_sq[i];
        m = m + (g - m) * (1.0f - b1);
        v = v * b2 + (1.0f - b2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        const float p = param[i] - step_size * (m / denom);
        param[i] = p;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int64_t j = i - sp.off[s];
            if (j >= 0 && j < nk) {
                float hi, lo;
                split_tf32_opt(p, hi, lo);
                sp.hi[s][j] = hi;
                sp.lo[s][j] = lo;
                const int row = (int)(j / sp.K), k = (int)(j - (int64_t)row * sp.K);
                sp.thi[(int64_t)k * sp.ldt + sp.toff[s] + row] = hi;
                sp.tlo[(int64_t)k * sp.ldt + sp.toff[s] + row] = lo;
            }
        }
    }
}

}  // namespace xb

using namespace xb;

extern "C" int xb_clip_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 int64_t* step_dev, float lr0, float lr_end_factor, int64_t lr_total_iters, float beta1,
                                 float beta2, float eps, float max_norm, float grad_scale, double* workspace,
                                 float* lr_out, float* gnorm_out, xb_stream_t stream) {
    if (n <= 0 || !param || !grad || !exp_avg || !exp_avg_sq || !step_dev || !workspace) return XB_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    AdamHyper h{lr0, lr_end_factor, beta1, beta2, eps, max_norm, grad_scale, lr_total_iters};
    int grid = grid_for(n, kOptBlock, 4);
    if (grid > kOptMaxGrid) grid = kOptMaxGrid;
    grad_norm_kernel<<<grid, kOptBlock, 0, s>>>(grad, n, step_dev, h, workspace, lr_out, gnorm_out);
    XB_LAUNCH_CHECK();
    adam_apply_kernel<<<grid, kOptBlock, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, h, workspace);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_adam_apply(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float beta1,
                             float beta2, float eps, float grad_scale, const double* workspace, xb_stream_t stream) {
    if (n <= 0 || !param || !grad || !exp_avg || !exp_avg_sq || !workspace) return XB_E_BADARG;
    AdamHyper h{0.0f, 0.0f, beta1, beta2, eps, 0.0f, grad_scale, 0};
    int grid = grid_for(n, kOptBlock, 4);
    if (grid > kOptMaxGrid) grid = kOptMaxGrid;
    XB_CUDA(launch_pdl(adam_apply_kernel, dim3(grid), dim3(kOptBlock), 0, (cudaStream_t)stream, true, param, grad, exp_avg, exp_avg_sq, n, h,
                       (const double*)workspace));
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_adam_apply_split(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float beta1,
                                   float beta2, float eps, float grad_scale, const double* workspace, int64_t w_off0,
                                   float* hi0, float* lo0, int64_t w_off1, float* hi1, float* lo1, int N, int K, float* thi,
                                   float* tlo, xb_stream_t stream) {
    if (n <= 0 || !param || !grad || !exp_avg || !exp_avg_sq || !workspace) return XB_E_BADARG;
    if (!hi0 || !lo0 || !hi1 || !lo1 || !thi || !tlo || N <= 0 || K <= 0) return XB_E_BADARG;
    if (w_off0 < 0 || w_off1 < 0 || w_off0 + (int64_t)N * K > n || w_off1 + (int64_t)N * K > n) return XB_E_BADARG;
    AdamHyper h{0.0f, 0.0f, beta1, beta2, eps, 0.0f, grad_scale, 0};
    AdamSplit sp{{w_off0, w_off1}, {hi0, hi1}, {lo0, lo1}, {0, N}, thi, tlo, N, K, 2 * N};
    int grid = grid_for(n, kOptBlock, 4);
    if (grid > kOptMaxGrid) grid = kOptMaxGrid;
    XB_CUDA(launch_pdl(adam_apply_split_kernel, dim3(grid), dim3(kOptBlock), 0, (cudaStream_t)stream, true, param, grad, exp_avg, exp_avg_sq,
                       n, h, workspace, sp));
    XB_LAUNCH_CHECK();
    return 0;
}
