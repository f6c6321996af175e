// optim.cuh — the scalars of one clipped Adam step, shared by every kernel that finishes a gradient-norm reduction
// (optim.cu grad_norm_kernel, peer_comm.cu peer_allreduce_grad_norm_kernel, dense_tc.cu wgrad_reduce_kernel).
// Replaces clip_grad_norm_'s coefficient, Adam's bias corrections and LinearLR.step() of
// xuance/torch/learners/policy_gradient/ppoclip_learner.py:48-51 (runner_drl.py:71-73).
#pragma once
#include "common.cuh"

namespace xb {

struct AdamHyper {
    float lr0, lr_end_factor, beta1, beta2, eps, max_norm, grad_scale;
    int64_t lr_total_iters;
};

// workspace layout (doubles): [0] grad norm, [1] clip coefficient, [2] lr, [3] bias_correction1,
// [4] sqrt(bias_correction2), [5] ticket (as bits), [8 .. 8+1024) per-CTA partial sums of squares
constexpr int kOptMaxGrid = 1024;

// One thread: from the total sum of squares of (grad * grad_scale) to the step scalars; advances the update counter.
__device__ __forceinline__ void adam_step_scalars(double sumsq, int64_t* step_dev, const AdamHyper& h, double* ws,
                                                  float* lr_out, float* gnorm_out) {
    const double norm = sqrt(sumsq);
    double clip = 1.0;
    if (h.max_norm > 0.0f) {  // clip_grad_norm_: min(1, max_norm / (norm + 1e-6))
        clip = (double)h.max_norm / (norm + 1e-6);
        if (clip > 1.0) clip = 1.0;
    }
    const int64_t it = *step_dev;  // updates done so far
    const int64_t capped = it < h.lr_total_iters ? it : h.lr_total_iters;
    double factor = 1.0;
    if (h.lr_total_iters > 0) factor = 1.0 + ((double)h.lr_end_factor - 1.0) * (double)capped / (double)h.lr_total_iters;
    const double lr = (double)h.lr0 * factor;
    const double t = (double)(it + 1);
    ws[0] = norm;
    ws[1] = clip;
    ws[2] = lr;
    ws[3] = 1.0 - pow((double)h.beta1, t);
    ws[4] = sqrt(1.0 - pow((double)h.beta2, t));
    *step_dev = it + 1;
    if (lr_out) *lr_out = (float)lr;
    if (gnorm_out) *gnorm_out = (float)norm;
}

// Every thread of the CTA calls it with its partial sum of squares: per-CTA partial -> ticket -> the last CTA adds the
// partials in a fixed order (deterministic) and derives the scalars.  smem: >= 32 doubles.
__device__ __forceinline__ void grad_norm_finish(double sq, int64_t* step_dev, const AdamHyper& h, double* ws, float* lr_out,
                                                 float* gnorm_out, double* smem, bool* is_last_smem) {
    double acc[1] = {sq};
    block_sum<1>(acc, smem);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + 5);
    if (threadIdx.x == 0) {
        ws[8 + blockIdx.x] = acc[0];
        __threadfence();
        *is_last_smem = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (*is_last_smem) {
        __threadfence();
        double tot[1] = {0.0};
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) tot[0] += ws[8 + b];
        block_sum<1>(tot, smem);
        if (threadIdx.x == 0) {
            adam_step_scalars(tot[0], step_dev, h, ws, lr_out, gnorm_out);
            *ticket = 0u;
        }
    }
}

}  // namespace xb
