// ppo_loss.cu — PPO-Clip loss forward + backward w.r.t. the network outputs, fused with the minibatch
// gather of the per-transition scalars and with the on-the-fly advantage normalisation.
//
// Replaces, in PPOCLIP_Learner.update (xuance/torch/learners/policy_gradient/ppoclip_learner.py):
//     :33     a_dist.log_prob(act_batch)                 (distributions.py:51-52 / :83-84)
//     :36-39  ratio, clipped surrogate, a_loss
//     :41     c_loss = mse_loss(v_pred, ret)             (value_batch is unused there; opt-in value clip below)
//     :43     e_loss = a_dist.entropy().mean()           (distributions.py:54-55 / :86-87)
//     :44-46  loss, and the part of loss.backward() between the loss and the network outputs
//     :54     clip_ratio,  :61 predict_value
// and the four scalar fancy-index gathers + advantage normalisation of DummyOnPolicyBuffer.sample
// (xuance/common/memory_tools.py:236-243).  Closed-form gradients: SURVEY.md App. C.
//
// One thread per sample.  HBM traffic per sample (Categorical, A=2): 8 B index + 4x4 B gathered scalars +
// 8 B logits + 4 B v read, 8 B dlogits + 4 B dv written = 48 B.  The scalar reductions (loss terms for the log, the
// log-std gradient) are warp-shuffle -> per-CTA partial -> fixed-order sum in the last CTA: deterministic, no atomics on
// the results, nothing to zero beforehand.
#include "common.cuh"

namespace xb {

struct LossCommon {
    const int64_t* idx;  // nullable
    int64_t B, T, N;
    int64_t stride;      // element stride (floats) of the dense per-sample arrays when idx == NULL (4 = packed float4)
    const float* v_pred;
    const float* act;
    const float* ret;
    const float* adv;
    const float* old_logp;
    const float* val_old;     // nullable
    const double* adv_stats;  // nullable
    double inv_adv_count;
    float clip_range, vf_coef, ent_coef, value_clip, inv_batch;
    float* dv;
    double* scalars;
};

__device__ __forceinline__ int64_t sample_row(const LossCommon& c, int64_t i) {
    if (!c.idx) return i * c.stride;
    int64_t k = c.idx[i];
    int64_t env = k / c.T;
    return (k - env * c.T) * c.N + env;
}

struct AdvNorm {
    float mean, denom;
    bool on;
};
__device__ __forceinline__ AdvNorm load_adv_norm(const LossCommon& c) {
    AdvNorm n{0.0f, 1.0f, false};
    if (c.adv_stats) {
        double mean = c.adv_stats[0] * c.inv_adv_count;
        double var = c.adv_stats[1] * c.inv_adv_count - mean * mean;
        n.mean = (float)mean;
        n.denom = (float)sqrt(var > 0.0 ? var : 0.0) + 1e-8f;
        n.on = true;
    }
    return n;
}

// The per-sample scalars {act, old_logp, adv, ret}.  Packed minibatch rows (xb_gather_records: one float4 per sample, the
// four arrays are the lanes of the same float4) are fetched with ONE 16-byte load instead of four strided 4-byte ones
// (ncu at 2 M samples: the scalar form ran at 0.30 of the HBM peak with 4x the load instructions per sample).
struct Scal4 {
    float act, old_logp, adv, ret;
};
__device__ __forceinline__ bool scalars_packed(const LossCommon& c) {
    return !c.idx && c.stride == 4 && c.old_logp == c.act + 1 && c.adv == c.act + 2 && c.ret == c.act + 3 &&
           ((uintptr_t)c.act & 15u) == 0;
}
__device__ __forceinline__ Scal4 load_scalars(const LossCommon& c, int64_t row, bool packed) {
    if (packed) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(c.act + row));
        return Scal4{q.x, q.y, q.z, q.w};
    }
    return Scal4{c.act[row], c.old_logp ? c.old_logp[row] : 0.0f, c.adv[row], c.ret[row]};
}

// surrogate + value terms shared by both policy heads.  Returns dL/dlogp; writes dv; accumulates log scalars.
__device__ __forceinline__ float surrogate_and_value(const LossCommon& c, const AdvNorm& nrm, int64_t i, int64_t row,
                                                     const Scal4& sc, float logp, double (&acc)[5]) {
    float A = sc.adv;
    if (nrm.on) A = (A - nrm.mean) / nrm.denom;
    float m, dlogp, ratio = 1.0f;
    const float lo = 1.0f - c.clip_range, hi = 1.0f + c.clip_range;
    if (c.clip_range > 0.0f) {   // PPO-Clip surrogate (ppoclip_learner.py:36-39)
        ratio = expf(logp - sc.old_logp);
        const float s1 = fminf(fmaxf(ratio, lo), hi) * A;
        const float s2 = A * ratio;
        m = fminf(s1, s2);
        const bool inactive = (A > 0.0f && ratio > hi) || (A < 0.0f && ratio < lo);
        dlogp = inactive ? 0.0f : -c.inv_batch * A * ratio;
    } else {                     // A2C / PG surrogate: -(adv * log_prob).mean()  (a2c_learner.py:28, pg_learner.py:24)
        m = A * logp;
        dlogp = -c.inv_batch * A;
    }

    const float v = c.v_pred[i], R = sc.ret;
    float verr = v - R;
    float vloss = verr * verr;
    float dvl = 2.0f * verr;
    if (c.value_clip > 0.0f) {  // opt-in: max((v-R)^2, (v_old + clip(v - v_old, +-c) - R)^2)
        const float vo = c.val_old[row];
        const float dlt = v - vo;
        const float dc = fminf(fmaxf(dlt, -c.value_clip), c.value_clip);
        const float e2 = vo + dc - R;
        const float l2 = e2 * e2;
        if (l2 > vloss) {
            vloss = l2;
            dvl = (dlt == dc) ? 2.0f * e2 : 0.0f;
        }
    }
    c.dv[i] = c.vf_coef * c.inv_batch * dvl;

    acc[0] += (double)m;
    acc[1] += (double)vloss;
    acc[3] += (double)v;
    acc[4] += (c.clip_range > 0.0f && (ratio < lo || ratio > hi)) ? 1.0 : 0.0;
    return dlogp;
}

constexpr int kLossBlock = 256;
constexpr int kLossMaxGrid = kNumSMs * 8;
constexpr int kLossSlots = 16;   // doubles per CTA partial: 5 log scalars + up to 8 log-std gradient sums

// Per-CTA partial sums -> ticket -> the last CTA adds them in CTA order: deterministic, no atomics on the results and no
// zeroing launch in front of the kernel (the outputs are overwritten).  One loss kernel runs at a time per device.
__device__ double g_loss_partials[kLossMaxGrid * kLossSlots];
__device__ unsigned int g_loss_ticket;

// v[0..4] -> scalars[0..4] (scalars[5..7] = 0); v[5..5+A) -> dlogstd[0..A) when dlogstd != NULL.
template <int K>
__device__ __forceinline__ void finish_sums(double (&v)[K], double* smem, double* __restrict__ scalars,
                                            double* __restrict__ dlogstd, int A) {
    __shared__ bool is_last;
    block_sum<K>(v, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) g_loss_partials[blockIdx.x * kLossSlots + k] = v[k];
        __threadfence();
        is_last = (atomicAdd(&g_loss_ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < K; k += kLossBlock / 32) {   // one warp per output, CTA partials added in a fixed order
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += g_loss_partials[b * kLossSlots + k];
        s = warp_sum(s);
        if (lane == 0) {
            if (k < 5) scalars[k] = s;
            else if (dlogstd && k - 5 < A) dlogstd[k - 5] = s;
        }
    }
    if (threadIdx.x < 3) scalars[5 + threadIdx.x] = 0.0;
    if (threadIdx.x == 0) g_loss_ticket = 0u;
}

// ------------------------------------------------------------------------------------------------ Categorical
template <int A_STATIC>
__global__ void __launch_bounds__(kLossBlock)
    loss_categorical_kernel(LossCommon c, const float* __restrict__ logits, int A_rt, float* __restrict__ dlogits) {
    __shared__ double smem[5 * 32];
    const int A = A_STATIC > 0 ? A_STATIC : A_rt;
    const AdvNorm nrm = load_adv_norm(c);
    const bool packed = scalars_packed(c);
    double acc[5] = {0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.B; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = sample_row(c, i);
        const Scal4 sc = load_scalars(c, row, packed);
        float zl[A_STATIC > 0 ? A_STATIC : 1];
        if (A_STATIC == 2) {     // one 8-byte load / store per sample
            const float2 q = __ldg(reinterpret_cast<const float2*>(logits) + i);
            zl[0] = q.x;
            zl[1] = q.y;
        }
        const float* z = A_STATIC == 2 ? zl : logits + i * A;
        float dzl[A_STATIC > 0 ? A_STATIC : 1];
        float* dz = A_STATIC == 2 ? dzl : dlogits + i * A;
        const int a = (int)sc.act;  // stored as float32 (memory_tools.py:173); torch casts with .long()
        float zmax = z[0];
        for (int j = 1; j < A; ++j) zmax = fmaxf(zmax, z[j]);
        float se = 0.0f;
        for (int j = 0; j < A; ++j) se += expf(z[j] - zmax);
        const float lse = zmax + logf(se);
        float H = 0.0f;
        for (int j = 0; j < A; ++j) {
            const float lp = z[j] - lse;
            H -= expf(lp) * lp;
        }
        const float logp = z[a] - lse;
        const float dlogp = surrogate_and_value(c, nrm, i, row, sc, logp, acc);
        acc[2] += (double)H;
        const float ge = c.ent_coef * c.inv_batch;  // d(-ent_coef * mean H)/dz_j = +ge * p_j (logp_j + H)
        for (int j = 0; j < A; ++j) {
            const float lp = z[j] - lse;
            const float pj = expf(lp);
            dz[j] = dlogp * ((j == a ? 1.0f : 0.0f) - pj) + ge * pj * (lp + H);
        }
        if (A_STATIC == 2) reinterpret_cast<float2*>(dlogits)[i] = make_float2(dzl[0], dzl[1]);
    }
    finish_sums<5>(acc, smem, c.scalars, nullptr, 0);
}

// ------------------------------------------------------------------------------------------------ Gaussian
constexpr int kMaxGaussA = 8;
constexpr float kHalfLog2Pi = 0.9189385332046727f;

__global__ void __launch_bounds__(kLossBlock)
    loss_gaussian_kernel(LossCommon c, const float* __restrict__ mu, const float* __restrict__ logstd, int A,
                         float* __restrict__ dmu, double* __restrict__ dlogstd_acc) {
    __shared__ double smem[(5 + kMaxGaussA) * 32];
    const AdvNorm nrm = load_adv_norm(c);
    float ls[kMaxGaussA], inv_var[kMaxGaussA];
    float H = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxGaussA; ++k) {
        ls[k] = k < A ? logstd[k] : 0.0f;
        const float sd = expf(ls[k]);
        inv_var[k] = 1.0f / (sd * sd);
        if (k < A) H += 0.5f + kHalfLog2Pi + ls[k];
    }
    double acc[5] = {0, 0, 0, 0, 0};
    double gls[kMaxGaussA];
#pragma unroll
    for (int k = 0; k < kMaxGaussA; ++k) gls[k] = 0.0;
    const float ge = c.ent_coef * c.inv_batch;
    const bool packed = scalars_packed(c);       // (stride 4 implies A == 1)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.B; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = sample_row(c, i);
        const Scal4 sc = load_scalars(c, row, packed);      // (sc.act is only used in the packed, A == 1 form)
        float logp = 0.0f;
        float diff[kMaxGaussA];
#pragma unroll
        for (int k = 0; k < kMaxGaussA; ++k) {
            if (k < A) {
                diff[k] = (packed ? sc.act : c.act[row * A + k]) - mu[i * A + k];
                logp += -(diff[k] * diff[k]) * (0.5f * inv_var[k]) - ls[k] - kHalfLog2Pi;
            }
        }
        const float dlogp = surrogate_and_value(c, nrm, i, row, sc, logp, acc);
        acc[2] += (double)H;
#pragma unroll
        for (int k = 0; k < kMaxGaussA; ++k) {
            if (k < A) {
                dmu[i * A + k] = dlogp * diff[k] * inv_var[k];
                gls[k] += (double)(dlogp * (diff[k] * diff[k] * inv_var[k] - 1.0f)) - (double)ge;
            }
        }
    }
    double all[5 + kMaxGaussA];
#pragma unroll
    for (int k = 0; k < 5; ++k) all[k] = acc[k];
#pragma unroll
    for (int k = 0; k < kMaxGaussA; ++k) all[5 + k] = gls[k];
    finish_sums<5 + kMaxGaussA>(all, smem, c.scalars, dlogstd_acc, A);
}

static int check_common(const int64_t* idx, int64_t B, int64_t T, int64_t N, const float* v_pred, const float* act,
                        const float* ret, const float* adv, const float* old_logp, const float* val_old,
                        const double* adv_stats, int64_t adv_count, float value_clip, float* dv, double* scalars,
                        float clip_range, int64_t stride, int A) {
    if (clip_range > 0.0f && !old_logp) return XB_E_BADARG;
    if (stride < 1 || (idx && stride != 1) || (stride > 1 && A > 1)) return XB_E_BADARG;
    if (B <= 0 || !v_pred || !act || !ret || !adv || !dv || !scalars) return XB_E_BADARG;
    if (idx && (T <= 0 || N <= 0)) return XB_E_BADARG;
    if (adv_stats && adv_count <= 0) return XB_E_BADARG;
    if (value_clip > 0.0f && !val_old) return XB_E_BADARG;
    return 0;
}

}  // namespace xb

using namespace xb;

extern "C" int xb_ppo_loss_categorical(const int64_t* idx, int64_t B, int64_t T, int64_t N, const float* logits, int A,
                                       const float* v_pred, const float* act, const float* ret, const float* adv,
                                       const float* old_logp, const float* val_old, const double* adv_stats,
                                       int64_t adv_count, float clip_range, float vf_coef, float ent_coef,
                                       float value_clip, float inv_batch, int64_t stride, float* dlogits, float* dv,
                                       double* scalars, xb_stream_t stream) {
    int rc = check_common(idx, B, T, N, v_pred, act, ret, adv, old_logp, val_old, adv_stats, adv_count, value_clip, dv, scalars, clip_range, stride, 1);
    if (rc) return rc;
    if (!logits || !dlogits || A < 2) return XB_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    LossCommon c{idx, B, T, N, stride, v_pred, act, ret, adv, old_logp, val_old, adv_stats,
                 adv_stats ? 1.0 / (double)adv_count : 0.0, clip_range, vf_coef, ent_coef, value_clip, inv_batch, dv, scalars};
    int grid = grid_for(B, kLossBlock, 8);
    if (A == 2)
        loss_categorical_kernel<2><<<grid, kLossBlock, 0, s>>>(c, logits, A, dlogits);
    else
        loss_categorical_kernel<0><<<grid, kLossBlock, 0, s>>>(c, logits, A, dlogits);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_ppo_loss_gaussian(const int64_t* idx, int64_t B, int64_t T, int64_t N, const float* mu,
                                    const float* logstd, int A, const float* v_pred, const float* act, const float* ret,
                                    const float* adv, const float* old_logp, const float* val_old,
                                    const double* adv_stats, int64_t adv_count, float clip_range, float vf_coef,
                                    float ent_coef, float value_clip, float inv_batch, int64_t stride, float* dmu,
                                    double* dlogstd_acc, float* dv, double* scalars, xb_stream_t stream) {
    int rc = check_common(idx, B, T, N, v_pred, act, ret, adv, old_logp, val_old, adv_stats, adv_count, value_clip, dv, scalars, clip_range, stride, A);
    if (rc) return rc;
    if (!mu || !logstd || !dmu || !dlogstd_acc || A < 1) return XB_E_BADARG;
    if (A > kMaxGaussA) return XB_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    LossCommon c{idx, B, T, N, stride, v_pred, act, ret, adv, old_logp, val_old, adv_stats,
                 adv_stats ? 1.0 / (double)adv_count : 0.0, clip_range, vf_coef, ent_coef, value_clip, inv_batch, dv, scalars};
    loss_gaussian_kernel<<<grid_for(B, kLossBlock, 8), kLossBlock, 0, s>>>(c, mu, logstd, A, dmu, dlogstd_acc);
    XB_LAUNCH_CHECK();
    return 0;
}
