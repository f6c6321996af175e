// dist_loss.cu — policy-gradient losses that carry the OLD ACTION DISTRIBUTION (not just old_logp), forward +
// backward w.r.t. the network outputs in one launch.  SURVEY.md §8 row f3: PPO-KL and PPG share the env /
// buffer / GAE path with PPO-Clip; only this kernel differs.
//
// One generic objective covers the four reference updates (B samples, A action dims / classes):
//     L = surr_coef * a_loss  +  kl_w * mean KL(new || old)  -  ent_coef * mean H  +  vf_coef * mean (v - R)^2
//         +  aux_coef * mean (aux_v - stopgrad(v))^2
//     a_loss = -mean(ratio * adv)                               clip_range <= 0   (ppokl_learner.py:33-34)
//            = -mean(min(clamp(ratio, 1 +- c) * adv, adv * ratio))  clip_range > 0 (ppg_learner.py:34-37)
//     ratio  = exp(log_prob(act) - old_dist.log_prob(act))      (ppokl_learner.py:30-33, ppg_learner.py:28-33)
//   PPOKL_Learner.update            ppokl_learner.py:21-61  surr 1, kl_w = the adaptive kl_coef, vf, ent
//   PPG_Learner.update_policy       ppg_learner.py:23-55    surr 1 (clipped), ent
//   PPG_Learner.update_critic       ppg_learner.py:57-68    vf 1 only
//   PPG_Learner.update_auxiliary    ppg_learner.py:70-88    kl_w = kl_beta, vf 1, aux 1
// KL is torch.distributions.kl_divergence(new, old) (distributions.py:63-66,97-100): Categorical sums over the
// classes; Normal is element-wise and the reference's `.mean()` averages over B*A elements.
//
// Closed forms (p = softmax(z), q = softmax(z_old)):
//     KL_cat = sum_j p_j (log p_j - log q_j)             dKL/dz_k = p_k ((log p_k - log q_k) - KL)
//     KL_norm = 0.5 (s + t - 1 - log s), s = (sd/sd_old)^2, t = ((mu-mu_old)/sd_old)^2
//                                                        dKL/dmu = (mu-mu_old)/sd_old^2,  dKL/dlogstd = s - 1
// One thread per sample; all per-sample inputs are dense [B] minibatch arrays (what sample() returns).
// HBM bytes per sample (Categorical A=2, PPO-KL): act, ret, adv 12 + logits 8 + old logits 8 + v 4 read,
// dlogits 8 + dv 4 written = 44 B.
#include "common.cuh"

namespace xb {

struct DistLossArgs {
    int64_t B;
    const float* v_pred;
    const float* aux_v;  // nullable
    const float* act;
    const float* ret;
    const float* adv;
    const float* kl_coef_dev;  // nullable: device-resident multiplier of kl_coef (the adaptive PPO-KL coefficient)
    float clip_range, surr_coef, kl_coef, vf_coef, ent_coef, aux_coef, inv_batch;
    float* dv;
    float* daux;  // nullable
    double* scalars;
};

constexpr int kDistBlock = 256;
constexpr int kDistScalars = 7;  // {surrogate, value loss, entropy, v_pred, clip count, KL, aux loss}

// surrogate + value + auxiliary-value terms.  Returns dL/dlogp.
__device__ __forceinline__ float dist_surrogate_value(const DistLossArgs& c, int64_t i, float logp, float old_logp,
                                                      double (&acc)[kDistScalars]) {
    const float A = c.adv[i];
    const float ratio = expf(logp - old_logp);
    const float lo = 1.0f - c.clip_range, hi = 1.0f + c.clip_range;
    float m, dlogp;
    if (c.clip_range > 0.0f) {
        const float s1 = fminf(fmaxf(ratio, lo), hi) * A;
        const float s2 = A * ratio;
        m = fminf(s1, s2);
        const bool inactive = (A > 0.0f && ratio > hi) || (A < 0.0f && ratio < lo);
        dlogp = inactive ? 0.0f : -c.inv_batch * A * ratio;
        acc[4] += (ratio < lo || ratio > hi) ? 1.0 : 0.0;
    } else {
        m = ratio * A;
        dlogp = -c.inv_batch * A * ratio;
    }
    const float v = c.v_pred[i];
    const float verr = v - c.ret[i];
    c.dv[i] = c.vf_coef * c.inv_batch * 2.0f * verr;  // the auxiliary term sees v detached (ppg_learner.py:77)
    acc[0] += (double)m;
    acc[1] += (double)(verr * verr);
    acc[3] += (double)v;
    if (c.aux_v) {
        const float e = c.aux_v[i] - v;
        acc[6] += (double)(e * e);
        if (c.daux) c.daux[i] = c.aux_coef * c.inv_batch * 2.0f * e;
    }
    return c.surr_coef * dlogp;
}

__device__ __forceinline__ void dist_flush(double (&acc)[kDistScalars], double* scalars, double* smem) {
    block_sum<kDistScalars>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kDistScalars; ++k) atomicAdd(&scalars[k], acc[k]);
    }
}

constexpr int kMaxCatA = 16;

__global__ void __launch_bounds__(kDistBlock)
    dist_loss_categorical_kernel(DistLossArgs c, const float* __restrict__ logits, const float* __restrict__ old_logits,
                                 int A, float* __restrict__ dlogits) {
    __shared__ double smem[kDistScalars * 32];
    const float kl_w = c.kl_coef * (c.kl_coef_dev ? *c.kl_coef_dev : 1.0f) * c.inv_batch;
    const float ge = c.ent_coef * c.inv_batch;
    double acc[kDistScalars] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.B; i += (int64_t)gridDim.x * blockDim.x) {
        const float* z = logits + i * A;
        const float* zo = old_logits + i * A;
        const int a = (int)c.act[i];
        float zmax = z[0], omax = zo[0];
        for (int j = 1; j < A; ++j) {
            zmax = fmaxf(zmax, z[j]);
            omax = fmaxf(omax, zo[j]);
        }
        float se = 0.0f, so = 0.0f;
        for (int j = 0; j < A; ++j) {
            se += expf(z[j] - zmax);
            so += expf(zo[j] - omax);
        }
        const float lse = zmax + logf(se), lso = omax + logf(so);
        float H = 0.0f, KL = 0.0f;
        for (int j = 0; j < A; ++j) {
            const float lp = z[j] - lse, lq = zo[j] - lso;
            const float pj = expf(lp);
            H -= pj * lp;
            KL += pj * (lp - lq);
        }
        const float dlogp = dist_surrogate_value(c, i, z[a] - lse, zo[a] - lso, acc);
        acc[2] += (double)H;
        acc[5] += (double)KL;
        float* dz = dlogits + i * A;
        for (int j = 0; j < A; ++j) {
            const float lp = z[j] - lse, lq = zo[j] - lso;
            const float pj = expf(lp);
            dz[j] = dlogp * ((j == a ? 1.0f : 0.0f) - pj) + ge * pj * (lp + H) + kl_w * pj * ((lp - lq) - KL);
        }
    }
    dist_flush(acc, c.scalars, smem);
}

constexpr int kMaxDistGaussA = 8;
constexpr float kHalfLog2PiD = 0.9189385332046727f;

// old_std: [B][A] (old_std_stride = A) or one shared row [A] (stride 0): merge_distributions (operations.py:82-92)
// concatenates one std row per sample, but every sample of a rollout carries the same row.
__global__ void __launch_bounds__(kDistBlock)
    dist_loss_gaussian_kernel(DistLossArgs c, const float* __restrict__ mu, const float* __restrict__ logstd,
                              const float* __restrict__ old_mu, const float* __restrict__ old_std, int64_t old_std_stride,
                              int A, float* __restrict__ dmu, double* __restrict__ dlogstd_acc) {
    __shared__ double smem[(kMaxDistGaussA > kDistScalars ? kMaxDistGaussA : kDistScalars) * 32];
    const float kl_w = c.kl_coef * (c.kl_coef_dev ? *c.kl_coef_dev : 1.0f) * c.inv_batch / (float)A;  // mean over B*A
    const float ge = c.ent_coef * c.inv_batch;
    float ls[kMaxDistGaussA], sd[kMaxDistGaussA], inv_var[kMaxDistGaussA];
    float H = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxDistGaussA; ++k) {
        ls[k] = k < A ? logstd[k] : 0.0f;
        sd[k] = expf(ls[k]);
        inv_var[k] = 1.0f / (sd[k] * sd[k]);
        if (k < A) H += 0.5f + kHalfLog2PiD + ls[k];
    }
    double acc[kDistScalars] = {0, 0, 0, 0, 0, 0, 0};
    double gls[kMaxDistGaussA];
#pragma unroll
    for (int k = 0; k < kMaxDistGaussA; ++k) gls[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.B; i += (int64_t)gridDim.x * blockDim.x) {
        float logp = 0.0f, old_logp = 0.0f, KL = 0.0f;
        float diff[kMaxDistGaussA], dkl_mu[kMaxDistGaussA], s_ratio[kMaxDistGaussA];
#pragma unroll
        for (int k = 0; k < kMaxDistGaussA; ++k) {
            if (k < A) {
                const float x = c.act[i * A + k], m = mu[i * A + k], mo = old_mu[i * A + k];
                const float so = old_std[i * old_std_stride + k];
                const float inv_var_o = 1.0f / (so * so);
                diff[k] = x - m;
                logp += -(diff[k] * diff[k]) * (0.5f * inv_var[k]) - ls[k] - kHalfLog2PiD;
                const float d_old = x - mo;
                old_logp += -(d_old * d_old) * (0.5f * inv_var_o) - logf(so) - kHalfLog2PiD;
                const float r = sd[k] / so;
                s_ratio[k] = r * r;
                const float dm = m - mo;
                const float t1 = (dm / so) * (dm / so);
                KL += 0.5f * (s_ratio[k] + t1 - 1.0f - logf(s_ratio[k]));
                dkl_mu[k] = dm * inv_var_o;
            }
        }
        const float dlogp = dist_surrogate_value(c, i, logp, old_logp, acc);
        acc[2] += (double)H;
        acc[5] += (double)KL;
#pragma unroll
        for (int k = 0; k < kMaxDistGaussA; ++k) {
            if (k < A) {
                dmu[i * A + k] = dlogp * diff[k] * inv_var[k] + kl_w * dkl_mu[k];
                gls[k] += (double)(dlogp * (diff[k] * diff[k] * inv_var[k] - 1.0f)) - (double)ge +
                          (double)(kl_w * (s_ratio[k] - 1.0f));
            }
        }
    }
    block_sum<kMaxDistGaussA>(gls, smem);
    if (threadIdx.x == 0) {
        for (int k = 0; k < A; ++k) atomicAdd(&dlogstd_acc[k], gls[k]);
    }
    __syncthreads();
    dist_flush(acc, c.scalars, smem);
}

// kl = scalars[5] * inv_count;  > 1.5 target: x2;  < 0.5 target: /2;  clip to [0.1, 20]  (ppokl_learner.py:39-43).
__global__ void kl_coef_adapt_kernel(const double* __restrict__ scalars, float* __restrict__ kl_coef, float target_kl,
                                     double inv_count) {
    const float kl = (float)(scalars[5] * inv_count);
    float k = *kl_coef;
    if (kl > target_kl * 1.5f)
        k = k * 2.0f;
    else if (kl < target_kl * 0.5f)
        k = k / 2.0f;
    *kl_coef = fminf(fmaxf(k, 0.1f), 20.0f);
}

static int check_dist(int64_t B, const float* v_pred, const float* act, const float* ret, const float* adv, float* dv,
                      double* scalars, const float* aux_v, float aux_coef, float* daux) {
    if (B <= 0 || !v_pred || !act || !ret || !adv || !dv || !scalars) return XB_E_BADARG;
    if (aux_coef != 0.0f && (!aux_v || !daux)) return XB_E_BADARG;
    return 0;
}

}  // namespace xb

using namespace xb;

extern "C" int xb_dist_loss_categorical(int64_t B, const float* logits, const float* old_logits, int A,
                                        const float* v_pred, const float* aux_v, const float* act, const float* ret,
                                        const float* adv, float clip_range, float surr_coef, float kl_coef,
                                        const float* kl_coef_dev, float vf_coef, float ent_coef, float aux_coef,
                                        float inv_batch, float* dlogits, float* dv, float* daux, double* scalars,
                                        xb_stream_t stream) {
    int rc = check_dist(B, v_pred, act, ret, adv, dv, scalars, aux_v, aux_coef, daux);
    if (rc) return rc;
    if (!logits || !old_logits || !dlogits || A < 2) return XB_E_BADARG;
    if (A > kMaxCatA) return XB_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    XB_CUDA(cudaMemsetAsync(scalars, 0, 8 * sizeof(double), s));
    DistLossArgs c{B, v_pred, aux_v, act, ret, adv, kl_coef_dev, clip_range, surr_coef, kl_coef, vf_coef, ent_coef, aux_coef, inv_batch, dv, daux, scalars};
    dist_loss_categorical_kernel<<<grid_for(B, kDistBlock, 8), kDistBlock, 0, s>>>(c, logits, old_logits, A, dlogits);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_dist_loss_gaussian(int64_t B, const float* mu, const float* logstd, const float* old_mu,
                                     const float* old_std, int old_std_per_sample, int A, const float* v_pred,
                                     const float* aux_v, const float* act, const float* ret, const float* adv,
                                     float clip_range, float surr_coef, float kl_coef, const float* kl_coef_dev,
                                     float vf_coef, float ent_coef, float aux_coef, float inv_batch, float* dmu,
                                     double* dlogstd_acc, float* dv, float* daux, double* scalars, xb_stream_t stream) {
    int rc = check_dist(B, v_pred, act, ret, adv, dv, scalars, aux_v, aux_coef, daux);
    if (rc) return rc;
    if (!mu || !logstd || !old_mu || !old_std || !dmu || !dlogstd_acc || A < 1) return XB_E_BADARG;
    if (A > kMaxDistGaussA) return XB_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    XB_CUDA(cudaMemsetAsync(scalars, 0, 8 * sizeof(double), s));
    XB_CUDA(cudaMemsetAsync(dlogstd_acc, 0, A * sizeof(double), s));
    DistLossArgs c{B, v_pred, aux_v, act, ret, adv, kl_coef_dev, clip_range, surr_coef, kl_coef, vf_coef, ent_coef, aux_coef, inv_batch, dv, daux, scalars};
    dist_loss_gaussian_kernel<<<grid_for(B, kDistBlock, 8), kDistBlock, 0, s>>>(c, mu, logstd, old_mu, old_std,
                                                                                  old_std_per_sample ? A : 0, A, dmu, dlogstd_acc);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_kl_coef_adapt(const double* scalars, float* kl_coef_dev, float target_kl, int64_t count,
                                xb_stream_t stream) {
    if (!scalars || !kl_coef_dev || count <= 0) return XB_E_BADARG;
    kl_coef_adapt_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(scalars, kl_coef_dev, target_kl, 1.0 / (double)count);
    XB_LAUNCH_CHECK();
    return 0;
}
