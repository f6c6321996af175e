// sample.cuh — per-env action sampling + log-prob, shared by the stand-alone sampling kernels (sample.cu) and the fused
// rollout-step kernel (env_classic.cu).  Formulas: xuance/torch/utils/distributions.py:51-58 (Categorical),
// :83-90 (DiagGaussian); random numbers: Philox4x32-10 keyed by `seed`, counter = (env index, device counter + step).
#pragma once
#include "common.cuh"

namespace xb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// (0, 1]: never 0, so log() is finite
__device__ __forceinline__ float u01_open0(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// [0, 1)
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }


struct Philox {
    uint2 key;
    uint64_t pos;
};
__device__ __forceinline__ Philox philox_setup(uint64_t seed, const uint64_t* counter_dev, uint64_t offset) {
    return Philox{make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), (counter_dev ? *counter_dev : 0ULL) + offset};
}

// inverse-CDF draw from softmax(z[0..A)); returns the action, writes its normalised log-probability
__device__ __forceinline__ int sample_categorical_one(const float* __restrict__ z, int A, int64_t e, const Philox& ph,
                                                      float* logp) {
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)ph.pos, (uint32_t)(ph.pos >> 32)),
                                    ph.key);
    float zmax = z[0];
    for (int j = 1; j < A; ++j) zmax = fmaxf(zmax, z[j]);
    float se = 0.0f;
    for (int j = 0; j < A; ++j) se += expf(z[j] - zmax);
    const float lse = zmax + logf(se);
    const float u = u01(rnd.x);
    int a = A - 1;
    float cum = 0.0f;
    for (int j = 0; j < A - 1; ++j) {
        cum += expf(z[j] - lse);
        if (u < cum) {
            a = j;
            break;
        }
    }
    *logp = z[a] - lse;
    return a;
}

// x[k] = mu[k] + exp(logstd[k]) * n_k for k < A; returns the log-pdf of the (rounded) action as Normal.log_prob sees it
__device__ __forceinline__ float sample_gaussian_one(const float* __restrict__ mu, const float* __restrict__ logstd, int A,
                                                     int64_t e, const Philox& ph, float* __restrict__ act_out) {
    float logp = 0.0f;
    for (int k0 = 0; k0 < A; k0 += 2) {
        // one Philox block gives two Box-Muller pairs; use one pair (2 normals) per block, block index in ctr.y high bits
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32) ^ ((uint32_t)(k0 >> 1) << 24),
                                                   (uint32_t)ph.pos, (uint32_t)(ph.pos >> 32)), ph.key);
        const float rad = sqrtf(-2.0f * logf(u01_open0(rnd.x)));
        float sn, cs;
        sincospif(2.0f * u01(rnd.y), &sn, &cs);
        const float zn[2] = {rad * cs, rad * sn};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = k0 + h;
            if (k < A) {
                const float ls = logstd[k];
                const float sd = expf(ls);
                const float m = mu[k];
                const float x = m + sd * zn[h];
                act_out[k] = x;
                const float d = x - m;
                logp += -(d * d) / (2.0f * sd * sd) - ls - 0.9189385332046727f;
            }
        }
    }
    return logp;
}

}  // namespace xb
