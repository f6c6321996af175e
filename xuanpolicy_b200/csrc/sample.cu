// sample.cu — action sampling fused with log-prob, right after the actor GEMM.
//
// Replaces, in PPOCLIP_Agent._action (xuance/torch/agents/policy_gradient/ppoclip_agent.py:50-57):
//     dists.stochastic_sample()   torch Categorical.sample / Normal.sample  (distributions.py:57-58, 89-90)
//     dists.log_prob(acts)        (distributions.py:51-52, 83-84)
// i.e. softmax + multinomial + gather (Categorical) or randn + affine + log-pdf (Gaussian): 6-10 torch kernels
// become one.  Random numbers: Philox4x32-10 keyed by `seed`, counter = (env index, *counter_dev + offset), so
// a CUDA graph that captured this launch draws fresh numbers on every replay once the device counter has been
// advanced (xb_counter_add).  Action *values* are not a parity target (the reference uses torch's RNG); the
// log-prob of the drawn action is, and is checked against the reference formulas.
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

__global__ void __launch_bounds__(128)
    sample_categorical_kernel(const float* __restrict__ logits, int A, uint64_t seed,
                              const uint64_t* __restrict__ counter_dev, uint64_t offset, int64_t* __restrict__ act_out,
                              float* __restrict__ logp_out, int64_t N) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const uint64_t pos = (counter_dev ? *counter_dev : 0ULL) + offset;
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)pos, (uint32_t)(pos >> 32)),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float* z = logits + e * A;
    float zmax = z[0];
    for (int j = 1; j < A; ++j) zmax = fmaxf(zmax, z[j]);
    float se = 0.0f;
    for (int j = 0; j < A; ++j) se += expf(z[j] - zmax);
    const float lse = zmax + logf(se);
    // inverse-CDF draw
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
    act_out[e] = (int64_t)a;
    logp_out[e] = z[a] - lse;
}

__global__ void __launch_bounds__(128)
    sample_gaussian_kernel(const float* __restrict__ mu, const float* __restrict__ logstd, int A, uint64_t seed,
                           const uint64_t* __restrict__ counter_dev, uint64_t offset, float* __restrict__ act_out,
                           float* __restrict__ logp_out, int64_t N) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const uint64_t pos = (counter_dev ? *counter_dev : 0ULL) + offset;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    float logp = 0.0f;
    for (int k0 = 0; k0 < A; k0 += 2) {
        // one Philox block gives two Box-Muller pairs; use one pair (2 normals) per block, block index in ctr.y high bits
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)e, (uint32_t)(e >> 32) ^ ((uint32_t)(k0 >> 1) << 24),
                                                   (uint32_t)pos, (uint32_t)(pos >> 32)), key);
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
                const float m = mu[e * A + k];
                const float x = m + sd * zn[h];
                act_out[e * A + k] = x;
                const float d = x - m;  // log-pdf of the rounded action, as Normal.log_prob(acts) would see it
                logp += -(d * d) / (2.0f * sd * sd) - ls - 0.9189385332046727f;
            }
        }
    }
    logp_out[e] = logp;
}

__global__ void counter_add_kernel(uint64_t* counter, uint64_t inc) { *counter += inc; }

}  // namespace xb

using namespace xb;

extern "C" int xb_sample_categorical(const float* logits, int A, uint64_t seed, const uint64_t* counter_dev,
                                     uint64_t offset, int64_t* act_out, float* logp_out, int64_t N, xb_stream_t stream) {
    if (N <= 0 || A < 2 || !logits || !act_out || !logp_out) return XB_E_BADARG;
    sample_categorical_kernel<<<ceil_div_i64(N, 128), 128, 0, (cudaStream_t)stream>>>(logits, A, seed, counter_dev, offset,
                                                                                      act_out, logp_out, N);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_sample_gaussian(const float* mu, const float* logstd, int A, uint64_t seed, const uint64_t* counter_dev,
                                  uint64_t offset, float* act_out, float* logp_out, int64_t N, xb_stream_t stream) {
    if (N <= 0 || A < 1 || A > 512 || !mu || !logstd || !act_out || !logp_out) return XB_E_BADARG;
    sample_gaussian_kernel<<<ceil_div_i64(N, 128), 128, 0, (cudaStream_t)stream>>>(mu, logstd, A, seed, counter_dev, offset,
                                                                                   act_out, logp_out, N);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_counter_add(uint64_t* counter_dev, uint64_t inc, xb_stream_t stream) {
    if (!counter_dev) return XB_E_BADARG;
    counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, inc);
    XB_LAUNCH_CHECK();
    return 0;
}
