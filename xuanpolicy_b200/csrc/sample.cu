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
#include "sample.cuh"

namespace xb {

__global__ void __launch_bounds__(128)
    sample_categorical_kernel(const float* __restrict__ logits, int A, uint64_t seed,
                              const uint64_t* __restrict__ counter_dev, uint64_t offset, int64_t* __restrict__ act_out,
                              float* __restrict__ logp_out, int64_t N) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const Philox ph = philox_setup(seed, counter_dev, offset);
    float logp;
    act_out[e] = (int64_t)sample_categorical_one(logits + e * A, A, e, ph, &logp);
    logp_out[e] = logp;
}

__global__ void __launch_bounds__(128)
    sample_gaussian_kernel(const float* __restrict__ mu, const float* __restrict__ logstd, int A, uint64_t seed,
                           const uint64_t* __restrict__ counter_dev, uint64_t offset, float* __restrict__ act_out,
                           float* __restrict__ logp_out, int64_t N) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const Philox ph = philox_setup(seed, counter_dev, offset);
    logp_out[e] = sample_gaussian_one(mu + e * A, logstd, A, e, ph, act_out + e * A);
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
