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

// ---------------------------------------------------------------------------------------------- device permutation
// A keyed pseudo-random BIJECTION of [0, n) evaluated per index — no sort, no scratch, one 8-byte store per index
// (np.random.shuffle(indexes) of ppoclip_agent.py:76-78; torch.randperm costs 7 launches incl. a radix sort).
// Construction: k = ceil(log2 n) bits split into a high part (a = k/2 bits) and a low part (b = k-a bits); eight
// swap-free Feistel rounds alternately xor one part with a keyed 64-bit mix of the other (each round is a bijection
// of [0, 2^k) whatever a and b are); values that land in [n, 2^k) are walked along their cycle until they fall back
// into [0, n) (cycle walking keeps the map a bijection of [0, n); 2^k < 2n, so < 2 evaluations on average).
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

constexpr int kPermRounds = 8;

__device__ __forceinline__ uint64_t feistel_bijection(uint64_t x, int a, int b, const uint64_t (&key)[kPermRounds]) {
    const uint64_t mask_a = (1ULL << a) - 1ULL, mask_b = (1ULL << b) - 1ULL;
    uint64_t hi = x >> b, lo = x & mask_b;
#pragma unroll
    for (int r = 0; r < kPermRounds; r += 2) {
        hi ^= mix64(lo + key[r]) & mask_a;
        lo ^= mix64(hi + key[r + 1]) & mask_b;
    }
    return (hi << b) | lo;
}

__global__ void __launch_bounds__(256)
    random_permutation_kernel(int64_t* __restrict__ out, int64_t n, int bits, uint64_t seed,
                              const uint64_t* __restrict__ counter_dev, uint64_t offset) {
    const int a = bits / 2, b = bits - a;
    uint64_t key[kPermRounds];
    const uint64_t base = mix64(seed ^ 0x9E3779B97F4A7C15ULL) + (counter_dev ? *counter_dev : 0ULL) + offset;
#pragma unroll
    for (int r = 0; r < kPermRounds; ++r) key[r] = mix64(base * 0xD1342543DE82EF95ULL + (uint64_t)(r + 1) * 0x9E3779B97F4A7C15ULL);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t j = feistel_bijection((uint64_t)i, a, b, key);
        while (j >= (uint64_t)n) j = feistel_bijection(j, a, b, key);
        out[i] = (int64_t)j;
    }
}

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

extern "C" int xb_random_permutation(int64_t* out, int64_t n, uint64_t seed, const uint64_t* counter_dev, uint64_t offset,
                                     xb_stream_t stream) {
    if (!out || n <= 0 || n > (1LL << 40)) return XB_E_BADARG;
    int bits = 2;                                   // at least one bit per Feistel half
    while ((1LL << bits) < n) ++bits;
    random_permutation_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(out, n, bits, seed, counter_dev, offset);
    XB_LAUNCH_CHECK();
    return 0;
}
