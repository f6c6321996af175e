// host_utils.cu — HOST-side helper of the input feeder (no device code).
//
// xb_host_permutation draws a uniform random permutation of 0..n-1 straight into a (pinned) host buffer.
// It stands in for `np.random.shuffle(indexes)` in PPOCLIP_Agent.train (ppoclip_agent.py:76-78), which costs
// ~35 ms per 5e5 indices in numpy and would otherwise bound the whole loop once everything else is on the GPU.
// Inside-out Fisher-Yates with xoshiro256** (seeded through splitmix64) and Lemire's multiply-shift bounded
// integers: ~2 ms per 5e5 indices on one core.  Any uniform permutation is equivalent for PPO; the exact
// stream of numpy's legacy MT19937 shuffle is not a parity target.
#include <stdint.h>

#include <vector>

#include "../../include/xb200.h"

namespace {
struct Xoshiro {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    explicit Xoshiro(uint64_t seed) {
        for (int i = 0; i < 4; ++i) s[i] = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    // uniform in [0, bound), bound >= 1 (Lemire 2019, unbiased)
    uint64_t below(uint64_t bound) {
        unsigned __int128 m = (unsigned __int128)next() * bound;
        uint64_t lo = (uint64_t)m;
        if (lo < bound) {
            const uint64_t thresh = (0 - bound) % bound;
            while (lo < thresh) {
                m = (unsigned __int128)next() * bound;
                lo = (uint64_t)m;
            }
        }
        return (uint64_t)(m >> 64);
    }
};
}  // namespace

extern "C" int xb_host_permutation(int64_t* out, int64_t n, uint64_t seed) {
    if (!out || n <= 0) return XB_E_BADARG;
    Xoshiro rng(seed);
    for (int64_t i = 0; i < n; ++i) {  // inside-out Fisher-Yates
        const int64_t j = (int64_t)rng.below((uint64_t)i + 1);
        out[i] = out[j];
        out[j] = i;
    }
    return 0;
}

// 32-bit flavour for large index sets (what the e2e feeder uses: half the pinned-memory and H2D bytes).  Above 2^16 indices a
// plain Fisher-Yates walks randomly over tens of megabytes (~12 ns per index, cache-miss bound); here one Rao-Sandelius
// scatter pass first deals the indices into power-of-two many buckets of ~2^14 (each index draws its bucket uniformly and
// independently; unbiased because the bucket count is a power of two), then every bucket — now cache resident — is
// Fisher-Yates shuffled in place.  Concatenating uniformly shuffled buckets of independently dealt elements is a uniform
// random permutation (Rao 1961, Sandelius 1962); ~3 ns per index.
extern "C" int xb_host_permutation32(int32_t* out, int64_t n, uint64_t seed) {
    if (!out || n <= 0 || n > 0x7fffffffLL) return XB_E_BADARG;
    Xoshiro rng(seed);
    if (n <= (1 << 16)) {
        for (int64_t i = 0; i < n; ++i) {  // inside-out Fisher-Yates
            const int64_t j = (int64_t)rng.below((uint64_t)i + 1);
            out[i] = out[j];
            out[j] = (int32_t)i;
        }
        return 0;
    }
    int log_b = 1;
    while (((n + (1 << 14) - 1) >> 14) > (1LL << log_b)) ++log_b;
    if (log_b > 16) log_b = 16;
    const int64_t B = 1LL << log_b;
    const uint64_t mask = (uint64_t)B - 1;
    std::vector<uint16_t> bid((size_t)n);
    std::vector<int64_t> pos((size_t)B + 1, 0);
    for (int64_t i = 0; i < n; i += 4) {   // four 16-bit bucket draws per 64-bit output
        uint64_t r = rng.next();
        for (int64_t k = i; k < i + 4 && k < n; ++k, r >>= 16) {
            const uint16_t b = (uint16_t)(r & mask);
            bid[(size_t)k] = b;
            ++pos[(size_t)b + 1];
        }
    }
    for (int64_t b = 0; b < B; ++b) pos[(size_t)b + 1] += pos[(size_t)b];
    std::vector<int64_t> cur(pos.begin(), pos.end() - 1);
    for (int64_t i = 0; i < n; ++i) out[cur[bid[(size_t)i]]++] = (int32_t)i;
    for (int64_t b = 0; b < B; ++b) {
        int32_t* p = out + pos[(size_t)b];
        const int64_t m = pos[(size_t)b + 1] - pos[(size_t)b];
        for (int64_t i = m - 1; i > 0; --i) {
            const int64_t j = (int64_t)rng.below((uint64_t)i + 1);
            const int32_t t = p[i];
            p[i] = p[j];
            p[j] = t;
        }
    }
    return 0;
}
