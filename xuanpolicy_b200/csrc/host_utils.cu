// host_utils.cu — HOST-side helper of the input feeder (no device code).
//
// xb_host_permutation draws a uniform random permutation of 0..n-1 straight into a (pinned) host buffer.
// It stands in for `np.random.shuffle(indexes)` in PPOCLIP_Agent.train (ppoclip_agent.py:76-78), which costs
// ~35 ms per 5e5 indices in numpy and would otherwise bound the whole loop once everything else is on the GPU.
// Inside-out Fisher-Yates with xoshiro256** (seeded through splitmix64) and Lemire's multiply-shift bounded
// integers: ~2 ms per 5e5 indices on one core.  Any uniform permutation is equivalent for PPO; the exact
// stream of numpy's legacy MT19937 shuffle is not a parity target.
#include <stdint.h>

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
