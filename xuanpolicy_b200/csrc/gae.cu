// gae.cu — GAE(gamma, lambda) / discounted-return reverse scan over T, parallel over envs.
//
// Replaces DummyOnPolicyBuffer.finish_path (xuance/common/memory_tools.py:206-229: a per-env Python loop,
// ~1.5 us per element) and discount_cumsum (xuance/common/common_tools.py:199-200) with the batched form of
// every env and every path segment of a rollout in one launch.
//
//     for t = T-1 .. 0:   seg_end = (t == T-1) | term[t] | trunc[t]
//         if seg_end: nextv = term[t] ? 0 : boot;  last = 0          (ppoclip_agent.py:71-75,96-100)
//         delta = rew + (1-term)*gamma*nextv - val ; last = delta + (1-term)*gamma*lam*last
//         adv[t] = last ; ret[t] = last + val ; nextv = val
//
// The recurrence is carried in fp64 registers and rounded once on store (matches the fp64 oracle to 0.5 ulp
// of fp32; the reference itself mixes fp32/fp64 depending on the numpy version).
//
// Roofline: HBM.  Algorithmic traffic 20 B/element (read rew, val, term; write adv, ret), +1 B with a
// truncation mask.  Time-major [T][N] storage makes thread-per-env column access unit-stride across a warp.
//
// Two variants behind one entry point:
//   LDG  — each thread prefetches a TILE_T-deep register tile of its column one tile ahead (plain coalesced
//          loads with L1::no_allocate), so ~2*TILE_T*12 B per thread are in flight.
//   TMA  — a producer thread streams [TILE_T x TILE_N] boxes of rew/val/term through a STAGES-deep shared
//          memory ring with cp.async.bulk.tensor + mbarrier complete_tx; 4 consumer warps scan the columns
//          out of shared memory (conflict-free: consecutive threads, consecutive banks) and store adv/ret
//          coalesced.  No registers are spent on prefetch and the copy engine does the address generation.
#include <cuda.h>

#include "common.cuh"

namespace xb {

struct GaeParams {
    const float* rew;
    const float* val;
    const float* term;
    const uint8_t* trunc;   // nullable
    const float* boot;      // nullable unless trunc
    const float* boot_last;
    float* adv;
    float* ret;
    double* stats;          // nullable: (sum adv, sum adv^2)
    int64_t T, N;
    double gamma, lam;
    int use_gae;
};

struct ScanState {
    double last, nextv, run;
};

// one element of the reverse scan
__device__ __forceinline__ void gae_element(ScanState& s, float r32, float v32, float d32, bool seg_end, double boot,
                                            double gamma, double gl, int use_gae, float& adv_out, float& ret_out) {
    if (seg_end) {
        double b = d32 != 0.0f ? 0.0 : boot;
        s.nextv = b;
        s.last = 0.0;
        s.run = b;
    }
    double r = (double)r32, v = (double)v32;
    if (use_gae) {
        double nt = 1.0 - (double)d32;
        double delta = r + nt * gamma * s.nextv - v;
        s.last = delta + nt * gl * s.last;
        adv_out = (float)s.last;
        ret_out = (float)(s.last + v);
    } else {
        s.run = r + gamma * s.run;
        adv_out = (float)(r + gamma * s.nextv - v);
        ret_out = (float)s.run;
    }
    s.nextv = v;
}

__device__ __forceinline__ void accumulate_stats(double sum, double sumsq, double* stats, double* smem) {
    double v[2] = {sum, sumsq};
    block_sum<2>(v, smem);
    if (threadIdx.x == 0) {
        atomicAdd(&stats[0], v[0]);
        atomicAdd(&stats[1], v[1]);
    }
}

// ------------------------------------------------------------------------------------------------ LDG variant
template <int TILE_T>
struct ColTile {
    float r[TILE_T], v[TILE_T], d[TILE_T];
    uint8_t tr[TILE_T];
};

template <int TILE_T>
__device__ __forceinline__ void ldg_load_tile(ColTile<TILE_T>& c, const GaeParams& p, int64_t tile, int64_t e) {
    const int64_t t0 = tile * TILE_T;
#pragma unroll
    for (int i = 0; i < TILE_T; ++i) {
        const int64_t t = t0 + i;
        if (t < p.T) {
            const int64_t k = t * p.N + e;
            c.r[i] = ld_stream(p.rew + k);
            c.v[i] = ld_stream(p.val + k);
            c.d[i] = ld_stream(p.term + k);
            c.tr[i] = p.trunc ? p.trunc[k] : (uint8_t)0;
        }
    }
}

template <int TILE_T>
__device__ __forceinline__ void ldg_scan_tile(const ColTile<TILE_T>& c, const GaeParams& p, int64_t tile, int64_t e,
                                              ScanState& s, double boot_last, double gamma, double gl, double& sum,
                                              double& sumsq) {
    const int64_t t0 = tile * TILE_T;
#pragma unroll
    for (int i = TILE_T - 1; i >= 0; --i) {
        const int64_t t = t0 + i;
        if (t < p.T) {
            const int64_t k = t * p.N + e;
            const bool last_row = (t == p.T - 1);
            const bool seg_end = last_row || c.d[i] != 0.0f || c.tr[i] != 0;
            double boot = 0.0;
            if (seg_end) boot = last_row ? boot_last : (p.boot ? (double)p.boot[k] : 0.0);
            float a, q;
            gae_element(s, c.r[i], c.v[i], c.d[i], seg_end, boot, gamma, gl, p.use_gae, a, q);
            st_stream(p.adv + k, a);
            st_stream(p.ret + k, q);
            sum += (double)a;
            sumsq += (double)a * (double)a;
        }
    }
}

template <int TILE_T>
__global__ void __launch_bounds__(128) gae_ldg_kernel(GaeParams p) {
    __shared__ double smem[64];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double gamma = p.gamma, gl = p.gamma * p.lam;
    double sum = 0.0, sumsq = 0.0;

    if (e < p.N) {
        ScanState s{0.0, 0.0, 0.0};
        const double boot_last = (double)p.boot_last[e];
        ColTile<TILE_T> a, b;  // two register tiles: one being scanned, one in flight
        int64_t tile = (p.T + TILE_T - 1) / TILE_T - 1;
        ldg_load_tile(a, p, tile, e);
        while (true) {
            if (tile > 0) ldg_load_tile(b, p, tile - 1, e);
            ldg_scan_tile(a, p, tile, e, s, boot_last, gamma, gl, sum, sumsq);
            if (--tile < 0) break;
            if (tile > 0) ldg_load_tile(a, p, tile - 1, e);
            ldg_scan_tile(b, p, tile, e, s, boot_last, gamma, gl, sum, sumsq);
            if (--tile < 0) break;
        }
    }
    if (p.stats) accumulate_stats(sum, sumsq, p.stats, smem);
}

// ------------------------------------------------------------------------------------------------ TMA variant
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_load_2d_if(uint32_t elected, void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred pe;\n"
        "setp.ne.b32 pe, %5, 0;\n"
        "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
        "}\n" ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "r"(elected)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_if(uint32_t elected, uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n"
        ".reg .pred pe;\n"
        "setp.ne.b32 pe, %2, 0;\n"
        "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(bytes), "r"(elected)
        : "memory");
}

constexpr int kTileN = 128;   // envs per CTA (512 B rows)
constexpr int kTileT = 8;     // time steps per stage
constexpr int kStages = 4;
constexpr int kConsumers = kTileN;            // 4 consumer warps
constexpr int kTmaThreads = kConsumers + 32;  // + 1 producer warp
constexpr int kTileFloats = kTileT * kTileN;
constexpr uint32_t kStageBytes = 3u * kTileFloats * sizeof(float);

struct __align__(128) GaeSmem {
    float tile[kStages][3][kTileFloats];  // rew, val, term boxes: [TILE_T][TILE_N]
    uint64_t full[kStages];
    uint64_t empty[kStages];
    double red[64];
};

__global__ void __launch_bounds__(kTmaThreads)
    gae_tma_kernel(const __grid_constant__ CUtensorMap map_rew, const __grid_constant__ CUtensorMap map_val,
                   const __grid_constant__ CUtensorMap map_term, GaeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    GaeSmem& sm = *reinterpret_cast<GaeSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5;
    const int64_t T = p.T, N = p.N;
    const int64_t e0 = (int64_t)blockIdx.x * kTileN;
    const int ntiles = (int)((T + kTileT - 1) / kTileT);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double sum = 0.0, sumsq = 0.0;
    if (warp == kConsumers / 32) {
        // ===== producer: the whole warp walks the loop, one elected lane streams tiles from the last time tile down to the first
        // (warp-uniform operands: inside `if (lane == 0)` ptxas wraps every UTMALDG in an ELECT / R2UR.BROADCAST waterfall) =====
        {
            uint32_t elected;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "elect.sync _|p, 0xffffffff;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(elected));
            int stage = 0;
            uint32_t phase = 0;
            for (int j = ntiles - 1; j >= 0; --j) {
                mbar_wait(&sm.empty[stage], phase ^ 1);
                mbar_arrive_expect_tx_if(elected, &sm.full[stage], kStageBytes);
                tma_load_2d_if(elected, sm.tile[stage][0], &map_rew, (int)e0, j * kTileT, &sm.full[stage]);
                tma_load_2d_if(elected, sm.tile[stage][1], &map_val, (int)e0, j * kTileT, &sm.full[stage]);
                tma_load_2d_if(elected, sm.tile[stage][2], &map_term, (int)e0, j * kTileT, &sm.full[stage]);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== consumers: thread c scans column e0 + c out of shared memory =====
        const int c = threadIdx.x;
        const int64_t e = e0 + c;
        const bool active = e < N;
        const double gamma = p.gamma, gl = p.gamma * p.lam;
        ScanState s{0.0, 0.0, 0.0};
        const double boot_last = active ? (double)p.boot_last[e] : 0.0;
        int stage = 0;
        uint32_t phase = 0;
        uint8_t tr_cur[kTileT], tr_next[kTileT];
#pragma unroll
        for (int i = 0; i < kTileT; ++i) tr_cur[i] = tr_next[i] = 0;
        auto load_trunc = [&](uint8_t (&dst)[kTileT], int j) {
#pragma unroll
            for (int i = 0; i < kTileT; ++i) {
                const int64_t t = (int64_t)j * kTileT + i;
                dst[i] = (active && t < T) ? p.trunc[t * N + e] : (uint8_t)0;
            }
        };
        if (p.trunc) load_trunc(tr_cur, ntiles - 1);
        for (int j = ntiles - 1; j >= 0; --j) {
            if (p.trunc && j > 0) load_trunc(tr_next, j - 1);
            mbar_wait(&sm.full[stage], phase);
            const float* tr_ = sm.tile[stage][0];
            const float* tv_ = sm.tile[stage][1];
            const float* td_ = sm.tile[stage][2];
            float r[kTileT], v[kTileT], d[kTileT];
#pragma unroll
            for (int i = 0; i < kTileT; ++i) {
                r[i] = tr_[i * kTileN + c];
                v[i] = tv_[i * kTileN + c];
                d[i] = td_[i * kTileN + c];
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&sm.empty[stage]);  // slot is in registers now
            if (active) {
#pragma unroll
                for (int i = kTileT - 1; i >= 0; --i) {
                    const int64_t t = (int64_t)j * kTileT + i;
                    if (t < T) {
                        const int64_t k = t * N + e;
                        const bool last_row = (t == T - 1);
                        const bool seg_end = last_row || d[i] != 0.0f || tr_cur[i] != 0;
                        double boot = 0.0;
                        if (seg_end) boot = last_row ? boot_last : (p.boot ? (double)p.boot[k] : 0.0);
                        float a, q;
                        gae_element(s, r[i], v[i], d[i], seg_end, boot, gamma, gl, p.use_gae, a, q);
                        st_stream(p.adv + k, a);
                        st_stream(p.ret + k, q);
                        sum += (double)a;
                        sumsq += (double)a * (double)a;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kTileT; ++i) tr_cur[i] = tr_next[i];
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    }
    if (p.stats) accumulate_stats(sum, sumsq, p.stats, sm.red);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

static bool make_map_f32(CUtensorMap* map, const float* base, int64_t T, int64_t N) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T};
    cuuint64_t strides[1] = {(cuuint64_t)N * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kTileN, (cuuint32_t)kTileT};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static bool tma_eligible(const GaeParams& p) {
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
    return p.N % 4 == 0 && p.N >= kTileN && p.N < (1LL << 31) && p.T < (1LL << 31) && al16(p.rew) && al16(p.val) &&
           al16(p.term);
}

}  // namespace xb

using namespace xb;

extern "C" int xb_gae(const float* rew, const float* val, const float* term, const uint8_t* trunc, const float* boot,
                      const float* boot_last, float* adv, float* ret, double* stats, int64_t T, int64_t N,
                      double gamma, double lam, int use_gae, int variant, xb_stream_t stream) {
    if (T <= 0 || N <= 0 || !rew || !val || !term || !boot_last || !adv || !ret || (trunc && !boot)) return XB_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    GaeParams p{rew, val, term, trunc, boot, boot_last, adv, ret, stats, T, N, gamma, lam, use_gae};
    if (stats) XB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double), s));

    bool use_tma = false;
    if (variant == XB_GAE_TMA) {
        if (!tma_eligible(p)) return XB_E_UNSUPPORTED;
        use_tma = true;
    } else if (variant == XB_GAE_AUTO) {
        use_tma = tma_eligible(p) && get_encode_fn() != nullptr;  // measured 2x faster at T=2048 x N=2^20 (profiles/)
    } else if (variant != XB_GAE_LDG) {
        return XB_E_BADARG;
    }

    if (use_tma) {
        CUtensorMap m_rew, m_val, m_term;
        if (!make_map_f32(&m_rew, rew, T, N) || !make_map_f32(&m_val, val, T, N) || !make_map_f32(&m_term, term, T, N))
            return XB_E_DRIVER;
        XB_CUDA(cudaFuncSetAttribute(gae_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GaeSmem)));   // per device, cheap: set on every launch
        int grid = ceil_div_i64(N, kTileN);
        gae_tma_kernel<<<grid, kTmaThreads, sizeof(GaeSmem), s>>>(m_rew, m_val, m_term, p);
    } else {
        int grid = ceil_div_i64(N, 128);
        gae_ldg_kernel<8><<<grid, 128, 0, s>>>(p);
    }
    XB_LAUNCH_CHECK();
    return 0;
}
