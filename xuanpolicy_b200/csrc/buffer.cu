// buffer.cu — device-resident rollout buffer: per-step store and minibatch gather.
//
// Replaces  store_element / DummyOnPolicyBuffer.store      xuance/common/memory_tools.py:39-54,196-204
//           sample_batch / DummyOnPolicyBuffer.sample      xuance/common/memory_tools.py:57-72,231-245
//
// Layout: time-major.  The reference keeps env-major [n_envs, n_size] numpy arrays, which makes the per-step
// write `memory[:, ptr] = data` a stride-T scatter.  Here a step is one contiguous row of N entries per field
// (observations: N float4), so the store is a pure streaming write and the GAE scan (gae.cu) reads columns
// with unit stride across threads.  The reference's flat sample index k = env*T + step (memory_tools.py:234)
// is honoured by the gather: row = (k % T) * N + k / T.
//
// HBM traffic: store 36 B/transition written (+36 B read of the step's staging vectors);
// obs gather 8 B index + 16 B row read + 4*obs_dim B written (+4 B advantage for the statistics).
#include "common.cuh"

namespace xb {

__global__ void __launch_bounds__(256)
    store_kernel(const float4* __restrict__ obs, const void* __restrict__ act, int act_is_i64, int act_dim,
                 const float* __restrict__ rew, const float* __restrict__ val, const uint8_t* __restrict__ term,
                 const uint8_t* __restrict__ trunc, const float* __restrict__ logp, float4* __restrict__ obs_row,
                 float* __restrict__ act_row, float* __restrict__ rew_row, float* __restrict__ val_row,
                 float* __restrict__ term_row, uint8_t* __restrict__ trunc_row, float* __restrict__ logp_row,
                 const float* __restrict__ rew_scale, float rew_clip, int obs_vec, int64_t N) {
    const float std = rew_scale ? *rew_scale : 1.0f;   // rew_scale carries rew_std: rewards are DIVIDED by it
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N; e += (int64_t)gridDim.x * blockDim.x) {
        obs_row[e * obs_vec] = obs[e * obs_vec];
        if (obs_vec == 2) obs_row[e * 2 + 1] = obs[e * 2 + 1];   // wide observations (5..8 floats): two float4 per row
        if (act_is_i64) {
            act_row[e] = (float)((const int64_t*)act)[e];
        } else {
            for (int k = 0; k < act_dim; ++k) act_row[e * act_dim + k] = ((const float*)act)[e * act_dim + k];
        }
        float r = rew[e];
        if (rew_scale) {
            r = r / std;
            r = fminf(fmaxf(r, -rew_clip), rew_clip);
        }
        rew_row[e] = r;
        val_row[e] = val[e];
        term_row[e] = term[e] ? 1.0f : 0.0f;
        if (trunc_row) trunc_row[e] = trunc ? trunc[e] : (uint8_t)0;
        logp_row[e] = logp[e];
    }
}

__device__ __forceinline__ int64_t flat_to_row(int64_t k, int64_t T, int64_t N) {
    int64_t env = k / T;
    int64_t step = k - env * T;
    return step * N + env;
}

// Last-block-done reduction of per-block (sum, sumsq) partials into stats[0..1]; deterministic order.
__device__ __forceinline__ void finish_stats(double s, double ss, double* partials, unsigned int* ticket,
                                             double* stats, double* smem) {
    __shared__ bool is_last;
    double v[2] = {s, ss};
    block_sum<2>(v, smem);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = v[0];
        partials[2 * blockIdx.x + 1] = v[1];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double a[2] = {0.0, 0.0};
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
            a[0] += partials[2 * b];
            a[1] += partials[2 * b + 1];
        }
        block_sum<2>(a, smem);
        if (threadIdx.x == 0) {
            stats[0] = a[0];
            stats[1] = a[1];
            *ticket = 0u;  // re-arm for the next launch
        }
    }
}

constexpr int kGatherBlock = 256;
constexpr int kGatherMaxGrid = kNumSMs * 8;

// scratch for the statistics reduction (one per process; launches on one stream are serialised)
__device__ double g_partials[2 * kGatherMaxGrid];
__device__ unsigned int g_ticket = 0;

// wide observation rows (obs_dim 5..8): two float4 per transition
__global__ void __launch_bounds__(kGatherBlock)
    gather_obs_wide_kernel(const int64_t* __restrict__ idx, int64_t B, int64_t T, int64_t N, const float4* __restrict__ b_obs,
                           int obs_dim, const float* __restrict__ b_adv, float* __restrict__ obs_out,
                           double* __restrict__ stats) {
    __shared__ double smem[64];
    double s = 0.0, ss = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = flat_to_row(idx[i], T, N);
        const float4 a = b_obs[2 * row], b = b_obs[2 * row + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        for (int k = 0; k < obs_dim; ++k) obs_out[i * obs_dim + k] = v[k];
        if (stats) {
            const double adv = (double)b_adv[row];
            s += adv;
            ss += adv * adv;
        }
    }
    if (stats) finish_stats(s, ss, g_partials, &g_ticket, stats, smem);
}

template <int OBS_DIM>
__global__ void __launch_bounds__(kGatherBlock)
    gather_obs_kernel(const int64_t* __restrict__ idx, int64_t B, int64_t T, int64_t N,
                      const float4* __restrict__ b_obs, const float* __restrict__ b_adv, float* __restrict__ obs_out,
                      double* __restrict__ stats) {
    __shared__ double smem[64];
    double s = 0.0, ss = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = flat_to_row(idx[i], T, N);
        float4 o = b_obs[row];
        if (OBS_DIM == 4) {
            reinterpret_cast<float4*>(obs_out)[i] = o;
        } else {
            float* dst = obs_out + i * OBS_DIM;
            dst[0] = o.x;
            if (OBS_DIM > 1) dst[1] = o.y;
            if (OBS_DIM > 2) dst[2] = o.z;
        }
        if (stats) {
            double a = (double)b_adv[row];
            s += a;
            ss += a * a;
        }
    }
    if (stats) finish_stats(s, ss, g_partials, &g_ticket, stats, smem);
}

__global__ void __launch_bounds__(kGatherBlock)
    gather_batch_kernel(const int64_t* __restrict__ idx, int64_t B, int64_t T, int64_t N,
                        const float4* __restrict__ b_obs, int obs_dim, const float* __restrict__ b_act, int act_dim,
                        const float* __restrict__ b_ret, const float* __restrict__ b_val,
                        const float* __restrict__ b_adv, const float* __restrict__ b_logp,
                        float* __restrict__ obs_out, float* __restrict__ act_out, float* __restrict__ ret_out,
                        float* __restrict__ val_out, float* __restrict__ adv_out, float* __restrict__ logp_out,
                        double* __restrict__ stats) {
    __shared__ double smem[64];
    double s = 0.0, ss = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = flat_to_row(idx[i], T, N);
        if (obs_out) {
            if (obs_dim <= 4) {
                float4 o = b_obs[row];
                float v[4] = {o.x, o.y, o.z, o.w};
                for (int k = 0; k < obs_dim; ++k) obs_out[i * obs_dim + k] = v[k];
            } else {   // wide rows: two float4 per transition
                const float4 a = b_obs[2 * row], b = b_obs[2 * row + 1];
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                for (int k = 0; k < obs_dim; ++k) obs_out[i * obs_dim + k] = v[k];
            }
        }
        if (act_out)
            for (int k = 0; k < act_dim; ++k) act_out[i * act_dim + k] = b_act[row * act_dim + k];
        if (ret_out) ret_out[i] = b_ret[row];
        if (val_out) val_out[i] = b_val[row];
        if (logp_out) logp_out[i] = b_logp[row];
        if (adv_out || stats) {
            float a = b_adv[row];
            if (adv_out) adv_out[i] = a;
            s += (double)a;
            ss += (double)a * (double)a;
        }
    }
    if (stats) finish_stats(s, ss, g_partials, &g_ticket, stats, smem);
}

// ---- packed 32-byte transition records -------------------------------------------------------------------
// A random minibatch gather over SoA arrays touches one 32 B DRAM sector per FIELD per sample (obs row, act, logp,
// adv, ret: 5 sectors = 160 B for 32 useful bytes).  Once per rollout, after the GAE scan, the fields a PPO update
// needs are packed into one sector-sized, sector-aligned record per transition: {obs[4], act, old_logp, adv, ret}.
// The minibatch gather then reads exactly one sector per sample (100 % sector efficiency) and emits the MLP input
// and a compact float4 {act, old_logp, adv, ret} per sample, which the loss kernel streams fully coalesced.
struct __align__(32) Record {
    float4 obs;
    float4 s;  // act, old_logp, adv, ret
};

__global__ void __launch_bounds__(256)
    pack_records_kernel(const float4* __restrict__ b_obs, const float* __restrict__ b_act,
                        const float* __restrict__ b_logp, const float* __restrict__ b_adv,
                        const float* __restrict__ b_ret, Record* __restrict__ rec, int64_t TN) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < TN; i += (int64_t)gridDim.x * blockDim.x) {
        Record r;
        r.obs = b_obs[i];
        r.s = make_float4(b_act[i], b_logp[i], b_adv[i], b_ret[i]);
        rec[i] = r;
    }
}

template <int OBS_DIM>
__global__ void __launch_bounds__(kGatherBlock)
    gather_records_kernel(const int64_t* __restrict__ idx, int64_t B, int64_t T, int64_t N,
                          const Record* __restrict__ rec, float* __restrict__ obs_out, float4* __restrict__ scal_out,
                          double* __restrict__ stats) {
    __shared__ double smem[64];
    double s = 0.0, ss = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const Record r = rec[flat_to_row(idx[i], T, N)];
        if (OBS_DIM == 4) {
            reinterpret_cast<float4*>(obs_out)[i] = r.obs;
        } else {
            float* dst = obs_out + i * OBS_DIM;
            dst[0] = r.obs.x;
            if (OBS_DIM > 1) dst[1] = r.obs.y;
            if (OBS_DIM > 2) dst[2] = r.obs.z;
        }
        scal_out[i] = r.s;
        const double a = (double)r.s.z;
        s += a;
        ss += a * a;
    }
    if (stats) finish_stats(s, ss, g_partials, &g_ticket, stats, smem);
}

__global__ void __launch_bounds__(256)
    normalize_adv_kernel(float* __restrict__ adv, const double* __restrict__ stats, double inv_count, int64_t B) {
    double mean = stats[0] * inv_count;
    double var = stats[1] * inv_count - mean * mean;
    float std = (float)sqrt(var > 0.0 ? var : 0.0);
    float m = (float)mean;
    float denom = std + 1e-8f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
        adv[i] = (adv[i] - m) / denom;
}

// ------------------------------------------------------------------------------------------------ gather + trunk layer
// xb_gather_records fused with the MLP's first layer (mlp_trunk.cu trunk_fwd_kernel: Linear(obs_dim, H) + LeakyReLU,
// xuance/torch/representations/mlp.py:40-51): the gathered observation never makes a round trip through HBM before the
// layer that consumes it, and the launch-latency-bound gather hides behind the store-bound layer (B x H floats written).
// Warp-centric: a warp first gathers 32 rows (lane l loads the index and the 32-byte record of row base + l: 32
// independent sector reads in flight, emits obs_out / scal_out and carries the advantage statistics), then walks those 32
// rows, broadcasting each row's observation with shuffles while every lane computes its 4 output features with exactly
// trunk_fwd_kernel's arithmetic and the warp stores one coalesced 512-byte piece of h1 per row.  H = 128: one warp covers a
// row; H = 256: two warps (feature halves) walk the same rows, only the first emits the gathered outputs.
template <int OBS_DIM>
__global__ void __launch_bounds__(256)
    gather_trunk_fwd_kernel(const int64_t* __restrict__ idx, int64_t B, int64_t T, int64_t N, const Record* __restrict__ rec,
                            const float* __restrict__ W0, const float* __restrict__ b0, float slope, int H,
                            float* __restrict__ obs_out, float4* __restrict__ scal_out, double* __restrict__ stats,
                            float4* __restrict__ h1, uint32_t* __restrict__ h1_signs) {
    __shared__ double smem[64];
    pdl_wait();                                 // (launched with the programmatic-serialization attribute, common.cuh)
    pdl_trigger();
    const int tpr = H >> 2;                     // float4 outputs per row
    const int wpr = tpr >> 5;                   // warps per row (1 or 2)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = warp % wpr;                // which 128-feature slice of the row this warp computes
    const int q = half * 32 + lane;
    float w[4][OBS_DIM], bias[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        bias[e] = b0[4 * q + e];
#pragma unroll
        for (int i = 0; i < OBS_DIM; ++i) w[e][i] = W0[(4 * q + e) * OBS_DIM + i];
    }
    double s = 0.0, ss = 0.0;
    const int64_t n_chunks = (B + 31) >> 5;
    const int64_t streams = (int64_t)gridDim.x * (blockDim.x >> 5) / wpr;
    for (int64_t c = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) / wpr; c < n_chunks; c += streams) {
        const int64_t b = c * 32 + lane;
        Record r;
        r.obs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B) {
            r = rec[flat_to_row(idx[b], T, N)];
            if (half == 0) {
                if (OBS_DIM == 4) {
                    reinterpret_cast<float4*>(obs_out)[b] = r.obs;
                } else {
                    float* dst = obs_out + b * OBS_DIM;
                    dst[0] = r.obs.x;
                    if (OBS_DIM > 1) dst[1] = r.obs.y;
                    if (OBS_DIM > 2) dst[2] = r.obs.z;
                }
                scal_out[b] = r.s;
                const double a = (double)r.s.z;
                s += a;
                ss += a * a;
            }
        }
        const int rows = (int)((B - c * 32) < 32 ? (B - c * 32) : 32);
#pragma unroll 4
        for (int j = 0; j < rows; ++j) {
            float x[4];
            x[0] = __shfl_sync(0xffffffffu, r.obs.x, j);
            if (OBS_DIM > 1) x[1] = __shfl_sync(0xffffffffu, r.obs.y, j);
            if (OBS_DIM > 2) x[2] = __shfl_sync(0xffffffffu, r.obs.z, j);
            if (OBS_DIM > 3) x[3] = __shfl_sync(0xffffffffu, r.obs.w, j);
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a = bias[e];
#pragma unroll
                for (int i = 0; i < OBS_DIM; ++i) a += x[i] * w[e][i];
                o[e] = a > 0.f ? a : a * slope;
            }
            h1[(c * 32 + j) * tpr + q] = make_float4(o[0], o[1], o[2], o[3]);
            if (h1_signs) {     // sign words of the row's 128-feature slice: bit l of word e = (h1[128 half + 4 l + e] > 0)
                const uint32_t b0_ = __ballot_sync(0xffffffffu, o[0] > 0.f), b1_ = __ballot_sync(0xffffffffu, o[1] > 0.f);
                const uint32_t b2_ = __ballot_sync(0xffffffffu, o[2] > 0.f), b3_ = __ballot_sync(0xffffffffu, o[3] > 0.f);
                if (lane < 4)
                    h1_signs[(c * 32 + j) * (H >> 5) + half * 4 + lane] = lane == 0 ? b0_ : (lane == 1 ? b1_ : (lane == 2 ? b2_ : b3_));
            }
        }
    }
    if (stats) finish_stats(s, ss, g_partials, &g_ticket, stats, smem);
}

}  // namespace xb

using namespace xb;

extern "C" int xb_store(const float* obs, const void* act, int act_is_i64, int act_dim, const float* rew,
                        const float* val, const uint8_t* term, const uint8_t* trunc, const float* logp,
                        float* obs_row, float* act_row, float* rew_row, float* val_row, float* term_row,
                        uint8_t* trunc_row, float* logp_row, const float* rew_scale, float rew_clip, int obs_vec,
                        int64_t N, xb_stream_t stream) {
    if (N <= 0 || act_dim < 1 || !obs || !act || !rew || !val || !term || !logp || !obs_row || !act_row ||
        !rew_row || !val_row || !term_row || !logp_row)
        return XB_E_BADARG;
    if (obs_vec != 1 && obs_vec != 2) return XB_E_UNSUPPORTED;
    store_kernel<<<grid_for(N, 256, 4), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)obs, act, act_is_i64, act_dim, rew, val, term, trunc, logp, (float4*)obs_row, act_row, rew_row,
        val_row, term_row, trunc_row, logp_row, rew_scale, rew_clip, obs_vec, N);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_gather_obs(const int64_t* idx, int64_t B, int64_t T, int64_t N, const float* b_obs, int obs_dim,
                             const float* b_adv, float* obs_out, double* stats, xb_stream_t stream) {
    if (B <= 0 || T <= 0 || N <= 0 || !idx || !b_obs || !obs_out || (stats && !b_adv)) return XB_E_BADARG;
    if (obs_dim < 1 || obs_dim > 8) return XB_E_UNSUPPORTED;
    int grid = grid_for(B, kGatherBlock, 8);
    cudaStream_t s = (cudaStream_t)stream;
    const float4* o = (const float4*)b_obs;
    if (obs_dim > 4) {
        gather_obs_wide_kernel<<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, o, obs_dim, b_adv, obs_out, stats);
        XB_LAUNCH_CHECK();
        return 0;
    }
    switch (obs_dim) {
        case 4: gather_obs_kernel<4><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, o, b_adv, obs_out, stats); break;
        case 3: gather_obs_kernel<3><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, o, b_adv, obs_out, stats); break;
        case 2: gather_obs_kernel<2><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, o, b_adv, obs_out, stats); break;
        default: gather_obs_kernel<1><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, o, b_adv, obs_out, stats); break;
    }
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_gather_batch(const int64_t* idx, int64_t B, int64_t T, int64_t N, const float* b_obs, int obs_dim,
                               const float* b_act, int act_dim, const float* b_ret, const float* b_val,
                               const float* b_adv, const float* b_logp, float* obs_out, float* act_out,
                               float* ret_out, float* val_out, float* adv_out, float* logp_out, double* stats,
                               xb_stream_t stream) {
    if (B <= 0 || T <= 0 || N <= 0 || !idx) return XB_E_BADARG;
    if (obs_dim < 1 || obs_dim > 8 || act_dim < 1) return XB_E_UNSUPPORTED;
    if ((obs_out && !b_obs) || (act_out && !b_act) || (ret_out && !b_ret) || (val_out && !b_val) ||
        ((adv_out || stats) && !b_adv) || (logp_out && !b_logp))
        return XB_E_BADARG;
    gather_batch_kernel<<<grid_for(B, kGatherBlock, 8), kGatherBlock, 0, (cudaStream_t)stream>>>(
        idx, B, T, N, (const float4*)b_obs, obs_dim, b_act, act_dim, b_ret, b_val, b_adv, b_logp, obs_out, act_out,
        ret_out, val_out, adv_out, logp_out, stats);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_pack_records(const float* b_obs, const float* b_act, const float* b_logp, const float* b_adv,
                               const float* b_ret, float* rec, int64_t TN, xb_stream_t stream) {
    if (TN <= 0 || !b_obs || !b_act || !b_logp || !b_adv || !b_ret || !rec) return XB_E_BADARG;
    if (((uintptr_t)rec & 31u) || ((uintptr_t)b_obs & 15u)) return XB_E_BADARG;
    pack_records_kernel<<<grid_for(TN, 256, 8), 256, 0, (cudaStream_t)stream>>>((const float4*)b_obs, b_act, b_logp, b_adv,
                                                                               b_ret, (Record*)rec, TN);
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_gather_records(const int64_t* idx, int64_t B, int64_t T, int64_t N, const float* rec, int obs_dim,
                                 float* obs_out, float* scal_out, double* stats, xb_stream_t stream) {
    if (B <= 0 || T <= 0 || N <= 0 || !idx || !rec || !obs_out || !scal_out) return XB_E_BADARG;
    if (obs_dim < 1 || obs_dim > 4) return XB_E_UNSUPPORTED;
    if (((uintptr_t)rec & 31u) || ((uintptr_t)scal_out & 15u) || (obs_dim == 4 && ((uintptr_t)obs_out & 15u))) return XB_E_BADARG;
    int grid = grid_for(B, kGatherBlock, 8);
    cudaStream_t s = (cudaStream_t)stream;
    const Record* r = (const Record*)rec;
    float4* so = (float4*)scal_out;
    switch (obs_dim) {
        case 4: gather_records_kernel<4><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, r, obs_out, so, stats); break;
        case 3: gather_records_kernel<3><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, r, obs_out, so, stats); break;
        case 2: gather_records_kernel<2><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, r, obs_out, so, stats); break;
        default: gather_records_kernel<1><<<grid, kGatherBlock, 0, s>>>(idx, B, T, N, r, obs_out, so, stats); break;
    }
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_gather_trunk_fwd(const int64_t* idx, int64_t B, int64_t T, int64_t N, const float* rec, int obs_dim,
                                   const float* W0, const float* b0, float slope, int H, float* obs_out, float* scal_out,
                                   double* stats, float* h1, uint32_t* h1_signs, xb_stream_t stream) {
    if (B <= 0 || T <= 0 || N <= 0 || !idx || !rec || !W0 || !b0 || !obs_out || !scal_out || !h1) return XB_E_BADARG;
    if (obs_dim < 1 || obs_dim > 4 || (H != 128 && H != 256)) return XB_E_UNSUPPORTED;
    if (((uintptr_t)rec & 31u) || ((uintptr_t)scal_out & 15u) || ((uintptr_t)h1 & 15u) ||
        (obs_dim == 4 && ((uintptr_t)obs_out & 15u)))
        return XB_E_BADARG;
    const int64_t warps = (B + 31) / 32 * (H / 128);          // one warp per (32-row chunk, 128-feature slice)
    int grid = grid_for(warps * 32, 256, 8);
    cudaStream_t s = (cudaStream_t)stream;
    const Record* r = (const Record*)rec;
    float4* so = (float4*)scal_out;
    float4* h = (float4*)h1;
    switch (obs_dim) {
        case 4: XB_CUDA(launch_pdl(gather_trunk_fwd_kernel<4>, dim3(grid), dim3(256), 0, s, true, idx, B, T, N, r, W0, b0, slope, H, obs_out, so, stats, h, h1_signs)); break;
        case 3: XB_CUDA(launch_pdl(gather_trunk_fwd_kernel<3>, dim3(grid), dim3(256), 0, s, true, idx, B, T, N, r, W0, b0, slope, H, obs_out, so, stats, h, h1_signs)); break;
        case 2: XB_CUDA(launch_pdl(gather_trunk_fwd_kernel<2>, dim3(grid), dim3(256), 0, s, true, idx, B, T, N, r, W0, b0, slope, H, obs_out, so, stats, h, h1_signs)); break;
        default: XB_CUDA(launch_pdl(gather_trunk_fwd_kernel<1>, dim3(grid), dim3(256), 0, s, true, idx, B, T, N, r, W0, b0, slope, H, obs_out, so, stats, h, h1_signs)); break;
    }
    XB_LAUNCH_CHECK();
    return 0;
}

extern "C" int xb_normalize_adv(float* adv, const double* stats, int64_t count, int64_t B, xb_stream_t stream) {
    if (B <= 0 || count <= 0 || !adv || !stats) return XB_E_BADARG;
    normalize_adv_kernel<<<grid_for(B, 256, 4), 256, 0, (cudaStream_t)stream>>>(adv, stats, 1.0 / (double)count, B);
    XB_LAUNCH_CHECK();
    return 0;
}
