// normalize.cuh — device-side RunningMeanStd pieces shared by normalize.cu, env_classic.cu (fused rollout step) and
// dense_tc.cu (rollout forward from raw observations).
//
// Reference: RunningMeanStd.update / update_from_moments  xuance/common/statistic_tools.py:63-112 (Chan merge)
//            Agent._process_observation / _process_reward xuance/torch/agents/agent.py:104-123
//            the per-env discounted-return tracker        ppoclip_agent.py:87,91-92
// State of an observation normaliser: fp64 [2*D + 1] = mean[D], var[D], count, D = floats per observation ROW (4, or 8
// for wide rows); mean / var hold float32 values (the reference keeps float32 arrays, statistic_tools.py:46-47).
// State of the return normaliser: fp64 [3] = mean, var, count (it evolves in float64 in the reference, see normalize.cu).
#pragma once
#include "common.cuh"

namespace xb {

constexpr float kNormEps = 1e-8f;   // EPS of xuance/torch/agents/agent.py

// Chan et al. merge, in float32 like numpy does with float32 arrays and weak python scalars.
__device__ __forceinline__ void chan_merge(float mean, float var, double count, float b_mean, float b_var,
                                           double b_count, float& new_mean, float& new_var, double& new_count) {
    const double tot = count + b_count;
    const float fc = (float)count, fb = (float)b_count, ft = (float)tot;
    const float delta = b_mean - mean;
    new_mean = mean + delta * fb / ft;
    const float m_a = var * fc, m_b = b_var * fb;
    const float m2 = m_a + m_b + delta * delta * fc * fb / ft;
    new_var = m2 / ft;
    new_count = tot;
}

// np.clip((obs - mean) / (std + EPS), -clip, clip)   (agent.py:112-113), float32 like the reference's arrays
__device__ __forceinline__ float norm_apply(float v, float mean, float den, float clip) {
    return fminf(fmaxf((v - mean) / den, -clip), clip);
}
__device__ __forceinline__ void norm_coeffs(const double* __restrict__ state, int D, int d, float& mean, float& den) {
    mean = (float)state[d];
    den = sqrtf((float)state[D + d]) + kNormEps;
}

// ---- statistics carried by the fused rollout step (env_classic.cu) ---------------------------------------------------
// One launch per vector step also (a) merges the NEXT observations' batch moments into the observation normaliser
// (`obs_rms.update(obs)` of the following loop iteration, ppoclip_agent.py:62), (b) advances the per-env return tracker and
// merges the finished episodes' returns into the return normaliser (:87-92), publishing the reward divisor of the next
// step.  Per-CTA partial sums -> ticket -> the last CTA adds them in CTA order (deterministic) and publishes.
struct StepStats {
    const double* obs_state_in;   // nullable: S_t, read by every thread (normalises the stored observation)
    double* obs_state_out;        // S_{t+1} = S_t merged with the next observations' moments (a different buffer)
    int dim;                      // observation floats (<= 4 * row float4s)
    float obs_clip;
    double* ret_state;            // nullable: (mean, var, count) of the return normaliser
    float* rew_std;               // in: divisor of this step's rewards; out: clip(sqrt(var), 0.1, 100) for the next step
    double* returns;              // per-env discounted-return tracker (fp64 [N])
    double gamma;
    int mask_terminal;            // PPO drops the running return on a terminal (:87); A2C does not (a2c_agent.py:85)
    double* partials;             // scratch [grid][kStepStatSlots]
    unsigned int* ticket;         // scratch, self-resetting
    double* sums_out;             // nullable (env-sharded form): the last CTA writes THIS RANK's totals [2D + 4] = sum x[D],
                                  // sum x^2[D], N, (sum R, sum R^2, n finished) there instead of merging; the ranks' sums are
                                  // then exchanged (peer_comm.cu) and merged by xb_rms_merge_sums
};
constexpr int kStepStatSlots = 20;   // 8 sums + 8 sums of squares + (sum R, sum R^2, n finished) + pad

// Merges batch sums into the two normalisers (threads 0..D-1 of one warp): observation state_in -> state_out (Chan, float32,
// statistic_tools.py:101-112) from (sum x, sum x^2) over n rows; return state in place (fp64) from (sum R, sum R^2, n_ret),
// publishing the reward divisor clip(sqrt(var), 0.1, 100) (agent.py:119-120).
template <int D>
__device__ __forceinline__ void merge_step_stats(const double* obs_state_in, double* obs_state_out, int dim, double* ret_state,
                                                 float* rew_std, const double* sum, const double* sumsq, double n, double r_sum,
                                                 double r_sumsq, double r_n, int tid) {
    if (tid < D && obs_state_in) {
        const int d = tid;
        float nm = (float)obs_state_in[d], nv = (float)obs_state_in[D + d];
        double new_count = obs_state_in[2 * D];
        if (d < dim) {
            const double bm = sum[d] / n;
            double bv = sumsq[d] / n - bm * bm;           // np.square(np.std(x, axis=0))
            bv = bv > 0.0 ? bv : 0.0;
            chan_merge(nm, nv, obs_state_in[2 * D], (float)bm, (float)bv, n, nm, nv, new_count);
        }
        obs_state_out[d] = (double)nm;
        obs_state_out[D + d] = (double)nv;
        if (d == 0) obs_state_out[2 * D] = obs_state_in[2 * D] + n;
    }
    if (tid == 0 && ret_state) {
        if (r_n > 0.0) {
            const double bm = r_sum / r_n;
            double bv = r_sumsq / r_n - bm * bm;
            bv = bv > 0.0 ? bv : 0.0;
            const double count = ret_state[2], tot = count + r_n, delta = bm - ret_state[0];   // Chan merge in fp64
            const double m2 = ret_state[1] * count + bv * r_n + delta * delta * count * r_n / tot;
            ret_state[0] = ret_state[0] + delta * r_n / tot;
            ret_state[1] = m2 / tot;
            ret_state[2] = tot;
        }
        *rew_std = (float)fmin(fmax(sqrt(ret_state[1]), 0.1), 100.0);     // agent.py:120
    }
}

// acc: [0, D) sum x_d | [D, 2D) sum x_d^2 | [2D, 2D+3) finished-return sums.  Called by EVERY thread of the CTA.
template <int D>
__device__ __forceinline__ void step_stats_finish(const StepStats& s, double (&acc)[2 * D + 3], int64_t N, double* smem,
                                                  bool* flag) {
    constexpr int K = 2 * D + 3;
    block_sum<K>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) s.partials[(int64_t)blockIdx.x * kStepStatSlots + k] = acc[k];
        __threadfence();
        *flag = (atomicAdd(s.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!*flag) return;
    __threadfence();
    {   // every thread of the last CTA adds the partials of CTAs tid, tid + blockDim, ... — the loads of U rows in flight at
        // once (one L2 round trip for up to U * blockDim CTAs) — then a fixed-order block reduction: the association order
        // depends only on the launch geometry, so every replay gives the same bits
        constexpr int U = D <= 4 ? 4 : 2;
        const int nthr = blockDim.x, grid = gridDim.x;
        double t[K];
#pragma unroll
        for (int k = 0; k < K; ++k) t[k] = 0.0;
        for (int b0 = threadIdx.x; b0 < grid; b0 += U * nthr) {
            double u[U][K];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int b = b0 + j * nthr;
                const double* p = s.partials + (int64_t)(b < grid ? b : 0) * kStepStatSlots;
#pragma unroll
                for (int k = 0; k < K; ++k) u[j][k] = b < grid ? __ldcg(p + k) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < U; ++j)
#pragma unroll
                for (int k = 0; k < K; ++k) t[k] += u[j][k];
        }
        block_sum<K>(t, smem);                  // (uniform: the whole CTA took this branch)
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) smem[k] = t[k];
        }
        __syncthreads();
    }
    if (threadIdx.x >= 32) return;              // the rest happens in warp 0
    __syncwarp();
    if (s.sums_out) {
        if (threadIdx.x < 2 * D) s.sums_out[threadIdx.x] = smem[threadIdx.x];
        if (threadIdx.x == 0) {
            s.sums_out[2 * D] = (double)N;
            s.sums_out[2 * D + 1] = smem[2 * D];
            s.sums_out[2 * D + 2] = smem[2 * D + 1];
            s.sums_out[2 * D + 3] = smem[2 * D + 2];
        }
    } else {
        merge_step_stats<D>(s.obs_state_in, s.obs_state_out, s.dim, s.ret_state, s.rew_std, smem, smem + D, (double)N,
                            smem[2 * D], smem[2 * D + 1], smem[2 * D + 2], threadIdx.x);
    }
    if (threadIdx.x == 0) *s.ticket = 0u;
}

}  // namespace xb
