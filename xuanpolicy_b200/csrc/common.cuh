// common.cuh — shared helpers for the libxb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/xb200.h"

#define XB_LAUNCH_CHECK()                        \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

#define XB_CUDA(call)                            \
    do {                                         \
        cudaError_t e__ = (call);                \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

namespace xb {

constexpr int kNumSMs = 148;  // B200

static inline int ceil_div_i64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// grid for a grid-stride kernel: enough CTAs to cover n, capped at `waves` CTAs per SM.
static inline int grid_for(int64_t n, int block, int ctas_per_sm) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// A kernel launched with `launch_pdl(..., true, ...)` may become resident while the launch before it on the stream is still
// running; `pdl_wait()` blocks until that launch has completed and its memory is visible (a no-op in a normally launched
// kernel), `pdl_trigger()` lets the NEXT launch become resident.  Every kernel here triggers only after its own wait, so a
// kernel's pre-wait prologue can overlap nothing older than its immediate predecessor.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("XB_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#ifdef XB_STEP_TS   // debug build only (tools/profile/step_timeline.py): %globaltimer stamps of the rollout's two kernels
__device__ __forceinline__ void step_ts(unsigned long long* ts, unsigned long long tag) {
    if (!ts) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long i = atomicAdd(ts, 1ull);
    if (i < 30000ull) { ts[1 + 2 * i] = tag; ts[2 + 2 * i] = t; }
}
#define XB_STEP_STAMP(ts, tag) step_ts(ts, tag)
#else
#define XB_STEP_STAMP(ts, tag) do { } while (0)
#endif

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of K doubles per thread; result valid in thread 0.  smem must hold K * 32 doubles.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = lane < nwarp ? smem[k * 32 + lane] : 0.0;
            v[k] = warp_sum(x);
        }
    }
    __syncthreads();
}

// streaming (read-once / write-once) accesses: keep them out of L1.
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace xb
