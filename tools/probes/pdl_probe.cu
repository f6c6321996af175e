// pdl_probe.cu — how much does programmatic dependent launch save per kernel boundary inside a CUDA graph on B200?
// Chain of K dependent launches of a small kernel (grid x block, ~work iterations each), captured into a graph,
// replayed with and without the programmatic-stream-serialization attribute.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu && ./pdl_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void work_kernel(float* __restrict__ buf, int n, int iters, int use_pdl, int early_trigger) {
    if (use_pdl && early_trigger) asm volatile("griddepcontrol.launch_dependents;");
    // "prologue" that does not depend on the previous kernel
    __shared__ float s[256];
    s[threadIdx.x] = (float)threadIdx.x;
    __syncthreads();
    if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v = i < n ? buf[i] : 0.f;
    for (int k = 0; k < iters; ++k) v = v * 1.0000001f + s[(threadIdx.x + k) & 255] * 1e-9f;
    if (i < n) buf[i] = v;
}

static float run(int K, int grid, int block, int iters, int pdl, int early, int reps) {
    float* buf;
    int n = grid * block;
    cudaMalloc(&buf, n * sizeof(float));
    cudaMemset(buf, 0, n * sizeof(float));
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int k = 0; k < K; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(block);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, work_kernel, buf, n, iters, pdl, early);
        if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (e != cudaSuccess) { printf("capture error %s\n", cudaGetErrorString(e)); exit(1); }
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int r = 0; r < 3; ++r) cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(b, s);
    cudaStreamSynchronize(s);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaFree(buf);
    return ms * 1e3f / (reps * K);
}

int main() {
    const int K = 200, reps = 20;
    int grids[] = {32, 148, 592};
    int iters[] = {0, 2000, 10000};
    printf("us per launch inside a %d-kernel graph chain (plain | pdl wait-only | pdl early trigger)\n", K);
    for (int g : grids)
        for (int it : iters) {
            float t0 = run(K, g, 256, it, 0, 0, reps), t1 = run(K, g, 256, it, 1, 0, reps), t2 = run(K, g, 256, it, 1, 1, reps);
            printf("grid %4d iters %6d : %7.2f | %7.2f | %7.2f\n", g, it, t0, t1, t2);
        }
    return 0;
}
