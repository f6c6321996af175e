"""CUDA-event timings of the tcgen05 dense kernels next to the cuBLAS fp32 calls they replace (L2 flushed between runs).

    python tools/bench_dense.py [M] [H]
"""
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from xuanpolicy_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(600000)   # keep the GPU busy while the host enqueues: no launch latency in the window
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


x = torch.randn(M, H, device=dev)
W = torch.randn(H, H, device=dev) / H ** 0.5
b = torch.randn(H, device=dev)
hw, hb = torch.randn(1, H, device=dev), torch.randn(1, device=dev)
hi, lo = torch.empty_like(W), torch.empty_like(W)
thi, tlo = torch.zeros(H, 2 * H, device=dev), torch.zeros(H, 2 * H, device=dev)
ops.dense_split_weights(W, hi, lo, thi, tlo, 0)
ops.dense_split_weights(W, hi, lo, thi, tlo, H)
y, ho = torch.empty(M, H, device=dev), torch.empty(M, 1, device=dev)
dout = torch.randn(M, 1, device=dev)
dz1 = torch.empty(M, H, device=dev)
bytes_fwd = 2 * M * H * 4
for res in (True, False):
    med, mn = timeit(lambda: ops.dense_fwd(x, hi, lo, b, 0.01, y, hw, hb, ho, b_resident=res))
    print("dense_fwd resident=%s: %.1f us (min %.1f)  %.0f GB/s algorithmic" % (res, med, mn, bytes_fwd / med / 1e3))
med, mn = timeit(lambda: torch.nn.functional.leaky_relu(torch.addmm(b, x, W.t()), 0.01))
print("cuBLAS fp32 addmm + leaky_relu: %.1f us (min %.1f)" % (med, mn))
med, mn = timeit(lambda: ops.dense_dgrad(y, dout, hw, y, dout, hw, thi, tlo, x, 0.01, dz1))
print("dense_dgrad (actor+critic): %.1f us (min %.1f)  %.0f GB/s algorithmic" % (med, mn, 4 * M * H * 4 / med / 1e3))
med, mn = timeit(lambda: (y @ W) + (y @ W))
print("cuBLAS fp32 2 x dgrad mm + add: %.1f us (min %.1f)" % (med, mn))
