import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xuanpolicy_b200 import ops
H = 128
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
W = torch.randn(H, H, device=dev) / H ** 0.5; b = torch.randn(H, device=dev)
hw, hb = torch.randn(1, H, device=dev), torch.randn(1, device=dev)
hi, lo = torch.empty_like(W), torch.empty_like(W); ops.dense_split_weights(W, hi, lo)
def timeit(fn, n=10, do_flush=True):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        if do_flush: flush.zero_()
        torch.cuda._sleep(600000)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b_.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b_) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for mult in (1, 2, 4, 8, 16, 32):
    M = 128 * 148 * mult
    x = torch.randn(M, H, device=dev); y, ho = torch.empty(M, H, device=dev), torch.empty(M, 1, device=dev)
    for dbg in (0, 7):
        os.environ["XB_DENSE_DEBUG"] = str(dbg)
        print("tiles/CTA=%d dbg=%d: flush %.1f us  hot %.1f us" % (mult, dbg,
              timeit(lambda: ops.dense_fwd(x, hi, lo, b, 0.01, y, hw, hb, ho), do_flush=True),
              timeit(lambda: ops.dense_fwd(x, hi, lo, b, 0.01, y, hw, hb, ho), do_flush=False)))
# empty-ish kernel for launch floor
z = torch.zeros(1024, device=dev)
print("tiny torch kernel:", timeit(lambda: z.add_(1.0), do_flush=False))
