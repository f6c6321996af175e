import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xuanpolicy_b200 import ops, _lib
H = 128; dev = "cuda"; M = 128 * 148 * 4
W = torch.randn(H, H, device=dev) / H ** 0.5; b = torch.randn(H, device=dev)
hw, hb = torch.randn(1, H, device=dev), torch.randn(1, device=dev)
hi, lo = torch.empty_like(W), torch.empty_like(W); ops.dense_split_weights(W, hi, lo)
x = torch.randn(M, H, device=dev); y, ho = torch.empty(M, H, device=dev), torch.empty(M, 1, device=dev)
lib = ctypes.CDLL(_lib.LIB_PATH)
for res in (True, False):
    lib.xb_dense_debug_set_ts(ctypes.c_void_p(0))
    for _ in range(3): ops.dense_fwd(x, hi, lo, b, 0.01, y, hw, hb, ho, b_resident=res)
    ts = torch.zeros(4 * 64 * 8, dtype=torch.int64, device=dev)
    lib.xb_dense_debug_set_ts(ctypes.c_void_p(ts.data_ptr()))
    ops.dense_fwd(x, hi, lo, b, 0.01, y, hw, hb, ho, b_resident=res)
    torch.cuda.synchronize()
    t = ts.cpu().view(4, 64, 8)
    t0 = int(t[t > 0].min())
    f = lambda v: [int(a) - t0 if a > 0 else -1 for a in v]
    print("=== resident", res)
    for it in range(16):
        print("it %2d  prod[empty]=%s  conv[full,loempty,work,arrive]=%s  mma[top,conv,issued,committed]=%s" % (it, f(t[0, it, :1]), f(t[2, it, :4]), f(t[1, it, :4])))
    for lt in range(4):
        print("tile", lt, "epi[tfull,c0,c1,c2,c3,-,-,done]=", f(t[3, lt, :8]))
