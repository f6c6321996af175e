import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import xuanpolicy_b200 as xb
from xuanpolicy_b200 import ops
from xuanpolicy_b200.fused_mlp import FusedActorCritic
from xuanpolicy_b200.policies import make_policy
B = 65536
obs_space, act_space = xb.make_spaces("Pendulum-v1")
policy = make_policy(obs_space, act_space, hidden=(128,), device="cuda", seed=5)
fused = FusedActorCritic(policy)
obs = torch.randn(B, 4, device="cuda")[:, :3]
fused.forward(obs)
b = fused._buf[B]
dact = torch.randn(B, 1, device="cuda") / B
dv2 = torch.randn(B, 1, device="cuda") / B
b["dz1"] = torch.empty(B, 128, device="cuda")
b["h1s"] = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
ops.mlp_trunk_fwd(obs, fused.l0.weight.data, fused.l0.bias.data, fused.slope, b["h1"], h1_signs=b["h1s"])
flush = torch.empty(64 << 20, device="cuda")
def timeit(fn, n=30):
    for _ in range(5): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]
for name, signs, h1s in (("signs+h1 signs", b["signs"], b["h1s"]), ("signs", b["signs"], None), ("tiles", None, None)):
    f = lambda: ops.dense_dgrad(b["ya"], dact, fused.la2.weight.data, b["yc"], dv2, fused.lc2.weight.data, fused.wtm_hi,
                                fused.wtm_lo, b["h1"], fused.slope, b["dz1"], wt_form=1, signs=signs, h1_signs=h1s)
    print("dgrad", name, "median %.1f us  min %.1f us" % timeit(f))
print("fwd2 (+signs) median %.1f min %.1f" % timeit(lambda: fused.stage_hidden(b)))
print("fwd2 (signs, no Y) median %.1f min %.1f" % timeit(lambda: fused.stage_hidden(b, keep_y=False)))
fused.stage_hidden(b)
print("fwd2 from obs (signs, no Y, h1 out) median %.1f min %.1f" % timeit(lambda: fused.stage_hidden(b, keep_y=False, obs=obs)))
fused.stage_hidden(b)
dv = dv2
print("wgrad (binary) median %.1f min %.1f" % timeit(lambda: fused.stage_wgrad(b, dact, dv2)))
fused.bin_wgrad = False
print("wgrad (3 MMA) median %.1f min %.1f" % timeit(lambda: fused.stage_wgrad(b, dact, dv2)))
