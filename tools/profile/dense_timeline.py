"""clock64 timelines of the tcgen05 dense kernels (CTA 0; -DXB_DENSE_TS build): where a tile's time goes per role.

    python tools/profile/dense_timeline.py build      (build container: compiles tools/profile/_dbg/libxb200_ts.so)
    XB200_LIB=tools/profile/_dbg/libxb200_ts.so python tools/profile/dense_timeline.py run [fwd2|dgrad|wgrad]   (B200)

Events (xuanpolicy_b200/csrc/dense_tc.cu XB_TS): role 0 = TMA producer (0: slot free), 1 = MMA issuer (0: before the operand
wait, 1: after it, 3: after issue + commits), 2 = operand warps (0: raw tile landed, 1: transformed, 2: TMEM slot free,
3: arrived), 3 = epilogue per tile (0: accumulator ready, 1..4: chunk starts, 7: accumulator released).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
DBG = os.path.join(HERE, "_dbg")


def build():
    from xuanpolicy_b200.csrc import build as kbuild
    print(kbuild.build(out=os.path.join(DBG, "libxb200_ts.so"), obj_dir=os.path.join(DBG, "obj"), extra=["-DXB_DENSE_TS"]))


def run(which):
    import ctypes as C
    import numpy as np
    import torch
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import _lib
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.learner import FlatAdamState
    from xuanpolicy_b200.policies import make_policy
    lib = _lib.load()
    B, H = 65536, 128
    obs_space, act_space = xb.make_spaces("Pendulum-v1")
    policy = make_policy(obs_space, act_space, hidden=(H,), device="cuda", seed=1)
    FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
    fused = FusedActorCritic(policy)
    obs = torch.randn(B, 4, device="cuda")[:, :3]
    fused.forward(obs)
    b = fused._last[1]
    b["dz1"] = torch.empty(B, H, device="cuda")
    dact, dv2 = torch.randn(B, 1, device="cuda") / B, torch.randn(B, 1, device="cuda") / B
    ts = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
    lib.xb_dense_debug_set_ts.argtypes = [C.c_void_p]
    xr = torch.randn(8192, 4, device="cuda")            # the rollout forward at C2: [obs ; terminal obs] of 4096 envs
    fn = {"fwd2": lambda: fused.stage_hidden(b), "dgrad": lambda: fused.stage_dgrad(b, dact, dv2),
          "wgrad": lambda: fused.stage_wgrad(b, dact, dv2),
          "infer": lambda: fused.forward_inference(xr[:, :3])}[which]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    lib.xb_dense_debug_set_ts(C.c_void_p(ts.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib.xb_dense_debug_set_ts(C.c_void_p(0))
    t = ts.cpu().numpy().reshape(4, 64, 8)
    t0 = t[0, 62, 0]
    rel = lambda v: int(v - t0) if v else -1
    print("== %s: kernel start 0, setup done %d, before final sync %d, after %d" % (which, rel(t[0, 62, 1]), rel(t[0, 63, 0]), rel(t[0, 63, 1])))
    n_it = int((t[1, :62, 0] != 0).sum())
    print("k-block iterations recorded (MMA role): %d" % n_it)
    print(" it | producer slot-free | operand: landed, transformed, tmem-free, arrived | mma: wait-start, wait-end, issued")
    for i in range(min(n_it, 40)):
        print("%3d | %7d | %7d %7d %7d %7d | %7d %7d [%7d %7d %7d] %7d" % (i, rel(t[0, i, 0]), rel(t[2, i, 0]), rel(t[2, i, 1]), rel(t[2, i, 2]),
                                                         rel(t[2, i, 3]), rel(t[1, i, 0]), rel(t[1, i, 1]), rel(t[1, i, 4]), rel(t[1, i, 5]), rel(t[1, i, 6]), rel(t[1, i, 3])))
    mma = np.array([rel(t[1, i, 3]) for i in range(n_it)])
    if n_it > 8:
        print("MMA issue-to-issue per k-block: mean %.0f cycles (all), median %.0f" % (np.diff(mma).mean(), np.median(np.diff(mma))))
    n_t = int((t[3, :62, 0] != 0).sum())
    print(" tile | epilogue: acc ready, chunk starts..., released")
    for i in range(min(n_t, 10)):
        print("%3d | %s" % (i, " ".join("%7d" % rel(v) for v in t[3, i])))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        for w in (sys.argv[2:] or ["fwd2", "dgrad", "wgrad"]):
            run(w)
