"""Stand-alone loss / gather kernels at the C3 update shape (CartPole, 65 536 envs x 256 steps, minibatch 2 097 152):
the launch sequence ncu captures for profiles/r2_ncu_loss_gather_2m_summary.txt.  Run on a B200:
    python tools/profile/loss_gather_2m.py            (plain run; prints CUDA-event timings)
    ncu --set full --clock-control none --import-source on -k regex:'loss_|gather_' -c 12 -o gpurun_out/r2_loss_gather python tools/profile/loss_gather_2m.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import xuanpolicy_b200 as xb  # noqa: E402
from xuanpolicy_b200 import ops  # noqa: E402

N, T, B = 65536, 256, 2097152
dev = "cuda"
obs_space, act_space = xb.make_spaces("CartPole-v1")
mem = xb.DummyOnPolicyBuffer(obs_space, act_space, {"old_logp": ()}, N, T, device=dev, native=True)
gen = torch.Generator(device=dev).manual_seed(7)
for t in (mem._obs, mem._rew, mem._val, mem._logp, mem._adv, mem._ret):
    t.copy_(torch.randn(t.shape, device=dev, generator=gen))
mem._act.copy_(torch.randint(0, 2, mem._act.shape, device=dev, generator=gen).float())
idx = torch.randperm(N * T, device=dev)[:B].contiguous()
obs_out = torch.empty((B, 4), device=dev)
stats = torch.zeros(2, dtype=torch.float64, device=dev)
logits = torch.randn((B, 2), device=dev)
vp = torch.randn(B, device=dev)
dl, dv = torch.empty_like(logits), torch.empty_like(vp)
scal64 = torch.zeros(8, dtype=torch.float64, device=dev)
rec = torch.zeros((N * T, 8), device=dev)
scal = torch.empty((B, 4), device=dev)
flush = torch.zeros(64 * 1024 * 1024, device=dev)

cases = {
    "pack_records": lambda: ops.pack_records(mem._obs, mem._act, mem._logp, mem._adv, mem._ret, rec),
    "gather_obs": lambda: ops.gather_obs(idx, T, N, mem._obs, 4, obs_out, b_adv=mem._adv, stats=stats),
    "gather_records": lambda: ops.gather_records(idx, T, N, rec, 4, obs_out, scal, stats=stats),
    "ppo_loss_idx": lambda: ops.ppo_loss_categorical(logits, vp, mem._act, mem._ret, mem._adv, mem._logp, dl, dv, scal64, 0.2, 0.25,
                                                     0.01, 1.0 / B, idx=idx, T=T, N=N, adv_stats=stats, adv_count=B),
    "ppo_loss_packed": lambda: ops.ppo_loss_categorical(logits, vp, None, None, None, None, dl, dv, scal64, 0.2, 0.25, 0.01,
                                                        1.0 / B, adv_stats=stats, adv_count=B, packed=scal),
}
for name, fn in cases.items():
    ms = []
    for i in range(4):
        flush.add_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    print("%-18s %.4f ms (min of %d after 1 warm-up)" % (name, min(ms[1:]), len(ms) - 1))
