"""One launch each of the tcgen05 kernels at the C2 update shape (65 536 x 128), for ncu source-level captures:
    ncu --set full --import-source on -k regex:"dense_kmajor_ts|dense_wgrad" -c 6 python tools/profile/ncu_dense_one.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import xuanpolicy_b200 as xb
from xuanpolicy_b200.fused_mlp import FusedActorCritic
from xuanpolicy_b200.learner import FlatAdamState
from xuanpolicy_b200.policies import make_policy

B = 65536
obs_space, act_space = xb.make_spaces("Pendulum-v1")
policy = make_policy(obs_space, act_space, hidden=(128,), device="cuda", seed=5)
FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
fused = FusedActorCritic(policy)
obs = torch.randn(B, 4, device="cuda")[:, :3]
dact = torch.randn(B, 1, device="cuda") / B
dv = torch.randn(B, device="cuda") / B
for _ in range(2):
    fused.forward(obs)
    fused.backward(dact, dv)
torch.cuda.synchronize()
print("done")
