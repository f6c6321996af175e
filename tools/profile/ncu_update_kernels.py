"""One eager C2 PPO iteration (Pendulum, 4096 envs x 128 steps, MLP 128, normalisers on) for ncu captures of the update
kernels:  ncu --set full --import-source on -k regex:dense_kmajor_ts -s 140 -c 2 python tools/profile/ncu_update_kernels.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from xuanpolicy_b200.configs import build_ppo

agent = build_ppo("Pendulum-v1", parallels=4096, n_steps=128, n_epoch=1, n_minibatch=8, use_obsnorm=True, use_rewnorm=True,
                  shuffle="device", seed=1, use_cuda_graphs=False)
agent.train(128)
torch.cuda.synchronize()
print("done", agent.last_info.get("actor-loss"))
