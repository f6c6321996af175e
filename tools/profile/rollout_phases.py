"""Rollout-graph time at C2 (Pendulum, 4096 envs x 128 steps) for the normaliser combinations: where the per-step time goes."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from xuanpolicy_b200.configs import build_ppo

flush = torch.zeros(64 * 1024 * 1024, device="cuda")
env_id = sys.argv[1] if len(sys.argv) > 1 else "Pendulum-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
for obsn, rewn in ((False, False), (True, False), (False, True), (True, True)):
    agent = build_ppo(env_id, parallels=n, n_steps=128, n_epoch=1, n_minibatch=8, use_obsnorm=obsn, use_rewnorm=rewn,
                      shuffle="device", seed=1, representation_hidden_size=[128], actor_hidden_size=[128], critic_hidden_size=[128])
    with torch.cuda.device(agent.device):
        agent._capture()
    ms = []
    for _ in range(8):
        flush.add_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); agent._rollout_graph.replay(); e.record(); torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    print("obsnorm=%s rewnorm=%s fused_norm=%s rollout_graph_ms mean %.4f min %.4f (%.2f us/step)" % (
        obsn, rewn, agent._fused_norm, np.mean(ms[2:]), np.min(ms), np.min(ms) * 1e3 / 128))
    del agent
