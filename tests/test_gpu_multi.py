"""GPU, 2 ranks (NCCL): env-sharded training keeps the replicated policy bit-identical across ranks and the
captured stage graphs + collectives produce finite, decreasing critic loss.  Skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, %r)
from xuanpolicy_b200 import dist as xd
from xuanpolicy_b200.configs import build_ppo
rank, local, world = xd.init_from_env("nccl")
agent = build_ppo("CartPole-v1", device="cuda", parallels=256, n_steps=32, n_epoch=2, n_minibatch=4, shuffle="host",
                  seed=1 + 1000 * rank, gamma=0.99)
for graphs in (True,):
    info = agent.train(3 * 32)
flat = agent.learner._flat.flat_param
gathered = [torch.empty_like(flat) for _ in range(world)]
torch.distributed.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
envs_differ = True
st = agent.envs._state.clone()
others = [torch.empty_like(st) for _ in range(world)]
torch.distributed.all_gather(others, st)
envs_differ = not torch.equal(others[0], others[1])
if rank == 0:
    print("RESULT same_params=%%s envs_differ=%%s critic=%%.4f steps=%%d" %% (same, envs_differ, info["critic-loss"], int(agent.learner._flat.step.item())))
torch.distributed.destroy_process_group()
''' % REPO


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_training_stays_in_sync(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    assert "same_params=True" in line and "envs_differ=True" in line and "steps=24" in line, line
