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
steps = int(agent.learner._flat.step.item())
peer = agent.learner._peer
peer_ok = "off"
if peer is not None:
    # kernel-level check of the fused exchange: rank-dependent gradients -> sum bit-identical on every rank,
    # equal to the NCCL all-reduce of the same data, norm equal to the norm of the sum
    fl = agent.learner._flat
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    fl.flat_grad.copy_(torch.randn(fl.n, device="cuda", generator=gen))
    ref = fl.flat_grad.clone()
    torch.distributed.all_reduce(ref)
    p0 = fl.flat_param.clone()
    fl.apply_peer(peer, 0.5, 1.0)
    torch.cuda.synchronize()
    sums = [torch.empty_like(fl.grad_sum) for _ in range(world)]
    torch.distributed.all_gather(sums, fl.grad_sum)
    bitwise = all(torch.equal(sums[0], x) for x in sums)
    close = torch.allclose(fl.grad_sum, ref, rtol=1e-6, atol=1e-6)
    norm_ok = abs(fl.gnorm.item() - ref.double().norm().item()) <= 1e-5 * ref.double().norm().item()
    moved = not torch.equal(p0, fl.flat_param)
    stats_in = torch.arange(6, dtype=torch.float64, device="cuda") * (rank + 1)
    peer.stats[:6].copy_(stats_in)
    out = torch.zeros(6, dtype=torch.float64, device="cuda")
    from xuanpolicy_b200 import ops
    ops.peer_allreduce_f64(peer, 6, out)
    torch.cuda.synchronize()
    stats_ok = torch.equal(out, torch.arange(6, dtype=torch.float64, device="cuda") * sum(range(1, world + 1)))
    peer_ok = "ok" if (bitwise and close and norm_ok and moved and stats_ok) else "BAD(%%s,%%s,%%s,%%s,%%s)" %% (bitwise, close, norm_ok, moved, stats_ok)
# sharded observation / reward normalisation: every rank must hold the same GLOBAL running statistics
norm_ok = "off"
if peer is not None:
    a2 = build_ppo("Pendulum-v1", device="cuda", parallels=96, n_steps=24, n_epoch=1, n_minibatch=2, shuffle="device",
                   seed=7 + 100 * rank, use_obsnorm=True, use_rewnorm=True)
    a2.train(2 * 24)
    torch.cuda.synchronize()
    st = torch.cat([a2._obs_rms[a2._rms_cur], a2._ret_rms, a2._rew_std.double()])
    allst = [torch.empty_like(st) for _ in range(world)]
    torch.distributed.all_gather(allst, st)
    same_stats = all(torch.equal(allst[0], x) for x in allst)
    merged_steps = 2 * 24 + (1 if a2._fused_norm else 0)   # fused path: the observations of the NEXT step are merged already
    count_ok = abs(float(st[8]) - (1e-4 + merged_steps * 96 * world)) < 1e-6       # every env of every rank merged every step
    flat2 = a2.learner._flat.flat_param
    g2 = [torch.empty_like(flat2) for _ in range(world)]
    torch.distributed.all_gather(g2, flat2)
    norm_ok = "ok" if (same_stats and count_ok and all(torch.equal(g2[0], x) for x in g2) and bool(torch.isfinite(st).all())) \
        else "BAD(%%s,%%s,%%s)" %% (same_stats, count_ok, float(st[8]))
if rank == 0:
    print("NORM %%s" %% norm_ok)
if rank == 0:
    print("RESULT same_params=%%s envs_differ=%%s critic=%%.4f steps=%%d peer=%%s" %% (same, envs_differ, info["critic-loss"], steps, peer_ok))
torch.distributed.destroy_process_group()
''' % REPO


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_rank_training_stays_in_sync(tmp_path, peer):
    """peer=1: both exchanges over NVLink peer memory (csrc/peer_comm.cu), the epoch is one CUDA graph;
    peer=0: the NCCL all-reduce path (per-stage graphs).  Either way the replicated policy stays bit-identical."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, XB_PEER_COMM=peer)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533" if peer == "1" else "29534", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    assert "same_params=True" in line and "envs_differ=True" in line and "steps=24" in line, line
    assert ("peer=ok" if peer == "1" else "peer=off") in line, line
    norm = [l for l in out.stdout.splitlines() if l.startswith("NORM")][-1]
    assert norm == ("NORM ok" if peer == "1" else "NORM off"), norm
