"""CPU, world_size 2, gloo: the env-sharding protocol (two collectives per update) reproduces the single-process
minibatch statistics and gradient.  The kernels are stood in for by the oracle's torch formulas — this checks
the host-side sharding logic, not the CUDA code."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_port
from xuanpolicy_b200 import dist as xdist
from xuanpolicy_b200 import policies, spaces


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(4)
    B = 256
    obs = torch.as_tensor(rng.standard_normal((B, 3)).astype(np.float32))
    act = torch.as_tensor(rng.standard_normal((B, 1)).astype(np.float32))
    ret = torch.as_tensor(rng.standard_normal(B).astype(np.float32))
    adv = torch.as_tensor((rng.standard_normal(B) * 2 + 0.5).astype(np.float32))
    old = torch.as_tensor((-rng.random(B)).astype(np.float32))
    return obs, act, ret, adv, old


def _policy():
    torch.manual_seed(0)
    return policies.make_policy(spaces.Box(-1, 1, (3,)), spaces.Box(-2.0, 2.0, (1,)), hidden=(32,), device="cpu")


def _grad(pol, obs, act, ret, adv_n, old, scale):
    for p in pol.parameters():
        p.grad = None
    _, d, v = pol(obs)
    loss, *_ = ref_port.ppo_clip_loss(d.log_prob(act), d.entropy(), v, ret, adv_n, old, 0.25, 0.01, 0.2)
    (loss * scale).backward()
    return torch.cat([p.grad.reshape(-1) for p in pol.parameters()])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, _, w = xdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    obs, act, ret, adv, old = _problem()
    lo, hi = xdist.shard_envs(obs.shape[0], world, rank)          # here: shard the minibatch rows like env shards
    sl = slice(lo, hi)
    stats = torch.stack([adv[sl].double().sum(), (adv[sl].double() ** 2).sum()])
    xdist.allreduce_adv_stats(stats)
    mean, std = xdist.mean_std_from_stats(stats, obs.shape[0])
    adv_n = (adv[sl] - mean.float()) / (std.float() + 1e-8)
    pol = _policy()
    flat = torch.zeros(sum(p.numel() for p in pol.parameters()))
    xdist.broadcast_parameters(flat)
    # local mean over B/W samples scaled by 1/W  ==  sum over local samples / (B_local * W)
    g = _grad(pol, obs[sl], act[sl], ret[sl], adv_n, old[sl], 1.0 / world)
    xdist.allreduce_flat_grad(g)
    if rank == 0:
        torch.save({"mean": mean, "std": std, "grad": g}, out)
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    obs, act, ret, adv, old = _problem()
    a = adv.numpy()
    assert abs(got["mean"].item() - a.mean(dtype=np.float64)) < 1e-12
    assert abs(got["std"].item() - a.astype(np.float64).std()) < 1e-10
    adv_n = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)          # memory_tools.py:241-242
    ref = _grad(_policy(), obs, act, ret, adv_n, old, 1.0)
    assert torch.allclose(got["grad"], ref, atol=2e-6, rtol=1e-4)


def test_shard_envs_partitions():
    assert [xdist.shard_envs(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 65536)]
    try:
        xdist.shard_envs(10, 4, 0)
    except ValueError:
        pass
    else:
        raise AssertionError("unequal shards must be rejected")
