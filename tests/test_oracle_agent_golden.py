"""CPU: pin the oracle's agent loop and normalisers (oracle/ref_port.PPOAgentPort, RunningMeanStdPort) to goldens recorded
from the LIVE reference (`PPOCLIP_Agent.train`, `RunningMeanStd`, `_process_observation/_process_reward`;
oracle/make_goldens_agent.py): with the recorded actions and minibatch permutations replayed, everything else — the
physics under the vec-env protocol, obs / reward normalisation, the per-env finish_path protocol, GAE, per-minibatch
advantage normalisation, the PPO-Clip update, Adam + LinearLR — must reproduce the reference's buffers, statistics and
parameters."""
import json

import numpy as np
import pytest
import torch

from oracle import ref_port
from tests.helpers import gae_close, load_golden
from xuanpolicy_b200 import policies


def _load(name):
    g = load_golden(name)
    return g, g["meta"]


def build_port_agent(g, m, device="cpu", **kw):
    torch.set_num_threads(1)
    envs = ref_port.VecEnvPort(m["env_id"], m["n_envs"], seed=m["seed"], trig=m["trig"])
    envs.reset()
    pol = policies.make_policy(envs.observation_space, envs.action_space, hidden=(m["hidden"],), device=device)
    pol.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")}, strict=True)
    opt = torch.optim.Adam(pol.parameters(), m["learning_rate"], eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=m["total_iters"])
    agent = ref_port.PPOAgentPort(envs, pol, opt, sched, m["n_steps"], m["n_epoch"], m["n_minibatch"], m["gamma"], m["gae_lambda"],
                                  vf_coef=m["vf_coef"], ent_coef=m["ent_coef"], clip_range=m["clip_range"],
                                  clip_grad_norm=m["clip_grad_norm"], use_obsnorm=True, use_rewnorm=True,
                                  obsnorm_range=m["obsnorm_range"], rewnorm_range=m["rewnorm_range"],
                                  action_tape=g["actions"], perm_tape=g["perms"], **kw)
    return agent, pol


def check_rollout(snap, g, r, tol=1e-5):
    """`snap`: dict with the same keys the golden holds for rollout r (env-major arrays)."""
    p = "r%d/" % r
    assert np.array_equal(snap["act"], g[p + "act"])
    assert np.array_equal(snap["term"], g[p + "term"])
    for k in ("obs", "rew", "val", "logp"):
        assert np.allclose(snap[k], g[p + k], rtol=tol, atol=tol), (r, k, np.abs(snap[k] - g[p + k]).max())
    for k in ("ret", "adv"):
        ok, err = gae_close(snap[k], g[p + k], rtol=2e-5)
        assert ok, (r, k, err)
    for k in ("obs_rms_mean", "obs_rms_var", "ret_rms_mean", "ret_rms_var", "returns_tracker"):
        assert np.allclose(snap[k], g[p + k], rtol=1e-5, atol=1e-6), (r, k, snap[k], g[p + k])
    for k in ("obs_rms_count", "ret_rms_count"):
        assert np.isclose(float(snap[k]), float(g[p + k]), rtol=1e-12), (r, k)


@pytest.mark.parametrize("name", ["agent_ppo_cartpole", "agent_ppo_pendulum"])
def test_port_agent_reproduces_the_live_reference_run(name):
    g, m = _load(name)
    agent, pol = build_port_agent(g, m)
    mem = agent.memory
    snaps = []
    orig_clear = mem.clear

    def clear():
        snaps.append(dict(obs=mem.observations.copy(), act=mem.actions.copy(), rew=mem.rewards.copy(), val=mem.values.copy(),
                          ret=mem.returns.copy(), adv=mem.advantages.copy(), term=mem.terminals.copy(), logp=mem.auxiliary_infos["old_logp"].copy(),
                          returns_tracker=np.asarray(agent.returns, np.float64).copy(),
                          obs_rms_mean=agent.obs_rms.mean, obs_rms_var=agent.obs_rms.var, obs_rms_count=agent.obs_rms.count,
                          ret_rms_mean=agent.ret_rms.mean, ret_rms_var=agent.ret_rms.var, ret_rms_count=agent.ret_rms.count,
                          params={k: v.detach().numpy().copy() for k, v in pol.state_dict().items()}, info=dict(agent.last_info)))
        orig_clear()
    mem.clear = clear
    agent.train(m["train_steps"])
    assert len(snaps) == m["n_rollouts"] and agent._tape_pos == len(g["actions"]) and agent._perm_pos == len(g["perms"])
    for r, snap in enumerate(snaps):
        check_rollout(snap, g, r)
        for k, v in snap["params"].items():                       # after n_epoch x n_minibatch clipped Adam steps
            ref = g["r%d/params/%s" % (r, k)]
            assert np.allclose(v, ref, rtol=1e-4, atol=2e-6), (r, k, np.abs(v - ref).max())
        info = json.loads(str(g["r%d/info" % r]))
        for k in ("actor-loss", "critic-loss", "entropy", "predict_value", "learning_rate"):
            assert np.isclose(float(snap["info"][k]), info[k], rtol=1e-3, atol=1e-5), (r, k, snap["info"][k], info[k])
    assert mem.ptr == int(g["end/buffer_ptr"]) and agent.current_step == int(g["end/current_step"])
    assert np.allclose(mem.observations[:, :mem.ptr], g["end/obs_rows"], rtol=1e-5, atol=1e-5)
    assert np.allclose(mem.rewards[:, :mem.ptr], g["end/rew_rows"], rtol=1e-5, atol=1e-5)
    assert np.allclose(agent.obs_rms.mean, g["end/obs_rms_mean"], rtol=1e-5, atol=1e-6)
    assert np.allclose(agent.ret_rms.var, g["end/ret_rms_var"], rtol=1e-5)
    assert np.allclose(agent.returns, g["end/returns_tracker"], rtol=1e-5, atol=1e-6)


def test_running_mean_std_port_matches_the_reference():
    g = load_golden("rms_reference")
    r = ref_port.RunningMeanStdPort((4,))
    for i in range(6):
        r.update(g["vec/batch%d" % i])
        assert np.allclose(r.mean, g["vec/mean%d" % i], rtol=1e-6, atol=1e-7)
        assert np.allclose(r.var, g["vec/var%d" % i], rtol=1e-6, atol=1e-7)
        assert np.isclose(r.count, float(g["vec/count%d" % i]), rtol=1e-12)
    s = ref_port.RunningMeanStdPort(())
    for i, x in enumerate(g["sc/batches"]):
        s.update(np.asarray([x], np.float64))
        assert np.isclose(float(s.mean), float(g["sc/mean%d" % i]), rtol=1e-6)
        assert np.isclose(float(s.var), float(g["sc/var%d" % i]), rtol=1e-6, atol=1e-9)
    # the agent's two normalisers (agent.py:104-123) with those statistics
    class _A(ref_port.PPOAgentPort):
        def __init__(self):
            self.use_obsnorm = self.use_rewnorm = True
            self.obsnorm_range, self.rewnorm_range = g["proc/ranges"]
            self.obs_rms, self.ret_rms = r, s
    a = _A()
    assert np.allclose(a._obs(g["proc/obs_in"]), g["proc/obs_out"], rtol=1e-6, atol=1e-6)
    assert np.allclose(a._rew(g["proc/rew_in"]), g["proc/rew_out"], rtol=1e-6, atol=1e-6)
    s.var = np.float64(1e-6)
    assert np.allclose(a._rew(g["proc/rew_in"]), g["proc/rew_out_floor"], rtol=1e-6, atol=1e-6)
