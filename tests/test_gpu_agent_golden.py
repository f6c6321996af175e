"""GPU: the live reference's `PPOCLIP_Agent.train` run (tests/golden/agent_ppo_*.npz, yaml defaults use_obsnorm /
use_rewnorm on; oracle/make_goldens_agent.py) reproduced — with the recorded actions and minibatch permutations replayed —
  (1) by the NATIVE device-resident agent (xuanpolicy_b200.PPOCLIP_Agent: device normalisers, batched GAE, fused updates),
  (2) by the UNMODIFIED reference agent (pip-installed copy in oracle/_ref, see oracle/build_ref.py) wired to the three
      drop-ins exactly as INTEGRATION.md §1b/§1c says: registry key for the vec env, module-global injection of the
      buffer and the learner (ppoclip_agent.py:25,37).
Bars: physics-derived integers / flags exact; float32 buffers 1e-5; returns / advantages 2e-5 of max(|x|, rms);
running statistics 1e-5; parameters after an update phase 1e-4 relative."""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import gae_close, load_golden
from tests.test_oracle_agent_golden import check_rollout

pytestmark = pytest.mark.gpu


class _TapedAgentMixin:
    """Replays recorded actions instead of sampling (the unfused rollout path: sample -> env step -> store)."""

    def set_tapes(self, actions, perms):
        self._tape, self._tape_pos = torch.as_tensor(actions, device=self.device), 0
        self._perm_tape, self._perm_pos = [torch.as_tensor(p, dtype=torch.int64).pin_memory() for p in perms], 0
        feeder = self._feeder
        agent = self

        def get(iteration, ep):
            p = agent._perm_tape[iteration * agent.n_epoch + ep]
            return p
        feeder.get = get
        feeder.prefetch = lambda it: None
        feeder.mark_consumed = lambda it: None

    def _sample(self, dist, offset):
        N = self.n_envs
        a = self._tape[self._tape_pos]
        self._tape_pos += 1
        prm = dist.get_param()
        if self.discrete:
            d = torch.distributions.Categorical(logits=prm[:N])
            self._act.copy_(a.reshape(N))
            self._logp.copy_(d.log_prob(a.reshape(N)))
        else:
            mu, std = prm
            std = std if std is not None else self.policy.actor.logstd.exp()
            d = torch.distributions.Normal(mu[:N], std)
            self._act.copy_(a.reshape(N, -1))
            self._logp.copy_(d.log_prob(a.reshape(N, -1)).sum(-1))


def _snapshot_native(agent):
    mem = agent.memory
    cpu = lambda t: t.detach().cpu().numpy()
    rms = cpu(agent._obs_rms[agent._rms_cur])
    od = mem.obs_dim
    return dict(obs=cpu(mem.observations), act=cpu(mem.actions), rew=cpu(mem.rewards), val=cpu(mem.values),
                ret=cpu(mem.returns), adv=cpu(mem.advantages), term=cpu(mem.terminals), logp=cpu(mem.auxiliary_infos["old_logp"]),
                obs_rms_mean=rms[:od], obs_rms_var=rms[4:4 + od], obs_rms_count=rms[8])


@pytest.mark.parametrize("name", ["agent_ppo_cartpole", "agent_ppo_pendulum"])
def test_native_agent_reproduces_the_live_reference_run(name, monkeypatch):
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200.configs import build_ppo
    monkeypatch.setenv("XB_FUSED_STEP", "0")
    g = load_golden(name)
    m = g["meta"]

    class Taped(_TapedAgentMixin, xb.PPOCLIP_Agent):
        pass
    h = [m["hidden"]]
    agent = build_ppo(m["env_id"], agent_class=Taped, parallels=m["n_envs"], n_steps=m["n_steps"], n_epoch=m["n_epoch"],
                      n_minibatch=m["n_minibatch"], seed=m["seed"], gamma=m["gamma"], gae_lambda=m["gae_lambda"],
                      representation_hidden_size=h, actor_hidden_size=h, critic_hidden_size=h, use_obsnorm=True,
                      use_rewnorm=True, shuffle="host", use_cuda_graphs=False, running_steps=m["total_iters"],
                      learning_rate=m["learning_rate"])
    agent.policy.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")}, strict=True)
    agent.learner._flat.total_iters = m["total_iters"]
    agent.set_tapes(g["actions"], g["perms"])
    snaps = []
    orig = agent._update_phase

    def update_phase():
        snaps.append(_snapshot_native(agent))
        orig()
        snaps[-1]["params"] = {k: v.detach().cpu().numpy().copy() for k, v in agent.policy.state_dict().items()}
    agent._update_phase = update_phase
    agent.train(m["train_steps"])
    assert len(snaps) == m["n_rollouts"] and agent._tape_pos == len(g["actions"])
    for r, snap in enumerate(snaps):
        p = "r%d/" % r
        assert np.array_equal(snap["act"], g[p + "act"]) and np.array_equal(snap["term"], g[p + "term"])
        for k in ("obs", "rew", "val", "logp"):
            assert np.allclose(snap[k], g[p + k], rtol=1e-5, atol=1e-5), (r, k, np.abs(snap[k] - g[p + k]).max())
        for k in ("ret", "adv"):
            ok, err = gae_close(snap[k], g[p + k], rtol=2e-5)
            assert ok, (r, k, err)
        for k in ("obs_rms_mean", "obs_rms_var"):
            assert np.allclose(snap[k], g[p + k], rtol=1e-5, atol=1e-6), (r, k, snap[k], g[p + k])
        assert np.isclose(snap["obs_rms_count"], float(g[p + "obs_rms_count"]), rtol=1e-9)
        for k, v in snap["params"].items():
            ref = g[p + "params/" + k]
            assert np.allclose(v, ref, rtol=1e-4, atol=5e-6), (r, k, np.abs(v - ref).max())
    # end state: the partial rollout's rows and both running statistics after every step
    mem = agent.memory
    ptr = int(g["end/buffer_ptr"])
    assert mem.ptr == ptr
    assert np.allclose(mem.observations[:, :ptr].cpu().numpy(), g["end/obs_rows"], rtol=1e-5, atol=1e-5)
    assert np.allclose(mem.rewards[:, :ptr].cpu().numpy(), g["end/rew_rows"], rtol=1e-5, atol=1e-5)
    ret = agent._ret_rms.cpu().numpy()
    assert np.isclose(ret[0], float(g["end/ret_rms_mean"]), rtol=1e-5) and np.isclose(ret[1], float(g["end/ret_rms_var"]), rtol=1e-5)
    assert np.isclose(ret[2], float(g["end/ret_rms_count"]), rtol=1e-9)
    assert np.allclose(agent._returns.cpu().numpy(), g["end/returns_tracker"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["agent_ppo_cartpole", "agent_ppo_pendulum"])
def test_unmodified_reference_agent_drives_the_dropins(name):
    """INTEGRATION.md §1b + §1c end to end under the LIVE reference code (oracle/_ref): `get_runner` -> `make_envs` ->
    REGISTRY_VEC_ENV["B200_Gym"]([_thunk] * parallels) builds our vec env from the reference's own closures;
    `PPOCLIP_Agent.__init__` resolves `DummyOnPolicyBuffer` / `PPOCLIP_Learner` from its module globals and builds ours;
    its unmodified `train()` (numpy in / out, list-of-dict infos, per-env finish_path, host-side normalisation) then
    reproduces the run recorded with the reference's own classes."""
    from oracle import build_ref, ref_agent, ref_loader
    if not build_ref.installed() and not ref_loader.available():
        pytest.skip("the reference is not installed in oracle/_ref (python -m oracle.build_ref)")
    import xuanpolicy_b200 as xb
    g = load_golden(name)
    m = g["meta"]
    ref_loader.load(trig="cr")
    import xuance.environment as E
    import xuance.torch.agents.policy_gradient.ppoclip_agent as ref_mod
    saved = (ref_mod.DummyOnPolicyBuffer, ref_mod.PPOCLIP_Learner, ref_mod.tqdm)
    E.REGISTRY_VEC_ENV["B200_Gym"] = xb.DummyVecEnv_Gym                                   # §1b
    ref_mod.DummyOnPolicyBuffer, ref_mod.PPOCLIP_Learner = xb.DummyOnPolicyBuffer, xb.PPOCLIP_Learner   # §1c
    ref_mod.tqdm = lambda x: x
    orig_shuffle = np.random.shuffle
    try:
        h = [m["hidden"]]
        runner = ref_agent.build_runner(m["env_id"], trig="cr", device="cuda:0", vectorize="B200_Gym", parallels=m["n_envs"],
                                        n_steps=m["n_steps"], seed=m["seed"], n_epoch=m["n_epoch"], n_minibatch=m["n_minibatch"],
                                        gamma=m["gamma"], representation_hidden_size=h, actor_hidden_size=h, critic_hidden_size=h)
        agent = runner.agent
        assert type(agent).__module__.startswith("xuance.") and type(agent).__name__ == "PPOCLIP_Agent"
        assert type(agent.envs) is xb.DummyVecEnv_Gym and agent.envs.num_envs == m["n_envs"] and agent.envs.env_id == m["env_id"]
        assert type(agent.memory) is xb.DummyOnPolicyBuffer and type(agent.learner) is xb.PPOCLIP_Learner
        agent.policy.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")}, strict=True)
        tape, perms = g["actions"], g["perms"]
        pos = {"a": 0, "p": 0, "main": True}

        def action(obs):                       # ppoclip_agent.py:50-57 with the recorded draw in place of stochastic_sample()
            _, dists, vs = agent.policy(obs)
            if pos["main"]:
                acts = torch.as_tensor(tape[pos["a"]], device=vs.device)
                pos["a"] += 1
                pos["main"] = False
            else:
                acts = dists.stochastic_sample()
            logps = dists.log_prob(acts)
            return acts.detach().cpu().numpy(), vs.detach().cpu().numpy(), logps.detach().cpu().numpy()
        agent._action = action
        rms_update = agent.obs_rms.update

        def upd(x):
            pos["main"] = True
            return rms_update(x)
        agent.obs_rms.update = upd

        def shuffle(x):
            x[:] = perms[pos["p"]]
            pos["p"] += 1
        np.random.shuffle = shuffle
        mem, snaps = agent.memory, []
        orig_clear = mem.clear

        def clear():
            snaps.append(dict(obs=mem.observations.copy(), act=mem.actions.copy(), rew=mem.rewards.copy(), val=mem.values.copy(),
                              ret=mem.returns.copy(), adv=mem.advantages.copy(), term=mem.terminals.copy(),
                              logp=mem.auxiliary_infos["old_logp"].copy(), returns_tracker=np.asarray(agent.returns, np.float64).copy(),
                              obs_rms_mean=agent.obs_rms.mean, obs_rms_var=agent.obs_rms.var, obs_rms_count=agent.obs_rms.count,
                              ret_rms_mean=agent.ret_rms.mean, ret_rms_var=agent.ret_rms.var, ret_rms_count=agent.ret_rms.count,
                              params={k: v.detach().cpu().numpy().copy() for k, v in agent.policy.state_dict().items()}))
            orig_clear()
        mem.clear = clear
        agent.train(m["train_steps"])
    finally:
        np.random.shuffle = orig_shuffle
        ref_mod.DummyOnPolicyBuffer, ref_mod.PPOCLIP_Learner, ref_mod.tqdm = saved
        del E.REGISTRY_VEC_ENV["B200_Gym"]
    assert len(snaps) == m["n_rollouts"] and pos["a"] == len(tape) and pos["p"] == len(perms)
    for r, snap in enumerate(snaps):
        check_rollout(snap, g, r)
        for k, v in snap["params"].items():
            ref = g["r%d/params/%s" % (r, k)]
            assert np.allclose(v, ref, rtol=1e-4, atol=5e-6), (r, k, np.abs(v - ref).max())
    assert agent.current_step == int(g["end/current_step"]) and mem.ptr == int(g["end/buffer_ptr"])
    assert np.allclose(np.asarray(agent.ret_rms.var), g["end/ret_rms_var"], rtol=1e-5)
    assert agent.learner.iterations == m["n_rollouts"] * m["n_epoch"] * m["n_minibatch"]
