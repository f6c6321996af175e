"""GPU, SURVEY.md §8 row f3: the native device-loop PG_Agent / PPG_Agent against the oracle's restated reference loops
(oracle/ref_port.PGAgentPort / PPGAgentPort — each pinned to the LIVE reference agent in tests/test_oracle_vs_reference.py).
The port runs first on the CPU (restated gym physics, torch CPU MLP, reference-shaped learner updates) and records its
action and permutation tapes; the native agent replays the tapes through the C ABI (device env step, device store,
batched GAE scan, fused loss launches).  Bars: actions / flags exact, float32 buffers 1e-5, returns / advantages 2e-5 of
max(|x|, rms), parameters after every update phase 1e-4 relative."""
import numpy as np
import pytest
import torch

from tests.helpers import gae_close
from tests.test_gpu_agent_golden import _TapedAgentMixin

pytestmark = pytest.mark.gpu


def _snap_port(port):
    m = port.memory
    return dict(obs=m.observations.copy(), act=m.actions.copy(), rew=m.rewards.copy(), val=m.values.copy(),
                ret=m.returns.copy(), adv=m.advantages.copy(), term=m.terminals.copy())


def _snap_native(agent):
    m = agent.memory
    cpu = lambda t: t.detach().cpu().numpy()
    return dict(obs=cpu(m.observations), act=cpu(m.actions), rew=cpu(m.rewards), val=cpu(m.values), ret=cpu(m.returns),
                adv=cpu(m.advantages), term=cpu(m.terminals))


def _record_rollouts(port, steps):
    """Runs the port; returns per-rollout (buffer snapshot, parameters after the update phase)."""
    snaps = []
    mem = port.memory
    clear = mem.clear

    def clear_and_snapshot():                       # end of an update phase
        assert snaps and "params" not in snaps[-1]
        snaps[-1]["params"] = {k: v.detach().clone().numpy() for k, v in port.policy.state_dict().items()}
        clear()
    mem.clear = clear_and_snapshot
    sample = mem.sample

    def sample_and_snapshot(idx):
        if not snaps or "params" in snaps[-1]:
            snaps.append(_snap_port(port))          # first minibatch of an update phase: the finished rollout
        return sample(idx)
    mem.sample = sample_and_snapshot
    port.train(steps)
    return snaps


def _compare(native, ported, what, skip_adv=False):
    assert len(native) == len(ported) and len(native) > 0, (what, len(native), len(ported))
    for r, (a, b) in enumerate(zip(native, ported)):
        assert np.array_equal(a["act"].reshape(b["act"].shape), b["act"]), (what, r, "act")
        assert np.array_equal(a["term"], b["term"]), (what, r, "term")
        for k in ("obs", "rew", "val"):
            assert np.allclose(a[k], b[k], rtol=1e-5, atol=1e-5), (what, r, k, np.abs(a[k] - b[k]).max())
        for k in ("ret",) if skip_adv else ("ret", "adv"):
            ok, err = gae_close(a[k], b[k], rtol=2e-5)
            assert ok, (what, r, k, err)
        for k, v in a["params"].items():
            assert np.allclose(v, b["params"][k], rtol=1e-4, atol=5e-6), (what, r, k, np.abs(v - b["params"][k]).max())


def _hook_update_snapshots(agent):
    snaps = []
    orig = agent._update_phase

    def update_phase():
        snaps.append(_snap_native(agent))
        orig()
        snaps[-1]["params"] = {k: v.detach().cpu().numpy().copy() for k, v in agent.policy.state_dict().items()}
    agent._update_phase = update_phase
    return snaps


@pytest.mark.parametrize("env_id,use_gae", [("CartPole-v1", False), ("Pendulum-v1", False), ("CartPole-v1", True)])
def test_native_pg_agent_reproduces_the_reference_loop(env_id, use_gae, monkeypatch):
    import xuanpolicy_b200 as xb
    from oracle import ref_port
    from xuanpolicy_b200 import policies
    from xuanpolicy_b200.configs import build_pg
    monkeypatch.setenv("XB_FUSED_STEP", "0")
    n, T, n_epoch, h, seed = 6, 20, 2, 32, 4
    steps = 3 * T + 5

    class Taped(_TapedAgentMixin, xb.PG_Agent):
        pass
    agent = build_pg(env_id, agent_class=Taped, parallels=n, n_steps=T, n_epoch=n_epoch, seed=seed, use_gae=use_gae,
                     representation_hidden_size=[h], actor_hidden_size=[h], shuffle="host", use_cuda_graphs=False,
                     running_steps=1000)
    cfg = agent.config
    assert agent.batch_size == n * T // n_epoch and not agent.memory.use_advnorm
    # ---- the restated reference loop on the CPU, same initial parameters
    torch.manual_seed(9)
    np.random.seed(9)
    envs = ref_port.VecEnvPort(env_id, n, seed=seed, trig="cr")
    envs.reset()
    rep = policies.MLPRepresentation(envs.observation_space.shape, [h], activation=torch.nn.ReLU, device="cpu")
    cls = policies.CategoricalActor if env_id == "CartPole-v1" else policies.GaussianActor
    pol = cls(envs.action_space, rep, [h], activation=torch.nn.ReLU, device="cpu")
    pol.load_state_dict({k: v.detach().cpu() for k, v in agent.policy.state_dict().items()}, strict=True)
    opt = torch.optim.Adam(pol.parameters(), cfg.learning_rate, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    port = ref_port.PGAgentPort(envs, pol, opt, sched, T, n_epoch, cfg.gamma, cfg.gae_lambda, ent_coef=cfg.ent_coef,
                                clip_grad=cfg.clip_grad, use_gae=use_gae, use_advnorm=False, use_obsnorm=True, use_rewnorm=True,
                                record=True)
    ported = _record_rollouts(port, steps)
    assert len(ported) == 3 and (port.episodes > 0 or env_id != "CartPole-v1")
    # ---- the native agent on the recorded draws
    agent.set_tapes(np.stack(port.recorded_actions), port.recorded_perms)
    native = _hook_update_snapshots(agent)
    info = agent.train(steps)
    assert set(info) >= {"actor-loss", "entropy", "learning_rate"} and "critic-loss" not in info and "clip_ratio" not in info
    assert np.all(native[0]["val"] == 0.0)                                   # value 0 stored (pg_agent.py:59)
    _compare(native, ported, "pg", skip_adv=True)                            # (the reference's PG never reads the advantages)
    ret = agent._ret_rms.cpu().numpy()
    assert np.isclose(ret[0], float(port.ret_rms.mean), rtol=1e-5, atol=1e-7) and np.isclose(ret[1], float(port.ret_rms.var), rtol=1e-5)
    assert np.isclose(ret[2], float(port.ret_rms.count), rtol=1e-9)
    assert agent.memory.ptr == port.memory.ptr == 5


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Pendulum-v1"])
def test_native_ppg_agent_reproduces_the_reference_loop(env_id, monkeypatch):
    import xuanpolicy_b200 as xb
    from oracle import ref_port
    from xuanpolicy_b200 import policies
    from xuanpolicy_b200.configs import build_ppg
    monkeypatch.setenv("XB_FUSED_STEP", "0")
    n, T, n_epoch, h, seed = 6, 16, 2, 32, 4
    nep = dict(policy_nepoch=2, value_nepoch=2, aux_nepoch=1)
    steps = 2 * T + 3

    class Taped(_TapedAgentMixin, xb.PPG_Agent):
        def set_tapes(self, actions, perms):
            self._tape, self._tape_pos = torch.as_tensor(actions, device=self.device), 0
            self._perm_tape, self._perm_pos = [torch.as_tensor(p, dtype=torch.int64, device=self.device) for p in perms], 0

        def _draw_permutation(self):
            self._perm_pos += 1
            return self._perm_tape[self._perm_pos - 1]
    agent = build_ppg(env_id, agent_class=Taped, parallels=n, n_steps=T, n_epoch=n_epoch, seed=seed, use_cuda_graphs=False,
                      representation_hidden_size=[h], actor_hidden_size=[h], critic_hidden_size=[h], running_steps=1000, **nep)
    cfg = agent.config
    assert agent.batch_size == n * T // n_epoch
    torch.manual_seed(9)
    np.random.seed(9)
    envs = ref_port.VecEnvPort(env_id, n, seed=seed, trig="cr")
    envs.reset()
    rep = policies.MLPRepresentation(envs.observation_space.shape, [h], activation=torch.nn.ReLU, device="cpu")
    cls = policies.CategoricalPPGActorCritic if env_id == "CartPole-v1" else policies.GaussianPPGActorCritic
    pol = cls(envs.action_space, rep, [h], [h], activation=torch.nn.ReLU, device="cpu")
    pol.load_state_dict({k: v.detach().cpu() for k, v in agent.policy.state_dict().items()}, strict=True)
    opt = torch.optim.Adam(pol.parameters(), cfg.learning_rate, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    act_shape = () if env_id == "CartPole-v1" else envs.action_space.shape
    mem = ref_port.OldDistBufferPort(envs.observation_space.shape, act_shape, n, T, True, True, cfg.gamma, cfg.gae_lambda)
    hp = dict(ent_coef=cfg.ent_coef, clip_range=cfg.clip_range, kl_beta=cfg.kl_beta)
    upd = {ph: (lambda o, a, r, ad, old, ph=ph: ref_port.ppg_update(ph, pol, opt, sched, (o, a, r, ad), old, **hp))
           for ph in ("policy", "critic", "aux")}
    port = ref_port.PPGAgentPort(envs, pol, mem, upd["policy"], upd["critic"], upd["aux"], n_steps=T, n_minibatch=n_epoch,
                                 use_obsnorm=True, use_rewnorm=True, record=True, **nep)
    ported = _record_rollouts(port, steps)
    assert len(ported) == 2
    agent.set_tapes(np.stack(port.recorded_actions), port.recorded_perms)
    native = _hook_update_snapshots(agent)
    info = agent.train(steps)
    assert set(info) >= {"actor-loss", "entropy", "learning_rate", "clip_ratio", "critic-loss", "kl-loss"}
    for k in ("actor-loss", "entropy", "critic-loss", "kl-loss"):
        assert np.isclose(float(info[k]), float(port.infos[k]), rtol=2e-3, atol=2e-5), (k, info[k], port.infos[k])
    assert agent._perm_pos == len(port.recorded_perms) == 2 * sum(nep.values())
    _compare(native, ported, "ppg")
    assert agent.learner.policy_iterations == 2 * nep["policy_nepoch"] * n_epoch
    ret = agent._ret_rms.cpu().numpy()
    assert ret[2] == 1e-4 and float(port.ret_rms.count) == 1e-4              # PPG_Agent never updates ret_rms
    assert agent.memory.ptr == mem.ptr == 3


@pytest.mark.parametrize("kind", ["pg", "ppg"])
def test_native_f3_agents_graph_rollout_equals_eager(kind):
    """The captured rollout graph (fused sample + env step + store + running statistics, one launch per vector step) against
    the same launches issued eagerly: identical buffers, statistics and parameters after three rollouts."""
    from xuanpolicy_b200.configs import build_pg, build_ppg
    out = []
    for graphs in (True, False):
        if kind == "pg":
            agent = build_pg("CartPole-v1", parallels=64, n_steps=32, n_epoch=2, seed=3, use_cuda_graphs=graphs, shuffle="device")
        else:
            agent = build_ppg("Pendulum-v1", parallels=64, n_steps=32, n_epoch=2, seed=3, use_cuda_graphs=graphs,
                              policy_nepoch=2, value_nepoch=1, aux_nepoch=1)
        assert agent._fused_step and agent._fused_norm
        info = agent.train(3 * 32)
        assert all(np.isfinite(float(v)) for k, v in info.items() if k not in ("mean_episode_score", "mean_episode_steps"))
        out.append((agent.memory._ret.clone(), agent.memory._obs.clone(), agent._obs_rms[0].clone(),
                    [p.detach().clone() for p in agent.policy.parameters()]))
    (ra, oa, sa, pa), (rb, ob, sb, pb) = out
    assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(sa, sb)
    for x, y in zip(pa, pb):
        assert torch.equal(x, y)
