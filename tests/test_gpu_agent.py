"""GPU: the device-resident PPO loop (rollout graph -> GAE -> fused updates) against the oracle, end to end."""
import numpy as np
import pytest
import torch

from tests.helpers import gae_close

pytestmark = pytest.mark.gpu


def _build(env_id, **kw):
    from xuanpolicy_b200.configs import build_ppo
    return build_ppo(env_id, **kw)


@pytest.mark.parametrize("env_id,n,T", [("CartPole-v1", 64, 40), ("Pendulum-v1", 48, 230), ("gym:MountainCar-v0", 40, 210),
                                       ("MountainCar-v0", 36, 210), ("Acrobot-v1", 24, 520)])
@pytest.mark.parametrize("graphs", [True, False])
def test_native_rollout_replays_bit_exact_through_the_oracle(env_id, n, T, graphs):
    """Take the actions the device loop drew, replay them through the C oracle from the same seed: the buffer's
    observations / rewards / terminals must be bit-identical, and its advantages / returns must equal the fp64
    oracle GAE fed with the stored values and the bootstrap values the loop computed (incl. truncation rows)."""
    from oracle import c_oracle
    agent = _build(env_id, parallels=n, n_steps=T, n_epoch=1, n_minibatch=2, use_cuda_graphs=graphs, shuffle="device", seed=5)
    mem, env = agent.memory, agent.envs
    ref = c_oracle.VecEnvC(env_id, n, seed=5, flavour="cr")
    for rollout in range(2):                      # second rollout: episodes continue across the rollout boundary
        with torch.cuda.device(agent.device):
            if graphs and agent._rollout_graph is None:
                agent._capture()
            agent._rollout_graph.replay() if graphs else agent._rollout()
        torch.cuda.synchronize()
        obs = mem._obs.cpu().numpy()
        act = mem._act.cpu().numpy()
        od = mem.obs_dim
        trunc_seen = 0
        for t in range(T):
            assert np.array_equal(obs[t, :, :od], ref.obs), (rollout, t)
            a = act[t, :, 0] if env_id == "Pendulum-v1" else act[t, :, 0].astype(np.int64)
            o = ref.step(a)
            assert np.array_equal(mem._rew[t].cpu().numpy(), o["rew"]), (rollout, t)
            assert np.array_equal(mem._term[t].cpu().numpy(), o["term"].astype(np.float32))
            assert np.array_equal(mem._trunc[t].cpu().numpy(), o["trunc"].astype(np.uint8))
            done = o["term"] | o["trunc"]
            trunc_seen += int(o["trunc"].sum())
            ref.obs[done] = o["reset_obs"][done]                      # obs[i] = infos[i]["reset_obs"] (:101)
        if env_id != "CartPole-v1":
            assert trunc_seen > 0
        # values stored are the critic's output on the stored observations
        with torch.no_grad():
            _, _, v = agent.policy(mem._obs.reshape(T * n, mem.obs_row)[:, :od])
        assert torch.allclose(v.reshape(T, n), mem._val, atol=1e-5, rtol=1e-5)
        adv64, ret64 = c_oracle.gae(mem._rew.cpu().numpy(), mem._val.cpu().numpy(), mem._term.cpu().numpy(),
                                    agent._boot_last.cpu().numpy(), agent.gamma, agent.gae_lam,
                                    segend=mem._trunc.cpu().numpy(), boot=mem._boot.cpu().numpy())
        assert gae_close(mem._adv.cpu().numpy(), adv64)[0] and gae_close(mem._ret.cpu().numpy(), ret64)[0]
        # old_logp stored == log-prob of the stored action under the rollout policy
        with torch.no_grad():
            _, dist, _ = agent.policy(mem._obs.reshape(T * n, mem.obs_row)[:, :od])
            a_t = mem._act.reshape(T * n, -1)
            lp = dist.log_prob(a_t if env_id == "Pendulum-v1" else a_t[:, 0])
        assert torch.allclose(lp.reshape(T, n), mem._logp, atol=2e-5, rtol=1e-5)


def test_graph_and_eager_training_agree():
    """The captured graphs replay exactly what the eager loop does (CartPole: no atomics on the gradient path)."""
    out = []
    for graphs in (True, False):
        agent = _build("CartPole-v1", parallels=32, n_steps=32, n_epoch=2, n_minibatch=4, use_cuda_graphs=graphs,
                       shuffle="device", seed=3)
        if graphs:
            with torch.cuda.device(agent.device):
                agent._capture()                  # the warm-up inside consumes CUDA generator state: do it before seeding
        torch.manual_seed(11)                     # device permutations come from torch's CUDA generator
        torch.cuda.manual_seed(11)
        info = agent.train(3 * 32)
        out.append((agent.learner._flat.flat_param.clone(), info))
        assert agent.learner.iterations == 3 * 2 * 4 and agent.current_step == 3 * 32 * 32
    assert torch.allclose(out[0][0], out[1][0], atol=2e-5, rtol=1e-4), (out[0][0] - out[1][0]).abs().max()
    assert int(agent.learner._flat.step.item()) == 24


@pytest.mark.parametrize("shuffle", ["host", "device"])
def test_cartpole_learns(shuffle):
    """Sanity that the whole path optimises: mean episode length grows well beyond the random-policy ~22 steps."""
    agent = _build("CartPole-v1", parallels=256, n_steps=64, n_epoch=4, n_minibatch=4, shuffle=shuffle, seed=1,
                   gamma=0.99, representation_hidden_size=[64], actor_hidden_size=[64], critic_hidden_size=[64])
    info0 = agent.train(64)
    st0 = agent.envs.ep_stats.cpu().numpy().copy()
    info = agent.train(64 * 30)
    st1 = agent.envs.ep_stats.cpu().numpy()
    first = st0[2] / max(st0[0], 1)
    late = (st1[2] - st0[2]) / max(st1[0] - st0[0], 1)
    assert np.isfinite(info["actor-loss"]) and np.isfinite(info["critic-loss"])
    assert first < 40 and late > 2.5 * first, (first, late)
    if shuffle == "host":
        assert agent.h2d_bytes == 31 * 4 * 256 * 64 * 4          # int32 permutations, one per epoch


@pytest.mark.parametrize("env_id,obsnorm", [("CartPole-v1", True), ("Pendulum-v1", False)])
def test_compat_dropins_under_the_reference_agent_loop(env_id, obsnorm):
    """the compat (numpy in/out) drop-ins driven by the reference's own agent loop — here its
    restatement oracle/ref_port.PPOAgentPort, which calls envs.step / memory.store / finish_path / sample /
    learner.update exactly like ppoclip_agent.py:59-111 (per-env finish_path calls, host-side obs/reward
    normalisation, list-of-dict infos)."""
    import xuanpolicy_b200 as xb
    from oracle import ref_port
    torch.manual_seed(0)
    np.random.seed(0)
    envs = xb.DummyVecEnv_Gym(xb.make_env_fns(env_id, 1, 12), device="cuda")
    envs.reset()
    policy = xb.make_policy(envs.observation_space, envs.action_space, hidden=(32,), device="cuda")
    opt = torch.optim.Adam(policy.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=10000)
    memory = xb.DummyOnPolicyBuffer(envs.observation_space, envs.action_space, {"old_logp": ()}, 12, 24, True, True, 0.98, 0.95)
    learner = xb.PPOCLIP_Learner(policy, opt, sched, "cuda", "/tmp", vf_coef=0.25, ent_coef=0.01, clip_range=0.2,
                                 clip_grad_norm=0.5, use_grad_clip=True)
    agent = ref_port.PPOAgentPort(envs, policy, opt, sched, 24, 2, 4, 0.98, 0.95, use_obsnorm=obsnorm, use_rewnorm=obsnorm,
                                  memory=memory, update_fn=learner.update)
    p0 = torch.cat([p.detach().reshape(-1).clone() for p in policy.parameters()])
    agent.train(24 * 3 + 5)
    p1 = torch.cat([p.detach().reshape(-1) for p in policy.parameters()])
    assert learner.iterations == 3 * 2 * 4 and agent.current_step == 12 * 77
    assert torch.isfinite(p1).all() and not torch.equal(p0, p1)
    assert set(agent.last_info) == {"actor-loss", "critic-loss", "entropy", "learning_rate", "predict_value", "clip_ratio"}
    assert memory.size == 5 and memory.ptr == 5 and not memory.full
    if env_id == "CartPole-v1":
        assert agent.episodes > 0


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Pendulum-v1"])
def test_ppg_dropins_under_the_reference_agent_loop(env_id):
    """Row f3 end to end: the reference's PPG_Agent.train (restated in oracle/ref_port.PPGAgentPort: rollout with stored
    old distributions, policy / critic phases, whole-buffer old-distribution refresh, auxiliary phase, ppg_agent.py:55-109)
    driving the drop-in env, buffer ({"old_dist": None} auxiliary on the device) and PPG_Learner."""
    import xuanpolicy_b200 as xb
    from oracle import ref_port
    from xuanpolicy_b200 import policies
    torch.manual_seed(0)
    np.random.seed(0)
    n, T = 10, 16
    envs = xb.DummyVecEnv_Gym(xb.make_env_fns(env_id, 1, n), device="cuda")
    envs.reset()
    rep = policies.MLPRepresentation(envs.observation_space.shape, [32], device="cuda")
    cls = policies.CategoricalPPGActorCritic if env_id == "CartPole-v1" else policies.GaussianPPGActorCritic
    policy = cls(envs.action_space, rep, [32], [32], device="cuda").cuda()
    opt = torch.optim.Adam(policy.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=10000)
    memory = xb.DummyOnPolicyBuffer(envs.observation_space, envs.action_space, {"old_dist": None}, n, T, True, True, 0.98, 0.95)
    learner = xb.PPG_Learner(policy, opt, sched, "cuda", "/tmp", ent_coef=0.01, clip_range=0.2, kl_beta=1.0)
    agent = ref_port.PPGAgentPort(envs, policy, memory, learner.update_policy, learner.update_critic, learner.update_auxiliary,
                                  n_steps=T, n_minibatch=4, policy_nepoch=2, value_nepoch=1, aux_nepoch=1)
    p0 = {k: v.detach().clone() for k, v in policy.named_parameters()}
    agent.train(2 * T + 3)
    assert learner.policy_iterations == 2 * 2 * 4 and learner.value_iterations == 2 * 1 * 4
    assert set(agent.infos) == {"actor-loss", "entropy", "learning_rate", "clip_ratio", "critic-loss", "kl-loss"}
    assert all(np.isfinite(float(v)) for v in agent.infos.values())
    for k, v in policy.named_parameters():
        assert torch.isfinite(v).all() and not torch.equal(v, p0[k]), k      # every sub-network was trained by some phase
    assert memory.size == 3 and not memory.full
    assert agent.current_step == n * (2 * T + 3)


@pytest.mark.parametrize("env_id,n,T", [("Pendulum-v1", 256, 16), ("CartPole-v1", 48, 16)])
def test_native_a2c_agent_uses_the_a2c_surrogate(env_id, n, T):
    """Row f3: A2C_Agent (a2c_agent.py:57-100) on the device-resident loop.  With one epoch of one minibatch the policy at
    update time is the rollout policy, so the logged actor loss must equal -(adv_normalised * stored log-prob).mean()
    (a2c_learner.py:28) computed from the buffer — at 4096 samples through the tensor-core MLP with the loss in its epilogue,
    at 768 through the torch MLP and the stand-alone loss kernel."""
    import xuanpolicy_b200 as xb
    agent = _build(env_id, parallels=n, n_steps=T, n_epoch=1, n_minibatch=1, shuffle="device", seed=3, agent_class=xb.A2C_Agent,
                   clip_grad=0.5)
    info = agent.train(T)
    assert "clip_ratio" not in info and agent.learner.clip_range == 0.0
    mem = agent.memory
    adv = mem._adv.double().reshape(-1)
    adv_n = (adv.float() - adv.mean().float()) / (adv.std(unbiased=False).float() + 1e-8)
    expect = float(-(adv_n.double() * mem._logp.double().reshape(-1)).mean())
    assert abs(info["actor-loss"] - expect) <= 1e-4 * max(1.0, abs(expect)), (info["actor-loss"], expect)
    expect_c = float(((mem._val - mem._ret).double() ** 2).mean())
    assert abs(info["critic-loss"] - expect_c) <= 1e-4 * max(1.0, abs(expect_c))
    info2 = agent.train(2 * T)
    assert np.isfinite(info2["actor-loss"]) and int(agent.learner._flat.step.item()) == 3


@pytest.mark.parametrize("env", [{"XB_ADAM_SPLIT": "0"}, {"XB_TAIL_NORM": "0"}, {}])
def test_graph_replay_resplits_weights_on_every_optimizer_path(env, monkeypatch):
    """the first minibatch of every captured epoch must see tf32 operand copies that match the weights, also
    when the optimiser launch does not rewrite them (XB_ADAM_SPLIT=0, XB_TAIL_NORM=0 = the path the NCCL fallback uses).
    Graph replay over several epochs and rollouts must equal the eager loop on the same path (tensor-core MLP active:
    minibatch 2048 rows)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    out = []
    for graphs in (True, False):
        agent = _build("CartPole-v1", parallels=64, n_steps=64, n_epoch=3, n_minibatch=2, use_cuda_graphs=graphs,
                       shuffle="device", seed=3)
        assert agent.learner._fused is not None and agent.batch_size >= agent.learner._fused.MIN_ROWS
        assert agent.learner.adam_resplits() == (not env)
        agent.train(2 * 64)
        out.append(agent.learner._flat.flat_param.clone())
    assert torch.allclose(out[0], out[1], atol=2e-5, rtol=1e-4), (out[0] - out[1]).abs().max()


def test_train_accepts_any_step_count_and_ragged_minibatches():
    """ppoclip_agent.py:59-61,79-83: `train(k)` for any k (the buffer position carries across calls) and a short last
    minibatch when n_envs * n_steps is not a multiple of n_minibatch."""
    agent = _build("CartPole-v1", parallels=30, n_steps=20, n_epoch=2, n_minibatch=7, shuffle="device", seed=2)
    assert agent.batch_size == 85 and agent.n_updates_per_epoch == 8           # 7 x 85 + one minibatch of 5
    agent.train(7)
    assert agent.memory.ptr == 7 and agent.learner.iterations == 0 and agent.current_step == 0
    agent.train(13 + 20 + 4)                                                   # finishes rollout 1 (eager), rollout 2 (graph), 4 into rollout 3
    assert agent.learner.iterations == 2 * 2 * 8 and agent.memory.ptr == 4
    assert int(agent.learner._flat.step.item()) == 32
    info = agent.train(16)
    assert agent.learner.iterations == 3 * 2 * 8 and agent.memory.ptr == 0 and agent.current_step == 3 * 30 * 20
    assert np.isfinite(info["actor-loss"]) and np.isfinite(info["critic-loss"])
    # the partial-rollout path is the graph's launches issued eagerly: same rollout either way
    a = _build("Pendulum-v1", parallels=16, n_steps=12, n_epoch=1, n_minibatch=2, shuffle="device", seed=4)
    b = _build("Pendulum-v1", parallels=16, n_steps=12, n_epoch=1, n_minibatch=2, shuffle="device", seed=4)
    a.train(12)
    b.train(5)
    b.train(7)
    assert torch.equal(a.envs._state, b.envs._state) and torch.equal(a.memory._adv, b.memory._adv)
    assert torch.allclose(a.learner._flat.flat_param, b.learner._flat.flat_param, atol=1e-6)


def test_learner_save_and_load_model(tmp_path):
    """Learner.save_model / load_model (xuance/torch/learners/learner.py:24-45): state_dict round trip through the
    reference's directory convention — `path/<...seed_{seed}...>/<sorted model files>`, newest (last sorted) wins,
    `obs_rms.npy` is skipped — and the reference's own state_dict keys."""
    import xuanpolicy_b200 as xb
    agent = _build("Pendulum-v1", parallels=64, n_steps=32, n_epoch=1, n_minibatch=1, shuffle="device", seed=2)
    lr = agent.learner
    run_dir = tmp_path / "seed_2_SatOct18" 
    run_dir.mkdir()
    (tmp_path / "seed_9_other").mkdir()
    lr.save_model(str(run_dir / "model_000.pth"))
    first = {k: v.clone() for k, v in agent.policy.state_dict().items()}
    agent.train(32)                                            # parameters move (flat-buffer Adam on the native path)
    lr.save_model(str(run_dir / "model_001.pth"))
    second = {k: v.clone() for k, v in agent.policy.state_dict().items()}
    assert any(not torch.equal(first[k], second[k]) for k in first)
    np.save(str(run_dir / "obs_rms.npy"), np.zeros(3))
    saved = torch.load(str(run_dir / "model_001.pth"))
    assert list(saved.keys()) == list(first.keys()) and "actor.logstd" in saved       # gaussian.py:25
    # a fresh policy restores the newest file of the seed's directory
    other = _build("Pendulum-v1", parallels=64, n_steps=32, n_epoch=1, n_minibatch=1, shuffle="device", seed=5, policy_seed=77)
    assert any(not torch.equal(second[k], v) for k, v in other.policy.state_dict().items())
    other.learner.load_model(str(tmp_path), seed=2)
    for k, v in other.policy.state_dict().items():
        assert torch.equal(v, second[k]), k
    if other.learner._fused is not None:
        assert not other.learner._fused.splits_fresh           # the tf32 operand copies must be rebuilt before the next forward
    info = other.train(32)                                     # and training continues from the loaded weights
    assert np.isfinite(info["critic-loss"])
    # the flat parameter buffer still backs the modules after load_state_dict (in-place copy)
    assert other.policy.actor.logstd.data_ptr() >= other.learner._flat.flat_param.data_ptr()


@pytest.mark.parametrize("env_id,native", [("CartPole-v1", True), ("CartPole-v1", False), ("Pendulum-v1", True)])
def test_agent_test_runs_evaluation_episodes(env_id, native):
    """PPOCLIP_Agent.test(env_fn, test_episode) (ppoclip_agent.py:113-165): fresh envs from `env_fn`, stochastic actions,
    returns the scores of the finished episodes (what Runner_DRL.run / benchmark consume)."""
    import xuanpolicy_b200 as xb
    agent = _build(env_id, parallels=16, n_steps=16, n_epoch=1, n_minibatch=1, shuffle="device", seed=2)
    made = []

    def env_fn():
        e = xb.DummyVecEnv_Gym(xb.make_env_fns(env_id, 3, 4), device="cuda", native=native)
        made.append(e)
        return e
    scores = agent.test(env_fn, 5)
    assert len(scores) >= 5 and all(np.isfinite(s) for s in scores) and made[0].closed
    if env_id == "CartPole-v1":
        assert all(8 <= s <= 500 for s in scores)              # one reward per step, random-ish policy
    else:
        assert all(-2000 < s < 0 for s in scores)              # 200 steps of negative cost
