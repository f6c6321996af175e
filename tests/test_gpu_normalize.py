"""GPU parity (row f1): device RunningMeanStd + observation / reward normalisation vs the oracle's restatement of
RunningMeanStd (statistic_tools.py:35-112) and Agent._process_observation/_process_reward (agent.py:104-123)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,N", [(3, 1000), (4, 4096), (2, 37)])
def test_obs_rms_update_and_normalize_vs_oracle(dim, N):
    from oracle import ref_port
    from xuanpolicy_b200 import ops
    rng = np.random.default_rng(dim)
    rms = ref_port.RunningMeanStdPort((dim,))
    f64 = dict(dtype=torch.float64, device="cuda")
    state = [torch.tensor([0, 0, 0, 0, 1, 1, 1, 1, 1e-4], **f64) for _ in range(2)]
    sums, ws = torch.zeros(9, **f64), torch.zeros(8 + 8 * 1184, **f64)
    out = torch.zeros((2 * N, 4), device="cuda")
    cur = 0
    prev_mean, prev_std = rms.mean.copy(), rms.std.copy()
    for it in range(6):
        x = (rng.standard_normal((2 * N, dim)) * (1 + it) + 3 * it).astype(np.float32)
        x4 = np.zeros((2 * N, 4), np.float32)
        x4[:, :dim] = x
        xd = torch.from_numpy(x4).cuda()
        ops.moments4(xd[:N], sums, ws)
        ops.rms_normalize(xd, dim, sums, state[cur], state[cur ^ 1], 5.0, out, N)
        cur ^= 1
        rms.update(x[:N])
        new = np.clip((x[:N] - rms.mean) / (rms.std + 1e-8), -5, 5)
        old = np.clip((x[N:] - prev_mean) / (prev_std + 1e-8), -5, 5)
        got = out.cpu().numpy()
        assert np.allclose(got[:N, :dim], new, rtol=2e-5, atol=2e-5), np.abs(got[:N, :dim] - new).max()
        assert np.allclose(got[N:, :dim], old, rtol=2e-5, atol=2e-5)
        assert np.all(got[:, dim:] == 0)
        st = state[cur].cpu().numpy()
        assert np.allclose(st[:dim], rms.mean, rtol=1e-5, atol=1e-6) and np.allclose(st[4:4 + dim], rms.var, rtol=1e-4)
        assert abs(st[8] - rms.count) < 1e-9 * rms.count
        prev_mean, prev_std = rms.mean.copy(), rms.std.copy()
    # n_merged_rows == 0: normalise only, state passes through unchanged
    ops.rms_normalize(xd, dim, None, state[cur], state[cur ^ 1], 5.0, out, 0)
    assert torch.equal(state[cur], state[cur ^ 1])


def test_returns_tracker_and_reward_normaliser_vs_oracle():
    from oracle import ref_port
    from xuanpolicy_b200 import ops
    N, gamma = 500, 0.98
    rng = np.random.default_rng(1)
    ret_rms = ref_port.RunningMeanStdPort(())
    returns = np.zeros(N, np.float32)
    f64 = dict(dtype=torch.float64, device="cuda")
    d_ret, d_state = torch.zeros(N, **f64), torch.tensor([0.0, 1.0, 1e-4], **f64)
    d_sums, d_ws, d_std = torch.zeros(3, **f64), torch.zeros(8 + 8 * 1184, **f64), torch.ones(1, device="cuda")
    for t in range(40):
        rew = (rng.standard_normal(N) * 3 - 1).astype(np.float32)
        term = rng.random(N) < 0.03
        trunc = rng.random(N) < 0.03
        # reference order: rewards are processed with the std BEFORE this step's episode-end updates (:68 then :87-92)
        std = np.clip(ret_rms.std, 0.1, 100)
        assert abs(float(d_std.item()) - float(std)) <= 1e-5 * float(std)
        returns = (1 - term) * gamma * returns + rew
        for i in range(N):
            if term[i] or trunc[i]:
                ret_rms.update(returns[i:i + 1])
                returns[i] = 0.0
        ops.returns_track(d_ret, torch.from_numpy(rew).cuda(), torch.from_numpy(term.astype(np.uint8)).cuda(),
                          torch.from_numpy(trunc.astype(np.uint8)).cuda(), gamma, d_sums, d_ws)
        ops.rms_merge_scalar(d_sums, d_state, d_std)
        assert returns.dtype == np.float64            # the reference's tracker is float64 after the first step
        assert np.allclose(d_ret.cpu().numpy(), returns, rtol=1e-12, atol=1e-12)
        st = d_state.cpu().numpy()
        assert abs(st[0] - ret_rms.mean) <= 1e-6 * max(1, abs(ret_rms.mean)) and abs(st[1] - ret_rms.var) <= 1e-5 * ret_rms.var
        assert abs(st[2] - ret_rms.count) < 1e-6


def test_native_rollout_with_obs_and_reward_normalisation():
    """The device loop with use_obsnorm/use_rewnorm: buffer observations and rewards equal the oracle's processed
    values when the same action tape is replayed through the C oracle env + the RunningMeanStd restatement."""
    from oracle import c_oracle, ref_port
    from xuanpolicy_b200.configs import build_ppo
    n, T, env_id = 64, 230, "Pendulum-v1"
    agent = build_ppo(env_id, parallels=n, n_steps=T, n_epoch=1, n_minibatch=2, shuffle="device", seed=5,
                      use_obsnorm=True, use_rewnorm=True)
    mem = agent.memory
    ref = c_oracle.VecEnvC(env_id, n, seed=5, flavour="cr")
    obs_rms, ret_rms = ref_port.RunningMeanStdPort((3,)), ref_port.RunningMeanStdPort(())
    returns = np.zeros(n, np.float32)
    for rollout in range(2):
        with torch.cuda.device(agent.device):
            if agent._rollout_graph is None:
                agent._capture()
            agent._rollout_graph.replay()
        torch.cuda.synchronize()
        obs, act, rew = mem._obs.cpu().numpy(), mem._act.cpu().numpy(), mem._rew.cpu().numpy()
        for t in range(T):
            raw = ref.obs.copy()
            obs_rms.update(raw)
            proc = np.clip((raw - obs_rms.mean) / (obs_rms.std + 1e-8), -5, 5)
            # the reference accumulates the batch mean sequentially in float32 (error ~1e-6 in observation units;
            # the kernel accumulates in fp64), and while all same-seed envs still share a state std is ~1e-3, so the
            # tolerance is expressed in observation units and divided by std
            tol = 1e-4 + 4e-6 / (obs_rms.std + 1e-8)
            assert np.all(np.abs(obs[t, :, :3] - proc) <= tol), (rollout, t, np.abs(obs[t, :, :3] - proc).max())
            o = ref.step(act[t, :, 0])
            std = np.clip(ret_rms.std, 0.1, 100)
            assert np.allclose(rew[t], np.clip(o["rew"] / std, -5, 5), rtol=1e-4, atol=1e-5), (rollout, t)
            returns = (1 - o["term"]) * agent.gamma * returns + o["rew"]
            done = o["term"] | o["trunc"]
            for i in np.nonzero(done)[0]:
                ret_rms.update(returns[i:i + 1])
                returns[i] = 0.0
            ref.obs[done] = o["reset_obs"][done]
    st = agent._obs_rms[0].cpu().numpy()
    if agent._fused_norm:      # the fused step already merged the observations the NEXT step will act on
        obs_rms.update(ref.obs)
    assert np.allclose(st[:3], obs_rms.mean, rtol=1e-4, atol=2e-5) and abs(st[8] - obs_rms.count) < 1e-6 * obs_rms.count
    info = agent.train(T)
    assert np.isfinite(info["critic-loss"])


@pytest.mark.parametrize("env_id,n", [("CartPole-v1", 96), ("Pendulum-v1", 1200), ("MountainCar-v0", 64), ("Acrobot-v1", 40)])
def test_fused_statistics_rollout_vs_oracle_normalisers(env_id, n):
    """The statistics carried by the fused rollout step (csrc/normalize.cuh: obs moments merged by the step that produces the
    observations, return tracker + return normaliser, reward divisor) and the normalisation done inside the rollout forward
    (Pendulum at 2400 rows: the one-launch tcgen05 forward; the others: xb_rms_apply + torch MLP; MountainCar / Acrobot:
    8-float rows) against oracle/ref_port.RunningMeanStdPort fed with the raw observations / finished returns of the same
    run, and against agent.py:104-123 applied with the oracle's statistics of each step."""
    from oracle import ref_port
    from xuanpolicy_b200.configs import build_ppo
    T = 24 if env_id != "Pendulum-v1" else 210
    agent = build_ppo(env_id, parallels=n, n_steps=T, n_epoch=1, n_minibatch=2, use_obsnorm=True, use_rewnorm=True,
                      use_cuda_graphs=False, shuffle="device", seed=3, gamma=0.98)
    assert agent._fused_norm and agent._fused_step
    od = agent._obs_dim
    raw, rews, terms, truncs = [], [], [], []
    orig = agent._rollout_step

    def step(t):
        raw.append(agent._x[agent._cur][:n, :od].cpu().numpy().copy())
        orig(t)
        rews.append(agent.envs._rew.cpu().numpy().copy())
        terms.append(agent.envs._term.cpu().numpy().astype(bool))
        truncs.append(agent.envs._trunc.cpu().numpy().astype(bool))
    agent._rollout_step = step
    obs_rms, ret_rms = ref_port.RunningMeanStdPort((od,)), ref_port.RunningMeanStdPort(())
    returns = np.zeros(n, np.float32)
    snaps = []
    orig_update = agent._update_phase
    agent._update_phase = lambda: (snaps.append((agent.memory._obs.cpu().numpy().copy(), agent.memory._rew.cpu().numpy().copy())),
                                   orig_update())
    agent.train(2 * T)
    assert len(raw) == 2 * T and len(snaps) == 2
    for t in range(2 * T):
        obs_rms.update(raw[t])                                                         # ppoclip_agent.py:62
        want = np.clip((raw[t] - obs_rms.mean) / (obs_rms.std + 1e-8), -5, 5)          # agent.py:112-113
        got = snaps[t // T][0][t % T][:, :od]
        # a float32 rounding of the mean is amplified by 1 / (std + 1e-8): negligible except while every env still holds
        # (nearly) the same observation — all envs are seeded alike — and std ~ 0
        tol = 2e-5 + 4e-6 / (obs_rms.std + 1e-8)
        assert np.all(np.abs(got - want) <= tol + 2e-5 * np.abs(want)), (t, np.abs(got - want).max())
        std = np.clip(ret_rms.std, 0.1, 100)                                           # agent.py:119-120, before this step's update
        want_r = np.clip(rews[t] / std, -5, 5)
        assert np.allclose(snaps[t // T][1][t % T], want_r, rtol=2e-5, atol=1e-6), t
        returns = (1 - terms[t]) * 0.98 * returns + rews[t]                            # ppoclip_agent.py:87
        for i in np.nonzero(terms[t] | truncs[t])[0]:
            ret_rms.update(returns[i:i + 1])
            returns[i] = 0.0
    st = agent._obs_rms[agent._rms_cur].cpu().numpy()
    D = (st.size - 1) // 2
    # the device state already holds the moments of the observations the NEXT step will act on
    obs_rms.update(agent._x[agent._cur][:n, :od].cpu().numpy())
    assert np.allclose(st[:od], obs_rms.mean, rtol=1e-5, atol=1e-6) and np.allclose(st[D:D + od], obs_rms.var, rtol=1e-5, atol=1e-7)
    assert np.isclose(st[2 * D], obs_rms.count, rtol=1e-12)
    rs = agent._ret_rms.cpu().numpy()
    if ret_rms.count > 1:
        assert np.isclose(rs[0], float(ret_rms.mean), rtol=1e-6) and np.isclose(rs[1], float(ret_rms.var), rtol=1e-5, atol=1e-9)
        assert np.isclose(rs[2], ret_rms.count, rtol=1e-12)
    assert np.allclose(agent._returns.cpu().numpy(), returns, rtol=1e-6, atol=1e-6)


def test_fused_statistics_path_equals_the_separate_launches(monkeypatch):
    """XB_FUSED_NORM=0 keeps the four separate launches per step (the path the env-sharded agent uses): same trajectory
    (CartPole: discrete actions), normalised observations / rewards and statistics within float32 rounding."""
    from xuanpolicy_b200.configs import build_ppo
    out = []
    for fused in ("1", "0"):
        monkeypatch.setenv("XB_FUSED_NORM", fused)
        agent = build_ppo("CartPole-v1", parallels=128, n_steps=32, n_epoch=1, n_minibatch=2, use_obsnorm=True, use_rewnorm=True,
                          shuffle="device", seed=5)
        assert agent._fused_norm == (fused == "1")
        with torch.cuda.device(agent.device):
            agent._capture()
            agent._rollout_graph.replay()
            agent._rollout_graph.replay()
        torch.cuda.synchronize()
        mem = agent.memory
        # (the observation normaliser's state itself is not comparable: on the fused path it already holds the moments of
        # the observations the next step will act on; the normalised observations below prove the statistics agree)
        out.append([t.clone() for t in (mem._act, mem._term, mem._obs, mem._rew, mem._val, agent._ret_rms, agent._returns,
                                        agent.envs._state)])
    a, b = out
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[7], b[7])
    for x, y in zip(a[2:7], b[2:7]):
        assert torch.allclose(x.double(), y.double(), rtol=1e-5, atol=1e-6), (x.double() - y.double()).abs().max()
