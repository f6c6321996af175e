"""GPU: degenerate and ragged sizes through the C ABI — one env, one step, one sample, odd widths — and argument errors.
The reference handles these implicitly (numpy slices of length 1); the kernels must too."""
import numpy as np
import pytest
import torch

from tests.helpers import gae_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env_id", ["CartPole-v1", "Pendulum-v1", "gym:MountainCar-v0", "MountainCar-v0", "Acrobot-v1"])
def test_single_env_steps_bit_exact(env_id):
    import xuanpolicy_b200 as xb
    from oracle import c_oracle
    envs = xb.DummyVecEnv_Gym(xb.make_env_fns(env_id, 7, 1), device="cuda")
    ref = c_oracle.VecEnvC(env_id, 1, seed=7, flavour="cr")
    obs, infos = envs.reset()
    assert obs.shape == (1,) + envs.obs_shape and np.array_equal(obs, ref.obs)
    rng = np.random.default_rng(1)
    for _ in range(30):
        a = rng.standard_normal((1, 1)).astype(np.float32) if env_id == "Pendulum-v1" else rng.integers(0, envs.action_space.n, 1)
        obs, rew, term, trunc, infos = envs.step(a)
        o = ref.step(a)
        assert np.array_equal(obs, o["obs"]) and np.array_equal(rew, o["rew"]) and np.array_equal(envs.get_state(), o["state"])
        assert infos[0]["episode_step"] == int(o["ep_step"][0])


def test_one_step_one_env_buffer_gae_and_single_sample():
    """T = 1, N = 1: finish_path on a one-element path, then a minibatch of one sample (std of one advantage = 0:
    (adv - mean) / (0 + 1e-8) = 0, as numpy gives at memory_tools.py:241-242)."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import spaces
    obs_space = spaces.Box(-np.ones(4, np.float32), np.ones(4, np.float32))
    buf = xb.DummyOnPolicyBuffer(obs_space, spaces.Discrete(2), {"old_logp": ()}, 1, 1, True, True, 0.99, 0.95)
    buf.store(np.array([[0.1, 0.2, 0.3, 0.4]], np.float32), np.array([1]), np.array([2.0], np.float32),
              np.array([0.5], np.float32), np.array([False]), {"old_logp": np.array([-0.7], np.float32)})
    assert buf.full
    buf.finish_path(0.25, 0)
    # delta = r + gamma * boot - v
    assert abs(float(buf.advantages[0, 0]) - (2.0 + 0.99 * 0.25 - 0.5)) < 1e-6
    assert abs(float(buf.returns[0, 0]) - (2.0 + 0.99 * 0.25)) < 1e-6
    obs, act, ret, val, adv, aux = buf.sample(np.array([0]))
    assert obs.shape == (1, 4) and act.shape == (1,) and adv.shape == (1,) and float(adv[0]) == 0.0
    assert float(aux["old_logp"][0]) == np.float32(-0.7) and float(act[0]) == 1.0


@pytest.mark.parametrize("T,N", [(1, 1), (3, 5), (2, 127), (5, 129)])
def test_gae_tiny_and_odd_shapes(T, N):
    """Shapes below / off the TMA variant's preconditions (N % 4, N < 128): the auto variant must fall back and agree."""
    from oracle import c_oracle
    from xuanpolicy_b200 import ops
    rng = np.random.default_rng(T * 100 + N)
    rew, val = rng.standard_normal((T, N)).astype(np.float32), rng.standard_normal((T, N)).astype(np.float32)
    term = (rng.random((T, N)) < 0.2).astype(np.float32)
    boot_last = rng.standard_normal(N).astype(np.float32)
    d = lambda a: torch.from_numpy(a).cuda()
    adv, ret = torch.empty((T, N), device="cuda"), torch.empty((T, N), device="cuda")
    ops.gae(d(rew), d(val), d(term), d(boot_last), adv, ret, 0.98, 0.9, variant="auto")
    adv64, ret64 = c_oracle.gae(rew, val, term, boot_last, 0.98, 0.9)
    assert gae_close(adv.cpu().numpy(), adv64)[0] and gae_close(ret.cpu().numpy(), ret64)[0]


def test_minibatch_statistics_of_a_whole_epoch_in_one_pass():
    """xb_adv_stats_minibatches (the env-sharded path's once-per-epoch statistics) equals per-minibatch numpy sums,
    for plain [T, N] advantages and for the adv lane of the packed 32-byte records."""
    from xuanpolicy_b200 import ops
    T, N, M = 16, 96, 6
    B = T * N // M
    rng = np.random.default_rng(5)
    adv = rng.standard_normal((T, N)).astype(np.float32)
    perm = rng.permutation(T * N).astype(np.int64)
    env, step = np.divmod(perm, T)
    a64 = adv[step, env].astype(np.float64).reshape(M, B)
    want = np.stack([a64.sum(1), (a64 * a64).sum(1)], axis=1).reshape(-1)
    stats = torch.zeros(2 * M, dtype=torch.float64, device="cuda")
    ops.adv_stats_minibatches(torch.from_numpy(perm).cuda(), M, B, T, N, torch.from_numpy(adv).cuda(), 1, stats)
    assert np.allclose(stats.cpu().numpy(), want, rtol=1e-12, atol=1e-9)
    rec = torch.zeros((T * N, 8), device="cuda")
    rec[:, 6] = torch.from_numpy(adv).cuda().reshape(-1)
    ops.adv_stats_minibatches(torch.from_numpy(perm).cuda(), M, B, T, N, rec.view(-1)[6:], 8, stats)
    assert np.allclose(stats.cpu().numpy(), want, rtol=1e-12, atol=1e-9)


def test_bad_arguments_raise_instead_of_corrupting():
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import ops
    with pytest.raises(xb.XB200Error):      # CPU tensor
        ops.random_permutation(torch.zeros(4, dtype=torch.int64), 1)
    with pytest.raises(xb.XB200Error):      # wrong dtype
        ops.random_permutation(torch.zeros(4, dtype=torch.int32, device="cuda"), 1)
    with pytest.raises(xb.XB200Error):      # unsupported observation width
        ops.gather_obs(torch.zeros(4, dtype=torch.int64, device="cuda"), 2, 2, torch.zeros((2, 2, 12), device="cuda"), 9,
                       torch.zeros((4, 9), device="cuda"))
    envs = xb.DummyVecEnv_Gym(xb.make_env_fns("CartPole-v1", 1, 4), device="cuda")
    with pytest.raises(xb.NotSteppingError):
        envs.step_wait()
    envs.step_async(np.zeros(4, np.int64))
    with pytest.raises(xb.AlreadySteppingError):
        envs.step_async(np.zeros(4, np.int64))
    with pytest.raises(NotImplementedError):
        xb.DummyVecEnv_Gym(xb.make_env_fns("LunarLander-v2", 1, 4), device="cuda")
