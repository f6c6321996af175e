"""GPU parity: batched CartPole/Pendulum/MountainCar kernels vs the oracle (bit-exact, Tier 1) through the drop-in API."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _envs(env_id, n, seed=1, **kw):
    import xuanpolicy_b200 as xb
    return xb.DummyVecEnv_Gym(xb.make_env_fns(env_id, seed, n), **kw)


def test_sincos_correctly_rounded_on_device():
    """The kernel's double-double sin/cos equals the correctly-rounded value (libquadmath oracle), 2e6 points."""
    from oracle import c_oracle
    from xuanpolicy_b200 import ops
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-0.5, 0.5, 700_000), rng.uniform(-100, 100, 1_000_000),
                        rng.uniform(-4, 4, 300_000), [0.0, -0.0, 1e-300, np.pi / 2, np.pi, 83.14, -83.14]])
    s, c = ops.sincos_f64(torch.from_numpy(x).cuda())
    s_ref, c_ref = c_oracle.sincos(x, "cr")
    assert np.array_equal(s.cpu().numpy().view(np.uint64), s_ref.view(np.uint64))
    assert np.array_equal(c.cpu().numpy().view(np.uint64), c_ref.view(np.uint64))


@pytest.mark.parametrize("name", ["physics_cartpole_cr", "physics_pendulum_cr"])
def test_env_tape_bit_exact_vs_golden(name):
    """Tier 1: identical action tape -> identical fp64 states, f32 obs/rewards, flags, counters, reset draws."""
    g = load_golden(name)
    m = g["meta"]
    envs = _envs(m["env_id"], m["n"], m["seed"])
    obs0, infos0 = envs.reset()
    assert np.array_equal(obs0, g["obs0"][:, :obs0.shape[1]]) and obs0.dtype == np.float32
    assert np.array_equal(envs.get_state(), g["state0"])
    assert infos0[0] == {"episode_step": 0}
    for t in range(m["steps"]):
        a = g["actions"][t]
        obs, rew, term, trunc, infos = envs.step(a if m["env_id"] == "CartPole-v1" else a.reshape(-1, 1))
        assert np.array_equal(obs, g["obs"][t]), t
        assert np.array_equal(rew, g["rew"][t]) and rew.dtype == np.float32, t
        assert np.array_equal(term, g["term"][t]) and np.array_equal(trunc, g["trunc"][t]), t
        assert np.array_equal(envs.get_state(), g["state"][t]), t
        assert [i["episode_step"] for i in infos] == g["ep_step"][t].tolist()
        assert [i["episode_score"] for i in infos] == g["ep_score"][t].tolist()
        for i, inf in enumerate(infos):
            if term[i] or trunc[i]:
                assert np.array_equal(inf["reset_obs"], g["reset_obs"][t][i])
            else:
                assert "reset_obs" not in inf
    assert np.array_equal(envs.buf_obs, g["obs"][-1])


@pytest.mark.parametrize("env_id,n,steps", [("CartPole-v1", 4099, 520), ("Pendulum-v1", 4096, 410),
                                            ("gym:MountainCar-v0", 2051, 430), ("MountainCar-v0", 1033, 430),
                                            ("Acrobot-v1", 1027, 560)])
def test_env_large_batch_bit_exact_vs_c_oracle(env_id, n, steps):
    """Ragged batch size, per-env divergent random actions, many resets: still bit-exact with the C oracle."""
    from oracle import c_oracle
    envs = _envs(env_id, n, 3)
    ref = c_oracle.VecEnvC(env_id, n, seed=3, flavour="cr")
    obs, _ = envs.reset()
    assert np.array_equal(obs, ref.obs)
    rng = np.random.default_rng(9)
    n_done = 0
    for t in range(steps):
        if env_id == "CartPole-v1":
            heur = (obs[:, 2] + 0.5 * obs[:, 3] > 0).astype(np.int64)
            a = np.where(rng.random(n) < 0.8, heur, rng.integers(0, 2, n))
        elif "MountainCar" in env_id:           # energy-pumping heuristic so that some cars reach the goal (terminated)
            heur = np.where(obs[:, -1] > 0, 2, 0).astype(np.int64)      # velocity of the newest frame
            a = np.where(rng.random(n) < 0.9, heur, rng.integers(0, 3, n))
        elif env_id == "Acrobot-v1":            # torque along the second joint's velocity: swings up (terminated) in ~70 steps
            heur = np.where(obs[:, 5] > 0, 2, 0).astype(np.int64)
            a = np.where(rng.random(n) < 0.9, heur, rng.integers(0, 3, n))
        else:
            a = (1.5 * rng.standard_normal((n, 1))).astype(np.float32)
        obs, rew, term, trunc, infos = envs.step(a)
        o = ref.step(a)
        assert np.array_equal(obs, o["obs"]) and np.array_equal(rew, o["rew"]), t
        assert np.array_equal(term, o["term"]) and np.array_equal(trunc, o["trunc"]), t
        assert np.array_equal(envs.get_state(), o["state"]), t
        done = term | trunc
        n_done += int(done.sum())
        if done.any():
            ro = np.stack([infos[i]["reset_obs"] for i in np.nonzero(done)[0]])
            assert np.array_equal(ro, o["reset_obs"][done])
    assert n_done > 0
    st = envs.ep_stats.cpu().numpy()
    assert st[0] == n_done


def test_mountaincar_is_the_references_four_frame_wrapper():
    """`make_envs` wraps MountainCar-v0 in `MountainCar(Gym_Env)` (xuance/environment/gym/gym_env.py:50-83): observation
    space (8,), observation = the last four frames oldest first, all four = the reset observation after a reset.  Golden
    recorded from the reference's own make_envs / DummyVecEnv_Gym over the restated physics (flavour "cr"): bit-exact."""
    g = load_golden("vecenv_mountaincar_stack")
    m = g["meta"]
    envs = _envs("MountainCar-v0", m["n"], m["seed"])
    assert envs.observation_space.shape == (8,) and envs.action_space.n == m["n_actions"]
    assert np.array_equal(envs.observation_space.low, g["space_low"].astype(np.float32))
    assert np.array_equal(envs.observation_space.high, g["space_high"].astype(np.float32))
    assert envs.max_episode_length == m["max_episode_length"] and envs.buf_obs.shape == (m["n"], 8)
    obs0, _ = envs.reset()
    assert np.array_equal(obs0, g["obs0"]) and np.array_equal(obs0[:, 0:2], obs0[:, 6:8])
    for t in range(m["steps"]):
        obs, rew, term, trunc, infos = envs.step(g["actions"][t])
        assert np.array_equal(obs, g["obs"][t]) and np.array_equal(rew, g["rew"][t]), t
        assert np.array_equal(term, g["term"][t]) and np.array_equal(trunc, g["trunc"][t]), t
        assert [i["episode_step"] for i in infos] == g["ep_step"][t].tolist()
        assert [i["episode_score"] for i in infos] == g["ep_score"][t].tolist()
        for i, inf in enumerate(infos):
            if term[i] or trunc[i]:
                assert np.array_equal(inf["reset_obs"], g["reset_obs"][t][i])
            else:
                assert "reset_obs" not in inf
    assert g["term"].sum() > 0 and g["trunc"].sum() > 0


@pytest.mark.parametrize("name", ["vecenv_cartpole", "vecenv_pendulum"])
def test_env_tier2_vs_reference_run_with_host_libm(name):
    """Tier 2: against the reference DummyVecEnv_Gym run over libm physics, the API-visible
    float32 observations / rewards / flags are expected to match; mismatches are counted and must be zero here."""
    g = load_golden(name)
    m = g["meta"]
    envs = _envs(m["env_id"], m["n"], m["seed"])
    obs0, _ = envs.reset()
    assert np.array_equal(obs0, g["obs0"])
    assert envs.max_episode_length == m["max_episode_length"]
    mism = 0
    for t in range(m["steps"]):
        obs, rew, term, trunc, infos = envs.step(g["actions"][t])
        mism += int((obs != g["obs"][t]).sum()) + int((rew != g["rew"][t]).sum())
        assert np.array_equal(term, g["term"][t]) and np.array_equal(trunc, g["trunc"][t])
        assert [i["episode_step"] for i in infos] == g["ep_step"][t].tolist()
    assert mism == 0, "f32 mismatches vs libm physics: %d" % mism


def test_vec_env_api_errors_and_native_mode():
    import xuanpolicy_b200 as xb
    envs = _envs("CartPole-v1", 8)
    with pytest.raises(xb.NotSteppingError):
        envs.step_wait()
    envs.step_async(np.zeros(8, np.int64))
    with pytest.raises(xb.AlreadySteppingError):
        envs.step_async(np.zeros(8, np.int64))
    envs.step_wait()
    envs.close()
    assert envs.closed
    nat = _envs("Pendulum-v1", 16, native=True)
    obs, infos = nat.reset()
    assert obs.is_cuda and obs.shape == (16, 3)
    o, r, d, tr, infos = nat.step(torch.zeros(16, 1, device="cuda"))
    assert o.is_cuda and r.is_cuda and d.dtype == torch.bool and len(infos) == 16
    assert infos[0]["episode_step"] == 1
