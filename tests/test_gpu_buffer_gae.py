"""GPU parity: device rollout buffer, batched GAE scan (LDG and TMA variants) and minibatch gather."""
import numpy as np
import pytest
import torch

from tests.helpers import gae_close, load_golden, replay_buffer_protocol

pytestmark = pytest.mark.gpu


def _spaces(obs_dim, discrete):
    from xuanpolicy_b200 import spaces
    obs = spaces.Box(-np.ones(obs_dim, np.float32), np.ones(obs_dim, np.float32))
    return obs, (spaces.Discrete(2) if discrete else spaces.Box(-2.0, 2.0, shape=(1,)))


@pytest.mark.parametrize("name", ["buffer_cat_gae", "buffer_box_gae_noadvnorm", "buffer_cat_nogae"])
def test_buffer_dropin_matches_reference_golden(name):
    """store / finish_path / sample through the reference's call protocol, against the reference's own outputs."""
    import xuanpolicy_b200 as xb
    g = load_golden(name)
    m = g["meta"]
    obs_space, act_space = _spaces(g["obs"].shape[2], m["discrete"])
    buf = xb.DummyOnPolicyBuffer(obs_space, act_space, {"old_logp": ()}, m["n_envs"], m["n_size"], m["use_gae"],
                                 m["use_advnorm"], m["gamma"], m["lam"])
    replay_buffer_protocol(buf, g)
    assert buf.full and buf.ptr == 0
    assert np.array_equal(buf.observations, g["observations"]) and np.array_equal(buf.actions, g["actions"])
    assert np.array_equal(buf.rewards, g["rew"].T) and np.array_equal(buf.values, g["val"].T)
    assert np.array_equal(buf.terminals, g["term"].T.astype(np.float32))
    ok, err = gae_close(buf.advantages, g["advantages"])   # 1e-5 relative (north_star)
    assert ok, err
    ok, err = gae_close(buf.returns, g["returns"])
    assert ok, err
    for k in range(g["idx"].shape[0]):
        o, a, r, v, adv, aux = buf.sample(g["idx"][k])
        assert np.array_equal(o, g["s_obs"][k]) and np.array_equal(a, g["s_act"][k])          # gathered rows: bit-exact
        assert np.array_equal(v, g["s_val"][k]) and np.array_equal(aux["old_logp"], g["s_logp"][k])
        assert gae_close(r, g["s_ret"][k])[0]
        assert np.allclose(adv, g["s_adv"][k], rtol=0, atol=2e-5), np.abs(adv - g["s_adv"][k]).max()
    buf.clear()
    assert buf.size == 0 and not buf.full and float(np.abs(buf.rewards).max()) == 0.0
    with pytest.raises(AssertionError):
        buf.sample(np.arange(4))


def _random_rollout(T, N, seed, p_term=0.01, p_trunc=0.01):
    rng = np.random.default_rng(seed)
    rew = rng.standard_normal((T, N)).astype(np.float32)
    val = rng.standard_normal((T, N)).astype(np.float32)
    term = (rng.random((T, N)) < p_term).astype(np.float32)
    trunc = (rng.random((T, N)) < p_trunc).astype(np.uint8)
    boot = rng.standard_normal((T, N)).astype(np.float32)
    boot_last = rng.standard_normal(N).astype(np.float32)
    return rew, val, term, trunc, boot, boot_last


@pytest.mark.parametrize("variant", ["ldg", "tma"])
@pytest.mark.parametrize("T,N,use_gae,with_trunc", [(1, 128, True, True), (7, 132, True, True), (257, 1000, True, True),
                                                     (128, 4096, True, False), (64, 640, False, True), (2048, 512, True, True)])
def test_gae_kernel_vs_fp64_oracle(variant, T, N, use_gae, with_trunc):
    from oracle import c_oracle
    from xuanpolicy_b200 import ops
    rew, val, term, trunc, boot, boot_last = _random_rollout(T, N, 17 + T)
    d = lambda a: torch.from_numpy(a).cuda()
    adv, ret = torch.empty((T, N), device="cuda"), torch.empty((T, N), device="cuda")
    stats = torch.zeros(2, dtype=torch.float64, device="cuda")
    ops.gae(d(rew), d(val), d(term), d(boot_last), adv, ret, 0.99, 0.95, trunc=d(trunc) if with_trunc else None,
            boot=d(boot) if with_trunc else None, stats=stats, use_gae=use_gae, variant=variant)
    adv64, ret64 = c_oracle.gae(rew, val, term, boot_last, 0.99, 0.95, segend=trunc if with_trunc else None,
                                boot=boot if with_trunc else None, use_gae=use_gae)
    a, r = adv.cpu().numpy(), ret.cpu().numpy()
    # carried in fp64 and rounded once: equal to the rounded fp64 oracle up to 1 ulp; tolerance of record is 1e-5
    ok, err = gae_close(a, adv64)
    assert ok, err
    ok, err = gae_close(r, ret64)
    assert ok, err
    assert np.abs(a - adv64.astype(np.float32)).max() <= 4e-6 * max(1.0, np.abs(adv64).max())
    s = stats.cpu().numpy()
    assert abs(s[0] - a.astype(np.float64).sum()) <= 1e-6 * max(1.0, abs(s[0]))
    assert abs(s[1] - (a.astype(np.float64) ** 2).sum()) <= 1e-9 * s[1]


def test_gae_full_size_slice_and_variant_agreement():
    """C4-shaped slice (T=2048) at N=65536: TMA and LDG variants agree bit-for-bit; 256 envs checked vs the oracle;
    size-independent property: scaling rewards, values and bootstraps by 2 scales adv/ret by exactly 2."""
    from oracle import c_oracle
    from xuanpolicy_b200 import ops
    T, N = 2048, 65536
    gen = torch.Generator(device="cuda").manual_seed(1234)
    rew = torch.randn((T, N), device="cuda", generator=gen)
    val = torch.randn((T, N), device="cuda", generator=gen)
    term = (torch.rand((T, N), device="cuda", generator=gen) < 1 / 200).float()
    boot_last = torch.randn(N, device="cuda", generator=gen)
    out = {}
    for variant in ("ldg", "tma"):
        adv, ret = torch.empty_like(rew), torch.empty_like(rew)
        ops.gae(rew, val, term, boot_last, adv, ret, 0.99, 0.95, variant=variant)
        out[variant] = (adv, ret)
    assert torch.equal(out["ldg"][0], out["tma"][0]) and torch.equal(out["ldg"][1], out["tma"][1])
    sl = slice(1000, 1256)
    adv64, ret64 = c_oracle.gae(rew[:, sl].cpu().numpy(), val[:, sl].cpu().numpy(), term[:, sl].cpu().numpy(),
                                boot_last[sl].cpu().numpy(), 0.99, 0.95)
    assert gae_close(out["tma"][0][:, sl].cpu().numpy(), adv64)[0]
    assert gae_close(out["tma"][1][:, sl].cpu().numpy(), ret64)[0]
    adv2, ret2 = torch.empty_like(rew), torch.empty_like(rew)
    ops.gae(rew * 2, val * 2, term, boot_last * 2, adv2, ret2, 0.99, 0.95, variant="tma")
    assert torch.equal(adv2, out["tma"][0] * 2) and torch.equal(ret2, out["tma"][1] * 2)


def test_gather_indices_bit_exact_and_stats():
    from xuanpolicy_b200 import ops
    T, N, B = 64, 300, 4096
    rng = np.random.default_rng(2)
    obs = rng.standard_normal((T, N, 4)).astype(np.float32)
    adv = rng.standard_normal((T, N)).astype(np.float32)
    idx = rng.permutation(T * N)[:B].astype(np.int64)
    env, step = np.divmod(idx, T)                     # memory_tools.py:234
    for obs_dim in (3, 4):
        out = torch.empty((B, obs_dim), device="cuda")
        stats = torch.zeros(2, dtype=torch.float64, device="cuda")
        ops.gather_obs(torch.from_numpy(idx).cuda(), T, N, torch.from_numpy(obs).cuda(), obs_dim, out,
                       b_adv=torch.from_numpy(adv).cuda(), stats=stats)
        assert np.array_equal(out.cpu().numpy(), obs[step, env, :obs_dim])
        a = adv[step, env].astype(np.float64)
        s = stats.cpu().numpy()
        assert abs(s[0] - a.sum()) < 1e-9 * B and abs(s[1] - (a * a).sum()) < 1e-9 * B


@pytest.mark.parametrize("name", ["buffer_cat_olddist", "buffer_box_olddist"])
@pytest.mark.parametrize("per_sample_objects", [False, True])
def test_buffer_old_dist_auxiliary_matches_reference_golden(name, per_sample_objects):
    """Row f3 (PPO-KL / PPG): the {"old_dist": None} auxiliary.  The reference stores numpy arrays of Python distribution
    objects (memory_tools.py:28-30); the drop-in keeps their parameters on the device.  store -> sample, then the
    whole-buffer reassignment of PPG_Agent.train (ppg_agent.py:90-93) -> sample: gathered parameters bit-exact."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import policies, spaces
    g = load_golden(name)
    m = g["meta"]
    T, N, A = m["n_size"], m["n_envs"], m["A"]
    if m["discrete"]:
        obs_space, act_space = spaces.Box(-1, 1, (4,)), spaces.Discrete(A)
    else:
        obs_space, act_space = spaces.Box(-1, 1, (3,)), spaces.Box(-2.0, 2.0, (A,))

    def wrap(p0, std=None):
        if m["discrete"]:
            d = policies.CategoricalDistribution(A)
            d.set_param(torch.as_tensor(p0))
        else:
            d = policies.DiagGaussianDistribution(A)
            d.set_param(torch.as_tensor(p0), torch.as_tensor(std))
        return d

    def split(p0, std=None):          # the shape split_distributions (operations.py:53-72) produces
        flat = p0.reshape(-1, A)
        objs = [wrap(row[None] if m["discrete"] else row, std) for row in flat]
        return np.array(objs, dtype=object).reshape(p0.shape[:-1])

    make = split if per_sample_objects else wrap
    buf = xb.DummyOnPolicyBuffer(obs_space, act_space, {"old_dist": None}, N, T, True, True, 0.99, 0.95)
    for t in range(T):
        buf.store(g["obs"][t], g["act"][t], g["rew"][t], g["val"][t], np.zeros(N, bool),
                  {"old_dist": make(g["p0"][t], None if m["discrete"] else g["std"][t])})
    for i in range(N):
        buf.finish_path(g["boot"][i], i)
    b1 = buf.sample(g["idx"])
    ok, err = gae_close(b1[4], g["s1_adv"], 2e-5)
    assert ok, err
    old = b1[5]["old_dist"].get_param()
    old = (old,) if m["discrete"] else old
    for k, t in enumerate(old):
        assert np.array_equal(t.cpu().numpy(), g["s1_old%d" % k])
    buf.auxiliary_infos["old_dist"] = make(g["new_p0"], None if m["discrete"] else g["new_std"])
    old = buf.sample(g["idx"])[5]["old_dist"].get_param()
    old = (old,) if m["discrete"] else old
    for k, t in enumerate(old):
        assert np.array_equal(t.cpu().numpy(), g["s2_old%d" % k])
    assert buf.auxiliary_infos["old_dist"].shape == (N, T)


@pytest.mark.parametrize("obs_dim", [5, 6, 8])
def test_wide_observation_rows_store_and_gather_bit_exact(obs_dim):
    """Observations of 5..8 floats (Acrobot-v1: 6) occupy two float4 per transition: store -> sample / gather_obs return
    exactly the stored rows under the reference's flat index k -> (env = k // T, step = k % T) (memory_tools.py:234)."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import ops, spaces
    T, N = 12, 37
    rng = np.random.default_rng(obs_dim)
    obs_space = spaces.Box(-np.ones(obs_dim, np.float32), np.ones(obs_dim, np.float32))
    buf = xb.DummyOnPolicyBuffer(obs_space, spaces.Discrete(3), {"old_logp": ()}, N, T, True, False, 0.99, 0.95)
    obs = rng.standard_normal((T, N, obs_dim)).astype(np.float32)
    act = rng.integers(0, 3, (T, N)).astype(np.int64)
    rew, val, logp = (rng.standard_normal((T, N)).astype(np.float32) for _ in range(3))
    for t in range(T):
        buf.store(obs[t], act[t], rew[t], val[t], np.zeros(N, bool), {"old_logp": logp[t]})
    for i in range(N):
        buf.finish_path(0.0, i)
    assert np.array_equal(buf.observations, obs.transpose(1, 0, 2))
    idx = rng.permutation(T * N)[:200].astype(np.int64)
    env, step = np.divmod(idx, T)
    s_obs, s_act, _, s_val, _, aux = buf.sample(idx)
    assert np.array_equal(s_obs, obs[step, env]) and np.array_equal(s_act, act[step, env].astype(np.float32))
    assert np.array_equal(s_val, val[step, env]) and np.array_equal(aux["old_logp"], logp[step, env])
    out = torch.empty((200, obs_dim), device="cuda")
    ops.gather_obs(torch.from_numpy(idx).cuda(), T, N, buf._obs, obs_dim, out)
    assert np.array_equal(out.cpu().numpy(), obs[step, env])
