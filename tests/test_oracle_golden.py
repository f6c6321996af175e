"""CPU: pin the oracle (oracle/ref_port.py, oracle/classic_control.c, oracle/gym_restated.py) against the
golden vectors generated from the live reference (oracle/make_goldens.py), and against numpy's PCG64."""
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle, gym_restated, ref_port
from tests.helpers import gae_close, load_golden, rel_close, replay_buffer_protocol
from xuanpolicy_b200 import policies, spaces


def test_pcg64_matches_numpy():
    for seed in (0, 1, 12345):
        rng = c_oracle.pcg64_state(seed)
        got = c_oracle.pcg64_uniform(rng, -0.05, 0.05, 64)
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        assert np.array_equal(got, g.uniform(-0.05, 0.05, size=64))
    rng = c_oracle.pcg64_state(1)
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(1)))
    high = np.array([math.pi, 1.0])
    for _ in range(5):
        ref = g.uniform(low=-high, high=high)
        got = [c_oracle.pcg64_uniform(rng, -math.pi, math.pi, 1)[0], c_oracle.pcg64_uniform(rng, -1.0, 1.0, 1)[0]]
        assert np.array_equal(ref, got)


def test_survey_kats():
    """SURVEY.md §8(c) known-answer vectors (self-derived from the published equations)."""
    e = gym_restated.CartPoleRestated("cr", with_spaces=False)
    e.state = (0.01, 0.02, 0.03, 0.04)
    e.step(1)
    assert [v.hex() for v in e.state] == ["0x1.54c985f06f694p-7", "0x1.b7a9b9e6f04bcp-3", "0x1.f8a0902de00d1p-6",
                                          "-0x1.f1ce03128d125p-3"]
    e.step(0)
    assert [v.hex() for v in e.state] == ["0x1.e17ab73018774p-7", "0x1.3971dd068cec8p-6", "0x1.a8fa7b3525a3cp-6",
                                          "0x1.e4b44a16b46c4p-5"]
    p = gym_restated.PendulumRestated("cr", with_spaces=False)
    p.state = (1.0, 0.5)
    _, r, _, _, _ = p.step(np.array([0.7], np.float32))
    assert [v.hex() for v in p.state] == ["0x1.0fd2768cd4a2ap+0", "0x1.3c7143009cb40p+0"]
    assert (-(-r)).hex() == "-0x1.0686833c4da90p+0"
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(1)))
    assert g.uniform(-0.05, 0.05, 4).tolist() == [0.0011821624700256717, 0.045046369632593536,
                                                   -0.03558403872803663, 0.04486494471372439]


def test_c_sincos_cr_matches_mpmath():
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-0.5, 0.5, 1500), rng.uniform(-90, 90, 1500), [0.0, -0.0, 1e-300, math.pi / 2]])
    s, c = c_oracle.sincos(x, "cr")
    for xi, si, ci in zip(x, s, c):
        assert si == gym_restated._sin_cr(float(xi)) and ci == gym_restated._cos_cr(float(xi))


@pytest.mark.parametrize("name", ["physics_cartpole_cr", "physics_pendulum_cr"])
def test_physics_tape_c_and_python_agree(name):
    g = load_golden(name)
    m = g["meta"]
    env = c_oracle.VecEnvC(m["env_id"], m["n"], seed=m["seed"], flavour="cr")
    assert np.array_equal(env.state, g["state0"]) and np.array_equal(env.obs, g["obs0"])
    pys = [gym_restated.make(m["env_id"], trig="cr") for _ in range(2)]
    for pe in pys:
        pe.reset(seed=m["seed"])
        pe.reset()
    for t in range(m["steps"]):
        o = env.step(g["actions"][t])
        for k in ("obs", "rew", "term", "trunc", "state", "ep_step", "ep_score"):
            assert np.array_equal(o[k], g[k][t]), (k, t)
        done = o["term"] | o["trunc"]
        assert np.array_equal(o["reset_obs"][done], g["reset_obs"][t][done])
        if t < 260:  # the (slow, mpmath) Python restatement follows the first two envs
            for i, pe in enumerate(pys):
                a = g["actions"][t][i]
                ob, rw, te, tr, _ = pe.step(int(a) if m["env_id"] == "CartPole-v1" else np.array([a], np.float32))
                assert np.array_equal(ob, g["obs"][t][i]) and np.float32(rw) == g["rew"][t][i]
                if te or tr:     # the tape's fp64 state is recorded after the auto-reset
                    ro, _ = pe.reset()
                    assert np.array_equal(ro, g["reset_obs"][t][i])
                assert np.array_equal(np.array(pe.env.state), g["state"][t][i])


@pytest.mark.parametrize("name", ["vecenv_cartpole", "vecenv_pendulum", "vecenv_mountaincar_stack"])
def test_vecenv_port_matches_reference_golden(name):
    """ref_port.VecEnvPort and the C oracle vs the reference DummyVecEnv_Gym run (libm flavour; the MountainCar golden —
    the reference's 4-frame `MountainCar` wrapper as built by its own make_envs — was recorded with the "cr" flavour)."""
    g = load_golden(name)
    m = g["meta"]
    port = ref_port.VecEnvPort(m["env_id"], m["n"], seed=m["seed"], trig=m["trig"])
    cenv = c_oracle.VecEnvC(m["env_id"], m["n"], seed=m["seed"], flavour=m["trig"])
    if "space_low" in g:
        assert port.observation_space.shape == (8,) and np.array_equal(port.observation_space.low, g["space_low"].astype(np.float32))
    obs0, _ = port.reset()
    assert np.array_equal(obs0, g["obs0"]) and np.array_equal(cenv.obs, g["obs0"])
    assert port.max_episode_length == m["max_episode_length"]
    for t in range(m["steps"]):
        a = g["actions"][t]
        o, r, d, tr, infos = port.step(a)
        co = cenv.step(a)
        assert np.array_equal(o, g["obs"][t]) and np.array_equal(r, g["rew"][t])
        assert np.array_equal(d, g["term"][t]) and np.array_equal(tr, g["trunc"][t])
        assert [i["episode_step"] for i in infos] == g["ep_step"][t].tolist()
        assert [i["episode_score"] for i in infos] == g["ep_score"][t].tolist()
        assert np.array_equal(co["obs"], g["obs"][t]) and np.array_equal(co["rew"], g["rew"][t])
        assert np.array_equal(co["term"], g["term"][t]) and np.array_equal(co["trunc"], g["trunc"][t])
        assert np.array_equal(co["ep_step"], g["ep_step"][t]) and np.array_equal(co["ep_score"], g["ep_score"][t])
        for i, inf in enumerate(infos):
            if d[i] or tr[i]:
                assert np.array_equal(inf["reset_obs"], g["reset_obs"][t][i])
                assert np.array_equal(co["reset_obs"][i], g["reset_obs"][t][i])
            else:
                assert "reset_obs" not in inf


@pytest.mark.parametrize("name", ["buffer_cat_gae", "buffer_box_gae_noadvnorm", "buffer_cat_nogae"])
def test_buffer_port_matches_reference_golden(name):
    g = load_golden(name)
    m = g["meta"]
    obs_shape = g["obs"].shape[2:]
    act_shape = g["act"].shape[2:]
    buf = ref_port.OnPolicyBufferPort(obs_shape, act_shape, m["n_envs"], m["n_size"], m["use_gae"], m["use_advnorm"],
                                      m["gamma"], m["lam"])
    replay_buffer_protocol(buf, g)
    assert np.array_equal(buf.observations, g["observations"]) and np.array_equal(buf.actions, g["actions"])
    ok, err = gae_close(buf.advantages, g["advantages"])
    assert ok, err
    ok, err = gae_close(buf.returns, g["returns"])
    assert ok, err
    # the batched fp64 form used as the GPU target
    adv64, ret64 = c_oracle.gae(g["rew"], g["val"], g["term"].astype(np.float32), g["boot"][-1], m["gamma"], m["lam"],
                                segend=g["trunc"].astype(np.uint8), boot=g["boot"], use_gae=m["use_gae"])
    ok, err = gae_close(adv64.T, g["advantages"])
    assert ok, err
    ok, err = gae_close(ret64.T, g["returns"])
    assert ok, err
    for k in range(g["idx"].shape[0]):
        o, a, r, v, adv, aux = buf.sample(g["idx"][k])
        assert np.array_equal(o, g["s_obs"][k]) and np.array_equal(a, g["s_act"][k])
        assert np.array_equal(v, g["s_val"][k]) and np.array_equal(aux["old_logp"], g["s_logp"][k])
        assert gae_close(r, g["s_ret"][k])[0]
        assert np.allclose(adv, g["s_adv"][k], rtol=0, atol=2e-5), np.abs(adv - g["s_adv"][k]).max()


def _policy_from_golden(g):
    m = g["meta"]
    if m["discrete"]:
        obs_space, act_space = spaces.Box(-1, 1, (4,)), spaces.Discrete(2)
    else:
        obs_space, act_space = spaces.Box(-1, 1, (3,)), spaces.Box(-2.0, 2.0, (1,))
    pol = policies.make_policy(obs_space, act_space, hidden=(m["hidden"],), device="cpu")
    sd = {k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")}
    pol.load_state_dict(sd, strict=True)     # same parameter names as the reference modules
    return pol


@pytest.mark.parametrize("name", ["loss_cat_h64", "loss_gauss_h128", "loss_cat_h128"])
def test_learner_port_matches_reference_golden(name):
    g = load_golden(name)
    m = g["meta"]
    for tag, clip in (("noclip", False), ("clip", True)):
        pol = _policy_from_golden(g)
        opt = torch.optim.Adam(pol.parameters(), 4e-4, eps=1e-5)
        sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
        info = ref_port.ppo_clip_update(pol, opt, sched, (g["obs"], g["act"], g["ret"], g["val"], g["adv"], g["old_logp"]),
                                        vf_coef=m["vf_coef"], ent_coef=m["ent_coef"], clip_range=m["clip_range"],
                                        clip_grad_norm=m["clip_grad_norm"], use_grad_clip=clip)
        for k in ("actor-loss", "critic-loss", "entropy", "learning_rate", "predict_value", "clip_ratio"):
            assert abs(float(info[k]) - float(g["info_%s/%s" % (tag, k)])) <= 1e-5 * max(1.0, abs(float(g["info_%s/%s" % (tag, k)]))), k
        for k, p in pol.named_parameters():
            ok, err = rel_close(p.grad.numpy(), g["grad_%s/%s" % (tag, k)], 1e-4)
            assert ok, (k, err)
            ok, err = rel_close(p.detach().numpy(), g["p1_%s/%s" % (tag, k)], 1e-5)
            assert ok, (k, err)


def _old_from_golden(g):
    return g["old_logits"] if "old_logits" in g else (g["old_mu"], g["old_std"])


@pytest.mark.parametrize("name", ["loss_ppokl_cat_h64", "loss_ppokl_gauss_h64"])
def test_ppokl_port_matches_reference_golden(name):
    """Row f3: the PPO-KL restatement (two updates, adaptive kl_coef) vs the reference's own PPOKL_Learner."""
    g = load_golden(name)
    m = g["meta"]
    pol = _policy_from_golden(g)
    opt = torch.optim.Adam(pol.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    state = {"kl_coef": 1.0}
    for it in (1, 2):
        info = ref_port.ppokl_update(pol, opt, sched, (g["obs"], g["act"], g["ret"], g["adv"]), _old_from_golden(g), state,
                                     vf_coef=m["vf_coef"], ent_coef=m["ent_coef"], target_kl=m["target_kl"])
        assert state["kl_coef"] == float(g["kl_coef%d" % it])
        for k, v in info.items():
            ref = float(g["info%d/%s" % (it, k)])
            assert abs(float(v) - ref) <= 1e-5 * max(1.0, abs(ref)), (it, k)
        for k, p in pol.named_parameters():
            ok, err = rel_close(p.grad.numpy(), g["grad%d/%s" % (it, k)], 1e-4)
            assert ok, (it, k, err)
            ok, err = rel_close(p.detach().numpy(), g["p%d/%s" % (it, k)], 1e-5)
            assert ok, (it, k, err)


def _ppg_policy_from_golden(g, device="cpu"):
    m = g["meta"]
    rep = policies.MLPRepresentation((4 if m["discrete"] else 3,), [m["hidden"]], device=device)
    if m["discrete"]:
        pol = policies.CategoricalPPGActorCritic(spaces.Discrete(2), rep, [m["hidden"]], [m["hidden"]], device=device)
    else:
        pol = policies.GaussianPPGActorCritic(spaces.Box(-2.0, 2.0, (1,)), rep, [m["hidden"]], [m["hidden"]], device=device)
    pol.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")}, strict=True)
    return pol


@pytest.mark.parametrize("name", ["loss_ppg_cat_h32", "loss_ppg_gauss_h64"])
def test_ppg_port_matches_reference_golden(name):
    """Row f3: the three PPG phase updates restated vs the reference's own PPG_Learner (run in sequence)."""
    g = load_golden(name)
    m = g["meta"]
    pol = _ppg_policy_from_golden(g)
    opt = torch.optim.Adam(pol.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    for phase in ("policy", "critic", "auxiliary"):
        info = ref_port.ppg_update(phase, pol, opt, sched, (g["obs"], g["act"], g["ret"], g["adv"]), _old_from_golden(g),
                                   ent_coef=m["ent_coef"], clip_range=m["clip_range"], kl_beta=m["kl_beta"])
        for k, v in info.items():
            ref = float(g["info_%s/%s" % (phase, k)])
            assert abs(float(v) - ref) <= 1e-5 * max(1.0, abs(ref)), (phase, k)
        for k, p in pol.named_parameters():
            key = "grad_%s/%s" % (phase, k)
            assert (p.grad is None) == (key not in g), (phase, k)
            if p.grad is not None:
                ok, err = rel_close(p.grad.numpy(), g[key], 1e-4)
                assert ok, (phase, k, err)
            ok, err = rel_close(p.detach().numpy(), g["p_%s/%s" % (phase, k)], 1e-5)
            assert ok, (phase, k, err)


def test_mountaincar_c_and_python_restatements_agree_and_kat():
    """MountainCar-v0: the C oracle and the Python restatement of gym 0.26.2's MountainCarEnv agree bit for
    bit (libm flavour) over random action tapes with resets, and on a hand-computable known answer: from
    (position, velocity) = (-0.5, 0) with action 2: velocity = 0.001 + cos(-1.5) * -0.0025, position = -0.5 + velocity."""
    import math
    from oracle import c_oracle, gym_restated
    n = 6
    ref = c_oracle.VecEnvC("gym:MountainCar-v0", n, seed=3, flavour="libm", n_warm_resets=1)
    envs = [gym_restated.make("MountainCar-v0", trig="libm") for _ in range(n)]
    obs = np.stack([e.reset(seed=3)[0] for e in envs])
    assert np.array_equal(obs, ref.obs)
    rng = np.random.default_rng(0)
    episodes = 0
    for t in range(450):
        a = rng.integers(0, 3, n)
        o = ref.step(a)
        for i, e in enumerate(envs):
            ob, r, term, trunc, _ = e.step(a[i])
            assert np.array_equal(ob, o["obs"][i]) and r == o["rew"][i] and term == o["term"][i] and trunc == o["trunc"][i]
            if term or trunc:
                assert np.array_equal(e.reset()[0], o["reset_obs"][i])
                episodes += 1
    assert episodes >= n
    e = gym_restated.MountainCarRestated("cr", with_spaces=False)
    e.state = (-0.5, 0.0)
    ob, r, term, trunc, _ = e.step(2)
    v = 0.001 + math.cos(-1.5) * -0.0025
    assert e.state == (-0.5 + v, v) and r == -1.0 and not term
    assert ob.dtype == np.float32 and ob[0] == np.float32(-0.5 + v)


ACROBOT_KAT = ["-0x1.171879311533ap-5", "0x1.ca4ce60785caap-7", "-0x1.e3f658adeead0p-6", "0x1.39e01fea84742p-2"]


def test_acrobot_c_and_python_restatements_agree_and_kat():
    """Acrobot-v1: the C oracle and the Python restatement (written expression by expression like gym
    0.26.2's AcrobotEnv: _dsdt, rk4, wrap, bound) agree bit for bit in both flavours over episodes that terminate and
    truncate; known-answer values (self-derived from the published equations, like the survey's other KATs)."""
    from oracle import c_oracle, gym_restated
    for fl in ("cr", "libm"):
        n = 5
        env = c_oracle.VecEnvC("Acrobot-v1", n, seed=1, flavour=fl, n_warm_resets=2)
        pys = []
        for e in range(n):
            p = gym_restated.make("Acrobot-v1", trig=fl)
            p.reset(seed=1)
            o, _ = p.reset()
            pys.append(p)
            assert np.array_equal(o, env.obs[e])
        rng = np.random.default_rng(3)
        n_term = n_trunc = 0
        for t in range(620):
            a = rng.integers(0, 3, n)
            a[:3] = np.where(env.state[:3, 3] > 0, 2, 0)          # torque along the second joint's velocity: swings up, terminates
            o = env.step(a)
            for e in range(n):
                ob, r, term, trunc, _ = pys[e].step(a[e])
                assert np.array_equal(ob, o["obs"][e]) and r == o["rew"][e], (fl, t, e)
                assert term == o["term"][e] and trunc == o["trunc"][e], (fl, t, e)
                n_term, n_trunc = n_term + int(term), n_trunc + int(trunc)
                if term or trunc:
                    ro, _ = pys[e].reset()
                    assert np.array_equal(ro, o["reset_obs"][e])
                assert np.array_equal(np.asarray(pys[e].env.state, np.float64), o["state"][e]), (fl, t, e)
        assert n_term > 0 and n_trunc > 0
    # KAT: PCG64(SeedSequence(1)), second reset draw (the first observation the agent sees), then action 2 ("cr" flavour)
    env = c_oracle.VecEnvC("Acrobot-v1", 1, seed=1, flavour="cr", n_warm_resets=2)
    draw = np.random.Generator(np.random.PCG64(np.random.SeedSequence(1)))
    draw.uniform(-0.1, 0.1, 4)
    assert np.array_equal(env.state[0], draw.uniform(-0.1, 0.1, 4).astype(np.float32).astype(np.float64))
    o = env.step(np.array([2]))
    assert o["rew"][0] == -1.0 and not o["term"][0]
    assert [float.hex(float(v)) for v in o["state"][0]] == ACROBOT_KAT
