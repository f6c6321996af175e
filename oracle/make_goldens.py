"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (/root/reference) on seeded inputs.

ORACLE / TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_goldens
The reference ships no golden vectors for this path, so these are produced by importing its
own classes behind stub gym/gymnasium/mpi4py modules (oracle/ref_loader.py):

    buffer_*.npz   DummyOnPolicyBuffer.store/finish_path/sample   xuance/common/memory_tools.py:143-245
                   driven with the exact call protocol of PPOCLIP_Agent.train (ppoclip_agent.py:68-100)
    loss_*.npz     PPOCLIP_Learner.update                          xuance/torch/learners/policy_gradient/ppoclip_learner.py:24-65
                   with the reference's own policy modules (policies/categorical.py, policies/gaussian.py)
    loss_ppokl_*   PPOKL_Learner.update                            .../ppokl_learner.py:21-61 (two updates: adaptive kl_coef)
    loss_ppg_*     PPG_Learner.update_policy/_critic/_auxiliary    .../ppg_learner.py:23-88
    vecenv_*.npz   DummyVecEnv_Gym.reset/step                      xuance/environment/gym/gym_vec_env.py:155-212
                   over the RESTATED physics (gym itself is absent: "parity unpinned" for the physics)
    physics_*.npz  action tapes through the C oracle, flavour "cr" (self-derived KATs, Tier-1 target)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

from oracle import ref_loader, c_oracle  # noqa: E402


def _spaces():
    import gym.spaces as S
    return S


def gen_buffer(name, seed, n_envs, n_size, discrete, use_gae, use_advnorm, gamma, lam, p_term=0.04, p_trunc=0.04):
    """Replays a synthetic rollout into the reference buffer with the agent's finish_path protocol."""
    from xuance.common import DummyOnPolicyBuffer
    S = _spaces()
    rng = np.random.default_rng(seed)
    obs_dim = 4 if discrete else 3
    obs_space = S.Box(-np.ones(obs_dim, np.float32), np.ones(obs_dim, np.float32))
    act_space = S.Discrete(2) if discrete else S.Box(-2.0, 2.0, shape=(1,))
    buf = DummyOnPolicyBuffer(obs_space, act_space, {"old_logp": ()}, n_envs, n_size, use_gae, use_advnorm, gamma, lam)
    T, N = n_size, n_envs
    obs = rng.standard_normal((T, N, obs_dim)).astype(np.float32)
    act = rng.integers(0, 2, (T, N)).astype(np.int64) if discrete else rng.standard_normal((T, N, 1)).astype(np.float32)
    rew = rng.standard_normal((T, N)).astype(np.float32)
    val = rng.standard_normal((T, N)).astype(np.float32)
    logp = (-rng.random((T, N))).astype(np.float32)
    term = rng.random((T, N)) < p_term
    trunc = rng.random((T, N)) < p_trunc
    boot = rng.standard_normal((T, N)).astype(np.float32)      # V(next/terminal obs) after step t
    for t in range(T):
        buf.store(obs[t], act[t], rew[t], val[t], term[t], {"old_logp": logp[t]})
        if buf.full:                                            # ppoclip_agent.py:69-75
            for i in range(N):
                buf.finish_path(0.0 if term[t, i] else boot[t, i], i)
            break
        for i in range(N):                                      # ppoclip_agent.py:89-100
            if term[t, i] or trunc[t, i]:
                buf.finish_path(0.0 if term[t, i] else boot[t, i], i)
    idx_all = rng.permutation(N * T).astype(np.int64)
    B = (N * T) // 4
    batches = [buf.sample(idx_all[k * B:(k + 1) * B]) for k in range(2)]
    out = dict(obs=obs, act=act, rew=rew, val=val, logp=logp, term=term, trunc=trunc, boot=boot,
               returns=buf.returns, advantages=buf.advantages, observations=buf.observations, actions=buf.actions,
               idx=idx_all[:2 * B].reshape(2, B),
               s_obs=np.stack([b[0] for b in batches]), s_act=np.stack([b[1] for b in batches]),
               s_ret=np.stack([b[2] for b in batches]), s_val=np.stack([b[3] for b in batches]),
               s_adv=np.stack([b[4] for b in batches]), s_logp=np.stack([b[5]["old_logp"] for b in batches]),
               meta=np.array(json.dumps(dict(n_envs=N, n_size=T, discrete=discrete, use_gae=use_gae,
                                             use_advnorm=use_advnorm, gamma=gamma, lam=lam))))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def _ref_policy(discrete, hidden, seed):
    from xuance.torch.representations import Basic_MLP
    from xuance.torch.policies import Categorical_AC_Policy, Gaussian_AC_Policy
    S = _spaces()
    torch.manual_seed(seed)
    act = torch.nn.LeakyReLU
    init = torch.nn.init.orthogonal_
    if discrete:
        rep = Basic_MLP((4,), [hidden], None, init, act, "cpu")
        return Categorical_AC_Policy(S.Discrete(2), rep, [hidden], [hidden], None, init, act, "cpu")
    rep = Basic_MLP((3,), [hidden], None, init, act, "cpu")
    return Gaussian_AC_Policy(S.Box(-2.0, 2.0, shape=(1,)), rep, [hidden], [hidden], None, init, act, "cpu")


def gen_loss(name, seed, discrete, hidden, B, hp):
    """Reference PPOCLIP_Learner.update: grads (clip off), and info + params after one clipped Adam step."""
    from xuance.torch.learners import PPOCLIP_Learner
    rng = np.random.default_rng(seed)
    obs_dim = 4 if discrete else 3
    obs = rng.standard_normal((B, obs_dim)).astype(np.float32)
    policy = _ref_policy(discrete, hidden, seed)
    with torch.no_grad():
        # perturb so logits / logstd are not at their symmetric init
        for p in policy.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd0 = {k: v.detach().clone().numpy() for k, v in policy.state_dict().items()}
    with torch.no_grad():
        _, dist, v0 = policy(obs)
        act_t = dist.stochastic_sample()
        logp0 = dist.log_prob(act_t).numpy()
    act = act_t.numpy().astype(np.float32)                     # the buffer stores actions as float32 (memory_tools.py:173)
    old_logp = (logp0 + 0.1 * rng.standard_normal(B)).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    ret = (v0.numpy() + rng.standard_normal(B)).astype(np.float32)
    val = v0.numpy().astype(np.float32)
    out = dict(obs=obs, act=act, ret=ret, val=val, adv=adv, old_logp=old_logp,
               meta=np.array(json.dumps(dict(discrete=discrete, hidden=hidden, B=B, **hp))))
    for k, v in sd0.items():
        out["p0/" + k] = v
    for clip in (False, True):
        policy.load_state_dict({k: torch.as_tensor(v) for k, v in sd0.items()})
        opt = torch.optim.Adam(policy.parameters(), 4e-4, eps=1e-5)
        sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
        learner = PPOCLIP_Learner(policy, opt, sched, "cpu", "/tmp/xb200_goldens_models", vf_coef=hp["vf_coef"],
                                  ent_coef=hp["ent_coef"], clip_range=hp["clip_range"],
                                  clip_grad_norm=hp["clip_grad_norm"], use_grad_clip=clip)
        info = learner.update(obs, act, ret, val, adv, old_logp)
        tag = "clip" if clip else "noclip"
        for k, v in info.items():
            out["info_%s/%s" % (tag, k)] = np.asarray(float(v))
        for k, p in policy.named_parameters():
            out["grad_%s/%s" % (tag, k)] = p.grad.detach().numpy().copy()
            out["p1_%s/%s" % (tag, k)] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def gen_loss_a2c_pg(name, seed, algo, discrete, hidden, B):
    """Reference A2C_Learner.update (a2c_learner.py:19-50) / PG_Learner.update (pg_learner.py:17-45): grads + step."""
    from xuance.torch.learners import A2C_Learner, PG_Learner
    from xuance.torch.representations import Basic_MLP
    from xuance.torch.policies import Categorical_Actor_Policy
    rng = np.random.default_rng(seed)
    obs_dim = 4 if discrete else 3
    obs = rng.standard_normal((B, obs_dim)).astype(np.float32)
    if algo == "a2c":
        policy = _ref_policy(discrete, hidden, seed)
    else:
        torch.manual_seed(seed)
        rep = Basic_MLP((4,), [hidden], None, torch.nn.init.orthogonal_, torch.nn.LeakyReLU, "cpu")
        policy = Categorical_Actor_Policy(_spaces().Discrete(2), rep, [hidden], None, torch.nn.init.orthogonal_,
                                          torch.nn.LeakyReLU, "cpu")
    with torch.no_grad():
        for p in policy.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd0 = {k: v.detach().clone().numpy() for k, v in policy.state_dict().items()}
    with torch.no_grad():
        a_dist = policy(obs)[1]
        act = a_dist.stochastic_sample().numpy().astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    ret = rng.standard_normal(B).astype(np.float32)
    out = dict(obs=obs, act=act, ret=ret, adv=adv,
               meta=np.array(json.dumps(dict(algo=algo, discrete=discrete, hidden=hidden, B=B, vf_coef=0.25, ent_coef=0.01,
                                             clip_grad=0.5))))
    for k, v in sd0.items():
        out["p0/" + k] = v
    opt = torch.optim.Adam(policy.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    if algo == "a2c":
        learner = A2C_Learner(policy, opt, sched, "cpu", "/tmp/xb200_goldens_models", vf_coef=0.25, ent_coef=0.01, clip_grad=0.5)
        info = learner.update(obs, act, ret, adv)
    else:
        learner = PG_Learner(policy, opt, sched, "cpu", "/tmp/xb200_goldens_models", ent_coef=0.01, clip_grad=0.5)
        info = learner.update(obs, act, ret)
    for k, v in info.items():
        out["info/" + k] = np.asarray(float(v))
    for k, p in policy.named_parameters():
        out["grad_clipped/" + k] = p.grad.detach().numpy().copy()      # these learners always clip (a2c_learner.py:36)
        out["p1/" + k] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def _old_dist_inputs(policy, obs, rng, ppg):
    """Actions + the OLD distribution (a perturbed copy of the policy) as the reference's per-sample object array."""
    from xuance.torch.utils.operations import split_distributions
    with torch.no_grad():
        out = policy(obs)
        act = out[1].stochastic_sample().numpy().astype(np.float32)
        saved = [p.detach().clone() for p in policy.parameters()]
        for p in policy.parameters():
            p.add_(0.03 * torch.randn_like(p))
        old = policy(obs)[1]
        old_dists = split_distributions(old)
        if hasattr(old, "logits"):
            old_params = dict(old_logits=old.logits.numpy().copy())
        else:
            old_params = dict(old_mu=old.mu.numpy().copy(), old_std=old.std.numpy().copy())
        for p, q in zip(policy.parameters(), saved):
            p.copy_(q)
    return act, old_dists, old_params


def gen_loss_ppokl(name, seed, discrete, hidden, B, target_kl):
    """Reference PPOKL_Learner.update (ppokl_learner.py:21-61), two consecutive updates (the second one runs with the
    adapted kl_coef): info, grads and parameters after each, kl_coef after each."""
    from xuance.torch.learners import PPOKL_Learner
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((B, 4 if discrete else 3)).astype(np.float32)
    policy = _ref_policy(discrete, hidden, seed)
    with torch.no_grad():
        for p in policy.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd0 = {k: v.detach().clone().numpy() for k, v in policy.state_dict().items()}
    act, old_dists, old_params = _old_dist_inputs(policy, obs, rng, False)
    adv = rng.standard_normal(B).astype(np.float32)
    with torch.no_grad():
        ret = (policy(obs)[2].numpy() + rng.standard_normal(B)).astype(np.float32)
    out = dict(obs=obs, act=act, ret=ret, adv=adv, **old_params,
               meta=np.array(json.dumps(dict(algo="ppokl", discrete=discrete, hidden=hidden, B=B, vf_coef=0.25,
                                             ent_coef=0.01, target_kl=target_kl))))
    for k, v in sd0.items():
        out["p0/" + k] = v
    opt = torch.optim.Adam(policy.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    learner = PPOKL_Learner(policy, opt, sched, "cpu", "/tmp/xb200_goldens_models", vf_coef=0.25, ent_coef=0.01,
                            target_kl=target_kl)
    for it in (1, 2):
        info = learner.update(obs, act, ret, adv, old_dists)
        for k, v in info.items():
            out["info%d/%s" % (it, k)] = np.asarray(float(v))
        out["kl_coef%d" % it] = np.asarray(float(learner.kl_coef))
        for k, p in policy.named_parameters():
            out["grad%d/%s" % (it, k)] = p.grad.detach().numpy().copy()
            out["p%d/%s" % (it, k)] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "kl_coef", float(out["kl_coef1"]), float(out["kl_coef2"]))


def gen_loss_ppg(name, seed, discrete, hidden, B):
    """Reference PPG_Learner.update_policy / update_critic / update_auxiliary (ppg_learner.py:23-88), run in that order
    on one policy (Categorical_PPG_Policy / Gaussian_PPG_Policy): info, grads (None -> absent) and parameters per phase."""
    from xuance.torch.learners import PPG_Learner
    from xuance.torch.representations import Basic_MLP
    from xuance.torch.policies import Categorical_PPG_Policy, Gaussian_PPG_Policy
    S = _spaces()
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    act_fn, init = torch.nn.LeakyReLU, torch.nn.init.orthogonal_
    obs_dim = 4 if discrete else 3
    rep = Basic_MLP((obs_dim,), [hidden], None, init, act_fn, "cpu")
    if discrete:
        policy = Categorical_PPG_Policy(S.Discrete(2), rep, [hidden], [hidden], None, init, act_fn, "cpu")
    else:
        policy = Gaussian_PPG_Policy(S.Box(-2.0, 2.0, shape=(1,)), rep, [hidden], [hidden], None, init, act_fn, "cpu")
    with torch.no_grad():
        for p in policy.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd0 = {k: v.detach().clone().numpy() for k, v in policy.state_dict().items()}
    obs = rng.standard_normal((B, obs_dim)).astype(np.float32)
    act, old_dists, old_params = _old_dist_inputs(policy, obs, rng, True)
    adv = rng.standard_normal(B).astype(np.float32)
    with torch.no_grad():
        ret = (policy(obs)[2].numpy() + rng.standard_normal(B)).astype(np.float32)
    out = dict(obs=obs, act=act, ret=ret, adv=adv, **old_params,
               meta=np.array(json.dumps(dict(algo="ppg", discrete=discrete, hidden=hidden, B=B, ent_coef=0.01,
                                             clip_range=0.2, kl_beta=1.5))))
    for k, v in sd0.items():
        out["p0/" + k] = v
    opt = torch.optim.Adam(policy.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    learner = PPG_Learner(policy, opt, sched, "cpu", "/tmp/xb200_goldens_models", ent_coef=0.01, clip_range=0.2, kl_beta=1.5)
    for phase in ("policy", "critic", "auxiliary"):
        info = getattr(learner, "update_" + phase)(obs, act, ret, adv, old_dists)
        for k, v in info.items():
            out["info_%s/%s" % (phase, k)] = np.asarray(float(v))
        for k, p in policy.named_parameters():
            if p.grad is not None:
                out["grad_%s/%s" % (phase, k)] = p.grad.detach().numpy().copy()
            out["p_%s/%s" % (phase, k)] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def gen_buffer_olddist(name, seed, n_envs, n_size, discrete):
    """Reference DummyOnPolicyBuffer with the PPO-KL / PPG auxiliary {"old_dist": None} (object arrays of per-sample
    distribution wrappers, memory_tools.py:28-30): store -> finish_path -> sample, then the whole-buffer reassignment
    PPG_Agent.train performs (ppg_agent.py:90-93) and a second sample."""
    from xuance.common import DummyOnPolicyBuffer
    from xuance.torch.utils.distributions import CategoricalDistribution, DiagGaussianDistribution
    from xuance.torch.utils.operations import split_distributions, merge_distributions
    S = _spaces()
    rng = np.random.default_rng(seed)
    T, N, A = n_size, n_envs, (3 if discrete else 2)
    obs_dim = 4 if discrete else 3
    obs_space = S.Box(-np.ones(obs_dim, np.float32), np.ones(obs_dim, np.float32))
    act_space = S.Discrete(A) if discrete else S.Box(-2.0, 2.0, shape=(A,))
    buf = DummyOnPolicyBuffer(obs_space, act_space, {"old_dist": None}, N, T, True, True, 0.99, 0.95)

    def batched(p0, std=None):
        if discrete:
            d = CategoricalDistribution(A)
            d.set_param(torch.as_tensor(p0))
        else:
            d = DiagGaussianDistribution(A)
            d.set_param(torch.as_tensor(p0), torch.as_tensor(std))
        return d

    def params(objs):
        m = merge_distributions(objs)
        return (m.logits.numpy().copy(),) if discrete else (m.mu.numpy().copy(), m.std.numpy().copy())

    obs = rng.standard_normal((T, N, obs_dim)).astype(np.float32)
    act = rng.integers(0, A, (T, N)).astype(np.int64) if discrete else rng.standard_normal((T, N, A)).astype(np.float32)
    rew, val = rng.standard_normal((T, N)).astype(np.float32), rng.standard_normal((T, N)).astype(np.float32)
    p0 = rng.standard_normal((T, N, A)).astype(np.float32)
    std = np.exp(0.3 * rng.standard_normal((T, A))).astype(np.float32)       # one std row per step (policy of that step)
    boot = rng.standard_normal(N).astype(np.float32)
    for t in range(T):
        buf.store(obs[t], act[t], rew[t], val[t], np.zeros(N, bool), {"old_dist": split_distributions(batched(p0[t], std[t]))})
    for i in range(N):
        buf.finish_path(boot[i], i)
    idx = rng.permutation(N * T).astype(np.int64)[:(N * T) // 2]
    b1 = buf.sample(idx)
    out = dict(obs=obs, act=act, rew=rew, val=val, p0=p0, std=std, boot=boot, idx=idx, s1_adv=b1[4], s1_ret=b1[2])
    for k, v in enumerate(params(b1[5]["old_dist"])):
        out["s1_old%d" % k] = v
    new_p0 = rng.standard_normal((N, T, A)).astype(np.float32)               # env-major, like memory.observations
    new_std = np.exp(0.3 * rng.standard_normal(A)).astype(np.float32)
    buf.auxiliary_infos["old_dist"] = split_distributions(batched(new_p0, new_std))
    b2 = buf.sample(idx)
    out.update(new_p0=new_p0, new_std=new_std)
    for k, v in enumerate(params(b2[5]["old_dist"])):
        out["s2_old%d" % k] = v
    out["meta"] = np.array(json.dumps(dict(n_envs=N, n_size=T, discrete=discrete, A=A)))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def gen_f3_dist():
    gen_loss_ppokl("loss_ppokl_cat_h64", 205, True, 64, 512, target_kl=0.01)
    gen_loss_ppokl("loss_ppokl_gauss_h64", 206, False, 64, 768, target_kl=1e-4)
    gen_loss_ppg("loss_ppg_cat_h32", 207, True, 32, 384)
    gen_loss_ppg("loss_ppg_gauss_h64", 208, False, 64, 640)
    gen_buffer_olddist("buffer_cat_olddist", 209, 5, 24, True)
    gen_buffer_olddist("buffer_box_olddist", 210, 4, 20, False)


def gen_vecenv(name, env_id, n, steps, seed=1):
    """Reference DummyVecEnv_Gym protocol over the restated physics (libm flavour, as gym would run on a host)."""
    from xuance.environment import DummyVecEnv_Gym, Gym_Env
    envs = DummyVecEnv_Gym([lambda: Gym_Env(env_id, seed, "rgb_array") for _ in range(n)])
    obs0, infos0 = envs.reset()
    rng = np.random.default_rng(7)
    rec = dict(obs0=obs0, actions=[], obs=[], rew=[], term=[], trunc=[], ep_step=[], ep_score=[], reset_obs=[])
    for t in range(steps):
        if env_id == "CartPole-v1":
            # alternate a stabilising heuristic (reaches the 500-step truncation) with random actions
            th, thd = envs.buf_obs[:, 2], envs.buf_obs[:, 3]
            a = np.where(rng.random(n) < 0.85, (th + 0.5 * thd > 0).astype(np.int64), rng.integers(0, 2, n))
            a[n // 2:] = rng.integers(0, 2, n - n // 2)
        else:
            a = (1.5 * rng.standard_normal((n, 1))).astype(np.float32)
        o, r, d, tr, infos = envs.step(a)
        ro = np.full_like(o, np.nan)
        for i, inf in enumerate(infos):
            if "reset_obs" in inf:
                ro[i] = inf["reset_obs"]
        rec["actions"].append(a); rec["obs"].append(o); rec["rew"].append(r); rec["term"].append(d)
        rec["trunc"].append(tr); rec["reset_obs"].append(ro)
        rec["ep_step"].append([inf["episode_step"] for inf in infos])
        rec["ep_score"].append([inf["episode_score"] for inf in infos])
    out = {k: (np.asarray(v) if k != "obs0" else v) for k, v in rec.items()}
    out["meta"] = np.array(json.dumps(dict(env_id=env_id, n=n, steps=steps, seed=seed, trig="libm",
                                           max_episode_length=int(envs.max_episode_length))))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "terminals", int(out["term"].sum()), "truncations", int(out["trunc"].sum()))


def gen_physics(name, env_id, n, steps, seed=1):
    """Tier-1 tapes: the C oracle with correctly-rounded trig; fp64 states recorded bit-for-bit."""
    env = c_oracle.VecEnvC(env_id, n, seed=seed, flavour="cr")
    rng = np.random.default_rng(11)
    rec = dict(obs0=env.obs.copy(), state0=env.state.copy(), actions=[], obs=[], rew=[], term=[], trunc=[],
               state=[], reset_obs=[], ep_step=[], ep_score=[])
    for t in range(steps):
        if env_id == "CartPole-v1":
            th, thd = env.obs[:, 2], env.obs[:, 3]
            a = np.where(rng.random(n) < 0.9, (th + 0.5 * thd > 0).astype(np.int64), rng.integers(0, 2, n))
            a[n // 2:] = rng.integers(0, 2, n - n // 2)
        else:
            a = (1.5 * rng.standard_normal(n)).astype(np.float32)
        o = env.step(a)
        rec["actions"].append(a)
        for k in ("obs", "rew", "term", "trunc", "state", "reset_obs", "ep_step", "ep_score"):
            rec[k].append(o[k])
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["meta"] = np.array(json.dumps(dict(env_id=env_id, n=n, steps=steps, seed=seed, trig="cr")))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "terminals", int(out["term"].sum()), "truncations", int(out["trunc"].sum()))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_loader.load(trig="libm")
    hp = dict(vf_coef=0.25, ent_coef=0.01, clip_range=0.2, clip_grad_norm=0.5)
    gen_buffer("buffer_cat_gae", 101, 6, 48, True, True, True, 0.99, 0.95)
    gen_buffer("buffer_box_gae_noadvnorm", 102, 5, 40, False, True, False, 0.98, 0.95)
    gen_buffer("buffer_cat_nogae", 103, 4, 32, True, False, True, 0.99, 0.95)
    gen_loss("loss_cat_h64", 201, True, 64, 512, hp)
    gen_loss("loss_gauss_h128", 202, False, 128, 1024, hp)
    gen_loss("loss_cat_h128", 211, True, 128, 1024, hp)          # width the tensor-core MLP path supports, Categorical head
    gen_loss_a2c_pg("loss_a2c_gauss_h64", 203, "a2c", False, 64, 640)
    gen_loss_a2c_pg("loss_pg_cat_h32", 204, "pg", True, 32, 384)
    gen_f3_dist()
    gen_vecenv("vecenv_cartpole", "CartPole-v1", 6, 700)
    gen_vecenv("vecenv_pendulum", "Pendulum-v1", 4, 450)
    gen_physics("physics_cartpole_cr", "CartPole-v1", 12, 1100)
    gen_physics("physics_pendulum_cr", "Pendulum-v1", 8, 450)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "a2c_pg":      # regenerate only the newer fixtures
        os.makedirs(OUT, exist_ok=True)
        ref_loader.load(trig="libm")
        gen_loss_a2c_pg("loss_a2c_gauss_h64", 203, "a2c", False, 64, 640)
        gen_loss_a2c_pg("loss_pg_cat_h32", 204, "pg", True, 32, 384)
    elif len(sys.argv) > 1 and sys.argv[1] == "f3_dist":
        os.makedirs(OUT, exist_ok=True)
        ref_loader.load(trig="libm")
        gen_f3_dist()
    else:
        main()
