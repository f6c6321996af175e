"""Goldens from the LIVE reference agent loop: tests/golden/agent_ppo_*.npz and rms_*.npz.

ORACLE / TEST INFRASTRUCTURE.  Build container only:  python -m oracle.make_goldens_agent
Runs the unmodified `PPOCLIP_Agent.train` (xuance/torch/agents/policy_gradient/ppoclip_agent.py:59-111) built by the
reference's own `get_runner` over `DummyVecEnv_Gym` / `DummyOnPolicyBuffer` / `PPOCLIP_Learner` with the yaml defaults
`use_obsnorm: True`, `use_rewnorm: True` (xuance/configs/ppo/classic_control/*.yaml:34-35), and records — without
touching reference code, by wrapping bound methods of the live objects —
    * every action the agent sampled (so that other implementations can replay the same trajectory),
    * every minibatch permutation `np.random.shuffle` produced (:76-78),
    * the rollout buffer right before each `memory.clear()` (observations as stored = normalised, rewards as stored =
      scaled, values, old_logp, terminals, returns, advantages), the bootstrap values of the buffer-full finish_path calls,
    * obs_rms / ret_rms (mean, var, count) and the per-env return tracker at those points and at the end (:87-92,
      statistic_tools.py:35-112, agent.py:104-123),
    * the policy parameters before training and after each update phase, and the learner's last info dict.
Physics: restated gym 0.26.2, flavour "cr" (correctly-rounded trig), so the device kernels reproduce the trajectory
bit for bit under the taped actions.

rms_*.npz: `RunningMeanStd.update` on a sequence of batches (vector and scalar shape, float32 and float64 inputs as the
agent feeds them) + `_process_observation` / `_process_reward` of a live agent holding those statistics.
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

from oracle import ref_agent, ref_loader  # noqa: E402


def _rms_state(r):
    return np.asarray(r.mean, np.float64).copy(), np.asarray(r.var, np.float64).copy(), float(r.count)


def gen_agent(name, env_id, n_envs, n_steps, train_steps, hidden, n_epoch, n_minibatch, seed, gamma):
    runner = ref_agent.build_runner(env_id, trig="cr", parallels=n_envs, n_steps=n_steps, seed=seed, n_epoch=n_epoch,
                                    n_minibatch=n_minibatch, gamma=gamma, representation_hidden_size=[hidden],
                                    actor_hidden_size=[hidden], critic_hidden_size=[hidden])
    agent = runner.agent
    import xuance.torch.agents.policy_gradient.ppoclip_agent as mod
    mod.tqdm = lambda x: x
    assert agent.use_obsnorm and agent.use_rewnorm
    rec = dict(actions=[], perms=[], rollouts=[], infos=[], boot=[])
    state0 = {k: v.detach().numpy().copy() for k, v in agent.policy.state_dict().items()}
    flags = {"main": True}

    orig_action = agent._action
    def action(obs):
        a, v, lp = orig_action(obs)
        if flags["main"]:
            rec["actions"].append(np.array(a).copy())
            flags["main"] = False
        return a, v, lp
    agent._action = action

    orig_rms_update = agent.obs_rms.update
    def rms_update(x):
        flags["main"] = True               # obs_rms.update(obs) opens every loop iteration (:62)
        return orig_rms_update(x)
    agent.obs_rms.update = rms_update

    orig_shuffle = np.random.shuffle
    def shuffle(x):
        orig_shuffle(x)
        rec["perms"].append(np.array(x).copy())
    orig_update = agent.learner.update
    def update(*a):
        info = orig_update(*a)
        rec["last_info"] = {k: float(v) for k, v in info.items()}
        return info
    agent.learner.update = update

    mem = agent.memory
    orig_clear = mem.clear
    def clear():
        snap = dict(obs=mem.observations.copy(), act=mem.actions.copy(), rew=mem.rewards.copy(), val=mem.values.copy(),
                    ret=mem.returns.copy(), adv=mem.advantages.copy(), term=mem.terminals.copy(),
                    logp=mem.auxiliary_infos["old_logp"].copy(), returns_tracker=np.asarray(agent.returns, np.float64).copy())
        for k, r in (("obs_rms", agent.obs_rms), ("ret_rms", agent.ret_rms)):
            m, v, c = _rms_state(r)
            snap[k + "_mean"], snap[k + "_var"], snap[k + "_count"] = m, v, np.float64(c)
        snap["params"] = {k: v.detach().numpy().copy() for k, v in agent.policy.state_dict().items()}
        snap["info"] = dict(rec["last_info"])
        rec["rollouts"].append(snap)
        return orig_clear()
    mem.clear = clear

    torch.manual_seed(seed + 100)
    np.random.seed(seed + 100)
    np.random.shuffle = shuffle
    try:
        agent.train(train_steps)
    finally:
        np.random.shuffle = orig_shuffle

    out = {"actions": np.asarray(rec["actions"]), "perms": np.asarray(rec["perms"])}
    for k, v in state0.items():
        out["p0/" + k] = v
    for i, snap in enumerate(rec["rollouts"]):
        for k, v in snap.items():
            if k == "params":
                for pk, pv in v.items():
                    out["r%d/params/%s" % (i, pk)] = pv
            elif k == "info":
                out["r%d/info" % i] = np.array(json.dumps(v))
            else:
                out["r%d/%s" % (i, k)] = v
    for k, r in (("obs_rms", agent.obs_rms), ("ret_rms", agent.ret_rms)):
        m, v, c = _rms_state(r)
        out["end/" + k + "_mean"], out["end/" + k + "_var"], out["end/" + k + "_count"] = m, v, np.float64(c)
    out["end/returns_tracker"] = np.asarray(agent.returns, np.float64)
    out["end/buffer_ptr"] = np.int64(mem.ptr)
    out["end/obs_rows"] = mem.observations[:, :mem.ptr].copy()
    out["end/rew_rows"] = mem.rewards[:, :mem.ptr].copy()
    out["end/current_step"] = np.int64(agent.current_step)
    cfg = agent.config
    out["meta"] = np.array(json.dumps(dict(
        env_id=env_id, n_envs=n_envs, n_steps=n_steps, train_steps=train_steps, hidden=hidden, n_epoch=n_epoch,
        n_minibatch=n_minibatch, seed=seed, gamma=gamma, gae_lambda=cfg.gae_lambda, vf_coef=cfg.vf_coef, ent_coef=cfg.ent_coef,
        clip_range=cfg.clip_range, clip_grad_norm=cfg.clip_grad_norm, learning_rate=cfg.learning_rate,
        use_obsnorm=True, use_rewnorm=True, obsnorm_range=cfg.obsnorm_range, rewnorm_range=cfg.rewnorm_range,
        running_steps=int(cfg.running_steps), total_iters=int(agent.learner.scheduler.total_iters), trig="cr",
        n_rollouts=len(rec["rollouts"]), rng_seed=seed + 100)))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    term = sum(int(s["term"].sum()) for s in rec["rollouts"])
    print("wrote", name, "rollouts", len(rec["rollouts"]), "terminals", term, "actions", out["actions"].shape,
          "perms", out["perms"].shape)


def gen_rms(name, seed):
    """RunningMeanStd.update sequences + the agent's _process_observation / _process_reward (live reference objects)."""
    from xuance.common import RunningMeanStd
    rng = np.random.default_rng(seed)
    out = {}
    # vector statistics, float32 batches of varying size (what obs_rms sees: envs.buf_obs, ppoclip_agent.py:60-62)
    r = RunningMeanStd(shape=(4,), comm=None, use_mpi=False)
    batches = [(rng.standard_normal((n, 4)) * np.array([1.0, 3.0, 0.1, 10.0]) + np.array([0.0, 1.0, -2.0, 5.0])).astype(np.float32)
               for n in (16, 16, 1, 7, 64, 3)]
    for i, b in enumerate(batches):
        r.update(b)
        out["vec/batch%d" % i] = b
        out["vec/mean%d" % i], out["vec/var%d" % i], out["vec/count%d" % i] = (np.asarray(r.mean, np.float64).copy(),
                                                                                np.asarray(r.var, np.float64).copy(), np.float64(r.count))
        out["vec/mean_dtype%d" % i] = np.array(str(np.asarray(r.mean).dtype))
    # scalar statistics, one finished-episode return at a time, float64 (what ret_rms sees: returns[i:i+1], :87-90)
    s = RunningMeanStd(shape=(), comm=None, use_mpi=False)
    rets = rng.standard_normal(12) * 20 - 50
    for i, x in enumerate(rets):
        s.update(np.asarray([x], np.float64))
        out["sc/mean%d" % i], out["sc/var%d" % i], out["sc/count%d" % i] = (np.float64(s.mean), np.float64(s.var), np.float64(s.count))
    out["sc/batches"] = rets
    # the agent's two normalisers with those statistics
    runner = ref_agent.build_runner("CartPole-v1", trig="cr", parallels=2, n_steps=8)
    agent = runner.agent
    agent.obs_rms.mean, agent.obs_rms.var, agent.obs_rms.count = r.mean, r.var, r.count
    agent.ret_rms.mean, agent.ret_rms.var, agent.ret_rms.count = s.mean, s.var, s.count
    obs = (rng.standard_normal((9, 4)) * np.array([2.0, 30.0, 0.2, 100.0])).astype(np.float32)
    rew = (rng.standard_normal(9) * 300).astype(np.float32)
    out["proc/obs_in"], out["proc/obs_out"] = obs, np.asarray(agent._process_observation(obs.copy()))
    out["proc/rew_in"], out["proc/rew_out"] = rew, np.asarray(agent._process_reward(rew.copy()))
    out["proc/ranges"] = np.array([agent.obsnorm_range, agent.rewnorm_range], np.float64)
    agent.ret_rms.var = np.float64(1e-6)         # std below the 0.1 floor of np.clip(std, 0.1, 100) (agent.py:119)
    out["proc/rew_out_floor"] = np.asarray(agent._process_reward(rew.copy()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def gen_vecenv_mountaincar(name, n, steps, seed=2):
    """The reference's `make_envs` on MountainCar-v0: DummyVecEnv_Gym over `MountainCar(Gym_Env)` (4-frame stack,
    xuance/environment/gym/gym_env.py:50-83, selected at xuance/environment/__init__.py:66-67) over the restated physics,
    flavour "cr" (the device kernel must match bit for bit)."""
    from argparse import Namespace
    import xuance.environment as E
    cfg = Namespace(env_name="Classic Control", env_id="MountainCar-v0", seed=seed, parallels=n, vectorize="Dummy_Gym",
                    render_mode="rgb_array")
    envs = E.make_envs(cfg)
    assert type(envs.envs[0]).__name__ == "MountainCar" and envs.observation_space.shape == (8,)
    obs0, _ = envs.reset()
    rng = np.random.default_rng(13)
    rec = dict(obs0=obs0, actions=[], obs=[], rew=[], term=[], trunc=[], ep_step=[], ep_score=[], reset_obs=[])
    for t in range(steps):
        # energy pumping (push along the velocity) on most envs so that some reach the goal; random actions on the rest
        heur = np.where(envs.buf_obs[:, 7] > 0, 2, 0).astype(np.int64)
        a = np.where(rng.random(n) < 0.9, heur, rng.integers(0, 3, n))
        a[n // 2:] = rng.integers(0, 3, n - n // 2)
        o, r, d, tr, infos = envs.step(a)
        ro = np.full_like(o, np.nan)
        for i, inf in enumerate(infos):
            if "reset_obs" in inf:
                ro[i] = np.asarray(inf["reset_obs"])
        rec["actions"].append(a); rec["obs"].append(o); rec["rew"].append(r); rec["term"].append(d)
        rec["trunc"].append(tr); rec["reset_obs"].append(ro)
        rec["ep_step"].append([inf["episode_step"] for inf in infos])
        rec["ep_score"].append([inf["episode_score"] for inf in infos])
    out = {k: (np.asarray(v) if k != "obs0" else v) for k, v in rec.items()}
    out["space_low"], out["space_high"] = np.asarray(envs.observation_space.low), np.asarray(envs.observation_space.high)
    out["meta"] = np.array(json.dumps(dict(env_id="MountainCar-v0", n=n, steps=steps, seed=seed, trig="cr",
                                           max_episode_length=int(envs.max_episode_length), n_actions=int(envs.action_space.n))))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "terminals", int(out["term"].sum()), "truncations", int(out["trunc"].sum()))


def main():
    assert ref_loader.source_tree_available(), "needs the reference source tree (/root/reference)"
    os.makedirs(OUT, exist_ok=True)
    ref_loader.load(trig="cr")
    gen_rms("rms_reference", 301)
    gen_vecenv_mountaincar("vecenv_mountaincar_stack", 8, 460)
    gen_agent("agent_ppo_cartpole", "CartPole-v1", n_envs=8, n_steps=32, train_steps=3 * 32 + 5, hidden=32, n_epoch=2,
              n_minibatch=4, seed=3, gamma=0.98)
    gen_agent("agent_ppo_pendulum", "Pendulum-v1", n_envs=4, n_steps=128, train_steps=2 * 128 + 9, hidden=32, n_epoch=2,
              n_minibatch=2, seed=4, gamma=0.98)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "mountaincar":
        ref_loader.load(trig="cr")
        gen_vecenv_mountaincar("vecenv_mountaincar_stack", 8, 460)
    else:
        main()
