"""CPU restatement ("port") of the reference's on-policy PPO hot path, numpy + torch-CPU.

ORACLE / TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product (xuanpolicy_b200/) never does.

The reference is Python and lives at /root/reference, which does not travel to the GPU box.  This file
restates — in the reference's own cost structure (per-env Python loops, host numpy buffers, torch eager
CPU) — the functions on the path, so that (i) GPU parity tests have a checker there and (ii) the CPU
baseline timed beside the B200 numbers is the reference's algorithm, not a vectorised rewrite:

    VecEnvPort          DummyVecEnv_Gym.reset/step_async/step_wait   xuance/environment/gym/gym_vec_env.py:155-212
                        + Gym_Env.reset/step bookkeeping              xuance/environment/gym/gym_env.py:36-49
    OnPolicyBufferPort  DummyOnPolicyBuffer.store/finish_path/sample  xuance/common/memory_tools.py:143-245
    ppo_clip_update     PPOCLIP_Learner.update                        xuance/torch/learners/policy_gradient/ppoclip_learner.py:24-65
    RunningMeanStdPort  RunningMeanStd                                xuance/common/statistic_tools.py:35-112
    PPOAgentPort.train  PPOCLIP_Agent.train                           xuance/torch/agents/policy_gradient/ppoclip_agent.py:59-111
    PGAgentPort.train   PG_Agent.train + PG_Learner.update            .../pg_agent.py:49-96, learners/policy_gradient/pg_learner.py:17-45
    PPGAgentPort.train  PPG_Agent.train                               .../ppg_agent.py:55-109

PINNING: tests/test_oracle_vs_reference.py runs these against the live reference (through
oracle/ref_loader.py) when /root/reference is present, and against tests/golden/*.npz (generated from the
live reference by oracle/make_goldens.py) everywhere.  The physics underneath (oracle/gym_restated.py) is
"parity unpinned" — see its header.
"""
import numpy as np
import torch

from . import gym_restated


# ------------------------------------------------------------------------------------------------ vec env
class MountainCarStackPort:
    """`MountainCar(Gym_Env)` restated (xuance/environment/gym/gym_env.py:50-83): a 4-frame stack over gym's env; the
    observation is `LazyFrames(list(frames))` -> np.concatenate(frames, axis=-1) (:227-254), shape (8,)."""

    def __init__(self, env):
        from collections import deque
        self.env, self._max_episode_steps, self.action_space = env, env._max_episode_steps, env.action_space
        self.frames = deque([], maxlen=4)
        lo, hi = np.array([-1.2, -0.07] * 4), np.array([0.6, 0.07] * 4)
        self.observation_space = type(env.observation_space)(lo.astype(np.float32), hi.astype(np.float32))

    def reset(self, **kw):
        obs, info = self.env.reset(**kw)
        for _ in range(4):
            self.frames.append(obs)
        return np.concatenate(list(self.frames), axis=-1), info

    def step(self, action):
        obs, rew, term, trunc, info = self.env.step(action)
        self.frames.append(obs)
        return np.concatenate(list(self.frames), axis=-1), rew, term, trunc, info


class VecEnvPort:
    """Serial vector env with xuance's auto-reset protocol (terminal obs in buf_obs, reset obs in infos)."""

    def __init__(self, env_id, num_envs, seed=1, trig="libm"):
        self.num_envs = num_envs
        self.envs = []
        for _ in range(num_envs):
            env = gym_restated.make(env_id.replace("gym:", ""), trig=trig)
            env.reset(seed=seed)                          # gym_env.py:19 (every env gets the same seed)
            if "MountainCar" in env_id and not env_id.startswith("gym:"):
                env = MountainCarStackPort(env)           # make_envs: environment/__init__.py:66-67
            self.envs.append(env)
        self.max_episode_length = self.envs[0]._max_episode_steps
        self.observation_space = self.envs[0].observation_space
        self.action_space = self.envs[0].action_space
        odim = self.observation_space.shape[0]
        self.buf_obs = np.zeros((num_envs, odim), np.float32)
        self.buf_rews = np.zeros(num_envs, np.float32)
        self.buf_dones = np.zeros(num_envs, bool)
        self.buf_trunctions = np.zeros(num_envs, bool)
        self.buf_infos = [{} for _ in range(num_envs)]
        self._ep_step = [0] * num_envs
        self._ep_score = [0.0] * num_envs

    def _reset_one(self, e):
        obs, info = self.envs[e].reset()
        self._ep_step[e], self._ep_score[e] = 0, 0.0
        info["episode_step"] = 0
        return obs, info

    def reset(self):
        for e in range(self.num_envs):
            self.buf_obs[e], self.buf_infos[e] = self._reset_one(e)
        return self.buf_obs.copy(), self.buf_infos.copy()

    def step(self, actions):
        for e in range(self.num_envs):
            obs, rew, term, trunc, info = self.envs[e].step(actions[e])
            self._ep_step[e] += 1
            self._ep_score[e] += rew
            info["episode_step"], info["episode_score"] = self._ep_step[e], self._ep_score[e]
            self.buf_rews[e], self.buf_dones[e], self.buf_trunctions[e], self.buf_infos[e] = rew, term, trunc, info
            if term or trunc:
                info["reset_obs"], _ = self._reset_one(e)
            self.buf_obs[e] = obs
        return (self.buf_obs.copy(), self.buf_rews.copy(), self.buf_dones.copy(), self.buf_trunctions.copy(),
                self.buf_infos.copy())


# ------------------------------------------------------------------------------------------------ buffer
class OnPolicyBufferPort:
    """Env-major [n_envs, n_size] float32 host buffer; GAE by a per-env reverse loop.

    Precision: the recurrence is carried in float64 and rounded once on store, which is what the pinned
    numpy 1.21.6 does (np.float32 * python-float promotes to float64, memory_tools.py:218-221); under
    numpy >= 2 the live reference mixes float32/float64.  Both lie within the 1e-5
    GAE tolerance of each other.
    """

    def __init__(self, obs_shape, act_shape, n_envs, n_size, use_gae=True, use_advnorm=True, gamma=0.99, gae_lam=0.95):
        self.obs_shape, self.act_shape = tuple(obs_shape), tuple(act_shape)
        self.n_envs, self.n_size = n_envs, n_size
        self.buffer_size = n_envs * n_size
        self.use_gae, self.use_advnorm, self.gamma, self.gae_lam = use_gae, use_advnorm, gamma, gae_lam
        self.start_ids = np.zeros(n_envs, np.int64)
        self.clear()

    def _zeros(self, shape=()):
        return np.zeros((self.n_envs, self.n_size) + tuple(shape), np.float32)

    def clear(self):
        self.ptr, self.size = 0, 0
        self.observations, self.actions = self._zeros(self.obs_shape), self._zeros(self.act_shape)
        self.rewards, self.returns, self.values = self._zeros(), self._zeros(), self._zeros()
        self.terminals, self.advantages = self._zeros(), self._zeros()
        self.auxiliary_infos = {"old_logp": self._zeros()}

    @property
    def full(self):
        return self.size >= self.n_size

    def store(self, obs, acts, rews, value, terminals, aux_info=None):
        p = self.ptr
        self.observations[:, p], self.actions[:, p] = obs, acts
        self.rewards[:, p], self.values[:, p], self.terminals[:, p] = rews, value, terminals
        if aux_info is not None:
            for k, v in aux_info.items():
                self.auxiliary_infos[k][:, p] = v
        self.ptr = (p + 1) % self.n_size
        self.size = min(self.size + 1, self.n_size)

    def finish_path(self, val, i):
        lo, hi = int(self.start_ids[i]), (self.n_size if self.full else self.ptr)
        r = self.rewards[i, lo:hi].astype(np.float64)
        v = np.append(self.values[i, lo:hi].astype(np.float64), float(val))
        if self.use_gae:
            d = self.terminals[i, lo:hi].astype(np.float64)
            adv = np.zeros(hi - lo, np.float64)
            acc = 0.0
            for t in range(hi - lo - 1, -1, -1):
                delta = r[t] + (1 - d[t]) * self.gamma * v[t + 1] - v[t]
                acc = delta + (1 - d[t]) * self.gamma * self.gae_lam * acc
                adv[t] = acc
            ret = adv + v[:-1]
        else:
            ret = np.zeros(hi - lo, np.float64)
            run = float(val)
            for t in range(hi - lo - 1, -1, -1):          # == discount_cumsum(append(r,[val]))[:-1]
                run = r[t] + self.gamma * run
                ret[t] = run
            adv = r + self.gamma * v[1:] - v[:-1]
        self.returns[i, lo:hi] = ret
        self.advantages[i, lo:hi] = adv
        self.start_ids[i] = self.ptr

    def sample(self, indexes):
        assert self.full, "Not enough transitions for on-policy buffer to random sample"
        env, step = np.divmod(np.asarray(indexes), self.n_size)
        adv = self.advantages[env, step]
        if self.use_advnorm:
            adv = (adv - np.mean(adv)) / (np.std(adv) + 1e-8)
        return (self.observations[env, step], self.actions[env, step], self.returns[env, step],
                self.values[env, step], adv, {k: a[env, step] for k, a in self.auxiliary_infos.items()})


class _OldDistRows(dict):
    """`auxiliary_infos` of OldDistBufferPort: assigning "old_dist" a distribution wrapper over the flattened env-major
    buffer (ppg_agent.py:93) stores its parameter rows."""

    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def __setitem__(self, key, value):
        if key == "old_dist" and hasattr(value, "get_param"):
            o = self._owner
            value = o._rows(value, o.buffer_size).reshape(o.n_envs, o.n_size, -1)
        dict.__setitem__(self, key, value)


class OldDistBufferPort(OnPolicyBufferPort):
    """The {"old_dist": None} buffer of PPG / PPO-KL (memory_tools.py:28-30 keeps one Python distribution object per
    transition; here their parameters: logits [A], or mean | std [2A]).  `sample` hands the parameters back in the form
    `ppg_update` / `ppokl_update` take as `old`."""

    def clear(self):
        super().clear()
        self.auxiliary_infos = _OldDistRows(self)

    @staticmethod
    def _rows(dist, rows):
        prm = dist.get_param()
        if isinstance(prm, (tuple, list)):
            mu = prm[0].detach().cpu().numpy().reshape(rows, -1)
            std = np.broadcast_to(prm[1].detach().cpu().numpy().reshape(-1, mu.shape[1]), mu.shape)
            return np.concatenate([mu, std], axis=1).astype(np.float32)
        return prm.detach().cpu().numpy().reshape(rows, -1).astype(np.float32)

    def store(self, obs, acts, rews, value, terminals, aux_info=None):
        rows = self._rows(aux_info["old_dist"], self.n_envs)
        if "old_dist" not in self.auxiliary_infos:
            dict.__setitem__(self.auxiliary_infos, "old_dist", self._zeros((rows.shape[1],)))
            self._gaussian = isinstance(aux_info["old_dist"].get_param(), (tuple, list))
        p = self.ptr
        super().store(obs, acts, rews, value, terminals, None)
        self.auxiliary_infos["old_dist"][:, p] = rows

    def sample(self, indexes):
        obs, act, ret, val, adv, aux = super().sample(indexes)
        w = torch.as_tensor(aux["old_dist"])
        if self._gaussian:
            A = w.shape[1] // 2
            w = (w[:, :A], w[:, A:])
        return obs, act, ret, val, adv, {"old_dist": w}


# ------------------------------------------------------------------------------------------------ learner
def policy_logp_entropy(policy, obs, act):
    """Runs a reference-shaped actor-critic module: returns (log_prob, entropy, v_pred)."""
    _, dist, v = policy(obs)
    return dist.log_prob(act), dist.entropy(), v


def ppo_clip_loss(logp, entropy, v_pred, ret, adv, old_logp, vf_coef, ent_coef, clip_range):
    ratio = (logp - old_logp).exp().float()
    unclipped = adv * ratio
    clipped = ratio.clamp(1.0 - clip_range, 1.0 + clip_range) * adv
    a_loss = -torch.minimum(clipped, unclipped).mean()
    c_loss = torch.nn.functional.mse_loss(v_pred, ret)
    e_loss = entropy.mean()
    return a_loss - ent_coef * e_loss + vf_coef * c_loss, a_loss, c_loss, e_loss, ratio


def ppo_clip_update(policy, optimizer, scheduler, batch, vf_coef=0.25, ent_coef=0.005, clip_range=0.25,
                    clip_grad_norm=0.25, use_grad_clip=True, device="cpu"):
    """One PPOCLIP_Learner.update step.  `batch` = (obs, act, ret, value, adv, old_logp) as sample() returns."""
    obs, act, ret, _value_unused, adv, old_logp = batch
    act, ret, adv, old_logp = (torch.as_tensor(a, device=device) for a in (act, ret, adv, old_logp))
    logp, ent, v_pred = policy_logp_entropy(policy, obs, act)
    loss, a_loss, c_loss, e_loss, ratio = ppo_clip_loss(logp, ent, v_pred, ret, adv, old_logp, vf_coef, ent_coef, clip_range)
    optimizer.zero_grad()
    loss.backward()
    if use_grad_clip:
        torch.nn.utils.clip_grad_norm_(policy.parameters(), clip_grad_norm)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    n_clipped = (ratio < 1 - clip_range).sum() + (ratio > 1 + clip_range).sum()
    return {"actor-loss": a_loss.item(), "critic-loss": c_loss.item(), "entropy": e_loss.item(),
            "learning_rate": optimizer.state_dict()["param_groups"][0]["lr"],
            "predict_value": v_pred.mean().item(), "clip_ratio": n_clipped / ratio.shape[0]}


def _kl(new, old):
    """kl_divergence(new, old) of two reference-shaped wrappers (distributions.py:63-66,97-100) from their parameters."""
    pn, po = new.get_param(), old
    if isinstance(pn, (tuple, list)):
        mu, std = pn
        omu, ostd = po
        var_ratio = (std / ostd) ** 2
        t1 = ((mu - omu) / ostd) ** 2
        return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())            # element-wise [B, A]
    lp = pn - pn.logsumexp(-1, keepdim=True)
    lq = po - po.logsumexp(-1, keepdim=True)
    return (lp.exp() * (lp - lq)).sum(-1)


def _old_logp(old, act):
    if isinstance(old, (tuple, list)):
        omu, ostd = old
        return (-((act - omu) ** 2) / (2 * ostd ** 2) - ostd.log() - 0.5 * np.log(2 * np.pi)).sum(-1)
    lq = old - old.logsumexp(-1, keepdim=True)
    return lq.gather(-1, act.long().unsqueeze(-1)).squeeze(-1)


def ppokl_update(policy, optimizer, scheduler, batch, old, state, vf_coef=0.25, ent_coef=0.005, target_kl=0.25):
    """One PPOKL_Learner.update step (ppokl_learner.py:21-61).  `old` = old logits, or (old_mu, old_std);
    `state` = {"kl_coef": float} carried between updates."""
    obs, act, ret, adv = batch
    act, ret, adv = (torch.as_tensor(a) for a in (act, ret, adv))
    old = tuple(torch.as_tensor(o) for o in old) if isinstance(old, (tuple, list)) else torch.as_tensor(old)
    _, dist, v_pred = policy(obs)
    logp = dist.log_prob(act)
    kl = _kl(dist, old).mean()
    ratio = (logp - _old_logp(old, act)).exp().float()
    a_loss = -(ratio * adv).mean() + state["kl_coef"] * kl
    c_loss = torch.nn.functional.mse_loss(v_pred, ret)
    e_loss = dist.entropy().mean()
    loss = a_loss - ent_coef * e_loss + vf_coef * c_loss
    if kl > target_kl * 1.5:
        state["kl_coef"] = state["kl_coef"] * 2.0
    elif kl < target_kl * 0.5:
        state["kl_coef"] = state["kl_coef"] / 2.0
    state["kl_coef"] = float(np.clip(state["kl_coef"], 0.1, 20))
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return {"actor-loss": a_loss.item(), "critic-loss": c_loss.item(), "entropy": e_loss.item(),
            "learning_rate": optimizer.state_dict()["param_groups"][0]["lr"], "kl": kl.item(),
            "predict_value": v_pred.mean().item()}


def ppg_update(phase, policy, optimizer, scheduler, batch, old, ent_coef=0.005, clip_range=0.25, kl_beta=1.0):
    """PPG_Learner.update_policy / update_critic / update_auxiliary (ppg_learner.py:23-88)."""
    obs, act, ret, adv = batch
    act, ret, adv = (torch.as_tensor(a) for a in (act, ret, adv))
    old = tuple(torch.as_tensor(o) for o in old) if isinstance(old, (tuple, list)) else torch.as_tensor(old)
    _, dist, v, aux_v = policy(obs)
    mse = torch.nn.functional.mse_loss
    if phase == "policy":
        ratio = (dist.log_prob(act) - _old_logp(old, act)).exp().float()
        a_loss = -torch.minimum(ratio.clamp(1.0 - clip_range, 1.0 + clip_range) * adv, adv * ratio).mean()
        e_loss = dist.entropy().mean()
        loss = a_loss - ent_coef * e_loss
    elif phase == "critic":
        loss = mse(v, ret)
    else:
        loss = mse(v.detach(), aux_v) + kl_beta * _kl(dist, old).mean() + mse(v, ret)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    if phase == "policy":
        if scheduler is not None:
            scheduler.step()
        cr = ((ratio < 1 - clip_range).sum() + (ratio > 1 + clip_range).sum()) / ratio.shape[0]
        return {"actor-loss": a_loss.item(), "entropy": e_loss.item(),
                "learning_rate": optimizer.state_dict()["param_groups"][0]["lr"], "clip_ratio": cr}
    return {"critic-loss": loss.item()} if phase == "critic" else {"kl-loss": loss.item()}


# ------------------------------------------------------------------------------------------------ normaliser
class RunningMeanStdPort:
    def __init__(self, shape, epsilon=1e-4):
        self.mean, self.var, self.count = np.zeros(shape, np.float32), np.ones(shape, np.float32), epsilon

    @property
    def std(self):
        return np.sqrt(self.var)

    def update(self, x):
        b_mean, b_var, b_n = np.mean(x, axis=0), np.square(np.std(x, axis=0)), x.shape[0]
        delta, tot = b_mean - self.mean, self.count + b_n
        m2 = self.var * self.count + b_var * b_n + np.square(delta) * self.count * b_n / tot
        self.mean, self.var, self.count = self.mean + delta * b_n / tot, m2 / tot, tot


# ------------------------------------------------------------------------------------------------ agent loop
class PPOAgentPort:
    """PPOCLIP_Agent.train restated (logging/Atari branches removed)."""

    def __init__(self, envs, policy, optimizer, scheduler, n_steps, n_epoch, n_minibatch, gamma, gae_lam,
                 vf_coef=0.25, ent_coef=0.01, clip_range=0.2, clip_grad_norm=0.5, use_grad_clip=True,
                 use_gae=True, use_advnorm=True, use_obsnorm=False, use_rewnorm=False, obsnorm_range=5, rewnorm_range=5,
                 memory=None, update_fn=None, action_tape=None, perm_tape=None):
        # action_tape [steps, N(, A)] / perm_tape [n_shuffles, buffer_size]: replay the random draws of a recorded run
        # (tests/golden/agent_ppo_*.npz from the live reference) instead of sampling / shuffling — everything else is
        # computed here, so the comparison does not hinge on the host's RNG stream or last-bit GEMM differences
        self.action_tape, self.perm_tape, self._tape_pos, self._perm_pos = action_tape, perm_tape, 0, 0
        self.envs, self.policy, self.optimizer, self.scheduler = envs, policy, optimizer, scheduler
        self.n_envs, self.n_steps, self.n_epoch = envs.num_envs, n_steps, n_epoch
        self.buffer_size = self.n_envs * n_steps
        self.batch_size = self.buffer_size // n_minibatch
        self.gamma = gamma
        act_shape = () if not hasattr(envs.action_space, "low") else envs.action_space.shape
        # `memory` / `update_fn` let a test drive OTHER implementations of the buffer / learner (the product's
        # drop-ins) through this restated agent loop, the way the unmodified PPOCLIP_Agent would drive them
        self.memory = memory if memory is not None else OnPolicyBufferPort(
            envs.observation_space.shape, act_shape, self.n_envs, n_steps, use_gae, use_advnorm, gamma, gae_lam)
        self.update_fn = update_fn
        self.hp = dict(vf_coef=vf_coef, ent_coef=ent_coef, clip_range=clip_range, clip_grad_norm=clip_grad_norm,
                       use_grad_clip=use_grad_clip)
        self.use_obsnorm, self.use_rewnorm = use_obsnorm, use_rewnorm
        self.obsnorm_range, self.rewnorm_range = obsnorm_range, rewnorm_range
        self.obs_rms = RunningMeanStdPort(envs.observation_space.shape)
        self.ret_rms = RunningMeanStdPort(())
        self.returns = np.zeros(self.n_envs, np.float32)
        self.current_step, self.episodes, self.last_info = 0, 0, {}

    def _obs(self, o):
        if not self.use_obsnorm:
            return o
        return np.clip((o - self.obs_rms.mean) / (self.obs_rms.std + 1e-8), -self.obsnorm_range, self.obsnorm_range)

    def _rew(self, r):
        if not self.use_rewnorm:
            return r
        return np.clip(r / np.clip(self.ret_rms.std, 0.1, 100), -self.rewnorm_range, self.rewnorm_range)

    def _action(self, obs, taped=None):
        _, dist, v = self.policy(obs)
        if taped is None:
            a = dist.stochastic_sample()
        else:
            a = torch.as_tensor(taped, device=v.device)
        lp = dist.log_prob(a)
        return a.detach().cpu().numpy(), v.detach().cpu().numpy(), lp.detach().cpu().numpy()

    def _shuffle(self, indexes):
        if self.perm_tape is None:
            np.random.shuffle(indexes)
        else:
            indexes[:] = self.perm_tape[self._perm_pos]
            self._perm_pos += 1

    def train(self, train_steps):
        obs = self.envs.buf_obs
        mem = self.memory
        for _ in range(train_steps):
            self.obs_rms.update(obs)
            obs = self._obs(obs)
            taped = None
            if self.action_tape is not None:
                taped = self.action_tape[self._tape_pos]
                self._tape_pos += 1
            acts, value, logps = self._action(obs, taped)
            next_obs, rewards, terminals, truncations, infos = self.envs.step(acts)
            mem.store(obs, acts, self._rew(rewards), value, terminals, {"old_logp": logps})
            if mem.full:
                _, vals, _ = self._action(self._obs(next_obs))
                for i in range(self.n_envs):
                    mem.finish_path(0.0 if terminals[i] else vals[i], i)
                indexes = np.arange(self.buffer_size)
                for _ in range(self.n_epoch):
                    self._shuffle(indexes)
                    for start in range(0, self.buffer_size, self.batch_size):
                        batch = mem.sample(indexes[start:start + self.batch_size])
                        if self.update_fn is not None:
                            self.last_info = self.update_fn(*batch[:5], batch[5]["old_logp"])
                        else:
                            self.last_info = ppo_clip_update(self.policy, self.optimizer, self.scheduler,
                                                             batch[:5] + (batch[5]["old_logp"],), **self.hp)
                mem.clear()
            self.returns = (1 - terminals) * self.gamma * self.returns + rewards
            obs = next_obs
            for i in range(self.n_envs):
                if terminals[i] or truncations[i]:
                    self.ret_rms.update(self.returns[i:i + 1])
                    self.returns[i] = 0.0
                    if terminals[i]:
                        mem.finish_path(0.0, i)
                    else:
                        _, vals, _ = self._action(self._obs(next_obs))
                        mem.finish_path(vals[i], i)
                    obs[i] = infos[i]["reset_obs"]
                    self.episodes += 1
            self.current_step += self.n_envs


def pg_update(policy, optimizer, scheduler, batch, ent_coef=0.005, clip_grad=0.5):
    """One PG_Learner.update step (pg_learner.py:17-45): a_loss = -(returns * log_prob).mean(), entropy bonus, always clipped."""
    obs, act, ret = batch
    dev = next(policy.parameters()).device
    act, ret = torch.as_tensor(act, device=dev), torch.as_tensor(ret, device=dev)
    _, dist = policy(obs)
    logp = dist.log_prob(act)
    a_loss = -(ret * logp).mean()
    e_loss = dist.entropy().mean()
    loss = a_loss - ent_coef * e_loss
    optimizer.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(policy.parameters(), clip_grad)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return {"actor-loss": a_loss.item(), "entropy": e_loss.item(),
            "learning_rate": optimizer.state_dict()["param_groups"][0]["lr"]}


class _TapeMixin:
    """Replay (or record) the random draws of a run: actions per vector step, minibatch permutations per shuffle."""

    def _init_tapes(self, action_tape, perm_tape, record):
        self.action_tape, self.perm_tape, self._tape_pos, self._perm_pos = action_tape, perm_tape, 0, 0
        self.recorded_actions, self.recorded_perms = ([], []) if record else (None, None)

    def _taped_action(self, dist):
        if self.action_tape is not None:
            a = torch.as_tensor(self.action_tape[self._tape_pos], device=self._device())
            self._tape_pos += 1
        else:
            a = dist.stochastic_sample()
        if self.recorded_actions is not None:
            self.recorded_actions.append(a.detach().cpu().numpy().copy())
        return a

    def _shuffle(self, indexes):
        if self.perm_tape is None:
            np.random.shuffle(indexes)
        else:
            indexes[:] = self.perm_tape[self._perm_pos]
            self._perm_pos += 1
        if self.recorded_perms is not None:
            self.recorded_perms.append(indexes.copy())

    def _device(self):
        return next(self.policy.parameters()).device


class PGAgentPort(_TapeMixin):
    """PG_Agent.train restated (xuance/torch/agents/policy_gradient/pg_agent.py:49-96; logging removed): actor-only policy,
    value 0 stored, `finish_path(processed_reward[i], i)` for every env when the buffer is full (:60-62), `finish_path(0, i)` at
    every episode end (:81), minibatch size = buffer_size // n_epoch (:28), the return tracker does not mask terminals (:74).
    `memory` / `update_fn` let a test drive the product's drop-in buffer / PG_Learner through this loop."""

    def __init__(self, envs, policy, optimizer, scheduler, n_steps, n_epoch, gamma, gae_lam, ent_coef=0.01, clip_grad=0.5,
                 use_gae=False, use_advnorm=False, use_obsnorm=True, use_rewnorm=True, obsnorm_range=5, rewnorm_range=5,
                 memory=None, update_fn=None, action_tape=None, perm_tape=None, record=False):
        self.envs, self.policy, self.optimizer, self.scheduler = envs, policy, optimizer, scheduler
        self.n_envs, self.n_steps, self.n_epoch, self.gamma = envs.num_envs, n_steps, n_epoch, gamma
        self.buffer_size = self.n_envs * n_steps
        self.batch_size = self.buffer_size // n_epoch
        act_shape = () if not hasattr(envs.action_space, "low") else envs.action_space.shape
        self.memory = memory if memory is not None else OnPolicyBufferPort(
            envs.observation_space.shape, act_shape, self.n_envs, n_steps, use_gae, use_advnorm, gamma, gae_lam)
        self.update_fn = update_fn
        self.hp = dict(ent_coef=ent_coef, clip_grad=clip_grad)
        self.use_obsnorm, self.use_rewnorm = use_obsnorm, use_rewnorm
        self.obsnorm_range, self.rewnorm_range = obsnorm_range, rewnorm_range
        self.obs_rms, self.ret_rms = RunningMeanStdPort(envs.observation_space.shape), RunningMeanStdPort(())
        self.returns = np.zeros(self.n_envs, np.float32)
        self.current_step, self.episodes, self.last_info = 0, 0, {}
        self._init_tapes(action_tape, perm_tape, record)

    _obs = PPOAgentPort._obs
    _rew = PPOAgentPort._rew

    def train(self, train_steps):
        obs = self.envs.buf_obs
        mem = self.memory
        for _ in range(train_steps):
            self.obs_rms.update(obs)
            obs = self._obs(obs)
            _, dist = self.policy(obs)
            acts = self._taped_action(dist).detach().cpu().numpy()
            next_obs, rewards, terminals, truncations, infos = self.envs.step(acts)
            mem.store(obs, acts, self._rew(rewards), 0, terminals)
            if mem.full:
                proc = self._rew(rewards)
                for i in range(self.n_envs):
                    mem.finish_path(proc[i], i)
                indexes = np.arange(self.buffer_size)
                for _ in range(self.n_epoch):
                    self._shuffle(indexes)
                    for start in range(0, self.buffer_size, self.batch_size):
                        obs_b, act_b, ret_b, _, _, _ = mem.sample(indexes[start:start + self.batch_size])
                        if self.update_fn is not None:
                            self.last_info = self.update_fn(obs_b, act_b, ret_b)
                        else:
                            self.last_info = pg_update(self.policy, self.optimizer, self.scheduler, (obs_b, act_b, ret_b), **self.hp)
                mem.clear()
            self.returns = self.gamma * self.returns + rewards
            obs = next_obs
            for i in range(self.n_envs):
                if terminals[i] or truncations[i]:
                    self.ret_rms.update(self.returns[i:i + 1])
                    self.returns[i] = 0.0
                    obs[i] = infos[i]["reset_obs"]
                    mem.finish_path(0, i)
                    self.episodes += 1
            self.current_step += self.n_envs


class PPGAgentPort(_TapeMixin):
    """PPG_Agent.train restated (xuance/torch/agents/policy_gradient/ppg_agent.py:55-109; logging removed): rollout with
    the old action distributions stored per transition, then the policy phase, the critic phase, the refresh of every
    stored old distribution from the current policy (:90-93) and the auxiliary phase.  `memory` and the three `update_*`
    callables let a test drive OTHER implementations (the product's drop-in buffer and PPG_Learner) through this loop.
    The reference wraps the distributions with split_distributions (one Python object per sample); the batched wrapper is
    passed as is here — the drop-in buffer accepts both.  Observation / reward normalisation as in the reference loop
    (:58-59,63): obs_rms is updated every step; ret_rms is never updated by PPG_Agent, so `_process_reward` only clips."""

    def __init__(self, envs, policy, memory, update_policy, update_critic, update_auxiliary, n_steps, n_minibatch=4,
                 policy_nepoch=1, value_nepoch=1, aux_nepoch=1, use_obsnorm=False, use_rewnorm=False, obsnorm_range=5,
                 rewnorm_range=5, action_tape=None, perm_tape=None, record=False):
        self.envs, self.policy, self.memory = envs, policy, memory
        self.update_policy, self.update_critic, self.update_auxiliary = update_policy, update_critic, update_auxiliary
        self.n_envs, self.n_steps = envs.num_envs, n_steps
        self.buffer_size = self.n_envs * n_steps
        self.batch_size = self.buffer_size // n_minibatch
        self.policy_nepoch, self.value_nepoch, self.aux_nepoch = policy_nepoch, value_nepoch, aux_nepoch
        self.use_obsnorm, self.use_rewnorm = use_obsnorm, use_rewnorm
        self.obsnorm_range, self.rewnorm_range = obsnorm_range, rewnorm_range
        self.obs_rms, self.ret_rms = RunningMeanStdPort(envs.observation_space.shape), RunningMeanStdPort(())
        self.current_step, self.episodes, self.infos = 0, 0, {}
        self._init_tapes(action_tape, perm_tape, record)

    _obs = PPOAgentPort._obs
    _rew = PPOAgentPort._rew

    def _action(self, obs, sample=True):
        _, dists, vs, _ = self.policy(obs)
        acts = self._taped_action(dists) if sample else dists.stochastic_sample()
        return acts.detach().cpu().numpy(), vs.detach().cpu().numpy(), dists

    def _phase(self, n_epoch, update, indexes):
        for _ in range(n_epoch):
            self._shuffle(indexes)
            for start in range(0, self.buffer_size, self.batch_size):
                obs_b, act_b, ret_b, _, adv_b, aux_b = self.memory.sample(indexes[start:start + self.batch_size])
                self.infos.update(update(obs_b, act_b, ret_b, adv_b, aux_b["old_dist"]))

    def train(self, train_steps):
        obs = self.envs.buf_obs
        mem = self.memory
        for _ in range(train_steps):
            self.obs_rms.update(obs)
            obs = self._obs(obs)
            acts, rets, dists = self._action(obs)
            next_obs, rewards, terminals, truncations, infos = self.envs.step(acts)
            mem.store(obs, acts, self._rew(rewards), rets, terminals, {"old_dist": dists})
            if mem.full:
                _, vals, _ = self._action(self._obs(next_obs), sample=False)
                for i in range(self.n_envs):
                    mem.finish_path(vals[i], i)
                indexes = np.arange(self.buffer_size)
                self._phase(self.policy_nepoch, self.update_policy, indexes)
                self._phase(self.value_nepoch, self.update_critic, indexes)
                buffer_obs = mem.observations                                     # [n_envs, n_size, obs_dim]
                if torch.is_tensor(buffer_obs):
                    buffer_obs = buffer_obs.reshape(self.buffer_size, -1)
                else:
                    buffer_obs = np.asarray(buffer_obs).reshape(self.buffer_size, -1)
                _, new_dist, _, _ = self.policy(buffer_obs)
                mem.auxiliary_infos["old_dist"] = new_dist                         # ppg_agent.py:93
                self._phase(self.aux_nepoch, self.update_auxiliary, indexes)
                mem.clear()
            obs = next_obs
            for i in range(self.n_envs):
                if terminals[i] or truncations[i]:
                    obs[i] = infos[i]["reset_obs"]
                    mem.finish_path(0, i)
                    self.episodes += 1
            self.current_step += self.n_envs
