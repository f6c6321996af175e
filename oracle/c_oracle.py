"""ctypes front-end of oracle/classic_control.c (ORACLE / TEST INFRASTRUCTURE — see that file's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
FLAVOURS = {"cr": 0, "libm": 1}


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "classic_control.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pcg64_state(seed):
    """[state_hi, state_lo, inc_hi, inc_lo] of numpy's PCG64(SeedSequence(seed)) (computed by numpy itself)."""
    st = np.random.PCG64(np.random.SeedSequence(seed)).state["state"]
    m = (1 << 64) - 1
    return np.array([st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)


def sincos(x, flavour="cr"):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = np.empty_like(x)
    c = np.empty_like(x)
    lib().oc_sincos(_p(x), _p(s), _p(c), C.c_long(x.size), C.c_int(FLAVOURS[flavour]))
    return s, c


def pcg64_uniform(rng, low, high, n):
    out = np.empty(n, dtype=np.float64)
    lib().oc_pcg64_uniform(_p(rng), C.c_double(low), C.c_double(high - low), _p(out), C.c_long(n))
    return out


class VecEnvC:
    """N independent copies of one classic-control env, stepped by the C oracle (AoS fp64 state)."""

    SPEC = {"CartPole-v1": dict(sdim=4, odim=4, max_steps=500, act=np.int64),
            "Pendulum-v1": dict(sdim=2, odim=3, max_steps=200, act=np.float32),
            "gym:MountainCar-v0": dict(sdim=2, odim=2, max_steps=200, act=np.int64),      # gym's bare env
            "Acrobot-v1": dict(sdim=4, odim=6, max_steps=500, act=np.int64)}

    def __new__(cls, env_id, *a, **k):
        # "MountainCar-v0" = what the reference's make_envs builds: the 4-frame wrapper (gym_env.py:50-83)
        if env_id == "MountainCar-v0" and cls is VecEnvC:
            return object.__new__(MountainCarStackC)
        return object.__new__(cls)

    def __init__(self, env_id, n, seed=1, flavour="cr", n_warm_resets=2, seeds=None):
        sp = self.SPEC[env_id]
        self.env_id, self.n, self.flavour = env_id, n, FLAVOURS[flavour]
        self.sdim, self.odim, self.max_steps, self.act_dtype = sp["sdim"], sp["odim"], sp["max_steps"], sp["act"]
        if seeds is None:
            self.rng = np.tile(pcg64_state(seed), (n, 1))
        else:
            self.rng = np.stack([pcg64_state(int(s)) for s in seeds])
        self.rng = np.ascontiguousarray(self.rng, dtype=np.uint64)
        self.state = np.zeros((n, self.sdim), np.float64)
        self.elapsed = np.zeros(n, np.int32)
        self.ep_score = np.zeros(n, np.float64)
        self.obs = np.zeros((n, self.odim), np.float32)
        self.reset_all(n_warm_resets)

    def reset_all(self, n_draws=1):
        L = lib()
        if self.env_id == "CartPole-v1":
            L.oc_cartpole_reset(_p(self.state), _p(self.rng), _p(self.elapsed), _p(self.ep_score), _p(self.obs),
                                C.c_int(n_draws), C.c_long(self.n))
        elif self.env_id == "gym:MountainCar-v0":
            L.oc_mountaincar_reset(_p(self.state), _p(self.rng), _p(self.elapsed), _p(self.ep_score), _p(self.obs),
                                   C.c_int(n_draws), C.c_long(self.n))
        elif self.env_id == "Acrobot-v1":
            L.oc_acrobot_reset(_p(self.state), _p(self.rng), _p(self.elapsed), _p(self.ep_score), _p(self.obs),
                               C.c_int(n_draws), C.c_long(self.n), C.c_int(self.flavour))
        else:
            L.oc_pendulum_reset(_p(self.state), _p(self.rng), _p(self.elapsed), _p(self.ep_score), _p(self.obs),
                                C.c_int(n_draws), C.c_long(self.n), C.c_int(self.flavour))
        return self.obs.copy()

    def step(self, actions):
        n = self.n
        a = np.ascontiguousarray(np.asarray(actions).reshape(n), dtype=self.act_dtype)
        rew = np.empty(n, np.float32)
        term = np.empty(n, np.uint8)
        trunc = np.empty(n, np.uint8)
        reset_obs = np.zeros((n, self.odim), np.float32)
        ep_step = np.empty(n, np.int32)
        ep_score = np.empty(n, np.float64)
        fn = {"CartPole-v1": lib().oc_cartpole_step, "Pendulum-v1": lib().oc_pendulum_step,
              "gym:MountainCar-v0": lib().oc_mountaincar_step, "Acrobot-v1": lib().oc_acrobot_step}[self.env_id]
        fn(_p(self.state), _p(self.rng), _p(self.elapsed), _p(self.ep_score), _p(a), _p(self.obs), _p(rew), _p(term),
           _p(trunc), _p(reset_obs), _p(ep_step), _p(ep_score), C.c_int(self.max_steps), C.c_long(n), C.c_int(self.flavour))
        return dict(obs=self.obs.copy(), rew=rew, term=term.astype(bool), trunc=trunc.astype(bool),
                    reset_obs=reset_obs, ep_step=ep_step, ep_score=ep_score, state=self.state.copy())


class MountainCarStackC(VecEnvC):
    """`MountainCar(Gym_Env)` of the reference (xuance/environment/gym/gym_env.py:50-83) over the C oracle's bare env, batched:
    observation = np.concatenate(last 4 frames, oldest first) (LazyFrames, :227-254); reset() fills all four frames with the
    reset observation (:67-68); step() appends the new frame (:82).  `state` = [position, velocity, frame t-3, t-2, t-1]
    (fp64 [n][8]) — the layout of the device kernel's state."""

    def __init__(self, env_id, n, seed=1, flavour="cr", n_warm_resets=2, seeds=None):
        self.raw = VecEnvC("gym:MountainCar-v0", n, seed, flavour, n_warm_resets, seeds)
        self.env_id, self.n, self.sdim, self.odim, self.max_steps = env_id, n, 8, 8, self.raw.max_steps
        self.act_dtype = np.int64
        self._refill()

    def _refill(self):
        self.frames = np.repeat(self.raw.obs[:, None, :], 4, axis=1).copy()           # [n][4][2] float32
        self.obs = self.frames.reshape(self.n, 8).copy()

    @property
    def state(self):
        return np.concatenate([self.raw.state, self.frames[:, :3].reshape(self.n, 6).astype(np.float64)], axis=1)

    @property
    def rng(self):
        return self.raw.rng

    def reset_all(self, n_draws=1):
        self.raw.reset_all(n_draws)
        self._refill()
        return self.obs.copy()

    def step(self, actions):
        o = self.raw.step(actions)
        self.frames = np.concatenate([self.frames[:, 1:], o["obs"][:, None, :]], axis=1)
        self.obs = self.frames.reshape(self.n, 8).copy()
        done = o["term"] | o["trunc"]
        reset_obs = np.zeros((self.n, 8), np.float32)
        if done.any():
            self.frames[done] = np.repeat(o["reset_obs"][done][:, None, :], 4, axis=1)
            reset_obs[done] = self.frames[done].reshape(-1, 8)
        out = dict(o)
        out.update(obs=self.obs.copy(), reset_obs=reset_obs, state=self.state)
        return out


def gae(rew, val, term, boot_last, gamma, lam, segend=None, boot=None, use_gae=True):
    """fp64 batched finish_path over time-major [T, N] fp32 arrays -> (adv, ret) fp64 [T, N]."""
    rew = np.ascontiguousarray(rew, np.float32)
    val = np.ascontiguousarray(val, np.float32)
    term = np.ascontiguousarray(term, np.float32)
    T, N = rew.shape
    boot_last = np.ascontiguousarray(boot_last, np.float32)
    if segend is not None:
        segend = np.ascontiguousarray(segend, np.uint8)
        boot = np.ascontiguousarray(boot, np.float32)
    adv = np.empty((T, N), np.float64)
    ret = np.empty((T, N), np.float64)
    lib().oc_gae(_p(rew), _p(val), _p(term), _p(segend), _p(boot), _p(boot_last), _p(adv), _p(ret),
                 C.c_long(T), C.c_long(N), C.c_double(gamma), C.c_double(lam), C.c_int(1 if use_gae else 0))
    return adv, ret
