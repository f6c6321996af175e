"""CPU restatement of gym 0.26.2 classic-control physics (CartPole-v1, Pendulum-v1, MountainCar-v0).

ORACLE / TEST INFRASTRUCTURE — never imported by the product (xuanpolicy_b200/).

gym is a pinned third-party dependency of the reference (`/root/reference/setup.py:51`
"gym==0.26.2") that is NOT vendored in /root/reference and not installed here, so this file
restates the published algorithm of
    gym/envs/classic_control/cartpole.py   (CartPoleEnv.step / reset)
    gym/envs/classic_control/pendulum.py   (PendulumEnv.step / reset / _get_obs / angle_normalize)
    gym/wrappers/time_limit.py             (TimeLimit.step / reset)
    gym/utils/seeding.py                   (np_random -> Generator(PCG64(SeedSequence(seed))))
anchored on the reference's own call sites: `gym.make` (xuance/environment/gym/gym_env.py:17),
`env.reset(seed=seed)` (:19), `env._max_episode_steps` (:26), `env.reset()` (:37), `env.step` (:44).

PARITY UNPINNED by the reference: it ships no golden vectors for this path, and real
gym cannot be run in this container.  The KATs in tests/golden/physics_kat.json were derived from this
restatement and cross-checked against the independent C restatement (oracle/classic_control.c).

Arithmetic contract:
  * IEEE fp64, round-to-nearest, one rounding per operation (CPython floats never contract to FMA).
  * dtype promotion follows the PINNED numpy 1.21.6 (setup.py:48): the float32 Pendulum action is
    widened to fp64 before `3.0*u`, `u**2`, `0.001*u**2`.  Casts are explicit so the result does not
    depend on the numpy that happens to be installed.
  * trig flavour "cr"   : correctly rounded sin/cos (mpmath, 200 bit) and squares as v*v
                          -> the platform-independent Tier-1 target the CUDA kernel must match bit-for-bit.
    trig flavour "libm" : math.sin/math.cos and `v ** 2.0` (libm pow) -> what gym executes on this host
                          (Tier 2, host dependent; reported, not required).
"""
import math

import numpy as np

_TWO_PI = 2.0 * math.pi

try:  # mpmath is only needed for the "cr" flavour
    import mpmath as _mp

    _mp_ctx = _mp.mp.clone()
    _mp_ctx.prec = 200
except Exception:  # pragma: no cover
    _mp_ctx = None


def _sin_cr(x):
    return float(_mp_ctx.sin(_mp_ctx.mpf(x)))


def _cos_cr(x):
    return float(_mp_ctx.cos(_mp_ctx.mpf(x)))


def _sq_exact(v):
    return v * v


def _sq_pow(v):
    return v ** 2.0  # float.__pow__ -> libm pow, as gym's `v**2` does


class _Math:
    def __init__(self, trig):
        if trig == "cr":
            if _mp_ctx is None:
                raise RuntimeError("mpmath is required for the correctly-rounded oracle flavour")
            self.sin, self.cos, self.sq = _sin_cr, _cos_cr, _sq_exact
        elif trig == "libm":
            self.sin, self.cos, self.sq = math.sin, math.cos, _sq_pow
        else:
            raise ValueError(trig)
        self.trig = trig


def new_rng(seed):
    """gym.utils.seeding.np_random(seed)[0]."""
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class _Spaces:
    """Built lazily so this module works with either the stub gym or the product's spaces."""

    @staticmethod
    def _mod():
        try:
            import gym.spaces as m          # the stub (or a real gym) when the reference is being driven
        except ImportError:
            import xuanpolicy_b200.spaces as m
        return m

    @staticmethod
    def box(low, high, shape=None):
        return _Spaces._mod().Box(low, high, shape=shape, dtype=np.float32)

    @staticmethod
    def discrete(n):
        return _Spaces._mod().Discrete(n)


class CartPoleRestated:
    """gym 0.26.2 CartPoleEnv, euler integrator."""
    gravity = 9.8
    masscart = 1.0
    masspole = 0.1
    total_mass = masspole + masscart
    length = 0.5
    polemass_length = masspole * length
    force_mag = 10.0
    tau = 0.02
    theta_threshold_radians = 12 * 2 * math.pi / 360
    x_threshold = 2.4
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 50}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, trig="libm", with_spaces=True):
        self.m = _Math(trig)
        self.np_random = None
        self.state = None
        if with_spaces:
            high = np.array([self.x_threshold * 2, np.finfo(np.float32).max,
                             self.theta_threshold_radians * 2, np.finfo(np.float32).max], dtype=np.float32)
            self.observation_space = _Spaces.box(-high, high)
            self.action_space = _Spaces.discrete(2)

    def reset(self, seed=None):
        if seed is not None or self.np_random is None:
            self.np_random = new_rng(seed)
        self.state = tuple(float(v) for v in self.np_random.uniform(low=-0.05, high=0.05, size=(4,)))
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        m = self.m
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        c = m.cos(theta)
        s = m.sin(theta)
        temp = (force + self.polemass_length * m.sq(theta_dot) * s) / self.total_mass
        thetaacc = (self.gravity * s - c * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * m.sq(c) / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * c / self.total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = (x, x_dot, theta, theta_dot)
        terminated = bool(x < -self.x_threshold or x > self.x_threshold
                          or theta < -self.theta_threshold_radians or theta > self.theta_threshold_radians)
        return np.array(self.state, dtype=np.float32), 1.0, terminated, False, {}

    def close(self):
        pass

    def render(self):
        return None


class PendulumRestated:
    """gym 0.26.2 PendulumEnv (g=10.0)."""
    max_speed = 8
    max_torque = 2.0
    dt = 0.05
    g = 10.0
    mass = 1.0
    l = 1.0
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, trig="libm", with_spaces=True):
        self.m = _Math(trig)
        self.np_random = None
        self.state = None
        if with_spaces:
            high = np.array([1.0, 1.0, self.max_speed], dtype=np.float32)
            self.observation_space = _Spaces.box(-high, high)
            self.action_space = _Spaces.box(-self.max_torque, self.max_torque, shape=(1,))

    def _obs(self):
        th, thdot = self.state
        return np.array([self.m.cos(th), self.m.sin(th), thdot], dtype=np.float32)

    def reset(self, seed=None):
        if seed is not None or self.np_random is None:
            self.np_random = new_rng(seed)
        high = np.array([math.pi, 1.0])
        th, thdot = self.np_random.uniform(low=-high, high=high)
        self.state = (float(th), float(thdot))
        return self._obs(), {}

    @staticmethod
    def _angle_normalize(x):
        # ((x + pi) % (2*pi)) - pi with Python/numpy floor-modulo semantics
        r = math.fmod(x + math.pi, _TWO_PI)
        if r != 0.0:
            if r < 0.0:
                r += _TWO_PI
        else:
            r = 0.0
        return r - math.pi

    def step(self, u):
        m = self.m
        th, thdot = self.state
        u32 = np.float32(np.asarray(u, dtype=np.float32).reshape(-1)[0])
        u32 = np.float32(min(max(u32, np.float32(-self.max_torque)), np.float32(self.max_torque)))
        uf = float(u32)  # pinned numpy 1.21.6: float32 scalar (x) python float -> float64
        costs = m.sq(self._angle_normalize(th)) + 0.1 * m.sq(thdot) + 0.001 * (uf * uf)
        newthdot = thdot + (15.0 * m.sin(th) + 3.0 * uf) * self.dt
        newthdot = min(max(newthdot, -8.0), 8.0)
        newth = th + newthdot * self.dt
        self.state = (newth, newthdot)
        return self._obs(), -costs, False, False, {}

    def close(self):
        pass

    def render(self):
        return None


class MountainCarRestated:
    """gym 0.26.2 MountainCarEnv (gym/envs/classic_control/mountain_car.py), goal_velocity = 0."""
    min_position = -1.2
    max_position = 0.6
    max_speed = 0.07
    goal_position = 0.5
    goal_velocity = 0
    force = 0.001
    gravity = 0.0025
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, trig="libm", with_spaces=True):
        self.m = _Math(trig)
        self.np_random = None
        self.state = None
        if with_spaces:
            low = np.array([self.min_position, -self.max_speed], dtype=np.float32)
            high = np.array([self.max_position, self.max_speed], dtype=np.float32)
            self.observation_space = _Spaces.box(low, high)
            self.action_space = _Spaces.discrete(3)

    def reset(self, seed=None):
        if seed is not None or self.np_random is None:
            self.np_random = new_rng(seed)
        self.state = (float(self.np_random.uniform(low=-0.6, high=-0.4)), 0.0)
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        position, velocity = self.state
        velocity = velocity + ((int(action) - 1) * self.force + self.m.cos(3 * position) * (-self.gravity))
        velocity = min(max(velocity, -self.max_speed), self.max_speed)
        position = position + velocity
        position = min(max(position, self.min_position), self.max_position)
        if position == self.min_position and velocity < 0:
            velocity = 0.0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        self.state = (position, velocity)
        return np.array(self.state, dtype=np.float32), -1.0, terminated, False, {}

    def close(self):
        pass

    def render(self):
        return None


class AcrobotRestated:
    """gym 0.26.2 AcrobotEnv (gym/envs/classic_control/acrobot.py), book_or_nips = "book", torque_noise_max = 0.
    Written expression by expression like the original (`_dsdt`, `rk4`, `wrap`, `bound`); `v**2` goes through `self.m.sq`
    and cos/sin through `self.m` so the two oracle flavours share every other line.  Deviation (documented): right after
    reset gym's state is a float32 array and numpy evaluates the observation's cos/sin in float32; here the observation is
    float32(double cos), like after every step."""
    dt = 0.2
    LINK_LENGTH_1 = 1.0
    LINK_LENGTH_2 = 1.0
    LINK_MASS_1 = 1.0
    LINK_MASS_2 = 1.0
    LINK_COM_POS_1 = 0.5
    LINK_COM_POS_2 = 0.5
    LINK_MOI = 1.0
    MAX_VEL_1 = 4 * math.pi
    MAX_VEL_2 = 9 * math.pi
    AVAIL_TORQUE = [-1.0, 0.0, +1]
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 15}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, trig="libm", with_spaces=True):
        self.m = _Math(trig)
        self.np_random = None
        self.state = None
        if with_spaces:
            high = np.array([1.0, 1.0, 1.0, 1.0, self.MAX_VEL_1, self.MAX_VEL_2], dtype=np.float32)
            self.observation_space = _Spaces.box(-high, high)
            self.action_space = _Spaces.discrete(3)

    def _get_ob(self):
        s = self.state
        c, sn = self.m.cos, self.m.sin
        return np.array([c(float(s[0])), sn(float(s[0])), c(float(s[1])), sn(float(s[1])), s[2], s[3]], dtype=np.float32)

    def reset(self, seed=None):
        if seed is not None or self.np_random is None:
            self.np_random = new_rng(seed)
        self.state = self.np_random.uniform(low=-0.1, high=0.1, size=(4,)).astype(np.float32)
        return self._get_ob(), {}

    def _dsdt(self, s_augmented):
        cos, sin, sq, pi = self.m.cos, self.m.sin, self.m.sq, math.pi
        m1, m2, l1 = self.LINK_MASS_1, self.LINK_MASS_2, self.LINK_LENGTH_1
        lc1, lc2 = self.LINK_COM_POS_1, self.LINK_COM_POS_2
        I1 = I2 = self.LINK_MOI
        g = 9.8
        a = s_augmented[-1]
        theta1, theta2, dtheta1, dtheta2 = s_augmented[:4]
        d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2
        d2 = m2 * (lc2 ** 2 + l1 * lc2 * cos(theta2)) + I2
        phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.0)
        phi1 = (-m2 * l1 * lc2 * sq(dtheta2) * sin(theta2)
                - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2)
                + (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2)
                + phi2)
        ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * sq(dtheta1) * sin(theta2) - phi2) / (m2 * lc2 ** 2 + I2 - sq(d2) / d1)
        ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
        return [dtheta1, dtheta2, ddtheta1, ddtheta2, 0.0]

    def _rk4(self, y0):
        dt = self.dt - 0
        dt2 = dt / 2.0
        ax = lambda y, c, k: [yi + c * ki for yi, ki in zip(y, k)]          # y0 + c * k, element by element
        k1 = self._dsdt(y0)
        k2 = self._dsdt(ax(y0, dt2, k1))
        k3 = self._dsdt(ax(y0, dt2, k2))
        k4 = self._dsdt(ax(y0, dt, k3))
        c6 = dt / 6.0
        return [y + c6 * (((a + 2 * b) + 2 * c) + d) for y, a, b, c, d in zip(y0, k1, k2, k3, k4)][:4]

    @staticmethod
    def _wrap(x, m, M):
        diff = M - m
        while x > M:
            x = x - diff
        while x < m:
            x = x + diff
        return x

    def step(self, a):
        torque = float(self.AVAIL_TORQUE[int(a)])
        s_augmented = [float(v) for v in self.state] + [torque]            # np.append(float32 state, torque) -> float64
        ns = self._rk4(s_augmented)
        ns[0] = self._wrap(ns[0], -math.pi, math.pi)
        ns[1] = self._wrap(ns[1], -math.pi, math.pi)
        ns[2] = min(max(ns[2], -self.MAX_VEL_1), self.MAX_VEL_1)
        ns[3] = min(max(ns[3], -self.MAX_VEL_2), self.MAX_VEL_2)
        self.state = np.array(ns, dtype=np.float64)
        terminated = bool(-self.m.cos(ns[0]) - self.m.cos(ns[1] + ns[0]) > 1.0)
        reward = -1.0 if not terminated else 0.0
        return self._get_ob(), reward, terminated, False, {}

    def close(self):
        pass

    def render(self):
        return None


class TimeLimitRestated:
    """gym 0.26.2 wrappers.TimeLimit (outermost wrapper returned by gym.make)."""

    def __init__(self, env, max_episode_steps):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None
        self.metadata = env.metadata
        self.reward_range = env.reward_range

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def unwrapped(self):
        return self.env

    def step(self, action):
        obs, rew, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, rew, terminated, truncated, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def close(self):
        self.env.close()

    def render(self):
        return self.env.render()


SPECS = {"CartPole-v1": (CartPoleRestated, 500), "Pendulum-v1": (PendulumRestated, 200),
         "MountainCar-v0": (MountainCarRestated, 200), "Acrobot-v1": (AcrobotRestated, 500)}
DEFAULT_TRIG = "libm"


def make(env_id, render_mode=None, trig=None, **kwargs):
    cls, limit = SPECS[env_id]
    return TimeLimitRestated(cls(trig=trig or DEFAULT_TRIG), limit)
