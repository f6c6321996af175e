/*
 * oracle/classic_control.c — CPU restatement (plain C) of the PPO hot path's integer/fp64 pieces.
 *
 * ORACLE / TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (xuanpolicy_b200/) never links or calls it.
 *
 * What it restates (citations relative to /root/reference unless noted):
 *   - gym 0.26.2 CartPoleEnv.step/reset, PendulumEnv.step/reset/_get_obs/angle_normalize, MountainCarEnv.step/reset,
 *     AcrobotEnv.step/reset/_dsdt/rk4/wrap/bound,
 *     TimeLimit.step/reset (third party, pinned setup.py:51, NOT vendored -> published algorithm,
 *     SURVEY.md App. A), driven the way xuance drives it:
 *       Gym_Env.step/reset bookkeeping   xuance/environment/gym/gym_env.py:36-49
 *       auto-reset + reset_obs stash      xuance/environment/gym/gym_vec_env.py:201-212
 *   - numpy PCG64 + Generator.uniform (checked against numpy in tests).
 *   - the batched form of DummyOnPolicyBuffer.finish_path (xuance/common/memory_tools.py:206-229)
 *     evaluated in fp64, the "float64 evaluation of the reference recurrence"
 *     that the 1e-5 GAE tolerance is measured against.
 *
 * PARITY UNPINNED for the physics (no golden vectors in the reference; gym absent).  Pinned pieces:
 * PCG64/uniform against numpy, GAE against the live reference's finish_path (tests/test_oracle_*.py).
 *
 * Flavours: 0 = "cr"   correctly-rounded sin/cos (libquadmath sinq/cosq rounded once to double) and
 *                      squares as v*v                       -> Tier-1 target, platform independent
 *           1 = "libm" sin/cos/pow(v,2.0) from the host libm -> what gym executes on this host (Tier 2)
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction: one rounding per operation).
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <string.h>

#define OC_PI 3.141592653589793
typedef unsigned __int128 u128;

/* ---------------------------------------------------------------- trig ---------------------------------- */
static inline double sin_f(double x, int flavour) { return flavour == 0 ? (double)sinq((__float128)x) : sin(x); }
static inline double cos_f(double x, int flavour) { return flavour == 0 ? (double)cosq((__float128)x) : cos(x); }
static inline double sq_f(double v, int flavour) { return flavour == 0 ? v * v : pow(v, 2.0); }

void oc_sincos(const double* x, double* s, double* c, long n, int flavour) {
    for (long i = 0; i < n; ++i) { s[i] = sin_f(x[i], flavour); c[i] = cos_f(x[i], flavour); }
}

/* ---------------------------------------------------------------- PCG64 --------------------------------- */
/* rng = {state_hi, state_lo, inc_hi, inc_lo}; numpy: step, then XSL-RR output of the NEW state. */
static inline uint64_t pcg64_next(uint64_t* rng) {
    const u128 mult = ((u128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    u128 st = ((u128)rng[0] << 64) | rng[1];
    u128 inc = ((u128)rng[2] << 64) | rng[3];
    st = st * mult + inc;
    rng[0] = (uint64_t)(st >> 64);
    rng[1] = (uint64_t)st;
    uint64_t x = rng[0] ^ rng[1];
    unsigned rot = (unsigned)(rng[0] >> 58);
    return (x >> rot) | (x << ((-rot) & 63));
}
static inline double pcg64_double(uint64_t* rng) { return (double)(pcg64_next(rng) >> 11) * (1.0 / 9007199254740992.0); }
static inline double pcg64_uniform(uint64_t* rng, double low, double range) { return low + range * pcg64_double(rng); }

void oc_pcg64_uniform(uint64_t* rng, double low, double range, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = pcg64_uniform(rng, low, range);
}

/* ---------------------------------------------------------------- CartPole ------------------------------ */
static void cartpole_draw(double* st, uint64_t* rng) {
    for (int k = 0; k < 4; ++k) st[k] = pcg64_uniform(rng, -0.05, 0.05 - (-0.05));
}

/* state [n][4] fp64, rng [n][4] u64.  n_draws resets per env (xuance does 2 before the first step:
 * Gym_Env ctor gym_env.py:19 and Runner_Base envs.reset() runner_basic.py:12). */
void oc_cartpole_reset(double* state, uint64_t* rng, int32_t* elapsed, double* ep_score, float* obs,
                       int n_draws, long n) {
    for (long e = 0; e < n; ++e) {
        for (int d = 0; d < n_draws; ++d) cartpole_draw(state + 4 * e, rng + 4 * e);
        elapsed[e] = 0; ep_score[e] = 0.0;
        for (int k = 0; k < 4; ++k) obs[4 * e + k] = (float)state[4 * e + k];
    }
}

void oc_cartpole_step(double* state, uint64_t* rng, int32_t* elapsed, double* ep_score,
                      const int64_t* actions, float* obs, float* rew, uint8_t* term, uint8_t* trunc,
                      float* reset_obs, int32_t* ep_step_out, double* ep_score_out,
                      int max_steps, long n, int flavour) {
    const double gravity = 9.8, masspole = 0.1, total_mass = 0.1 + 1.0, length = 0.5;
    const double polemass_length = 0.1 * 0.5, force_mag = 10.0, tau = 0.02;
    const double theta_thr = 12 * 2 * OC_PI / 360, x_thr = 2.4;
    for (long e = 0; e < n; ++e) {
        double* st = state + 4 * e;
        double x = st[0], x_dot = st[1], theta = st[2], theta_dot = st[3];
        double force = actions[e] == 1 ? force_mag : -force_mag;
        double c = cos_f(theta, flavour), s = sin_f(theta, flavour);
        double temp = (force + polemass_length * sq_f(theta_dot, flavour) * s) / total_mass;
        double thetaacc = (gravity * s - c * temp) / (length * (4.0 / 3.0 - masspole * sq_f(c, flavour) / total_mass));
        double xacc = temp - polemass_length * thetaacc * c / total_mass;
        x = x + tau * x_dot;
        x_dot = x_dot + tau * xacc;
        theta = theta + tau * theta_dot;
        theta_dot = theta_dot + tau * thetaacc;
        st[0] = x; st[1] = x_dot; st[2] = theta; st[3] = theta_dot;
        int terminated = (x < -x_thr) || (x > x_thr) || (theta < -theta_thr) || (theta > theta_thr);
        elapsed[e] += 1;
        int truncated = elapsed[e] >= max_steps;
        ep_score[e] += 1.0;
        for (int k = 0; k < 4; ++k) obs[4 * e + k] = (float)st[k];
        rew[e] = 1.0f; term[e] = (uint8_t)terminated; trunc[e] = (uint8_t)truncated;
        ep_step_out[e] = elapsed[e]; ep_score_out[e] = ep_score[e];
        if (terminated || truncated) {
            cartpole_draw(st, rng + 4 * e);
            elapsed[e] = 0; ep_score[e] = 0.0;
            for (int k = 0; k < 4; ++k) reset_obs[4 * e + k] = (float)st[k];
        }
    }
}


/* ---------------------------------------------------------------- MountainCar-v0 ------------------------ */
/* gym 0.26.2 MountainCarEnv (gym/envs/classic_control/mountain_car.py), restated from the published algorithm:
 *   velocity += (action - 1) * force + math.cos(3 * position) * (-gravity); clip; position += velocity; clip;
 *   left-wall inelastic stop; terminated = position >= 0.5 and velocity >= 0; reward -1; reset position U(-0.6, -0.4). */
static void mountaincar_draw(double* st, uint64_t* rng) {
    st[0] = pcg64_uniform(rng, -0.6, -0.4 - (-0.6));
    st[1] = 0.0;
}

void oc_mountaincar_reset(double* state /*[n][2]*/, uint64_t* rng, int32_t* elapsed, double* ep_score, float* obs /*[n][2]*/,
                          int n_draws, long n) {
    for (long e = 0; e < n; ++e) {
        for (int d = 0; d < n_draws; ++d) mountaincar_draw(state + 2 * e, rng + 4 * e);
        elapsed[e] = 0; ep_score[e] = 0.0;
        obs[2 * e] = (float)state[2 * e]; obs[2 * e + 1] = (float)state[2 * e + 1];
    }
}

void oc_mountaincar_step(double* state, uint64_t* rng, int32_t* elapsed, double* ep_score,
                         const int64_t* actions, float* obs, float* rew, uint8_t* term, uint8_t* trunc,
                         float* reset_obs, int32_t* ep_step_out, double* ep_score_out,
                         int max_steps, long n, int flavour) {
    const double force = 0.001, gravity = 0.0025, max_speed = 0.07, min_pos = -1.2, max_pos = 0.6, goal_pos = 0.5;
    for (long e = 0; e < n; ++e) {
        double* st = state + 2 * e;
        double position = st[0], velocity = st[1];
        velocity = velocity + ((double)(actions[e] - 1) * force + cos_f(3.0 * position, flavour) * (-gravity));
        velocity = velocity < -max_speed ? -max_speed : (velocity > max_speed ? max_speed : velocity);
        position = position + velocity;
        position = position < min_pos ? min_pos : (position > max_pos ? max_pos : position);
        if (position == min_pos && velocity < 0.0) velocity = 0.0;
        st[0] = position; st[1] = velocity;
        int terminated = position >= goal_pos && velocity >= 0.0;
        elapsed[e] += 1;
        int truncated = elapsed[e] >= max_steps;
        ep_score[e] += -1.0;
        obs[2 * e] = (float)st[0]; obs[2 * e + 1] = (float)st[1];
        rew[e] = -1.0f; term[e] = (uint8_t)terminated; trunc[e] = (uint8_t)truncated;
        ep_step_out[e] = elapsed[e]; ep_score_out[e] = ep_score[e];
        if (terminated || truncated) {
            mountaincar_draw(st, rng + 4 * e);
            elapsed[e] = 0; ep_score[e] = 0.0;
            reset_obs[2 * e] = (float)st[0]; reset_obs[2 * e + 1] = (float)st[1];
        }
    }
}

/* ---------------------------------------------------------------- Acrobot-v1 ---------------------------- */
/* gym 0.26.2 AcrobotEnv (gym/envs/classic_control/acrobot.py), "book" dynamics, restated from the published algorithm:
 *   s_augmented = np.append(state, AVAIL_TORQUE[a]); ns = rk4(_dsdt, s_augmented, [0, dt=0.2]) (one RK4 step);
 *   ns[0:2] = wrap(., -pi, pi); ns[2] = bound(., +-4pi); ns[3] = bound(., +-9pi);
 *   terminated = -cos(s0) - cos(s1 + s0) > 1.0; reward = -1.0 (0.0 when terminated);
 *   obs = float32([cos s0, sin s0, cos s1, sin s1, s2, s3]); reset: uniform(-0.1, 0.1, 4).astype(float32).
 * Every expression keeps Python's left-to-right order; products with the unit constants (m = l = I = 1, lc = 0.5) are
 * exact and folded.  Deviation (documented): right after reset gym's state is a float32 array and numpy evaluates the
 * observation's cos/sin in FLOAT32; here it is float32(double cos), like after every step. */
static void acrobot_draw(double* st, uint64_t* rng) {
    for (int k = 0; k < 4; ++k) st[k] = (double)(float)pcg64_uniform(rng, -0.1, 0.1 - (-0.1));
}
static void acrobot_obs(const double* st, float* o, int flavour) {
    o[0] = (float)cos_f(st[0], flavour); o[1] = (float)sin_f(st[0], flavour);
    o[2] = (float)cos_f(st[1], flavour); o[3] = (float)sin_f(st[1], flavour);
    o[4] = (float)st[2]; o[5] = (float)st[3];
}
static void acrobot_dsdt(const double* y, double a, double* k, int flavour) {
    const double theta1 = y[0], theta2 = y[1], dtheta1 = y[2], dtheta2 = y[3];
    const double c2 = cos_f(theta2, flavour), s2 = sin_f(theta2, flavour);
    const double d1 = ((0.25 + (1.25 + c2)) + 1.0) + 1.0;
    const double d2 = (0.25 + 0.5 * c2) + 1.0;
    const double phi2 = (0.5 * 9.8) * cos_f((theta1 + theta2) - OC_PI / 2.0, flavour);
    const double phi1 = (((-0.5 * sq_f(dtheta2, flavour)) * s2 - (dtheta2 * dtheta1) * s2)
                         + (1.5 * 9.8) * cos_f(theta1 - OC_PI / 2.0, flavour)) + phi2;
    const double ddtheta2 = (((a + (d2 / d1) * phi1) - (0.5 * sq_f(dtheta1, flavour)) * s2) - phi2)
                            / (1.25 - sq_f(d2, flavour) / d1);
    const double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
    k[0] = dtheta1; k[1] = dtheta2; k[2] = ddtheta1; k[3] = ddtheta2;
}
static double acrobot_wrap(double x, double m, double M) {
    const double diff = M - m;
    while (x > M) x = x - diff;
    while (x < m) x = x + diff;
    return x;
}

void oc_acrobot_reset(double* state /*[n][4]*/, uint64_t* rng, int32_t* elapsed, double* ep_score, float* obs /*[n][6]*/,
                      int n_draws, long n, int flavour) {
    for (long e = 0; e < n; ++e) {
        for (int d = 0; d < n_draws; ++d) acrobot_draw(state + 4 * e, rng + 4 * e);
        elapsed[e] = 0; ep_score[e] = 0.0;
        acrobot_obs(state + 4 * e, obs + 6 * e, flavour);
    }
}

void oc_acrobot_step(double* state, uint64_t* rng, int32_t* elapsed, double* ep_score,
                     const int64_t* actions, float* obs, float* rew, uint8_t* term, uint8_t* trunc,
                     float* reset_obs, int32_t* ep_step_out, double* ep_score_out,
                     int max_steps, long n, int flavour) {
    const double dt = 0.2, dt2 = 0.2 / 2.0, dt6 = 0.2 / 6.0;
    for (long e = 0; e < n; ++e) {
        double* st = state + 4 * e;
        const double a = (double)(actions[e] - 1);
        double k1[4], k2[4], k3[4], k4[4], y[4];
        acrobot_dsdt(st, a, k1, flavour);
        for (int i = 0; i < 4; ++i) y[i] = st[i] + dt2 * k1[i];
        acrobot_dsdt(y, a, k2, flavour);
        for (int i = 0; i < 4; ++i) y[i] = st[i] + dt2 * k2[i];
        acrobot_dsdt(y, a, k3, flavour);
        for (int i = 0; i < 4; ++i) y[i] = st[i] + dt * k3[i];
        acrobot_dsdt(y, a, k4, flavour);
        for (int i = 0; i < 4; ++i) y[i] = st[i] + dt6 * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
        y[0] = acrobot_wrap(y[0], -OC_PI, OC_PI);
        y[1] = acrobot_wrap(y[1], -OC_PI, OC_PI);
        const double v1 = 4 * OC_PI, v2 = 9 * OC_PI;
        y[2] = fmin(fmax(y[2], -v1), v1);
        y[3] = fmin(fmax(y[3], -v2), v2);
        for (int i = 0; i < 4; ++i) st[i] = y[i];
        int terminated = (-cos_f(st[0], flavour) - cos_f(st[1] + st[0], flavour)) > 1.0;
        double reward = terminated ? 0.0 : -1.0;
        elapsed[e] += 1;
        int truncated = elapsed[e] >= max_steps;
        ep_score[e] += reward;
        acrobot_obs(st, obs + 6 * e, flavour);
        rew[e] = (float)reward; term[e] = (uint8_t)terminated; trunc[e] = (uint8_t)truncated;
        ep_step_out[e] = elapsed[e]; ep_score_out[e] = ep_score[e];
        if (terminated || truncated) {
            acrobot_draw(st, rng + 4 * e);
            elapsed[e] = 0; ep_score[e] = 0.0;
            acrobot_obs(st, reset_obs + 6 * e, flavour);
        }
    }
}

/* ---------------------------------------------------------------- Pendulum ------------------------------ */
static void pendulum_draw(double* st, uint64_t* rng) {
    st[0] = pcg64_uniform(rng, -OC_PI, OC_PI - (-OC_PI));
    st[1] = pcg64_uniform(rng, -1.0, 1.0 - (-1.0));
}
static void pendulum_obs(const double* st, float* o, int flavour) {
    o[0] = (float)cos_f(st[0], flavour); o[1] = (float)sin_f(st[0], flavour); o[2] = (float)st[1];
}
static double angle_normalize(double x) {
    const double two_pi = 2 * OC_PI;
    double r = fmod(x + OC_PI, two_pi);          /* numpy floor-mod: fmod, then fix the sign */
    if (r != 0.0) { if (r < 0.0) r += two_pi; } else r = 0.0;
    return r - OC_PI;
}

void oc_pendulum_reset(double* state /*[n][2]*/, uint64_t* rng, int32_t* elapsed, double* ep_score, float* obs /*[n][3]*/,
                       int n_draws, long n, int flavour) {
    for (long e = 0; e < n; ++e) {
        for (int d = 0; d < n_draws; ++d) pendulum_draw(state + 2 * e, rng + 4 * e);
        elapsed[e] = 0; ep_score[e] = 0.0;
        pendulum_obs(state + 2 * e, obs + 3 * e, flavour);
    }
}

void oc_pendulum_step(double* state, uint64_t* rng, int32_t* elapsed, double* ep_score,
                      const float* actions, float* obs, float* rew, uint8_t* term, uint8_t* trunc,
                      float* reset_obs, int32_t* ep_step_out, double* ep_score_out,
                      int max_steps, long n, int flavour) {
    const double dt = 0.05;
    for (long e = 0; e < n; ++e) {
        double* st = state + 2 * e;
        double th = st[0], thdot = st[1];
        float u32 = actions[e];
        u32 = u32 < -2.0f ? -2.0f : (u32 > 2.0f ? 2.0f : u32);
        double u = (double)u32;                                    /* pinned numpy 1.21.6 promotion */
        double costs = sq_f(angle_normalize(th), flavour) + 0.1 * sq_f(thdot, flavour) + 0.001 * (u * u);
        double newthdot = thdot + (15.0 * sin_f(th, flavour) + 3.0 * u) * dt;
        newthdot = newthdot < -8.0 ? -8.0 : (newthdot > 8.0 ? 8.0 : newthdot);
        double newth = th + newthdot * dt;
        st[0] = newth; st[1] = newthdot;
        double reward = -costs;
        elapsed[e] += 1;
        int truncated = elapsed[e] >= max_steps;
        ep_score[e] += reward;
        pendulum_obs(st, obs + 3 * e, flavour);
        rew[e] = (float)reward; term[e] = 0; trunc[e] = (uint8_t)truncated;
        ep_step_out[e] = elapsed[e]; ep_score_out[e] = ep_score[e];
        if (truncated) {
            pendulum_draw(st, rng + 4 * e);
            elapsed[e] = 0; ep_score[e] = 0.0;
            pendulum_obs(st, reset_obs + 3 * e, flavour);
        }
    }
}

/* ---------------------------------------------------------------- GAE (fp64, batched finish_path) ------- */
/* Time-major [T][N] fp32 inputs.  segend (may be NULL) marks truncations; boot [T][N] (may be NULL when
 * segend is NULL) holds V(terminal obs) where segend is set; boot_last [N] is the bootstrap at T-1.
 * use_gae=0 restates the discount_cumsum branch (memory_tools.py:222-225, common_tools.py:199-200).
 * Outputs are fp64 so tests can measure the CUDA kernel's fp32 rounding against an unrounded target. */
void oc_gae(const float* rew, const float* val, const float* term, const uint8_t* segend, const float* boot,
            const float* boot_last, double* adv, double* ret, long T, long N, double gamma, double lam, int use_gae) {
    for (long e = 0; e < N; ++e) {
        double last = 0.0, nextv = 0.0, run = 0.0;
        for (long t = T - 1; t >= 0; --t) {
            long k = t * N + e;
            int seg_end = (t == T - 1) || term[k] != 0.0f || (segend && segend[k]);
            if (seg_end) {
                double b = (t == T - 1) ? (double)boot_last[e] : (boot ? (double)boot[k] : 0.0);
                if (term[k] != 0.0f) b = 0.0;     /* ppoclip_agent.py:73,97: finish_path(0.0, i) on terminal */
                nextv = b; last = 0.0; run = b;
            }
            double r = rew[k], v = val[k];
            if (use_gae) {
                double nt = 1.0 - (double)term[k];
                double delta = r + nt * gamma * nextv - v;
                last = delta + nt * gamma * lam * last;
                adv[k] = last; ret[k] = last + v;
            } else {
                run = r + gamma * run;
                adv[k] = r + gamma * nextv - v; ret[k] = run;
            }
            nextv = v;
        }
    }
}
