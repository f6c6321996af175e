"""Actor-critic torch modules for the PPO path (the GEMMs stay in torch / cuBLAS by design).

These mirror the *interface and parameter naming* of the reference modules so a `state_dict` moves between
the two unchanged and `PPOCLIP_Learner.update` can drive either:

    MLPRepresentation      <-> Basic_MLP               xuance/torch/representations/mlp.py:21-51
    CategoricalActorCritic <-> Categorical_AC_Policy   xuance/torch/policies/categorical.py:16-85
    GaussianActorCritic    <-> Gaussian_AC_Policy      xuance/torch/policies/gaussian.py:8-77
    distributions          <-> xuance/torch/utils/distributions.py:39-101
    Categorical/GaussianPPGActorCritic <-> PPGActorCritic   categorical.py:110-141 / gaussian.py:103-131

`forward(obs)` returns `(outputs_dict, dist, value)` exactly like the reference.  When the drop-in classes are
used behind the reference's own runner, the reference's modules are used instead and these are not needed.
"""
import copy
import math
from typing import Sequence

import numpy as np
import torch
import torch.nn as nn

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def _split_k(B):
    S = 1
    while S < 64 and B % (2 * S) == 0 and B // (2 * S) >= 512:
        S *= 2
    return S


def _wgrad(dz, x):
    """dW = dz^T @ x for tall-skinny operands (M = N = H, K = B): cuBLAS's fp32 heuristics serve that shape with a
    handful of CTAs, so it is issued as a batched split-K product over S slabs of the batch and summed."""
    B, H = dz.shape
    S = _split_k(B)
    if S > 1 and x.is_contiguous():
        return torch.bmm(dz.view(S, B // S, H).transpose(1, 2), x.view(S, B // S, x.shape[1])).sum(0)
    return dz.t() @ x


def _colsum(dy):
    """Bias gradient of a narrow head (A or 1 columns): two-stage sum instead of torch's single-CTA reduction."""
    B = dy.shape[0]
    if B % 256 == 0 and B >= 4096:
        return dy.view(256, B // 256, -1).sum(1).sum(0)
    return dy.sum(0)


class _LinearLeakyReLU(torch.autograd.Function):
    """y = leaky_relu(x @ W^T + b).  The GEMM-shaped pieces stay cuBLAS calls (mm / bmm); everything torch would run
    around them is hand-written (csrc/mlp_epilogue.cu): forward bias + activation in one in-place pass (torch's
    addmm adds a separate 36 us bias kernel + a 10 us activation kernel at [65536,128]) and, in the backward,
    leaky_relu' fused with the bias-gradient column reduction (torch: 14 us + 82 us)."""

    @staticmethod
    def forward(ctx, x, weight, bias, slope, workspace):
        from . import ops
        y = x @ weight.t()
        ops.bias_act_fwd(y, bias, slope)
        ctx.save_for_backward(x, weight, y)
        ctx.slope, ctx.workspace = slope, workspace
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        x, weight, y = ctx.saved_tensors
        dy = dy.contiguous()
        dz = torch.empty_like(dy)
        db = torch.empty(dy.shape[1], dtype=dy.dtype, device=dy.device)
        ops.act_bias_bwd(dy, y, ctx.slope, dz, db, ctx.workspace)
        dx = dz @ weight if ctx.needs_input_grad[0] else None
        return dx, _wgrad(dz, x), db, None, None


class _LinearHead(torch.autograd.Function):
    """Final Linear of a head (no activation): same math as F.linear, cheaper narrow-output gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ weight if ctx.needs_input_grad[0] else None
        return dx, _wgrad(dy, x), _colsum(dy)


class _HiddenThenHead(torch.autograd.Function):
    """out = (leaky_relu(x @ W1^T + b1)) @ W2^T + b2 with a narrow head (<= 4 outputs): the tail
    [Linear, LeakyReLU, Linear] of the actor / critic (categorical.py:26-32,48-54; gaussian.py:17-24,41-48) as ONE
    autograd node.  cuBLAS does the H x H products; the head product (bandwidth-bound matrix-vector work) and the whole
    non-GEMM backward — head gradients, leaky_relu', both bias gradients — are two hand-written kernels."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, slope, workspace):
        from . import ops
        y = x @ w1.t()
        ops.bias_act_fwd(y, b1, slope)
        out = torch.empty((x.shape[0], w2.shape[0]), dtype=x.dtype, device=x.device)
        ops.head_fwd(y, w2, b2, out)
        ctx.save_for_backward(x, w1, y, w2)
        ctx.slope, ctx.workspace = slope, workspace
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import ops
        x, w1, y, w2 = ctx.saved_tensors
        dout = dout.contiguous()
        dz = torch.empty_like(y)
        db1 = torch.empty(y.shape[1], dtype=y.dtype, device=y.device)
        dw2, db2 = torch.empty_like(w2), torch.empty(w2.shape[0], dtype=y.dtype, device=y.device)
        ops.head_bwd_act(dout, y, w2, ctx.slope, dz, db1, dw2, db2, ctx.workspace)
        dx = dz @ w1 if ctx.needs_input_grad[0] else None
        return dx, _wgrad(dz, x), db1, dw2, db2, None, None


def _epilogue_ok(lin):
    """The float4 epilogue kernels need H % 4 == 0, 256 % (H/4) == 0 and 16-byte aligned parameters."""
    h = lin.out_features
    return (h % 4 == 0 and h <= 1024 and 256 % (h // 4) == 0 and lin.bias.data_ptr() % 16 == 0
            and lin.weight.data_ptr() % 16 == 0)


def _slope(mod):
    """Negative-side slope of a (Leaky)ReLU module — ReLU (the PG / PPG yamls' activation) is the slope-0 case of the
    same fused bias+activation kernels — or None for anything else."""
    if isinstance(mod, nn.LeakyReLU):
        return float(mod.negative_slope)
    return 0.0 if isinstance(mod, nn.ReLU) else None


class _DenseStack(nn.Sequential):
    """nn.Sequential of [Linear, LeakyReLU, ...] (same parameter names as the reference's Sequential).  On a CUDA
    device Linear+LeakyReLU pairs run as cuBLAS mm + the hand-written epilogue kernels (with a custom autograd
    function when gradients are recorded); everywhere else it is the plain module chain."""

    def _workspace(self, device):
        ws = getattr(self, "_xb_ws", None)
        if ws is None or ws.device != device:
            ws = torch.zeros(4 + 592 * 1024, dtype=torch.float32, device=device)
            object.__setattr__(self, "_xb_ws", ws)
        return ws

    def forward(self, x):
        from . import ops
        mods = list(self)
        on_gpu = x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
        k = 0
        while k < len(mods):
            m = mods[k]
            nxt = mods[k + 1] if k + 1 < len(mods) else None
            head = mods[k + 2] if k + 2 == len(mods) - 1 else None
            fusable = on_gpu and isinstance(m, nn.Linear) and _slope(nxt) is not None and _epilogue_ok(m)
            if (fusable and isinstance(head, nn.Linear) and head.out_features <= 4 and m.out_features <= 512
                    and head.weight.data_ptr() % 16 == 0):
                # tail [Linear, LeakyReLU, Linear(<= 4 outputs)]: one autograd node, hand-written head kernels
                slope = _slope(nxt)
                if torch.is_grad_enabled() and (m.weight.requires_grad or x.requires_grad):
                    x = _HiddenThenHead.apply(x, m.weight, m.bias, head.weight, head.bias, slope, self._workspace(x.device))
                else:
                    y = x @ m.weight.t()
                    ops.bias_act_fwd(y, m.bias, slope)
                    x = torch.empty((y.shape[0], head.out_features), dtype=y.dtype, device=y.device)
                    ops.head_fwd(y, head.weight, head.bias, x)
                k += 3
            elif fusable:
                if torch.is_grad_enabled() and (m.weight.requires_grad or x.requires_grad):
                    x = _LinearLeakyReLU.apply(x, m.weight, m.bias, _slope(nxt), self._workspace(x.device))
                else:
                    x = x @ m.weight.t()
                    ops.bias_act_fwd(x, m.bias, _slope(nxt))
                k += 2
            elif on_gpu and isinstance(m, nn.Linear) and torch.is_grad_enabled() and m.weight.requires_grad:
                x = _LinearHead.apply(x, m.weight, m.bias)
                k += 1
            else:
                x = m(x)
                k += 1
        return x


def _dense_stack(sizes: Sequence[int], act, init, device, last_plain=False, last_init=True):
    """[Linear, act, Linear, act, ...]; with last_plain the final Linear has no activation."""
    mods = []
    for k in range(len(sizes) - 1):
        lin = nn.Linear(sizes[k], sizes[k + 1], device=device)
        final = k == len(sizes) - 2
        if init is not None and (last_init or not final):
            init(lin.weight)
            nn.init.zeros_(lin.bias)
        mods.append(lin)
        if not (final and last_plain):
            mods.append(act())
    return _DenseStack(*mods)


class MLPRepresentation(nn.Module):
    def __init__(self, input_shape, hidden_sizes, normalize=None, initialize=nn.init.orthogonal_,
                 activation=nn.LeakyReLU, device=None):
        super().__init__()
        assert normalize is None, "normalisation layers are not on the PPO classic-control path"
        self.input_shape, self.hidden_sizes, self.device = tuple(input_shape), list(hidden_sizes), device
        self.output_shapes = {"state": (self.hidden_sizes[-1],)}
        self.model = _dense_stack([self.input_shape[0]] + self.hidden_sizes, activation, initialize, device)

    def forward(self, observations):
        x = torch.as_tensor(observations, dtype=torch.float32, device=self.device)
        return {"state": self.model(x)}


class CategoricalDistribution:
    def __init__(self, action_dim):
        self.action_dim = action_dim
        self.logits = None

    def set_param(self, logits):
        self.logits = logits
        self._logp_cache = None

    @property
    def _logp(self):  # normalised log-probabilities, computed only when a torch-side consumer asks for them
        key = torch.is_grad_enabled()
        if self._logp_cache is None or self._logp_cache[0] != key:
            self._logp_cache = (key, self.logits - self.logits.logsumexp(dim=-1, keepdim=True))
        return self._logp_cache[1]

    def get_param(self):
        return self.logits

    def log_prob(self, x):
        return self._logp.gather(-1, x.long().unsqueeze(-1)).squeeze(-1)

    def entropy(self):
        return -(self._logp.exp() * self._logp).sum(-1)

    def stochastic_sample(self):
        return torch.multinomial(self._logp.exp(), 1).squeeze(-1)

    def deterministic_sample(self):
        return torch.argmax(self._logp, dim=1)


class DiagGaussianDistribution:
    def __init__(self, action_dim):
        self.action_dim = action_dim
        self.mu = self.std = None

    def set_param(self, mu, std):
        self.mu, self.std = mu, std

    def get_param(self):
        return self.mu, self.std

    def log_prob(self, x):
        var = self.std ** 2
        return (-((x - self.mu) ** 2) / (2 * var) - self.std.log() - _HALF_LOG_2PI).sum(-1)

    def entropy(self):
        return (0.5 + _HALF_LOG_2PI + self.std.log()).expand_as(self.mu).sum(-1)

    def stochastic_sample(self):
        with torch.no_grad():
            return torch.normal(self.mu, self.std.expand_as(self.mu))

    def deterministic_sample(self):
        return self.mu


def old_dist_params(old_dists, device):
    """Parameters of the OLD action distribution of a minibatch as CUDA tensors: ('categorical', logits [B, A], None)
    or ('gaussian', mu [B, A], std [B, A] | [A]).

    Accepts what the reference hands its learners — a numpy object array of per-sample distribution wrappers
    (`split_distributions` / `merge_distributions`, xuance/torch/utils/operations.py:53-92) — as well as one batched
    wrapper (anything with `get_param()`), which is what the device buffer returns from `sample`."""
    f32 = lambda t: torch.as_tensor(t, device=device).detach().to(torch.float32).contiguous()
    if hasattr(old_dists, "get_param"):
        p = old_dists.get_param()
        if isinstance(p, (tuple, list)):
            return "gaussian", f32(p[0]), f32(p[1])
        return "categorical", f32(p), None
    flat = np.asarray(old_dists, dtype=object).reshape(-1)
    first = flat[0].get_param()
    if isinstance(first, (tuple, list)):
        mu = torch.stack([torch.as_tensor(d.get_param()[0]).reshape(-1) for d in flat])
        std = torch.stack([torch.as_tensor(d.get_param()[1]).reshape(-1) for d in flat])
        return "gaussian", f32(mu), f32(std)
    logits = torch.cat([torch.as_tensor(d.get_param()).reshape(1, -1) for d in flat], dim=0)
    return "categorical", f32(logits), None


class OldDistBatch:
    """The old action distributions of a batch of transitions as two device tensors — what the device buffer keeps
    and returns for the auxiliary key "old_dist" instead of the reference's numpy array of per-sample Python objects
    (memory_tools.py:28-30, operations.py:53-72).  `get_param()` has the wrappers' meaning: logits, or (mu, std)."""

    def __init__(self, kind, p0, p1=None):
        self.kind, self.p0, self.p1 = kind, p0, p1

    def get_param(self):
        return self.p0 if self.kind == "categorical" else (self.p0, self.p1)

    @property
    def shape(self):
        return tuple(self.p0.shape[:-1])

    def __len__(self):
        return self.p0.shape[0]


class _CategoricalActor(nn.Module):
    def __init__(self, state_dim, action_dim, hidden, act, init, device):
        super().__init__()
        self.model = _dense_stack([state_dim] + list(hidden) + [action_dim], act, init, device, last_plain=True)
        self.dist = CategoricalDistribution(action_dim)

    def forward(self, x):
        self.dist.set_param(self.model(x))
        return self.dist


class _GaussianActor(nn.Module):
    def __init__(self, state_dim, action_dim, hidden, act, init, device):
        super().__init__()
        self.mu = _dense_stack([state_dim] + list(hidden) + [action_dim], act, init, device, last_plain=True)
        self.logstd = nn.Parameter(-torch.ones((action_dim,), device=device))
        self.dist = DiagGaussianDistribution(action_dim)

    def forward(self, x):
        self.dist.set_param(self.mu(x), self.logstd.exp())
        return self.dist


class _Critic(nn.Module):
    def __init__(self, state_dim, hidden, act, init, device, last_init=True):
        super().__init__()
        self.model = _dense_stack([state_dim] + list(hidden) + [1], act, init, device, last_plain=True,
                                  last_init=last_init)

    def forward(self, x):
        return self.model(x)[:, 0]


class _ActorCritic(nn.Module):
    def forward(self, observation):
        outputs = self.representation(observation)
        return outputs, self.actor(outputs["state"]), self.critic(outputs["state"])


class CategoricalActorCritic(_ActorCritic):
    def __init__(self, action_space, representation, actor_hidden_size, critic_hidden_size, normalize=None,
                 initialize=nn.init.orthogonal_, activation=nn.LeakyReLU, device=None):
        super().__init__()
        self.device, self.action_dim, self.representation = device, action_space.n, representation
        self.representation_info_shape = representation.output_shapes
        sd = representation.output_shapes["state"][0]
        self.actor = _CategoricalActor(sd, self.action_dim, actor_hidden_size, activation, initialize, device)
        self.critic = _Critic(sd, critic_hidden_size, activation, initialize, device)


class GaussianActorCritic(_ActorCritic):
    def __init__(self, action_space, representation, actor_hidden_size, critic_hidden_size, normalize=None,
                 initialize=nn.init.orthogonal_, activation=nn.LeakyReLU, device=None):
        super().__init__()
        self.device, self.action_dim, self.representation = device, action_space.shape[0], representation
        self.representation_info_shape = representation.output_shapes
        sd = representation.output_shapes["state"][0]
        self.actor = _GaussianActor(sd, self.action_dim, actor_hidden_size, activation, initialize, device)
        # the reference leaves the Gaussian critic's last layer at torch's default init (gaussian.py:47)
        self.critic = _Critic(sd, critic_hidden_size, activation, initialize, device, last_init=False)


class CategoricalActor(nn.Module):
    """Actor-only policy (Categorical_Actor_Policy, xuance/torch/policies/categorical.py:88-107), used by PG."""

    def __init__(self, action_space, representation, actor_hidden_size, normalize=None, initialize=nn.init.orthogonal_,
                 activation=nn.LeakyReLU, device=None):
        super().__init__()
        self.action_dim, self.representation = action_space.n, representation
        self.representation_info_shape = representation.output_shapes
        self.actor = _CategoricalActor(representation.output_shapes["state"][0], self.action_dim, actor_hidden_size,
                                       activation, initialize, device)

    def forward(self, observation):
        outputs = self.representation(observation)
        return outputs, self.actor(outputs["state"])


class GaussianActor(nn.Module):
    """Actor-only policy for Box actions (Gaussian_Actor: ActorPolicy, xuance/torch/policies/gaussian.py:80-100), used by PG."""

    def __init__(self, action_space, representation, actor_hidden_size, normalize=None, initialize=nn.init.orthogonal_,
                 activation=nn.LeakyReLU, device=None):
        super().__init__()
        self.action_dim, self.representation = action_space.shape[0], representation
        self.representation_info_shape = representation.output_shapes
        self.actor = _GaussianActor(representation.output_shapes["state"][0], self.action_dim, actor_hidden_size,
                                    activation, initialize, device)

    def forward(self, observation):
        outputs = self.representation(observation)
        return outputs, self.actor(outputs["state"])


class CategoricalPPGActorCritic(nn.Module):
    """PPGActorCritic for Discrete actions (xuance/torch/policies/categorical.py:110-141): three copies of the
    representation (actor / critic / auxiliary critic) and `forward -> (policy_outputs, a_dist, v, aux_v)`."""

    def __init__(self, action_space, representation, actor_hidden_size, critic_hidden_size, normalize=None,
                 initialize=nn.init.orthogonal_, activation=nn.LeakyReLU, device=None):
        super().__init__()
        self.action_dim = action_space.n
        self.actor_representation = representation
        self.critic_representation = copy.deepcopy(representation)
        self.aux_critic_representation = copy.deepcopy(representation)
        self.representation_info_shape = representation.output_shapes
        sd = representation.output_shapes["state"][0]
        self.actor = _CategoricalActor(sd, self.action_dim, actor_hidden_size, activation, initialize, device)
        self.critic = _Critic(sd, critic_hidden_size, activation, initialize, device)
        self.aux_critic = _Critic(sd, critic_hidden_size, activation, initialize, device)

    def forward(self, observation):
        policy_outputs = self.actor_representation(observation)
        critic_outputs = self.critic_representation(observation)
        aux_outputs = self.aux_critic_representation(observation)
        return (policy_outputs, self.actor(policy_outputs["state"]), self.critic(critic_outputs["state"]),
                self.aux_critic(aux_outputs["state"]))


class GaussianPPGActorCritic(nn.Module):
    """PPGActorCritic for Box actions (xuance/torch/policies/gaussian.py:103-131): the auxiliary critic reads the
    ACTOR's representation (:130), the critic has its own copy."""

    def __init__(self, action_space, representation, actor_hidden_size, critic_hidden_size, normalize=None,
                 initialize=nn.init.orthogonal_, activation=nn.LeakyReLU, device=None):
        super().__init__()
        self.action_dim = action_space.shape[0]
        self.actor_representation = representation
        self.critic_representation = copy.deepcopy(representation)
        self.representation_info_shape = representation.output_shapes
        sd = representation.output_shapes["state"][0]
        self.actor = _GaussianActor(sd, self.action_dim, actor_hidden_size, activation, initialize, device)
        self.critic = _Critic(sd, critic_hidden_size, activation, initialize, device, last_init=False)
        self.aux_critic = _Critic(sd, critic_hidden_size, activation, initialize, device, last_init=False)

    def forward(self, observation):
        policy_outputs = self.actor_representation(observation)
        critic_outputs = self.critic_representation(observation)
        return (policy_outputs, self.actor(policy_outputs["state"]), self.critic(critic_outputs["state"]),
                self.aux_critic(policy_outputs["state"]))


def make_policy(observation_space, action_space, hidden=(128,), device=None, seed=None):
    """Builds the `representation_hidden_size=actor_hidden_size=critic_hidden_size=hidden` policy of the yaml configs."""
    from .spaces import is_discrete
    if seed is not None:
        torch.manual_seed(seed)
        np.random.seed(seed)
    rep = MLPRepresentation(observation_space.shape, list(hidden), device=device)
    cls = CategoricalActorCritic if is_discrete(action_space) else GaussianActorCritic
    return cls(action_space, rep, list(hidden), list(hidden), device=device)
