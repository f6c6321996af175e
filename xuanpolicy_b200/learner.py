"""PPOCLIP_Learner drop-in: the loss + its backward run in one hand-written kernel; the MLP stays in torch.

Mirrors PPOCLIP_Learner (xuance/torch/learners/policy_gradient/ppoclip_learner.py:4-65) and its base Learner
(xuance/torch/learners/learner.py:10-52): same constructor, `update(obs_batch, act_batch, ret_batch, value_batch,
adv_batch, old_logp) -> info`, `save_model`, `load_model`, `iterations`.

`update` (compat): forwards the policy with torch, calls the fused loss kernel on the network outputs
(csrc/ppo_loss.cu) to get dL/dlogits|dmu, dL/dlogstd, dL/dv, back-propagates them through the torch MLP, then
uses the caller's torch optimizer / scheduler exactly like the reference.
`update_from_buffer` (native): minibatch gather fused into the loss kernel, gradients in one flat buffer,
clip + Adam + LinearLR in one fused device step (csrc/optim.cu), optional NCCL all-reduces for env-sharded
data parallelism, and no host synchronisation (info scalars stay on the device until asked for).
"""
import os

import numpy as np
import torch

from . import ops
from .fused_mlp import FusedActorCritic


def _dist_params(a_dist):
    """('categorical', logits) or ('gaussian', mu, std) from a reference-shaped distribution wrapper
    (xuance/torch/utils/distributions.py:39-101: CategoricalDistribution.get_param / DiagGaussianDistribution.get_param)."""
    p = a_dist.get_param()
    if isinstance(p, (tuple, list)):
        return "gaussian", p[0], p[1]
    return "categorical", p, None


class FlatAdamState:
    """Flat fp32 views of the policy's parameters / gradients + Adam moments for csrc/optim.cu.

    Parameters are re-pointed at slices of one buffer (values preserved), `.grad` at slices of another, so
    torch autograd accumulates straight into the flat gradient and one NCCL all-reduce covers every tensor.
    Hyper-parameters are read from the torch optimizer / LinearLR scheduler the caller built
    (xuance/torch/runners/runner_drl.py:71-73)."""

    def __init__(self, policy, optimizer, scheduler):
        params = [p for p in policy.parameters() if p.requires_grad]
        dev = params[0].device
        pad4 = lambda k: (k + 3) // 4 * 4          # every tensor starts 16-byte aligned (float4 epilogue kernels)
        n = sum(pad4(p.numel()) for p in params)
        self.n = n
        self.n_params = sum(p.numel() for p in params)
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p.data)
            p.grad = self.flat_grad[off:off + k].view_as(p.data)
            off += pad4(k)
        self.params = params
        self.grad_views = [p.grad for p in params]
        group = optimizer.param_groups[0]
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
            raise NotImplementedError("fused Adam supports plain Adam (no weight decay / amsgrad / maximize)")
        self.lr0 = float(group.get("initial_lr", group["lr"]))
        self.beta1, self.beta2 = (float(b) for b in group["betas"])
        self.eps = float(group["eps"])
        self.end_factor, self.total_iters = 1.0, 0
        if scheduler is not None:
            if not isinstance(scheduler, torch.optim.lr_scheduler.LinearLR) or scheduler.start_factor != 1.0:
                raise NotImplementedError("fused Adam supports LinearLR(start_factor=1.0) as built by the reference runner")
            self.end_factor, self.total_iters = float(scheduler.end_factor), int(scheduler.total_iters)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.workspace = torch.zeros(8 + 1024, dtype=torch.float64, device=dev)
        self.lr = torch.full((1,), self.lr0, dtype=torch.float32, device=dev)
        self.gnorm = torch.zeros(1, dtype=torch.float32, device=dev)

    def backward_into(self, outputs, grad_outputs):
        """d(loss)/d(params) straight into the flat gradient: `torch.autograd.grad` (fresh gradient tensors, no
        per-parameter accumulate kernels) followed by one multi-tensor copy.  Parameters that are not reached from
        `outputs` (e.g. logstd, whose gradient the loss kernel produces directly) get zero."""
        grads = torch.autograd.grad(outputs, self.params, grad_outputs, allow_unused=True)
        dst = [v for v, g in zip(self.grad_views, grads) if g is not None]
        src = [g for g in grads if g is not None]
        torch._foreach_copy_(dst, src)
        for v, g in zip(self.grad_views, grads):
            if g is None:
                v.zero_()

    def apply(self, max_norm, grad_scale=1.0):
        ops.clip_adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step, self.lr0,
                           self.end_factor, self.total_iters, self.beta1, self.beta2, self.eps, max_norm, grad_scale,
                           self.workspace, lr_out=self.lr, gnorm_out=self.gnorm)


class PPOCLIP_Learner:
    def __init__(self, policy, optimizer, scheduler=None, device=None, model_dir="./", vf_coef=0.25, ent_coef=0.005,
                 clip_range=0.25, clip_grad_norm=0.25, use_grad_clip=True, value_clip=None):
        self.policy, self.optimizer, self.scheduler = policy, optimizer, scheduler
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("xuanpolicy_b200.PPOCLIP_Learner runs on a CUDA device only (no CPU fallback)")
        self.model_dir = model_dir
        self.iterations = 0
        self.vf_coef, self.ent_coef, self.clip_range = vf_coef, ent_coef, clip_range
        self.clip_grad_norm, self.use_grad_clip = clip_grad_norm, use_grad_clip
        self.value_clip = float(value_clip) if value_clip else 0.0   # opt-in; the reference has none
        self._scalars = torch.zeros(8, dtype=torch.float64, device=self.device)
        self._flat = None
        self._fused = None          # FusedActorCritic: the MLP on the tcgen05 dense kernels (native path, large batches)
        self.use_fused_mlp = os.environ.get("XB_FUSED_MLP", "1") != "0"
        self._mb = {}
        self.world_size, self.process_group = 1, None

    # ---------------------------------------------------------------------------------------------- checkpoints
    def save_model(self, model_path):
        torch.save(self.policy.state_dict(), model_path)

    def load_model(self, path, seed=1):
        for f in os.listdir(path):
            if "seed_%s" % seed in f:
                path = os.path.join(path, f)
                break
        names = sorted(n for n in os.listdir(path) if n != "obs_rms.npy")
        state = torch.load(os.path.join(path, names[-1]), map_location=self.device)
        self.policy.load_state_dict(state)

    # ---------------------------------------------------------------------------------------------- loss kernel
    def _loss_backward(self, a_dist, v_pred, act, ret, adv, old_logp, val_old, inv_batch, idx=None, T=0, N=0,
                       adv_stats=None, adv_count=0, packed=None, flat=None, fused=None):
        """Fused loss fwd+bwd on the network outputs, then the MLP backward: torch autograd (into `.grad`, or — with
        `flat` — straight into the flat gradient buffer) or, with `fused`, the hand-written dgrad/wgrad kernels."""
        backward = torch.autograd.backward if flat is None else flat.backward_into
        if fused is not None:
            backward = lambda outs, grads: fused.backward(grads[0], grads[1])
        kind, p0, p1 = _dist_params(a_dist)
        v = v_pred.detach().contiguous()
        dv = torch.empty_like(v)
        common = dict(clip_range=self.clip_range, vf_coef=self.vf_coef, ent_coef=self.ent_coef, inv_batch=inv_batch,
                      idx=idx, T=T, N=N, val_old=val_old if self.value_clip > 0 else None, adv_stats=adv_stats,
                      adv_count=adv_count, value_clip=self.value_clip, packed=packed)
        if kind == "categorical":
            logits = p0.detach().contiguous()
            dlogits = torch.empty_like(logits)
            ops.ppo_loss_categorical(logits, v, act, ret, adv, old_logp, dlogits, dv, self._scalars, **common)
            backward([p0, v_pred], [dlogits, dv])
        else:
            mu = p0.detach().contiguous()
            std = p1
            param = getattr(getattr(self.policy, "actor", None), "logstd", None)   # gaussian.py:25
            direct = param is not None and param.requires_grad and param.numel() == mu.shape[1]
            logstd = param.detach() if direct else std.detach().log().contiguous()
            dmu = torch.empty_like(mu)
            dls = torch.empty(mu.shape[1], dtype=torch.float64, device=mu.device)
            ops.ppo_loss_gaussian(mu, logstd, v, act, ret, adv, old_logp, dmu, dls, dv, self._scalars, **common)
            if direct and fused is not None and flat is not None:   # fp64 -> fp32 log-std gradient inside the tail launch
                fused.backward(dmu, dv, dls, flat.grad_views[[id(q) for q in flat.params].index(id(param))])
            elif direct:           # the kernel's dL/dlogstd goes straight into the parameter's gradient
                backward([p0, v_pred], [dmu, dv])
                if flat is not None:
                    flat.grad_views[[id(q) for q in flat.params].index(id(param))].copy_(dls)
                elif param.grad is None:
                    param.grad = dls.to(param.dtype)
                else:
                    param.grad.add_(dls.to(param.dtype))
            elif std.requires_grad:  # d/dstd = d/dlogstd / std ; autograd carries it back to the logstd parameter
                backward([p0, v_pred, std], [dmu, dv, (dls / std.detach().double()).to(std.dtype)])
            else:
                backward([p0, v_pred], [dmu, dv])

    # ---------------------------------------------------------------------------------------------- compat update
    def update(self, obs_batch, act_batch, ret_batch, value_batch, adv_batch, old_logp):
        self.iterations += 1
        dev = self.device

        def dv_(x):
            return torch.as_tensor(x, device=dev).to(torch.float32).contiguous()

        with torch.cuda.device(dev):
            act, ret, val, adv = dv_(act_batch), dv_(ret_batch), dv_(value_batch), dv_(adv_batch)
            olp = dv_(old_logp) if old_logp is not None else None
            obs = torch.as_tensor(obs_batch, device=dev)
            B = ret.shape[0]
            out = self.policy(obs)
            a_dist = out[1]
            # actor-only policies (PG) have no critic: a zero value head that receives (and ignores) a zero gradient
            v_pred = out[2] if len(out) > 2 else torch.zeros(B, device=dev, requires_grad=True)
            self.optimizer.zero_grad()
            self._loss_backward(a_dist, v_pred, act.reshape(B, -1) if act.dim() > 1 else act, ret, adv, olp, val, 1.0 / B)
            if self.use_grad_clip:
                torch.nn.utils.clip_grad_norm_(self.policy.parameters(), self.clip_grad_norm)
            self.optimizer.step()
            if self.scheduler is not None:
                self.scheduler.step()
            lr = self.optimizer.state_dict()["param_groups"][0]["lr"]
            s = self._scalars.cpu().numpy() / B
        return {"actor-loss": float(-s[0]), "critic-loss": float(s[1]), "entropy": float(s[2]), "learning_rate": lr,
                "predict_value": float(s[3]), "clip_ratio": torch.tensor(s[4], dtype=torch.float32)}

    # ---------------------------------------------------------------------------------------------- native update
    def enable_fused_optimizer(self, process_group=None):
        """Switch to the flat-buffer fused clip+Adam+LinearLR step (native path).  The torch optimizer handed to
        the constructor is only read for its hyper-parameters from here on."""
        if self._flat is None:
            self._flat = FlatAdamState(self.policy, self.optimizer, self.scheduler)
            if self.use_fused_mlp and FusedActorCritic.plan(self.policy) is not None:
                self._fused = FusedActorCritic(self.policy)
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.process_group = process_group
            self.world_size = torch.distributed.get_world_size(process_group)
        return self._flat

    def _minibatch_buffers(self, B, obs_dim):
        key = (B, obs_dim)
        if key not in self._mb:
            self._mb[key] = dict(obs=torch.empty((B, obs_dim), dtype=torch.float32, device=self.device),
                                 scal=torch.empty((B, 4), dtype=torch.float32, device=self.device),
                                 stats=torch.zeros(2, dtype=torch.float64, device=self.device))
        return self._mb[key]

    def stage_gather(self, memory, idx):
        """Stage 1 of a native update: gather the MLP input rows and the minibatch advantage statistics."""
        mb = self._minibatch_buffers(idx.numel(), memory.obs_dim)
        if memory.packed and self.value_clip <= 0:   # one 32-byte record per sample: obs + {act, old_logp, adv, ret}
            ops.gather_records(idx, memory.n_size, memory.n_envs, memory._rec, memory.obs_dim, mb["obs"], mb["scal"],
                               stats=mb["stats"] if memory.use_advnorm else None)
        else:
            ops.gather_obs(idx, memory.n_size, memory.n_envs, memory._obs, memory.obs_dim, mb["obs"],
                           b_adv=memory._adv if memory.use_advnorm else None,
                           stats=mb["stats"] if memory.use_advnorm else None)
        return mb

    def stage_forward_backward(self, memory, idx, mb):
        """Stage 2: torch MLP forward, fused gather+loss+backward kernel, torch MLP backward into the flat gradient."""
        B = idx.numel()
        fused = self._fused if (self._fused is not None and B >= FusedActorCritic.MIN_ROWS) else None
        if fused is not None:                        # tcgen05 dense kernels; weights re-split after every Adam step
            act_out, v_pred = fused.forward(mb["obs"], refresh=True)
            a_dist = fused.dist_params(act_out)
        else:
            _, a_dist, v_pred = self.policy(mb["obs"])
        stats = mb["stats"] if memory.use_advnorm else None
        if memory.packed and self.value_clip <= 0:   # scalars already gathered, compact and coalesced
            self._loss_backward(a_dist, v_pred, None, None, None, None, None, 1.0 / (B * self.world_size),
                                adv_stats=stats, adv_count=B * self.world_size, packed=mb["scal"], flat=self._flat,
                                fused=fused)
        else:                                        # gather fused into the loss kernel
            self._loss_backward(a_dist, v_pred, memory._act, memory._ret, memory._adv, memory._logp, memory._val,
                                1.0 / (B * self.world_size), idx=idx, T=memory.n_size, N=memory.n_envs,
                                adv_stats=stats, adv_count=B * self.world_size, flat=self._flat, fused=fused)

    def stage_optimizer(self):
        """Stage 3: global-norm clip + Adam + LinearLR on the flat buffers (one fused device step)."""
        self._flat.apply(self.clip_grad_norm if self.use_grad_clip else 0.0, 1.0)

    def update_from_buffer(self, memory, idx):
        """One PPO-Clip SGD step on the minibatch `idx` (CUDA int64 flat indices) of a native buffer.
        Env-sharded data parallel: the two collectives are the (sum, sumsq) of advantages and the flat gradient."""
        self.iterations += 1
        if self._flat is None:
            self.enable_fused_optimizer()
        mb = self.stage_gather(memory, idx)
        if self.world_size > 1 and memory.use_advnorm:
            torch.distributed.all_reduce(mb["stats"], group=self.process_group)
        self.stage_forward_backward(memory, idx, mb)
        if self.world_size > 1:
            torch.distributed.all_reduce(self._flat.flat_grad, group=self.process_group)
        self.stage_optimizer()

    def info(self, batch_size):
        """Host copy of the last update's log scalars (one sync; the native loop calls it once per rollout)."""
        s = self._scalars.cpu().numpy() / batch_size
        lr = float(self._flat.lr.item()) if self._flat is not None else self.optimizer.param_groups[0]["lr"]
        if self._flat is not None and self._flat.total_iters > 0:   # the reference reports the lr AFTER scheduler.step()
            it = min(int(self._flat.step.item()), self._flat.total_iters)
            lr = self._flat.lr0 * (1.0 + (self._flat.end_factor - 1.0) * it / self._flat.total_iters)
        return {"actor-loss": float(-s[0]), "critic-loss": float(s[1]), "entropy": float(s[2]), "learning_rate": lr,
                "predict_value": float(s[3]), "clip_ratio": torch.tensor(s[4], dtype=torch.float32)}


class A2C_Learner(PPOCLIP_Learner):
    """A2C_Learner drop-in (xuance/torch/learners/policy_gradient/a2c_learner.py:4-50): same constructor and
    `update(obs_batch, act_batch, ret_batch, adv_batch)`.  Same fused loss kernel with the A2C surrogate
    (`clip_range <= 0`: a_loss = -(adv * log_prob).mean()); the reference always clips the gradient norm (:36)."""

    def __init__(self, policy, optimizer, scheduler=None, device=None, model_dir="./", vf_coef=0.25, ent_coef=0.005,
                 clip_grad=None):
        super().__init__(policy, optimizer, scheduler, device, model_dir, vf_coef=vf_coef, ent_coef=ent_coef,
                         clip_range=0.0, clip_grad_norm=clip_grad, use_grad_clip=clip_grad is not None)
        self.clip_grad = clip_grad

    def update(self, obs_batch, act_batch, ret_batch, adv_batch):
        info = super().update(obs_batch, act_batch, ret_batch, ret_batch, adv_batch, None)
        info.pop("clip_ratio")
        return info


class PG_Learner(PPOCLIP_Learner):
    """PG_Learner drop-in (xuance/torch/learners/policy_gradient/pg_learner.py:4-45): actor-only policy
    (`policy(obs) -> (outputs, dist)`), a_loss = -(returns * log_prob).mean(), no value term."""

    def __init__(self, policy, optimizer, scheduler=None, device=None, model_dir="./", ent_coef=0.005, clip_grad=None):
        super().__init__(policy, optimizer, scheduler, device, model_dir, vf_coef=0.0, ent_coef=ent_coef, clip_range=0.0,
                         clip_grad_norm=clip_grad, use_grad_clip=clip_grad is not None)
        self.clip_grad = clip_grad

    def update(self, obs_batch, act_batch, ret_batch):
        info = super().update(obs_batch, act_batch, ret_batch, ret_batch, ret_batch, None)
        return {k: info[k] for k in ("actor-loss", "entropy", "learning_rate")}
