"""The actor-critic MLP of the PPO path on the hand-written dense kernels (csrc/dense_tc.cu, csrc/mlp_trunk.cu).

The torch modules keep owning the parameters (same names / state_dict as the reference's
Basic_MLP + ActorNet + CriticNet, xuance/torch/representations/mlp.py:21-51, policies/categorical.py:16-85,
policies/gaussian.py:8-77); this class only *evaluates* them at large batch:

    forward   obs -> h1 = leaky(W0 obs + b0)                         SIMT (3/4-wide reduction)
              h1  -> ya = leaky(Wa1 h1 + ba1), act_out = Wa2 ya + ba2    tcgen05 3xTF32 GEMM + fused head
              h1  -> yc = leaky(Wc1 h1 + bc1), v = Wc2 yc + bc2          tcgen05 3xTF32 GEMM + fused head
    backward  (dL/dact_out, dL/dv) -> every parameter gradient, written straight into the flat gradient buffer:
              dgrad (dz_a | dz_c generated on the fly) -> dz1; wgrad (dWa1, dba1, dWa2, dba2, dWc1, ...); trunk wgrad.

It replaces torch autograd + cuBLAS SIMT sgemm for the supported shape (one hidden layer per block, width 128 or 256,
LeakyReLU, obs dim <= 8, action dim <= 2 — or a Discrete(3) head, folded onto the two-head kernels through the shift
invariance of the softmax (csrc/mlp_trunk.cu xb_head3_fold) — i.e. every classic-control config of the reference:
CartPole, Pendulum, MountainCar, Acrobot); anything else keeps the torch path.  No CPU path: construction requires CUDA
parameters.
"""
import os

import torch
import torch.nn as nn

from . import ops


def _stack(seq):
    return list(seq) if isinstance(seq, nn.Sequential) else None


class _OutParams:
    """What PPOCLIP_Agent._sample / the loss need from a distribution: `get_param()` (distributions.py:47,78)."""

    def __init__(self, logits=None, mu=None, std=None):
        self._logits, self._mu, self._std = logits, mu, std

    def get_param(self):
        return self._logits if self._logits is not None else (self._mu, self._std)


class FusedActorCritic:
    MIN_ROWS = 2048          # below this the launch-bound torch path is as good

    @staticmethod
    def plan(policy):
        """Returns the layer tuple if `policy` has the supported structure, else None."""
        try:
            rep = _stack(policy.representation.model)
            actor = _stack(getattr(policy.actor, "mu", None) if hasattr(policy.actor, "mu") else policy.actor.model)
            critic = _stack(policy.critic.model)
        except AttributeError:
            return None
        if rep is None or actor is None or critic is None or len(rep) != 2 or len(actor) != 3 or len(critic) != 3:
            return None
        l0, a0 = rep
        la1, aa, la2 = actor
        lc1, ac, lc2 = critic
        lins = (l0, la1, la2, lc1, lc2)
        if not all(isinstance(m, nn.Linear) and m.bias is not None for m in lins):
            return None
        if not all(isinstance(a, nn.LeakyReLU) for a in (a0, aa, ac)):
            return None
        if len({float(a.negative_slope) for a in (a0, aa, ac)}) != 1:
            return None
        H = l0.out_features
        if H not in (128, 256) or la1.in_features != H or la1.out_features != H or lc1.in_features != H \
                or lc1.out_features != H or la2.in_features != H or lc2.in_features != H:
            return None
        gaussian = hasattr(policy.actor, "logstd")
        if l0.in_features > 8 or la2.out_features > (2 if gaussian else 3) or lc2.out_features != 1:
            return None
        for m in lins:
            for t in (m.weight, m.bias):
                if not t.is_cuda or t.dtype != torch.float32:
                    return None
        return dict(l0=l0, la1=la1, la2=la2, lc1=lc1, lc2=lc2, H=H, obs_dim=l0.in_features, A=la2.out_features,
                    slope=float(a0.negative_slope), gaussian=hasattr(policy.actor, "logstd"))

    def __init__(self, policy):
        pl = self.plan(policy)
        if pl is None:
            raise ValueError("policy structure not supported by the fused dense kernels")
        self.policy = policy
        self.__dict__.update(pl)
        dev = self.l0.weight.device
        self.device = dev
        H = self.H
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        # hi/lo splits of the two H x H layers and their transposes concatenated along the reduction dim (dgrad operand)
        self.wa_hi, self.wa_lo, self.wc_hi, self.wc_lo = z(H, H), z(H, H), z(H, H), z(H, H)
        self.wt_hi, self.wt_lo = z(H, 2 * H), z(H, 2 * H)
        # mask-form dgrad operand (w2-scaled transposed weights), rewritten by every training forward (stage_hidden)
        self.wtm_hi, self.wtm_lo = z(H, 2 * H), z(H, 2 * H)
        self.mask_dgrad = os.environ.get("XB_MASK_DGRAD", "1") != "0"
        # activation sign words: the forward leaves one bit per hidden activation, dgrad reads those instead of ya / yc
        self.sign_bits = os.environ.get("XB_SIGN_BITS", "1") != "0"
        self.sign_dgrad = self.sign_bits and H == 128       # (the tensor-memory dgrad kernel; H = 256 runs the SS form on ya / yc)
        # binary-form weight gradients (rank-1 head gradients: 0/1 A operand from the sign words, 2 MMAs per k-step, no read of ya / yc)
        self.bin_wgrad = self.sign_bits and os.environ.get("XB_BIN_WGRAD", "1") != "0"
        self._bin_now = False
        self.h1_signs = self.sign_dgrad and os.environ.get("XB_H1_SIGNS", "0") == "1"
        self._mask_ready = False
        # Discrete(3): folded head parameters (w_j - w_2, b_j - b_2), two-column head outputs / gradients
        self.fold3 = (not self.gaussian) and self.A == 3
        self.nh_a = 2 if self.fold3 else self.A
        self.w2f, self.b2f = (z(2, H), z(2)) if self.fold3 else (None, None)
        self.ws_wgrad = ops.dense_wgrad_workspace(H, dev)
        self.ws_trunk = ops.mlp_trunk_wgrad_workspace(self.obs_dim, H, dev)
        self._buf = {}
        self._last = None
        # (FlatAdamState, max_norm) when the learner wants the gradient norm + Adam scalars out of the tail launch
        self.norm_sink, self.norm_done = None, False
        # True while the hi/lo operand copies are known to match the weights (the Adam launch rewrote them)
        self.splits_fresh = False
        self._loss_partials = torch.zeros(148 * 8, dtype=torch.float64, device=dev)     # scratch of the fused loss epilogue
        self._loss_ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.fork_wgrad = os.environ.get("XB_FORK_WGRAD", "0") == "1"
        self._side = None

    # every tensor below is read at launch time: parameters may have been re-pointed (FlatAdamState) since __init__
    def _head_a(self):
        """(weight, bias) of the actor head as the kernels see it."""
        return (self.w2f, self.b2f) if self.fold3 else (self.la2.weight.data, self.la2.bias.data)

    def fold_head(self):
        if self.fold3:
            ops.head3_fold(self.la2.weight.data, self.la2.bias.data, self.w2f, self.b2f)

    def refresh_weights(self):
        self.splits_fresh = True
        self.fold_head()
        ops.dense_split_weights2(self.la1.weight.data, self.wa_hi, self.wa_lo, self.lc1.weight.data, self.wc_hi, self.wc_lo,
                                 self.wt_hi, self.wt_lo)

    def _buffers(self, B):
        b = self._buf.get(B)
        if b is None:
            e = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
            b = dict(h1=e(B, self.H), ya=e(B, self.H), yc=e(B, self.H), act=e(B, self.A), v=e(B, 1), dz1=None)
            b["signs"] = torch.empty(B, 2 * self.H // 32, dtype=torch.int32, device=self.device) if self.sign_bits else None
            # trunk sign words (written by whichever kernel produces h1): the dgrad epilogue's mask without the h1 tiles.
            # Opt-in (XB_H1_SIGNS=1): measured at 65 536 x 128, dgrad 34.5 -> 32.5 us but the gather + trunk launch that has to
            # produce the words 16.4 -> 18.5 (shuffle-packed) / 20.5 us (ballots) — a net loss, the h1 mask tiles stay.
            b["h1s"] = torch.empty(B, self.H // 32, dtype=torch.int32, device=self.device) if self.h1_signs else None
            if self.fold3:      # the kernels write two logits; the reported third one is 0
                b["act2"] = e(B, 2)
                b["act"].zero_()
            self._buf[B] = b
        return b

    # ---------------------------------------------------------------------------------------------- stages (one launch each)
    def stage_trunk(self, obs, b):
        ops.mlp_trunk_fwd(obs, self.l0.weight.data, self.l0.bias.data, self.slope, b["h1"], h1_signs=b.get("h1s"))

    def can_skip_y(self, softmax_pair=False):
        """True if a backward with these head gradients reads the hidden activations only through their sign words (sign-word
        dgrad + binary-form wgrad): the training forward then need not write ya / yc at all (67 MB per 65 536-row minibatch)."""
        return self.sign_dgrad and self.bin_wgrad and self._rank1(softmax_pair)

    def stage_hidden(self, b, loss=None, keep_y=True, obs=None):
        """Actor + critic hidden layers and heads in one launch.  `loss` (dict: scal, adv_stats, adv_count, clip_range,
        vf_coef, ent_coef, inv_batch, logstd, scalars, dlogstd): also the PPO loss forward + backward, fused into the
        kernel's epilogue — dL/d(act_out) lands in b["dact"], dL/dv in b["dv"]."""
        hw, hb = self._head_a()
        b["y_valid"] = keep_y
        l0 = (self.wa_hi, self.wa_lo, self.la1.bias.data, b["ya"] if keep_y else None, hw, hb, b["act2"] if self.fold3 else b["act"])
        l1 = (self.wc_hi, self.wc_lo, self.lc1.bias.data, b["yc"] if keep_y else None, self.lc2.weight.data, self.lc2.bias.data, b["v"])
        # the same launch writes the w2-scaled weight operand of the mask-form dgrad that follows (csrc/dense_tc.cu KParams)
        prep = (self.la1.weight.data, self.lc1.weight.data, self.wtm_hi, self.wtm_lo) if self.mask_dgrad else None
        self._mask_ready = prep is not None
        if loss is not None and "dact" not in b:
            B = b["h1"].shape[0]
            b["dact"] = torch.empty(B, self.A, dtype=torch.float32, device=self.device)
            b["dv"] = torch.empty(B, 1, dtype=torch.float32, device=self.device)
        if obs is not None:
            # `obs`: the trunk layer is generated INSIDE this launch (the operand warps also store h1 for the backward kernels):
            # no first-layer launch, no read of h1 here
            lt = None if loss is None else (loss["scal"], loss["adv_stats"], loss["adv_count"], loss["clip_range"], loss["vf_coef"],
                                            loss["ent_coef"], loss["inv_batch"], loss["logstd"], b["dact"], b["dv"],
                                            self._loss_partials, self._loss_ticket, loss["scalars"], loss["dlogstd"])
            ops.mlp_fwd_from_obs_train(obs, self.l0.weight.data, self.l0.bias.data, self.slope, l0, l1, b["h1"], loss=lt,
                                       prep=prep, sign_out=b["signs"])
            if self.fold3 and loss is None:
                b["act"][:, :2].copy_(b["act2"])
            return
        if loss is None:
            ops.dense_fwd2(b["h1"], self.slope, l0, l1, prep=prep, sign_out=b["signs"])
            if self.fold3:
                b["act"][:, :2].copy_(b["act2"])
            return
        ops.dense_fwd2_loss(b["h1"], self.slope, l0, l1, loss["scal"], loss["adv_stats"], loss["adv_count"],
                            loss["clip_range"], loss["vf_coef"], loss["ent_coef"], loss["inv_batch"], loss["logstd"],
                            b["dact"], b["dv"], self._loss_partials, self._loss_ticket, loss["scalars"], loss["dlogstd"],
                            prep=prep, sign_out=b["signs"])

    def stage_dgrad(self, b, dact, dv2, softmax_pair=False):
        """softmax_pair: `dact` [B, 2] are the gradients w.r.t. the two logits of a softmax head (they are opposite), so
        the actor's head gradient is rank-1 like a one-head source and the mask-form operand applies."""
        assert b.get("y_valid", True) or (self.sign_dgrad and b.get("signs") is not None), "forward(keep_y=False) needs the sign-word dgrad"
        if self._mask_ready and (self.A == 1 or (self.A == 2 and softmax_pair)):
            ops.dense_dgrad(b["ya"], dact, self.la2.weight.data, b["yc"], dv2, self.lc2.weight.data, self.wtm_hi, self.wtm_lo,
                            b["h1"], self.slope, b["dz1"], wt_form=1, signs=b["signs"], h1_signs=b.get("h1s"))
            return
        ops.dense_dgrad(b["ya"], dact, self._head_a()[0], b["yc"], dv2, self.lc2.weight.data, self.wt_hi, self.wt_lo,
                        b["h1"], self.slope, b["dz1"], signs=b["signs"], h1_signs=b.get("h1s"))

    def _rank1(self, softmax_pair):
        return (not self.fold3) and (self.A == 1 or (self.A == 2 and softmax_pair))

    def stage_wgrad(self, b, dact, dv2, softmax_pair=False):
        """Per-CTA partial sums only (into ws_wgrad); `stage_tail` finishes them."""
        self._bin_now = self.bin_wgrad and self._rank1(softmax_pair) and b.get("signs") is not None
        if self._bin_now:
            ops.dense_wgrad_bin(b["signs"], dact, self.A, dv2, 1, b["h1"], self.H, self.ws_wgrad)
            return
        assert b.get("y_valid", True), "forward(keep_y=False) needs the binary-form wgrad (rank-1 head gradients)"
        ops.dense_wgrad(b["ya"], dact, self._head_a()[0], b["yc"], dv2, self.lc2.weight.data, b["h1"], self.slope,
                        self.ws_wgrad, None, None, None, None)

    def stage_trunk_wgrad(self, obs, b):
        """Per-block partial sums only (into ws_trunk); `stage_tail` finishes them."""
        ops.mlp_trunk_wgrad(b["dz1"], obs, self.ws_trunk, None, None)

    def stage_tail(self, dls64=None, dls32=None):
        """One launch: all partial sums -> the ten Linear gradients (+ the loss kernel's fp64 log-std gradient -> fp32)
        and, when `norm_sink` is armed, their global norm + the clipped-Adam step scalars (optim workspace)."""
        g = lambda p: p.grad
        norm, self.norm_done = None, False
        # (folded Discrete(3) head: its third gradient row is written after this launch, so the norm cannot be taken here)
        if self.norm_sink is not None and not self.fold3 and (not self.gaussian or dls64 is not None):   # every gradient passes through here
            fl, max_norm = self.norm_sink
            norm = (fl.workspace, fl.step, fl.lr0, fl.end_factor, fl.total_iters, fl.beta1, fl.beta2, max_norm, 1.0, fl.lr,
                    fl.gnorm)
            self.norm_done = True
        bin_form = None
        if self._bin_now:
            bin_form = (self.la1.weight.data, self.la1.bias.data, self.la2.weight.data, self.lc1.weight.data, self.lc1.bias.data,
                        self.lc2.weight.data, self.slope)
        ops.mlp_backward_tail(self.ws_wgrad, self.H, self.H, self.nh_a, 1,
                              (g(self.la1.weight), g(self.la1.bias), g(self.la2.weight), g(self.la2.bias)),
                              (g(self.lc1.weight), g(self.lc1.bias), g(self.lc2.weight), g(self.lc2.bias)),
                              self.ws_trunk, self.obs_dim, g(self.l0.weight), g(self.l0.bias), dls64, dls32, norm=norm,
                              bin_form=bin_form)
        if self.fold3:      # third head row from the two the kernels produced: dL/dz2 = -(dL/dz0 + dL/dz1)
            ops.head3_unfold_grads(g(self.la2.weight), g(self.la2.bias))

    def fwd_from_obs_ok(self):
        """True if the one-launch rollout forward (trunk generated in-kernel) covers this policy."""
        return self.obs_dim <= 4 and self.H <= 128

    def forward_inference(self, obs, norm=None, weights_stable=False):
        """Rollout forward (no activations kept): trunk + both hidden layers + heads in ONE launch for obs_dim <= 4,
        H <= 128 (the trunk layer is generated inside the kernel); otherwise the training forward without refresh.
        norm = (state_new, state_old, n_new_rows, clip): `obs` is raw, the kernel normalises it (one-launch form only).
        weights_stable: the launch right before this one writes no weights (lets the kernel's set-up overlap it)."""
        if not self.fwd_from_obs_ok():
            assert norm is None, "normalise the observations before the multi-launch forward"
            return self.forward(obs, refresh=False)
        B = obs.shape[0]
        b = self._buffers(B)
        ops.mlp_fwd_from_obs(obs, self.l0.weight.data, self.l0.bias.data, self.slope,
                             (self.wa_hi, self.wa_lo, self.la1.bias.data, None, *self._head_a(), b["act2"] if self.fold3 else b["act"]),
                             (self.wc_hi, self.wc_lo, self.lc1.bias.data, None, self.lc2.weight.data, self.lc2.bias.data, b["v"]),
                             norm=norm, weights_stable=weights_stable)
        if self.fold3:
            b["act"][:, :2].copy_(b["act2"])
        return b["act"], b["v"][:, 0]

    # ---------------------------------------------------------------------------------------------- forward / backward
    def forward(self, obs, refresh=True, trunk_done=False, loss=None, keep_y=True, trunk_in_kernel=False):
        """obs: CUDA fp32 [B, obs_dim] with contiguous rows (a column slice of the float4 observation rows is fine).
        Returns (act_out [B, A], v [B]); the activations stay in per-batch-size buffers for `backward`."""
        B = obs.shape[0]
        b = self._buffers(B)
        if refresh:
            self.refresh_weights()
        else:
            self.fold_head()                # the head parameters move with every optimiser step
        in_kernel = trunk_in_kernel and not trunk_done and self.fwd_from_obs_ok()
        if not trunk_done and not in_kernel:    # (trunk_done: the gather kernel already produced h1, xb_gather_trunk_fwd)
            self.stage_trunk(obs, b)
        self.stage_hidden(b, loss, keep_y, obs=obs if in_kernel else None)
        self._last = (obs, b)
        return b["act"], b["v"][:, 0]

    def dist_params(self, act_out):
        if self.gaussian:
            return _OutParams(mu=act_out, std=None)
        return _OutParams(logits=act_out)

    def backward(self, dact, dv, dls64=None, dls32=None, softmax_pair=False):
        """dact [B, A], dv [B] = dL/d(act_out), dL/d(v) for the most recent `forward`; writes .grad of the ten
        Linear parameters (views of the flat gradient buffer) — plain stores, no accumulation.  dls64 -> dls32: the
        loss kernel's fp64 log-std gradient, converted in the same tail launch."""
        obs, b = self._last
        B = obs.shape[0]
        if b["dz1"] is None:
            b["dz1"] = torch.empty(B, self.H, dtype=torch.float32, device=self.device)
        dv2 = dv.reshape(B, 1)
        if self.fold3:                      # (d0, d1) drive the folded head; d2 = -(d0 + d1) is implied
            if "dact2" not in b:
                b["dact2"] = torch.empty(B, 2, dtype=torch.float32, device=self.device)
            b["dact2"].copy_(dact[:, :2])
            dact = b["dact2"]
        if self.fork_wgrad:
            # opt-in experiment (XB_FORK_WGRAD=1): dgrad -> trunk wgrad and the (independent) hidden-layer wgrad as two graph
            # branches, so that wgrad CTAs could fill the SMs dgrad's last, partially filled round of tiles leaves idle.
            # Measured: no gain (epoch graph 1.2255 vs 1.2235 ms) — both kernels are one-CTA-per-SM persistent loops with
            # static work shares, so the branch that starts late still ends late.
            cur = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(cur)
            self.stage_dgrad(b, dact, dv2, softmax_pair)
            with torch.cuda.stream(self._side):
                self.stage_wgrad(b, dact, dv2, softmax_pair)
            self.stage_trunk_wgrad(obs, b)
            cur.wait_stream(self._side)
        else:
            self.stage_dgrad(b, dact, dv2, softmax_pair)
            self.stage_wgrad(b, dact, dv2, softmax_pair)
            self.stage_trunk_wgrad(obs, b)
        self.stage_tail(dls64, dls32)
