"""FusedActorCritic (tcgen05 dense kernels + SIMT trunk) against torch autograd on the same parameters.

The reference computes these with torch modules (xuance/torch/policies/gaussian.py:73-77, categorical.py:81-85) and
`loss.backward()` (ppoclip_learner.py:47).  Bars: north_star's 1e-4 relative for outputs and for every param.grad."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    ref = ref.double()
    rms = ref.pow(2).mean().sqrt().clamp_min(1e-30)
    return float(((a.double() - ref).abs() / torch.maximum(ref.abs(), rms)).max())


@pytest.mark.parametrize("env_id,hidden,B", [("Pendulum-v1", 128, 65536), ("CartPole-v1", 128, 8192),
                                              ("Pendulum-v1", 256, 4096), ("CartPole-v1", 128, 2048 + 77)])
def test_fused_forward_backward_match_torch_autograd(env_id, hidden, B):
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.learner import FlatAdamState
    from xuanpolicy_b200.policies import make_policy
    torch.backends.cuda.matmul.allow_tf32 = False
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(hidden,), device="cuda", seed=3)
    with torch.no_grad():                      # biases are zero-initialised by the reference; make them count
        for p in policy.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    flat = FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
    assert FusedActorCritic.plan(policy) is not None
    fused = FusedActorCritic(policy)
    g = torch.Generator(device="cuda").manual_seed(B)
    obs4 = torch.randn(B, 4, device="cuda", generator=g)          # float4 observation rows, as the rollout buffer holds them
    obs = obs4[:, :obs_space.shape[0]]
    act_out, v = fused.forward(obs)
    A = act_out.shape[1]
    dact = torch.randn(B, A, device="cuda", generator=g) / B
    dv = torch.randn(B, device="cuda", generator=g) / B
    # fp64 reference of the same network (the reference's module graph: mlp.py:40-51, gaussian.py:17-24,41-48).
    # LeakyReLU' is discontinuous at 0: an fp64 graph and an fp32 one disagree on the sign of the ~1e-6 fraction of
    # pre-activations with |z| < 1e-6, and ONE flipped unit moves a weight-gradient entry by ~1/sqrt(B) of its size.
    # The reference therefore takes the activation pattern (not the values) from the kernel's saved activations.
    buf = fused._buf[B]
    pat = iter([buf["h1"], buf["ya"], buf["yc"]])
    lk = lambda t: t * torch.where(next(pat) > 0, 1.0, 0.01).double()
    names = [n for n, _ in policy.named_parameters()]
    P = {n: q.detach().double().requires_grad_(True) for n, q in policy.named_parameters()}
    head = "actor.mu" if "actor.mu.0.weight" in P else "actor.model"
    h1 = lk(obs.double() @ P["representation.model.0.weight"].t() + P["representation.model.0.bias"])
    ya = lk(h1 @ P[head + ".0.weight"].t() + P[head + ".0.bias"])
    yc = lk(h1 @ P["critic.model.0.weight"].t() + P["critic.model.0.bias"])
    ref_act = ya @ P[head + ".2.weight"].t() + P[head + ".2.bias"]
    ref_v = (yc @ P["critic.model.2.weight"].t() + P["critic.model.2.bias"])[:, 0]
    assert _rel(act_out, ref_act) < 1e-4 and _rel(v, ref_v) < 1e-4
    grads = torch.autograd.grad([ref_act, ref_v], [P[n] for n in names], [dact.double(), dv.double()], allow_unused=True)
    fused.backward(dact, dv)
    torch.cuda.synchronize()
    mine = dict(policy.named_parameters())
    worst = 0.0
    for n, gref in zip(names, grads):
        if gref is None:
            continue                            # logstd: its gradient comes from the loss kernel, not from the MLP
        e = _rel(mine[n].grad, gref)
        print("   %-32s %.2e" % (n, e))
        worst = max(worst, e)
    for n, gref in zip(names, grads):
        if gref is not None:
            assert _rel(mine[n].grad, gref) < 1e-4, n
    print("%s H=%d B=%d: worst param.grad error %.2e" % (env_id, hidden, B, worst))


def test_fused_update_matches_torch_update_one_step():
    """One native update (gather -> forward -> loss -> backward -> clip+Adam) with and without the fused MLP."""
    from xuanpolicy_b200.configs import build_ppo
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    for fused in ("1", "0"):
        agent = build_ppo("Pendulum-v1", parallels=1024, n_steps=32, seed=5, use_cuda_graphs=False, shuffle="device")
        f = agent.learner._fused
        assert f is not None
        agent.learner._fused = None            # identical (torch) rollouts, so both variants update on the same buffer
        torch.manual_seed(11)
        agent._rollout()
        agent.learner._fused = f if fused == "1" else None
        agent.memory.ptr, agent.memory.size = 0, agent.n_steps
        torch.manual_seed(12)
        idx = torch.randperm(agent.buffer_size, device="cuda")[:agent.batch_size]
        agent.learner.update_from_buffer(agent.memory, idx)
        torch.cuda.synchronize()
        res[fused] = (agent.learner._flat.flat_grad.clone(), agent.learner._flat.flat_param.clone(),
                      agent.learner._scalars.clone())
    g1, p1, s1 = res["1"]
    g0, p0, s0 = res["0"]
    assert _rel(s1, s0) < 1e-4                                   # loss scalars
    # gradients: equal to 1e-4 except where a LeakyReLU unit sits within float rounding of 0 and the two forwards
    # (cuBLAS SIMT vs tensor-core split) put it on different sides — a handful of units per 65 536 x 384 activations
    err = (g1.double() - g0.double()).abs() / torch.maximum(g0.double().abs(), g0.double().pow(2).mean().sqrt())
    assert float((err < 1e-4).double().mean()) > 0.98, float((err < 1e-4).double().mean())
    assert float((g1 - g0).norm() / g0.norm()) < 1e-3
    # parameters after the clipped Adam step: the first Adam step is lr * g / (|g| + eps), i.e. sign-like, so a
    # flipped unit can move a near-zero-gradient parameter by a visible fraction of lr; bound the step difference by 5% of lr
    assert float((p1 - p0).abs().max()) < 0.05 * 4e-4


@pytest.mark.parametrize("env_id,B", [("Pendulum-v1", 8192), ("CartPole-v1", 4096 + 33)])
def test_rollout_forward_from_observations_matches_training_forward(env_id, B):
    """xb_mlp_fwd_from_obs (trunk generated inside the tensor-core kernel, heads only) == trunk kernel + dense_fwd2."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.policies import make_policy
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(128,), device="cuda", seed=4)
    with torch.no_grad():
        for p in policy.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    fused = FusedActorCritic(policy)
    obs = torch.randn(B, 4, device="cuda")[:, :obs_space.shape[0]]
    a1, v1 = fused.forward(obs)
    a1, v1 = a1.clone(), v1.clone()
    a2, v2 = fused.forward_inference(obs)
    torch.cuda.synchronize()
    assert _rel(a2, a1) < 2e-5 and _rel(v2, v1) < 2e-5


def test_tensor_core_update_matches_reference_golden():
    """The whole native update with the MLP on the tcgen05 kernels (forward, loss kernel, dgrad/wgrad, clip + Adam)
    against the REFERENCE's own `PPOCLIP_Learner.update` output (tests/golden/loss_gauss_h128.npz, generated by running
    the unmodified reference): info scalars, every param.grad before clipping (1e-4), parameters after the step (1e-5)."""
    import xuanpolicy_b200 as xb
    from tests.helpers import load_golden, rel_close
    from xuanpolicy_b200 import policies, spaces
    torch.backends.cuda.matmul.allow_tf32 = False
    g = load_golden("loss_gauss_h128")
    m = g["meta"]
    dev = "cuda"
    for tag, clip in (("noclip", False), ("clip", True)):
        pol = policies.make_policy(spaces.Box(-1, 1, (3,)), spaces.Box(-2.0, 2.0, (1,)), hidden=(m["hidden"],), device=dev)
        pol.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")})
        opt = torch.optim.Adam(pol.parameters(), 4e-4, eps=1e-5)
        sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
        learner = xb.PPOCLIP_Learner(pol, opt, sched, dev, "/tmp", vf_coef=m["vf_coef"], ent_coef=m["ent_coef"],
                                     clip_range=m["clip_range"], clip_grad_norm=m["clip_grad_norm"], use_grad_clip=clip)
        flat = learner.enable_fused_optimizer()
        fused = learner._fused
        assert fused is not None
        t = lambda k: torch.as_tensor(g[k], device=dev).float().contiguous()
        B = g["ret"].shape[0]
        act_out, v = fused.forward(t("obs"))
        learner._loss_backward(fused.dist_params(act_out), v, t("act").reshape(B, -1), t("ret"), t("adv"), t("old_logp"),
                               t("val"), 1.0 / B, flat=flat, fused=fused)
        torch.cuda.synchronize()
        for k, p in pol.named_parameters():
            ok, err = rel_close(p.grad.cpu().numpy(), g["grad_noclip/%s" % k], 1e-4)
            assert ok, (tag, k, err)
        learner.stage_optimizer()
        info = learner.info(B)
        for k in ("actor-loss", "critic-loss", "entropy", "predict_value"):
            ref = float(g["info_%s/%s" % (tag, k)])
            assert abs(float(info[k]) - ref) <= 1e-4 * max(1.0, abs(ref)), (k, float(info[k]), ref)
        for k, p in pol.named_parameters():
            ok, err = rel_close(p.detach().cpu().numpy(), g["p1_%s/%s" % (tag, k)], 1e-5)
            assert ok, (tag, k, err)


@pytest.mark.parametrize("obs_dim,H,B", [(3, 128, 65536), (4, 128, 5000), (3, 256, 4097), (2, 128, 2048)])
def test_gather_trunk_fwd_equals_gather_then_trunk(obs_dim, H, B):
    """xb_gather_trunk_fwd (gather + first MLP layer in one launch) is bit-identical to xb_gather_records followed by
    xb_mlp_trunk_fwd: gathered observations, packed scalars, advantage statistics and h1."""
    from xuanpolicy_b200 import ops
    T, N = 64, 1031
    gen = torch.Generator(device="cuda").manual_seed(obs_dim * 1000 + H)
    rec = torch.randn((T * N, 8), device="cuda", generator=gen)
    idx = torch.randperm(T * N, device="cuda", generator=gen)[:B].contiguous()
    w0, b0 = torch.randn((H, obs_dim), device="cuda", generator=gen), torch.randn(H, device="cuda", generator=gen)
    obs_a, scal_a = torch.empty((B, obs_dim), device="cuda"), torch.empty((B, 4), device="cuda")
    st_a, h_a = torch.zeros(2, dtype=torch.float64, device="cuda"), torch.empty((B, H), device="cuda")
    ops.gather_records(idx, T, N, rec, obs_dim, obs_a, scal_a, stats=st_a)
    ops.mlp_trunk_fwd(obs_a, w0, b0, 0.01, h_a)
    obs_b, scal_b = torch.empty_like(obs_a), torch.empty_like(scal_a)
    st_b, h_b = torch.zeros_like(st_a), torch.empty_like(h_a)
    hs_a = torch.zeros((B, H // 32), dtype=torch.int32, device="cuda")
    hs_b = torch.zeros_like(hs_a)
    ops.mlp_trunk_fwd(obs_a, w0, b0, 0.01, h_a, h1_signs=hs_a)
    ops.gather_trunk_fwd(idx, T, N, rec, obs_dim, w0, b0, 0.01, obs_b, scal_b, h_b, stats=st_b, h1_signs=hs_b)
    assert torch.equal(obs_a, obs_b) and torch.equal(scal_a, scal_b) and torch.equal(h_a, h_b)
    assert torch.allclose(st_a, st_b, rtol=1e-12, atol=1e-9)
    # both kernels' sign words of h1: bit l of word 4 s + e = h1[:, 128 s + 4 l + e] > 0
    bits = (((hs_b.long() & 0xFFFFFFFF)[:, :, None] >> torch.arange(32, device="cuda")) & 1).bool()      # [B, H/32, l]
    bits = bits.reshape(B, H // 128, 4, 32).permute(0, 1, 3, 2).reshape(B, H)                             # [s][l][e] -> feature
    assert torch.equal(hs_a, hs_b) and torch.equal(bits, h_b > 0)


@pytest.mark.parametrize("env_id,hidden,B,clip,advnorm", [("Pendulum-v1", 128, 65536, 0.2, True), ("CartPole-v1", 128, 5000, 0.2, True),
                                                          ("Pendulum-v1", 256, 4096, 0.2, False), ("CartPole-v1", 128, 2048, 0.0, True)])
def test_loss_fused_into_the_forward_epilogue_equals_the_loss_kernel(env_id, hidden, B, clip, advnorm):
    """xb_dense_fwd2_loss: the PPO loss forward + backward computed in the epilogue of the hidden-layer launch gives the same
    dL/d(head outputs), log scalars and log-std gradient as xb_dense_fwd2 followed by xb_ppo_loss_* on its outputs."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import ops
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.learner import FlatAdamState
    from xuanpolicy_b200.policies import make_policy
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(hidden,), device="cuda", seed=4)
    with torch.no_grad():
        for p in policy.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
    fused = FusedActorCritic(policy)
    g = torch.Generator(device="cuda").manual_seed(B + 1)
    obs = torch.randn(B, 4, device="cuda", generator=g)[:, :obs_space.shape[0]]
    act_out, v = fused.forward(obs)
    act_out, v = act_out.clone(), v.clone().contiguous()
    gauss = fused.gaussian
    scal = torch.empty(B, 4, device="cuda")                        # {act, old_logp, adv, ret}
    if gauss:
        logstd = policy.actor.logstd.detach()
        scal[:, 0] = act_out[:, 0] + logstd.exp() * torch.randn(B, device="cuda", generator=g)
        logp = -((scal[:, 0] - act_out[:, 0]) ** 2) / (2 * (2 * logstd).exp()) - logstd - 0.9189385332046727
    else:
        logstd = None
        scal[:, 0] = torch.randint(0, 2, (B,), device="cuda", generator=g).float()
        logp = torch.log_softmax(act_out, -1).gather(1, scal[:, :1].long())[:, 0]
    scal[:, 1] = logp + 0.1 * torch.randn(B, device="cuda", generator=g)
    scal[:, 2] = torch.randn(B, device="cuda", generator=g) * 2 + 0.3
    scal[:, 3] = v + torch.randn(B, device="cuda", generator=g)
    stats = torch.stack([scal[:, 2].double().sum(), (scal[:, 2].double() ** 2).sum()]) if advnorm else None
    kw = dict(clip_range=clip, vf_coef=0.25, ent_coef=0.01, inv_batch=1.0 / B, adv_stats=stats, adv_count=B, packed=scal)
    s_ref = torch.zeros(8, dtype=torch.float64, device="cuda")
    dv_ref = torch.empty(B, device="cuda")
    dact_ref = torch.empty_like(act_out)
    if gauss:
        dls_ref = torch.zeros(1, dtype=torch.float64, device="cuda")
        ops.ppo_loss_gaussian(act_out, logstd, v, None, None, None, None, dact_ref, dls_ref, dv_ref, s_ref, **kw)
    else:
        ops.ppo_loss_categorical(act_out, v, None, None, None, None, dact_ref, dv_ref, s_ref, **kw)
    s_new = torch.full((8,), 7.0, dtype=torch.float64, device="cuda")
    dls_new = torch.zeros(1, dtype=torch.float64, device="cuda")
    loss = dict(scal=scal, adv_stats=stats, adv_count=B, clip_range=clip, vf_coef=0.25, ent_coef=0.01, inv_batch=1.0 / B,
                logstd=logstd, scalars=s_new, dlogstd=dls_new if gauss else None)
    for _ in range(2):                                             # twice: the ticket re-arms itself
        fused.forward(obs, loss=loss)
    b = fused._last[1]
    torch.cuda.synchronize()
    assert torch.equal(b["act"], act_out) and torch.equal(b["v"][:, 0], v)
    assert torch.allclose(b["dact"], dact_ref, rtol=1e-6, atol=1e-12) and torch.allclose(b["dv"][:, 0], dv_ref, rtol=1e-6, atol=1e-12)
    # sums of B float terms whose last bits may differ between the two translation units (FMA contraction): 1e-6 of the scale
    scale = float(s_ref.abs().max())
    assert torch.allclose(s_new, s_ref, rtol=1e-6, atol=1e-6 * scale), (s_new, s_ref)
    if gauss:
        assert torch.allclose(dls_new, dls_ref, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", ["loss_gauss_h128", "loss_cat_h128"])
def test_fused_loss_epilogue_update_matches_reference_golden(name):
    """The native update with the loss INSIDE the forward kernel's epilogue (xb_dense_fwd2_loss), the norm inside the
    backward tail launch and the operand split inside the Adam launch, against the REFERENCE's own PPOCLIP_Learner.update
    (golden generated by running the unmodified reference): info scalars, every param.grad before clipping (1e-4),
    parameters after the clipped Adam step (1e-5)."""
    import xuanpolicy_b200 as xb
    from tests.helpers import load_golden, rel_close
    from xuanpolicy_b200 import policies, spaces
    torch.backends.cuda.matmul.allow_tf32 = False
    g = load_golden(name)
    m = g["meta"]
    dev = "cuda"
    t = lambda k: torch.as_tensor(g[k], device=dev).float().contiguous()
    B = g["ret"].shape[0]
    scal = torch.stack([t("act").reshape(B), t("old_logp"), t("adv"), t("ret")], dim=1).contiguous()   # packed minibatch rows
    for tag, clip in (("noclip", False), ("clip", True)):
        if m["discrete"]:
            pol = policies.make_policy(spaces.Box(-1, 1, (4,)), spaces.Discrete(2), hidden=(m["hidden"],), device=dev)
        else:
            pol = policies.make_policy(spaces.Box(-1, 1, (3,)), spaces.Box(-2.0, 2.0, (1,)), hidden=(m["hidden"],), device=dev)
        pol.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")})
        opt = torch.optim.Adam(pol.parameters(), 4e-4, eps=1e-5)
        sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
        learner = xb.PPOCLIP_Learner(pol, opt, sched, dev, "/tmp", vf_coef=m["vf_coef"], ent_coef=m["ent_coef"],
                                     clip_range=m["clip_range"], clip_grad_norm=m["clip_grad_norm"], use_grad_clip=clip)
        flat = learner.enable_fused_optimizer()
        fused = learner._fused
        gauss = not m["discrete"]
        dls64 = torch.zeros(1, dtype=torch.float64, device=dev) if gauss else None
        loss = dict(scal=scal, adv_stats=None, adv_count=B, clip_range=m["clip_range"], vf_coef=m["vf_coef"],
                    ent_coef=m["ent_coef"], inv_batch=1.0 / B, logstd=pol.actor.logstd.detach() if gauss else None,
                    scalars=learner._scalars, dlogstd=dls64)
        fused.norm_sink = (flat, m["clip_grad_norm"] if clip else 0.0)
        fused.forward(t("obs"), loss=loss)
        b = fused._last[1]
        if gauss:
            dls32 = flat.grad_views[[id(q) for q in flat.params].index(id(pol.actor.logstd))]
            fused.backward(b["dact"], b["dv"], dls64, dls32)
        else:
            fused.backward(b["dact"], b["dv"])
        assert fused.norm_done
        torch.cuda.synchronize()
        for k, p in pol.named_parameters():
            ok, err = rel_close(p.grad.cpu().numpy(), g["grad_noclip/%s" % k], 1e-4)
            assert ok, (tag, k, err)
        learner.stage_optimizer()                    # Adam (+ operand split) from the scalars the tail launch derived
        assert fused.splits_fresh
        info = learner.info(B)
        for k in ("actor-loss", "critic-loss", "entropy", "predict_value", "clip_ratio"):
            ref = float(g["info_%s/%s" % (tag, k)])
            assert abs(float(info[k]) - ref) <= 1e-4 * max(1.0, abs(ref)), (k, float(info[k]), ref)
        for k, p in pol.named_parameters():
            ok, err = rel_close(p.detach().cpu().numpy(), g["p1_%s/%s" % (tag, k)], 1e-5)
            assert ok, (tag, k, err)
        # the operand copies the Adam launch wrote equal a fresh split of the updated weights
        hi, lo = fused.wa_hi.clone(), fused.wa_lo.clone()
        fused.refresh_weights()
        assert torch.equal(hi, fused.wa_hi) and torch.equal(lo, fused.wa_lo)


@pytest.mark.parametrize("env_id,hidden,B", [("Pendulum-v1", 128, 65536), ("CartPole-v1", 128, 4096 + 33), ("Pendulum-v1", 256, 4096)])
def test_mask_form_dgrad_equals_general_form(env_id, hidden, B):
    """dgrad with the w2-scaled weight operand and the two-constants-per-row A operand (csrc/dense_tc.cu KParams::mask_form;
    one head per source, or a softmax pair of logit gradients) against the general form (dz generated element by element):
    same 3xTF32 product, only the association of the fp32 roundings differs."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.learner import FlatAdamState
    from xuanpolicy_b200.policies import make_policy
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(hidden,), device="cuda", seed=4)
    FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
    fused = FusedActorCritic(policy)
    g = torch.Generator(device="cuda").manual_seed(B)
    obs = torch.randn(B, 4, device="cuda", generator=g)[:, :obs_space.shape[0]]
    fused.forward(obs)
    assert fused._mask_ready
    A = fused.A
    d0 = torch.randn(B, 1, device="cuda", generator=g) / B
    dact = d0 if A == 1 else torch.cat([d0, -d0], dim=1).contiguous()
    dv2 = (torch.randn(B, 1, device="cuda", generator=g) / B).contiguous()
    b = fused._last[1]
    b["dz1"] = torch.empty(B, hidden, device="cuda")
    fused.stage_dgrad(b, dact, dv2, softmax_pair=True)
    masked = b["dz1"].clone()
    fused._mask_ready = False
    fused.stage_dgrad(b, dact, dv2, softmax_pair=True)
    general = b["dz1"].clone()
    # fp64 reference from the saved activations
    w = lambda m: m.weight.detach().double()
    slope = fused.slope
    mk = lambda y: torch.where(y > 0, 1.0, slope).double()
    dza = (dact.double() @ w(fused.la2)) * mk(b["ya"])
    dzc = (dv2.double() @ w(fused.lc2)) * mk(b["yc"])
    ref = (dza @ w(fused.la1) + dzc @ w(fused.lc1)) * mk(b["h1"])
    em, eg, ed = _rel(masked, ref), _rel(general, ref), _rel(masked, general.double())
    print("H=%d B=%d: mask form %.2e, general form %.2e vs fp64; mask vs general %.2e" % (hidden, B, em, eg, ed))
    bar = 2e-5 if hidden == 128 else 1e-4           # the tensor core truncates its accumulator: the error grows with K = 2H
    assert em < bar and eg < bar and ed < bar, (em, eg, ed)


@pytest.mark.parametrize("env_id,B", [("Acrobot-v1", 8192 + 5), ("MountainCar-v0", 4096)])
def test_three_logit_heads_run_folded_on_the_two_head_kernels(env_id, B):
    """Discrete(3) policies (Acrobot-v1, MountainCar-v0 with the reference's 8-float stacked observation): the actor head is
    folded onto the two-head tensor-core kernels through the softmax's shift invariance (csrc/mlp_trunk.cu xb_head3_fold).
    Log-probabilities equal the torch modules'; for gradients that come from a softmax loss (sum over the logits = 0) every
    parameter gradient equals torch autograd's, incl. the third head row reconstructed as -(row 0 + row 1)."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.learner import FlatAdamState
    from xuanpolicy_b200.policies import make_policy
    torch.backends.cuda.matmul.allow_tf32 = False
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(128,), device="cuda", seed=5)
    with torch.no_grad():
        for p in policy.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
    assert FusedActorCritic.plan(policy) is not None
    fused = FusedActorCritic(policy)
    assert fused.fold3 and fused.A == 3 and fused.obs_dim == obs_space.shape[0]
    g = torch.Generator(device="cuda").manual_seed(B)
    obs = torch.randn(B, 8, device="cuda", generator=g)[:, :obs_space.shape[0]]
    act_out, v = fused.forward(obs)
    assert act_out.shape == (B, 3) and float(act_out[:, 2].abs().max()) == 0.0
    buf = fused._buf[B]
    pat = iter([buf["h1"], buf["ya"], buf["yc"]])
    lk = lambda t: t * torch.where(next(pat) > 0, 1.0, 0.01).double()
    names = [n for n, _ in policy.named_parameters()]
    P = {n: q.detach().double().requires_grad_(True) for n, q in policy.named_parameters()}
    h1 = lk(obs.double() @ P["representation.model.0.weight"].t() + P["representation.model.0.bias"])
    ya = lk(h1 @ P["actor.model.0.weight"].t() + P["actor.model.0.bias"])
    yc = lk(h1 @ P["critic.model.0.weight"].t() + P["critic.model.0.bias"])
    ref_logits = ya @ P["actor.model.2.weight"].t() + P["actor.model.2.bias"]
    ref_v = (yc @ P["critic.model.2.weight"].t() + P["critic.model.2.bias"])[:, 0]
    assert _rel(torch.log_softmax(act_out.double(), 1), torch.log_softmax(ref_logits, 1)) < 1e-4 and _rel(v, ref_v) < 1e-4
    # a softmax-loss gradient: rows sum to zero
    d = torch.randn(B, 3, device="cuda", generator=g) / B
    dact = (d - d.mean(1, keepdim=True)).contiguous()
    dv = torch.randn(B, device="cuda", generator=g) / B
    grads = torch.autograd.grad([ref_logits, ref_v], [P[n] for n in names], [dact.double(), dv.double()])
    fused.backward(dact, dv, softmax_pair=True)
    torch.cuda.synchronize()
    mine = dict(policy.named_parameters())
    for n, gref in zip(names, grads):
        assert _rel(mine[n].grad, gref) < 1e-4, (n, _rel(mine[n].grad, gref))
    # the rollout forward (multi-launch for these observation widths) reports the same folded logits
    a_inf, v_inf = fused.forward_inference(obs)
    assert _rel(torch.log_softmax(a_inf.double(), 1), torch.log_softmax(ref_logits, 1)) < 1e-4 and _rel(v_inf, ref_v) < 1e-4


@pytest.mark.parametrize("env_id", ["Acrobot-v1", "MountainCar-v0"])
def test_native_agent_on_three_action_envs_uses_the_tensor_core_mlp(env_id):
    """Row f4 end to end: the device-resident PPO loop on the Discrete(3) envs with the yaml defaults (obs / reward
    normalisation on) — fused rollout step, folded tensor-core MLP for the rollout forward and the updates — against the same
    loop with the torch MLP (XB_FUSED_MLP=0): same trajectories (discrete actions), parameters within fp32 tolerance."""
    import os
    from xuanpolicy_b200.configs import build_ppo
    out = []
    for fused in ("1", "0"):
        os.environ["XB_FUSED_MLP"] = fused
        try:
            agent = build_ppo(env_id, parallels=1024, n_steps=8, n_epoch=2, n_minibatch=2, shuffle="device", seed=6,
                              use_obsnorm=True, use_rewnorm=True)
        finally:
            os.environ.pop("XB_FUSED_MLP", None)
        assert (agent.learner._fused is not None) == (fused == "1") and agent._fused_step and agent._fused_norm
        if fused == "1":
            assert agent.learner._fused.fold3 and agent.batch_size >= agent.learner._fused.MIN_ROWS
        info = agent.train(8)
        out.append((agent.memory._act.clone(), agent.memory._obs.clone(), agent.learner._flat.flat_param.clone(), info))
    assert torch.equal(out[0][0], out[1][0])                                   # same actions -> same trajectory
    assert torch.allclose(out[0][1], out[1][1], atol=1e-5, rtol=1e-5)
    assert torch.allclose(out[0][2], out[1][2], atol=2e-5, rtol=2e-4), (out[0][2] - out[1][2]).abs().max()
    assert abs(out[0][3]["critic-loss"] - out[1][3]["critic-loss"]) <= 1e-4 * max(1.0, abs(out[1][3]["critic-loss"]))


@pytest.mark.parametrize("env_id,B", [("Pendulum-v1", 65536), ("CartPole-v1", 2048 + 77)])
def test_sign_words_match_activations_and_dgrad_is_bit_identical(env_id, B):
    """The forward's activation sign words (one bit per hidden activation) are exactly (y > 0), and the dgrad kernel fed with
    them produces the bit-identical dZ1 of the kernel that TMA-loads the activation tiles (mask form and plain form); the same for
    the trunk's sign words in place of the h1 mask tiles."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import ops
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.policies import make_policy
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(128,), device="cuda", seed=5)
    fused = FusedActorCritic(policy)
    assert fused.sign_bits
    g = torch.Generator(device="cuda").manual_seed(B)
    obs = torch.randn(B, 4, device="cuda", generator=g)[:, :obs_space.shape[0]]
    act_out, v = fused.forward(obs)
    b = fused._buf[B]
    words = b["signs"].view(B, 2, 4).long() & 0xFFFFFFFF
    for s, y in enumerate((b["ya"], b["yc"])):
        bits = ((words[:, s, :, None] >> torch.arange(32, device="cuda")) & 1).reshape(B, 128).bool()
        assert torch.equal(bits, y > 0)
    A = act_out.shape[1]
    dact = torch.randn(B, A, device="cuda", generator=g) / B
    if A == 2:
        dact[:, 1] = -dact[:, 0]            # softmax pair (what the mask form assumes)
    dv2 = (torch.randn(B, device="cuda", generator=g) / B).reshape(B, 1)
    # trunk sign words (the dgrad epilogue's mask, opt-in): bit l of word e = h1[:, 4 l + e] > 0
    b["h1s"] = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
    ops.mlp_trunk_fwd(obs, fused.l0.weight.data, fused.l0.bias.data, fused.slope, b["h1"], h1_signs=b["h1s"])
    hwords = b["h1s"].view(B, 4).long() & 0xFFFFFFFF
    hbits = ((hwords[:, :, None] >> torch.arange(32, device="cuda")) & 1).bool().permute(0, 2, 1).reshape(B, 128)
    assert torch.equal(hbits, b["h1"] > 0)
    outs = []
    for signs, h1s in ((b["signs"], b["h1s"]), (b["signs"], None), (None, None)):
        for form in (1, 0):
            dz1 = torch.full((B, 128), float("nan"), device="cuda")
            if form:
                ops.dense_dgrad(b["ya"], dact, fused.la2.weight.data, b["yc"], dv2, fused.lc2.weight.data, fused.wtm_hi,
                                fused.wtm_lo, b["h1"], fused.slope, dz1, wt_form=1, signs=signs, h1_signs=h1s)
            else:
                ops.dense_dgrad(b["ya"], dact, fused.la2.weight.data, b["yc"], dv2, fused.lc2.weight.data, fused.wt_hi,
                                fused.wt_lo, b["h1"], fused.slope, dz1, signs=signs, h1_signs=h1s)
            outs.append(dz1)
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[4]) and torch.equal(outs[1], outs[5])
    assert torch.equal(outs[2], outs[4]) and torch.equal(outs[3], outs[5])


@pytest.mark.parametrize("env_id,B", [("Pendulum-v1", 65536), ("CartPole-v1", 8192 + 33), ("Pendulum-v1", 2048 + 5)])
def test_binary_form_wgrad_matches_the_three_mma_form(env_id, B):
    """xb_dense_wgrad_bin + xb_mlp_backward_tail_bin (0/1 A operand from the sign words, head-weight gradient rebuilt from the
    weight-gradient partials) against xb_dense_wgrad + xb_mlp_backward_tail on the same inputs: every hidden-layer and head
    gradient within 5e-5 of the gradient's scale (both are TF32-split sums in different orders; each is within 2e-5 of fp64)."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.learner import FlatAdamState
    from xuanpolicy_b200.policies import make_policy
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(128,), device="cuda", seed=11)
    with torch.no_grad():
        for p in policy.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
    fused = FusedActorCritic(policy)
    assert fused.bin_wgrad
    g = torch.Generator(device="cuda").manual_seed(B)
    obs = torch.randn(B, 4, device="cuda", generator=g)[:, :obs_space.shape[0]]
    act_out, v = fused.forward(obs)
    A = act_out.shape[1]
    dact = torch.randn(B, A, device="cuda", generator=g) / B
    if A == 2:
        dact[:, 1] = -dact[:, 0]
    dv = torch.randn(B, device="cuda", generator=g) / B
    names = [n for n, q in policy.named_parameters() if q.grad is not None and "logstd" not in n]
    out = {}
    for mode in (True, False):
        fused.bin_wgrad = mode
        for q in policy.parameters():
            if q.grad is not None:
                q.grad.fill_(float("nan"))
        fused.backward(dact, dv, softmax_pair=True)
        torch.cuda.synchronize()
        assert fused._bin_now == mode
        out[mode] = {n: dict(policy.named_parameters())[n].grad.clone() for n in names}
    for n in names:
        a, b = out[True][n].double(), out[False][n].double()
        assert torch.isfinite(a).all(), n
        scale = b.pow(2).mean().sqrt().clamp_min(1e-30)
        err = float((a - b).abs().max() / scale)
        print("   %-32s %.2e" % (n, err))
        assert err < 5e-5, (n, err)      # (measured against fp64: binary form <= 1e-5, three-MMA form <= 2.1e-5)


@pytest.mark.parametrize("env_id,B", [("Pendulum-v1", 65536), ("CartPole-v1", 4096 + 5)])
def test_training_forward_with_the_trunk_generated_in_the_kernel(env_id, B):
    """xb_mlp_fwd_from_obs_train (the first MLP layer generated by the hidden-layer launch's operand warps, which also store h1)
    against xb_mlp_trunk_fwd + xb_dense_fwd2: h1, head outputs, sign words and hidden activations bit-identical."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200.fused_mlp import FusedActorCritic
    from xuanpolicy_b200.policies import make_policy
    obs_space, act_space = xb.make_spaces(env_id)
    policy = make_policy(obs_space, act_space, hidden=(128,), device="cuda", seed=9)
    with torch.no_grad():
        for p in policy.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    fused = FusedActorCritic(policy)
    assert fused.fwd_from_obs_ok()
    g = torch.Generator(device="cuda").manual_seed(B)
    obs = torch.randn(B, obs_space.shape[0], device="cuda", generator=g)
    out = {}
    for mode in (False, True):
        b = fused._buffers(B)
        for k in ("h1", "ya", "yc", "act", "v"):
            b[k].fill_(float("nan"))
        b["signs"].zero_()
        a, v = fused.forward(obs, trunk_in_kernel=mode)
        torch.cuda.synchronize()
        out[mode] = {k: b[k].clone() for k in ("h1", "ya", "yc", "act", "v", "signs")}
    for k in out[True]:
        assert torch.equal(out[True][k], out[False][k]), k
