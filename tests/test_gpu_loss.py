"""GPU parity: fused PPO loss fwd/bwd, the learner drop-in, action sampling and the fused clip+Adam step."""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, rel_close

pytestmark = pytest.mark.gpu


def _policy(g, device):
    from xuanpolicy_b200 import policies, spaces
    m = g["meta"]
    if m["discrete"]:
        obs_space, act_space = spaces.Box(-1, 1, (4,)), spaces.Discrete(2)
    else:
        obs_space, act_space = spaces.Box(-1, 1, (3,)), spaces.Box(-2.0, 2.0, (1,))
    pol = policies.make_policy(obs_space, act_space, hidden=(m["hidden"],), device=device)
    pol.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")})
    return pol


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name", ["loss_cat_h64", "loss_gauss_h128"])
def test_learner_update_matches_reference_golden(name):
    """PPOCLIP_Learner.update drop-in vs the reference's own update: info scalars, param.grad (1e-4 rel, fp32),
    parameters after the clipped Adam step."""
    import xuanpolicy_b200 as xb
    g = load_golden(name)
    m = g["meta"]
    for tag, clip in (("noclip", False), ("clip", True)):
        pol = _policy(g, "cuda")
        opt = torch.optim.Adam(pol.parameters(), 4e-4, eps=1e-5)
        sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
        learner = xb.PPOCLIP_Learner(pol, opt, sched, "cuda", "/tmp", vf_coef=m["vf_coef"], ent_coef=m["ent_coef"],
                                     clip_range=m["clip_range"], clip_grad_norm=m["clip_grad_norm"], use_grad_clip=clip)
        info = learner.update(g["obs"], g["act"], g["ret"], g["val"], g["adv"], g["old_logp"])
        assert learner.iterations == 1 and torch.is_tensor(info["clip_ratio"])
        for k in ("actor-loss", "critic-loss", "entropy", "learning_rate", "predict_value", "clip_ratio"):
            ref = float(g["info_%s/%s" % (tag, k)])
            assert abs(float(info[k]) - ref) <= 1e-4 * max(1.0, abs(ref)), (k, float(info[k]), ref)
        for k, p in pol.named_parameters():
            ok, err = rel_close(p.grad.cpu().numpy(), g["grad_%s/%s" % (tag, k)], 1e-4)
            assert ok, (tag, k, err)
            ok, err = rel_close(p.detach().cpu().numpy(), g["p1_%s/%s" % (tag, k)], 1e-5)
            assert ok, (tag, k, err)


@pytest.mark.parametrize("discrete,B,A", [(True, 512, 2), (True, 4097, 5), (False, 1000, 1), (False, 777, 3)])
@pytest.mark.parametrize("value_clip", [0.0, 0.2])
def test_loss_kernel_vs_torch_autograd(discrete, B, A, value_clip):
    """Kernel gradients w.r.t. network outputs and log scalars vs a plain torch fp32 autograd evaluation of the
    reference formulas (oracle/ref_port.ppo_clip_loss), dense and through the fused index gather."""
    from oracle import ref_port
    from xuanpolicy_b200 import ops
    from xuanpolicy_b200.policies import CategoricalDistribution, DiagGaussianDistribution
    gen = torch.Generator(device="cuda").manual_seed(B + A)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=gen)
    clip, vf, ent = 0.2, 0.25, 0.01
    T, N = 16, (B + 15) // 16 + 3
    v = rnd(B).requires_grad_()
    ret, adv_raw, val_old = rnd(B), rnd(B) * 1.7 + 0.3, rnd(B)
    if discrete:
        logits = rnd(B, A).requires_grad_()
        dist = CategoricalDistribution(A)
        dist.set_param(logits)
        act = torch.randint(0, A, (B,), device="cuda", generator=gen).float()
    else:
        mu = rnd(B, A).requires_grad_()
        logstd = (rnd(A) * 0.3 - 0.5).requires_grad_()
        dist = DiagGaussianDistribution(A)
        dist.set_param(mu, logstd.exp())
        act = (mu.detach() + rnd(B, A) * logstd.detach().exp())
    with torch.no_grad():
        old_logp = dist.log_prob(act) + 0.1 * rnd(B)
    adv = (adv_raw - adv_raw.mean()) / (adv_raw.std(unbiased=False) + 1e-8)
    loss, a_loss, c_loss, e_loss, ratio = ref_port.ppo_clip_loss(dist.log_prob(act), dist.entropy(), v, ret, adv, old_logp, vf, ent, clip)
    if value_clip > 0:
        vc = val_old + (v - val_old).clamp(-value_clip, value_clip)
        c_loss = torch.max((v - ret) ** 2, (vc - ret) ** 2).mean()
        loss = a_loss - ent * e_loss + vf * c_loss
    loss.backward()
    stats = torch.stack([adv_raw.double().sum(), (adv_raw.double() ** 2).sum()])
    scal = torch.zeros(8, dtype=torch.float64, device="cuda")
    dvk = torch.empty(B, device="cuda")
    # scatter the minibatch into a [T, N] rollout at random flat indices so the fused gather path is exercised too
    perm = torch.randperm(T * N, device="cuda", generator=gen)[:B]
    row = (perm % T) * N + perm // T
    def scat(x, width=1):
        full = torch.zeros((T * N, width), device="cuda")
        full[row] = x.reshape(B, width)
        return full.reshape(-1) if width == 1 else full
    modes = [None, perm] + (["packed"] if (value_clip == 0 and (discrete or A == 1)) else [])
    for idx in modes:
        kw = dict(clip_range=clip, vf_coef=vf, ent_coef=ent, inv_batch=1.0 / B, adv_stats=stats, adv_count=B,
                  value_clip=value_clip)
        if idx is None:
            f = lambda x, w=1: x.contiguous()
            kw.update(val_old=val_old)
        elif isinstance(idx, str):      # compact float4 {act, old_logp, adv, ret} rows, as xb_gather_records emits them
            f = lambda x, w=1: None
            kw.update(packed=torch.stack([act.reshape(B), old_logp, adv_raw, ret], dim=1).contiguous())
        else:
            f = scat
            kw.update(idx=idx, T=T, N=N, val_old=scat(val_old))
        if discrete:
            dl = torch.empty(B, A, device="cuda")
            ops.ppo_loss_categorical(logits.detach(), v.detach(), f(act), f(ret), f(adv_raw), f(old_logp), dl, dvk, scal, **kw)
            grads = [(dl, logits.grad)]
        else:
            dm = torch.empty(B, A, device="cuda")
            dls = torch.empty(A, dtype=torch.float64, device="cuda")
            ops.ppo_loss_gaussian(mu.detach(), logstd.detach(), v.detach(), f(act, A), f(ret), f(adv_raw), f(old_logp), dm, dls,
                                  dvk, scal, **kw)
            grads = [(dm, mu.grad), (dls.float(), logstd.grad)]
        grads.append((dvk, v.grad))
        for got, ref in grads:
            ok, err = rel_close(got.cpu().numpy(), ref.cpu().numpy(), 1e-4)   # 1e-4 relative, fp32 (north_star)
            assert ok, err
        s = scal.cpu().numpy() / B
        n_clip = ((ratio < 1 - clip).sum() + (ratio > 1 + clip).sum()).item() / B
        for got, ref in ((-s[0], a_loss.item()), (s[1], c_loss.item()), (s[2], e_loss.item()), (s[3], v.mean().item()),
                         (s[4], n_clip)):
            assert abs(got - ref) <= 1e-4 * max(1.0, abs(ref)), (got, ref)


def test_sampling_kernels_logp_and_distribution():
    from xuanpolicy_b200 import ops
    N = 200_000
    gen = torch.Generator(device="cuda").manual_seed(3)
    ctr = torch.zeros(1, dtype=torch.int64, device="cuda")
    logits = torch.randn(3, device="cuda", generator=gen).repeat(N, 1).contiguous()
    act, logp = torch.empty(N, dtype=torch.int64, device="cuda"), torch.empty(N, device="cuda")
    ops.sample_categorical(logits, 7, ctr, 0, act, logp)
    ref_lp = torch.log_softmax(logits, -1)
    assert torch.allclose(logp, ref_lp.gather(1, act[:, None])[:, 0], atol=1e-6)
    freq = torch.bincount(act, minlength=3).double() / N
    assert torch.allclose(freq, ref_lp[0].exp().double(), atol=5e-3)
    act2 = torch.empty_like(act)
    ops.sample_categorical(logits, 7, ctr, 0, act2, logp)
    assert torch.equal(act, act2)                      # same (seed, counter, offset) -> same draw
    ops.counter_add(ctr, 1)
    ops.sample_categorical(logits, 7, ctr, 0, act2, logp)
    assert not torch.equal(act, act2) and int(ctr.item()) == 1
    mu = torch.randn(N, 2, device="cuda", generator=gen)
    logstd = torch.tensor([-1.0, 0.3], device="cuda")
    a, lp = torch.empty(N, 2, device="cuda"), torch.empty(N, device="cuda")
    ops.sample_gaussian(mu, logstd, 11, ctr, 5, a, lp)
    ref = torch.distributions.Normal(mu, logstd.exp()).log_prob(a).sum(-1)
    assert torch.allclose(lp, ref, atol=2e-5, rtol=1e-5)
    z = (a - mu) / logstd.exp()
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1.0) < 0.01
    assert abs((z[:, 0] * z[:, 1]).mean().item()) < 0.01


def test_fused_clip_adam_matches_torch():
    """csrc/optim.cu vs clip_grad_norm_ + torch.optim.Adam(eps=1e-5) + LinearLR over several steps."""
    from xuanpolicy_b200 import ops
    torch.manual_seed(0)
    n = 8835
    p_ref = torch.randn(n, device="cuda").requires_grad_()
    opt = torch.optim.Adam([p_ref], 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=50)
    p = p_ref.detach().clone()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    ws = torch.zeros(8 + 1024, dtype=torch.float64, device="cuda")
    lr_out, gn = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    for it in range(12):
        g = torch.randn(n, device="cuda") * (3.0 if it % 2 else 0.001)
        p_ref.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_([p_ref], 0.5)
        lr_used = opt.param_groups[0]["lr"]
        opt.step()
        sched.step()
        ops.clip_adam_step(p, g, m, v, step, 4e-4, 0.0, 50, 0.9, 0.999, 1e-5, 0.5, 1.0, ws, lr_out, gn)
        assert abs(gn.item() - norm.item()) <= 1e-5 * norm.item()
        assert abs(lr_out.item() - lr_used) <= 1e-6 * 4e-4
        ok, err = rel_close(p.cpu().numpy(), p_ref.detach().cpu().numpy(), 1e-6)
        assert ok, (it, err)
    assert int(step.item()) == 12


@pytest.mark.parametrize("name", ["loss_a2c_gauss_h64", "loss_pg_cat_h32"])
def test_a2c_and_pg_learners_match_reference_golden(name):
    """Row f3: A2C_Learner / PG_Learner drop-ins (same fused kernel, A2C surrogate) vs the reference's own updates."""
    import xuanpolicy_b200 as xb
    from xuanpolicy_b200 import policies, spaces
    g = load_golden(name)
    m = g["meta"]
    if m["algo"] == "a2c":
        pol = _policy(g, "cuda")
    else:
        rep = policies.MLPRepresentation((4,), [m["hidden"]], device="cuda")
        pol = policies.CategoricalActor(spaces.Discrete(2), rep, [m["hidden"]], device="cuda")
        pol.load_state_dict({k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("p0/")})
    opt = torch.optim.Adam(pol.parameters(), 4e-4, eps=1e-5)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=1000)
    if m["algo"] == "a2c":
        learner = xb.A2C_Learner(pol, opt, sched, "cuda", "/tmp", vf_coef=m["vf_coef"], ent_coef=m["ent_coef"], clip_grad=m["clip_grad"])
        info = learner.update(g["obs"], g["act"], g["ret"], g["adv"])
    else:
        learner = xb.PG_Learner(pol, opt, sched, "cuda", "/tmp", ent_coef=m["ent_coef"], clip_grad=m["clip_grad"])
        info = learner.update(g["obs"], g["act"], g["ret"])
    keys = [k[5:] for k in g if k.startswith("info/")]
    assert sorted(info) == sorted(keys)
    for k in keys:
        ref = float(g["info/" + k])
        assert abs(float(info[k]) - ref) <= 1e-4 * max(1.0, abs(ref)), (k, float(info[k]), ref)
    for k, p in pol.named_parameters():
        ok, err = rel_close(p.grad.cpu().numpy(), g["grad_clipped/" + k], 1e-4)
        assert ok, (k, err)
        ok, err = rel_close(p.detach().cpu().numpy(), g["p1/" + k], 1e-5)
        assert ok, (k, err)


def test_packed_records_gather_bit_exact():
    """pack_records + gather_records: one 32-byte record per transition; gathered rows equal the SoA gather bit for bit."""
    from xuanpolicy_b200 import ops
    T, N, B = 48, 301, 5000
    rng = np.random.default_rng(8)
    obs = rng.standard_normal((T, N, 4)).astype(np.float32)
    act, logp, adv, ret = (rng.standard_normal((T, N)).astype(np.float32) for _ in range(4))
    d = lambda a: torch.from_numpy(a).cuda()
    rec = torch.zeros((T * N, 8), device="cuda")
    ops.pack_records(d(obs), d(act), d(logp), d(adv), d(ret), rec)
    r = rec.cpu().numpy().reshape(T, N, 8)
    assert np.array_equal(r[..., :4], obs) and np.array_equal(r[..., 4], act) and np.array_equal(r[..., 5], logp)
    assert np.array_equal(r[..., 6], adv) and np.array_equal(r[..., 7], ret)
    idx = rng.permutation(T * N)[:B].astype(np.int64)
    env, step = np.divmod(idx, T)
    for obs_dim in (3, 4):
        obs_out, scal = torch.empty((B, obs_dim), device="cuda"), torch.empty((B, 4), device="cuda")
        stats = torch.zeros(2, dtype=torch.float64, device="cuda")
        ops.gather_records(d(idx), T, N, rec, obs_dim, obs_out, scal, stats=stats)
        assert np.array_equal(obs_out.cpu().numpy(), obs[step, env, :obs_dim])
        sc = scal.cpu().numpy()
        assert np.array_equal(sc[:, 0], act[step, env]) and np.array_equal(sc[:, 1], logp[step, env])
        assert np.array_equal(sc[:, 2], adv[step, env]) and np.array_equal(sc[:, 3], ret[step, env])
        a = adv[step, env].astype(np.float64)
        s = stats.cpu().numpy()
        assert abs(s[0] - a.sum()) < 1e-9 * B and abs(s[1] - (a * a).sum()) < 1e-9 * B


@pytest.mark.parametrize("B,H,A", [(4096, 128, 1), (1000, 64, 2), (65536, 128, 1), (3333, 256, 4)])
def test_mlp_epilogue_and_head_kernels_vs_torch(B, H, A):
    """csrc/mlp_epilogue.cu against the plain torch ops they replace."""
    from xuanpolicy_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(B + H)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=gen)
    y0, bias, slope = rnd(B, H), rnd(H), 0.01
    y = y0.clone()
    ops.bias_act_fwd(y, bias, slope)
    assert torch.equal(y, torch.nn.functional.leaky_relu(y0 + bias, slope))
    dy = rnd(B, H)
    dz, db = torch.empty_like(dy), torch.empty(H, device="cuda")
    ws = torch.zeros(4 + 592 * 1024, device="cuda")
    ops.act_bias_bwd(dy, y, slope, dz, db, ws)
    ref_dz = torch.where(y > 0, dy, dy * slope)
    assert torch.equal(dz, ref_dz)
    assert torch.allclose(db, ref_dz.double().sum(0).float(), rtol=1e-4, atol=1e-3)
    w2, b2 = rnd(A, H) * 0.1, rnd(A)
    out = torch.empty(B, A, device="cuda")
    ops.head_fwd(y, w2, b2, out)
    assert torch.allclose(out, torch.addmm(b2, y, w2.t()), rtol=1e-4, atol=1e-4)
    dout = rnd(B, A)
    db1, dw2, db2 = torch.empty(H, device="cuda"), torch.empty(A, H, device="cuda"), torch.empty(A, device="cuda")
    ops.head_bwd_act(dout, y, w2, slope, dz, db1, dw2, db2, ws)
    dh = dout @ w2
    ref_dz = torch.where(y > 0, dh, dh * slope)
    assert torch.allclose(dz, ref_dz, rtol=1e-5, atol=1e-6)
    scale = lambda t: t.abs().max().item() + 1e-6
    assert (db1 - ref_dz.double().sum(0).float()).abs().max().item() <= 2e-5 * scale(db1) * 10
    ref_dw2 = (dout.double().t() @ y.double()).float()
    assert (dw2 - ref_dw2).abs().max().item() <= 1e-4 * scale(ref_dw2)
    assert torch.allclose(db2, dout.double().sum(0).float(), rtol=1e-4, atol=1e-3)
    # second launch reuses the (self-resetting) workspace ticket
    ops.head_bwd_act(dout, y, w2, slope, dz, db1, dw2, db2, ws)
    assert (dw2 - ref_dw2).abs().max().item() <= 1e-4 * scale(ref_dw2)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 1000, 4096, 65536, 524288, 1000003])
def test_device_permutation_is_a_bijection(n):
    """xb_random_permutation (np.random.shuffle stand-in, ppoclip_agent.py:76-78): exactly the values 0..n-1, each once;
    a different key (seed or device counter) gives a different permutation; the same key reproduces it."""
    from xuanpolicy_b200 import ops
    ctr = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    ops.random_permutation(out, 123, ctr, 0)
    assert torch.equal(torch.sort(out).values, torch.arange(n, device="cuda"))
    again = torch.empty_like(out)
    ops.random_permutation(again, 123, ctr, 0)
    assert torch.equal(out, again)
    if n >= 64:
        ops.counter_add(ctr, 1)
        ops.random_permutation(again, 123, ctr, 0)
        assert torch.equal(torch.sort(again).values, torch.arange(n, device="cuda")) and not torch.equal(out, again)
        ops.random_permutation(again, 124, None, 0)
        assert not torch.equal(out, again)
        assert (out == torch.arange(n, device="cuda")).float().mean().item() < 0.05     # not the identity


def test_device_permutation_statistics():
    """Uniformity checks a shuffle must pass for minibatch SGD: over 400 keys, where element 0 lands and which element
    lands first are uniform over 64 buckets (chi-square, 63 dof, p ~ 1e-4 bound 113), minibatch membership of neighbours
    is independent, and the mean displacement matches n/3."""
    from xuanpolicy_b200 import ops
    n, K = 8192, 400
    ctr = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    first, pos0, same_mb, disp = [], [], [], []
    for k in range(K):
        ops.random_permutation(out, 99, ctr, k)
        p = out.cpu().numpy()
        first.append(p[0])
        pos0.append(int(np.nonzero(p == 0)[0][0]))
        mb = np.empty(n, np.int64)
        mb[p] = np.arange(n) // (n // 8)          # minibatch each element falls in
        same_mb.append(np.mean(mb[:-1] == mb[1:]))
        disp.append(np.mean(np.abs(p - np.arange(n))))
    for v in (first, pos0):
        counts = np.bincount(np.asarray(v) // (n // 64), minlength=64)
        chi2 = float(np.sum((counts - K / 64) ** 2 / (K / 64)))
        assert chi2 < 113.0, chi2
    assert abs(np.mean(same_mb) - 1 / 8) < 0.003, np.mean(same_mb)
    assert abs(np.mean(disp) / n - 1 / 3) < 0.005, np.mean(disp) / n
