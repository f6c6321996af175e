import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xuanpolicy_b200 as xb
from xuanpolicy_b200 import ops
from xuanpolicy_b200.fused_mlp import FusedActorCritic
from xuanpolicy_b200.learner import FlatAdamState
from xuanpolicy_b200.policies import make_policy
def rel(a, ref):
    ref = ref.double(); rms = ref.pow(2).mean().sqrt().clamp_min(1e-30)
    return float(((a.double() - ref).abs() / torch.maximum(ref.abs(), rms)).max())
B, hidden = 65536, 128
obs_space, act_space = xb.make_spaces("Pendulum-v1")
policy = make_policy(obs_space, act_space, hidden=(hidden,), device="cuda", seed=3)
flat = FlatAdamState(policy, torch.optim.Adam(policy.parameters(), 1e-3), None)
fused = FusedActorCritic(policy)
g = torch.Generator(device="cuda").manual_seed(B)
obs = torch.randn(B, 4, device="cuda", generator=g)[:, :3]
act_out, v = fused.forward(obs)
for scale in (1.0, 1.0 / B):
    dact = torch.randn(B, 1, device="cuda", generator=g) * scale
    dv = torch.randn(B, device="cuda", generator=g) * scale
    b = fused._buf[B]
    lk = lambda t: torch.nn.functional.leaky_relu(t, 0.01)
    mask = lambda y: torch.where(y > 0, 1.0, 0.01).double()
    dza = (dact.double() @ fused.la2.weight.double()) * mask(b["ya"])
    dzc = (dv.double()[:, None] @ fused.lc2.weight.double()) * mask(b["yc"])
    dh1 = dza @ fused.la1.weight.double() + dzc @ fused.lc1.weight.double()
    dz1_ref = dh1 * mask(b["h1"])
    fused.backward(dact, dv)
    torch.cuda.synchronize()
    print("scale", scale)
    print(" dz1", rel(b["dz1"], dz1_ref))
    print(" dWa1", rel(fused.la1.weight.grad, dza.t() @ b["h1"].double()), " dWc1", rel(fused.lc1.weight.grad, dzc.t() @ b["h1"].double()))
    print(" dba1", rel(fused.la1.bias.grad, dza.sum(0)), " dbc1", rel(fused.lc1.bias.grad, dzc.sum(0)))
    print(" dW0", rel(fused.l0.weight.grad, dz1_ref.t() @ obs.double()), " db0", rel(fused.l0.bias.grad, dz1_ref.sum(0)))
    # standalone wgrad into fresh outputs
    H = hidden
    outs = [torch.zeros(H, H, device="cuda"), torch.zeros(H, device="cuda"), torch.zeros(1, H, device="cuda"), torch.zeros(1, device="cuda"),
            torch.zeros(H, H, device="cuda"), torch.zeros(H, device="cuda"), torch.zeros(1, H, device="cuda"), torch.zeros(1, device="cuda")]
    ops.dense_wgrad(b["ya"], dact, fused.la2.weight.data, b["yc"], dv.reshape(B, 1), fused.lc2.weight.data, b["h1"], 0.01, fused.ws_wgrad, *outs)
    torch.cuda.synchronize()
    print(" standalone dWa1", rel(outs[0], dza.t() @ b["h1"].double()), " dWc1", rel(outs[4], dzc.t() @ b["h1"].double()))
    err = (outs[4].double() - dzc.t() @ b["h1"].double()).abs()
    print(" dWc1 err rows max:", err.max(1).values.topk(5), " cols:", err.max(0).values.topk(5).indices)
    print(" |lc2.w| min", float(fused.lc2.weight.abs().min()), "|la2.w| min", float(fused.la2.weight.abs().min()))
