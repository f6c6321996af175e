"""The algebra behind the binary-form weight gradient (csrc/dense_tc.cu dense_wgrad_bin_kernel + wgrad_reduce_kernel), in fp64
numpy against torch autograd of the reference's module graph (Linear + LeakyReLU -> head, xuance/torch/policies/gaussian.py:17-24,
categorical.py:26-32; `loss.backward()` at ppoclip_learner.py:47).  CPU only: it pins the identities the kernels rely on —

    leaky'(y) = slope + (1 - slope) [y > 0]
    Gm[m][n] = (1 - slope) sum_b [y > 0][b][m] e[b] x[b][n] + slope sum_b e[b] x[b][n],   gm[m] likewise with x -> 1
    dW = w2' Gm,   db = w2' gm,   dw2 = sum_n W Gm + b gm  (= sum_b e y),   db2 = sum_b e

for a one-output head and for the two logits of a softmax pair (opposite gradients, w2' = w2[0] - w2[1])."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("pair", [False, True])
def test_binary_form_identities_match_autograd(pair):
    rng = np.random.default_rng(7 + pair)
    B, Hin, Hout, slope = 257, 24, 16, 0.01
    x = rng.standard_normal((B, Hin))
    W = rng.standard_normal((Hout, Hin)) / 4
    b = rng.standard_normal(Hout) / 4
    nh = 2 if pair else 1
    w2 = rng.standard_normal((nh, Hout))
    e = rng.standard_normal(B) / B
    dout = np.stack([e, -e], 1) if pair else e[:, None]            # gradients w.r.t. the head outputs
    # autograd of the module graph
    tx, tW, tb, tw2 = (torch.tensor(a, dtype=torch.float64) for a in (x, W, b, w2))
    tW.requires_grad_(True), tb.requires_grad_(True), tw2.requires_grad_(True)
    tb2 = torch.zeros(nh, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.leaky_relu(tx @ tW.t() + tb, slope)
    out = y @ tw2.t() + tb2
    out.backward(torch.tensor(dout))
    # the binary form
    z = x @ W.T + b
    binm = (z > 0).astype(np.float64)                              # what the forward's sign words hold
    w2p = w2[0] - w2[1] if pair else w2[0]
    ex = e[:, None] * x
    A = binm.T @ ex                                                # the 2-MMA GEMM: 0/1 operand x (e x) split hi/lo
    G = ex.sum(0)
    S = binm.T @ e
    E = e.sum()
    Gm = (1 - slope) * A + slope * G[None, :]
    gm = (1 - slope) * S + slope * E
    dW, db = w2p[:, None] * Gm, w2p * gm
    dw2_0 = (W * Gm).sum(1) + b * gm
    np.testing.assert_allclose(dW, tW.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(db, tb.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dw2_0, tw2.grad.numpy()[0], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(E, tb2.grad.numpy()[0], rtol=1e-10, atol=1e-14)
    if pair:
        np.testing.assert_allclose(-dw2_0, tw2.grad.numpy()[1], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(-E, tb2.grad.numpy()[1], rtol=1e-10, atol=1e-14)


def test_sign_word_layout_round_trip():
    """bit j of word c of a layer = (y[:, 32 c + j] > 0): the layout xb_dense_fwd2 writes and dgrad / the binary wgrad read."""
    rng = np.random.default_rng(3)
    y = rng.standard_normal((37, 128))
    bits = (y > 0)
    words = (bits.reshape(37, 4, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)
    back = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(37, 128).astype(bool)
    assert (back == bits).all()
