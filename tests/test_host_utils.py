"""CPU: the host-side helper of the input feeder (xb_host_permutation, a HOST function of the C ABI: no GPU involved)
and host-only logic of the drop-ins."""
import numpy as np
import pytest
import torch


def _perm(n, seed):
    from xuanpolicy_b200 import _lib
    out = np.empty(n, np.int64)
    _lib.call("xb_host_permutation", out.ctypes.data, n, seed)
    return out


@pytest.mark.parametrize("n", [1, 2, 17, 4096, 100003])
def test_host_permutation_is_a_permutation_and_reproducible(n):
    """np.random.shuffle(indexes) stand-in (ppoclip_agent.py:76-78): exactly 0..n-1, same seed -> same draw."""
    p = _perm(n, 5)
    assert np.array_equal(np.sort(p), np.arange(n))
    assert np.array_equal(p, _perm(n, 5))
    if n > 16:
        assert not np.array_equal(p, _perm(n, 6))


def test_host_permutation_is_uniform():
    """Inside-out Fisher-Yates with unbiased bounded integers: where element 0 lands and which element lands first are
    uniform (chi-square over 32 buckets, 31 dof: p ~ 1e-4 bound 66), mean displacement n/3."""
    n, K = 1024, 640
    first, pos0, disp = [], [], []
    for k in range(K):
        p = _perm(n, 1000 + k)
        first.append(p[0])
        pos0.append(int(np.nonzero(p == 0)[0][0]))
        disp.append(np.mean(np.abs(p - np.arange(n))))
    for v in (first, pos0):
        counts = np.bincount(np.asarray(v) // (n // 32), minlength=32)
        chi2 = float(np.sum((counts - K / 32) ** 2 / (K / 32)))
        assert chi2 < 66.0, chi2
    assert abs(np.mean(disp) / n - 1 / 3) < 0.01


def _perm32(n, seed):
    from xuanpolicy_b200 import _lib
    out = np.empty(n, np.int32)
    _lib.call("xb_host_permutation32", out.ctypes.data, n, seed)
    return out


@pytest.mark.parametrize("n", [1, 3, 65536, 65537, 200003, 1 << 21])
def test_host_permutation32_is_a_permutation_and_reproducible(n):
    """The e2e feeder's 32-bit shuffle: plain Fisher-Yates up to 2^16 indices, bucket scatter + in-bucket Fisher-Yates above."""
    p = _perm32(n, 9)
    assert p.dtype == np.int32 and np.array_equal(np.sort(p), np.arange(n, dtype=np.int32))
    assert np.array_equal(p, _perm32(n, 9))
    if n > 16:
        assert not np.array_equal(p, _perm32(n, 10))


def test_host_permutation32_bucketed_regime_is_uniform():
    """Above 2^16 indices (Rao-Sandelius scatter): the position of a fixed element and the element at a fixed position are
    uniform over the whole range (chi-square over 32 cells), also ACROSS bucket boundaries, and the mean displacement is n/3."""
    n, K = 100000, 480
    first, pos7, disp = [], [], []
    for k in range(K):
        p = _perm32(n, 5000 + k)
        first.append(int(p[0]))
        pos7.append(int(np.nonzero(p == 7)[0][0]))
        if k < 40:
            disp.append(np.mean(np.abs(p.astype(np.int64) - np.arange(n))))
    for v in (first, pos7):
        counts = np.bincount(np.asarray(v) * 32 // n, minlength=32)
        chi2 = float(np.sum((counts - K / 32) ** 2 / (K / 32)))
        assert chi2 < 66.0, chi2
    assert abs(np.mean(disp) / n - 1 / 3) < 0.005
    # neighbours in the input do not stay neighbours: the fraction of i with |p[i+1] - p[i]| == 1 is ~2/n
    p = _perm32(n, 1)
    assert np.mean(np.abs(np.diff(p.astype(np.int64))) == 1) < 1e-3


def test_old_dist_params_accepts_the_reference_shapes():
    """learner / buffer helper for the {"old_dist": None} auxiliary: a batched wrapper, or the numpy object array of
    per-sample wrappers that split_distributions (xuance/torch/utils/operations.py:53-72) produces."""
    from xuanpolicy_b200 import policies
    logits = torch.randn(5, 3)
    d = policies.CategoricalDistribution(3)
    d.set_param(logits)
    kind, p0, p1 = policies.old_dist_params(d, "cpu")
    assert kind == "categorical" and torch.equal(p0, logits) and p1 is None
    objs = []
    for row in logits:
        w = policies.CategoricalDistribution(3)
        w.set_param(row.unsqueeze(0))
        objs.append(w)
    kind, p0, _ = policies.old_dist_params(np.array(objs, dtype=object), "cpu")
    assert kind == "categorical" and torch.equal(p0, logits)
    mu, std = torch.randn(4, 2), torch.rand(2) + 0.5
    objs = []
    for row in mu:
        w = policies.DiagGaussianDistribution(2)
        w.set_param(row, std)
        objs.append(w)
    kind, p0, p1 = policies.old_dist_params(np.array(objs, dtype=object).reshape(2, 2), "cpu")
    assert kind == "gaussian" and torch.equal(p0, mu) and torch.equal(p1, std.expand(4, 2))
    batch = policies.OldDistBatch("gaussian", mu, std.expand(4, 2))
    assert batch.shape == (4,) and len(batch) == 4 and torch.equal(batch.get_param()[0], mu)
