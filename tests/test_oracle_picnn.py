"""Pins for the restated PICNN f / df/dy (the reference's TF graph cannot run here):
finite differences and an independently written torch-autograd forward.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import picnn_np, synth


def torch_forward(p, x, y):
    """Independent forward written from multi-label-cls/icnn_ebundle.py:339-388 with autograd
    supplying dE/dy (the role tf.gradients plays at :146)."""
    t = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    x = t(x)
    L = p.L
    us, prev = [], x
    for i in range(L):
        u = prev @ t(p.Wu[i]) + t(p.bu[i])
        if i < L - 1:
            u = torch.relu(u)
        us.append(u)
        prev = u
    prevU, prevZ = x, None
    z = None
    for i in range(L + 1):
        terms = []
        if i > 0:
            zu_u = torch.relu(prevU @ t(p.Wzu[i]) + t(p.bzu[i]))
            terms.append((prevZ * zu_u) @ t(p.Wz[i]))
        yu_u = prevU @ t(p.Wyu[i]) + t(p.byu[i])
        terms.append((y * yu_u) @ t(p.Wy[i]))
        terms.append(prevU @ t(p.Wzx[i]) + t(p.bzx[i]))
        z = sum(terms)
        if i < L:
            z = torch.nn.functional.leaky_relu(z, p.alpha) if p.alpha else torch.relu(z)
        prevU = us[i] if i < L else None
        prevZ = z
    return z.reshape(-1)


@pytest.mark.parametrize("name,B", [("C1", 16), ("C3", 4), ("C4", 32)])
def test_fg_matches_autograd(name, B):
    p, x, _ = synth.make_inputs(name, B=B)
    y = np.random.RandomState(7).uniform(0.05, 0.95, size=(B, p.n))
    f, g = picnn_np.make_fg(p, x)(y)
    yt = torch.tensor(y, dtype=torch.float64, requires_grad=True)
    E = torch_forward(p, x, yt)
    (gt,) = torch.autograd.grad(E.sum(), yt)
    np.testing.assert_allclose(f, E.detach().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g, gt.numpy(), rtol=1e-11, atol=1e-12)


def test_fg_finite_difference():
    p, x, _ = synth.make_inputs("C1", B=8)
    fg = picnn_np.make_fg(p, x)
    y = np.random.RandomState(3).uniform(0.1, 0.9, size=(8, p.n))
    _, g = fg(y)
    eps = 1e-6
    for j in range(p.n):
        yp, ym = y.copy(), y.copy()
        yp[:, j] += eps
        ym[:, j] -= eps
        fd = (fg(yp)[0] - fg(ym)[0]) / (2 * eps)
        np.testing.assert_allclose(fd, g[:, j], atol=1e-6)


def test_affine_wrapper_and_convexity():
    p, x, _ = synth.make_inputs("C4", B=16)
    fg = picnn_np.make_fg(p, x, affine=True)
    fr = picnn_np.make_fg(p, x)
    y = np.random.RandomState(1).uniform(0, 1, size=(16, p.n))
    f, g = fg(y)
    f2, g2 = fr(2 * y - 1)
    np.testing.assert_allclose(f, f2)
    np.testing.assert_allclose(g, 2 * g2)
    # f is convex in y (W^z >= 0, convex non-decreasing activation): bundle rows under-estimate f
    y2 = np.random.RandomState(2).uniform(0, 1, size=(16, p.n))
    fy2, _ = fg(y2)
    assert np.all(f + np.sum(g * (y2 - y), axis=1) <= fy2 + 1e-9)


def test_momentum_gd_recurrence():
    p, x, y0 = synth.make_inputs("C1", B=4)
    fg = picnn_np.make_fg(p, x)
    y, f = picnn_np.momentum_gd(fg, y0, nIter=3, lr=0.01, momentum=0.3)
    # hand-unrolled multi-label-cls/icnn-back.py:120-131
    yi, vi = y0.copy(), 0.0
    for _ in range(3):
        prev = vi
        vi = 0.3 * prev - 0.01 * fg(yi)[1]
        yi = yi - 0.3 * prev + 1.3 * vi
    np.testing.assert_allclose(y, yi, atol=1e-14)
    np.testing.assert_allclose(f, fg(yi)[0], atol=1e-14)


def test_batchnorm_fold_is_exact_in_float64():
    """Inference-mode batch-norm on the u-path (multi-label-cls/icnn_ebundle.py:343-345) folded into the
    consumer weights (icnn_b200.workloads.fold_batchnorm -- what the device path does) gives the same gates."""
    from icnn_b200 import workloads
    from oracle import picnn_np
    p = workloads.synth_params(11, 20, 12, [24, 16, 10])
    rs = np.random.RandomState(3)
    for i in range(p.L - 1):
        w = p.hidden[i]
        p.bn[i] = workloads.bn_affine(rs.uniform(0.5, 1.5, w), rs.randn(w) * 0.1, rs.randn(w) * 0.2, rs.uniform(0.5, 2.0, w))
    x = rs.randn(7, 20)
    q = workloads.fold_batchnorm(p)
    assert all(b is None for b in q.bn)
    ga, gb = picnn_np.gates(p, x), picnn_np.gates(q, x)
    for la, lb in zip(ga, gb):
        for a, b in zip(la, lb):
            if a is not None:
                np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)
    # and BN really changes the gates (the test is not vacuous)
    p0 = workloads.fold_batchnorm(p)
    p0.Wu, p0.bu = p.Wu, p.bu
    assert np.abs(picnn_np.gates(p, x)[1][1] - picnn_np.gates(workloads.synth_params(11, 20, 12, [24, 16, 10]), x)[1][1]).max() > 1e-3
