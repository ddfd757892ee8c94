"""Pins oracle/gd_grad_np.py (d loss / d theta through the unrolled momentum-GD loop) against
torch autograd with create_graph=True on the same unrolled graph -- the role
``opt.compute_gradients(self.mse_, self.theta_)`` plays in the reference
(multi-label-cls/icnn-back.py:120-139).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import gd_grad_np, picnn_np, synth
from oracle.gd_grad_torch import torch_unrolled_grads

@pytest.mark.parametrize("name,B,nIter,lr,mom", [("C1", 12, 6, 0.05, 0.3), ("C1", 5, 1, 0.1, 0.9),
                                                ("C3", 6, 4, 0.02, 0.3), ("C4", 16, 5, 0.05, 0.5)])
def test_gd_backward_matches_autograd(name, B, nIter, lr, mom):
    p, x, y0 = synth.make_inputs(name, B=B)
    trueY = (np.random.RandomState(11).uniform(size=(B, p.n)) < 0.3).astype(np.float64)
    yN_t, _, G = torch_unrolled_grads(p, x, y0, trueY, nIter, lr, mom)
    gts = picnn_np.gates(p, x)
    yN, gr = gd_grad_np.gd_backward(p, gts, y0, nIter, lr, mom,
                                    lambda y: 2.0 * (y - trueY) / y.size)
    np.testing.assert_allclose(yN, yN_t, rtol=1e-12, atol=1e-12)
    xg = gd_grad_np.xpath_backward(p, x, gr["dcy"], gr["dcz"])

    def close(a, b, what):
        scale = max(np.abs(b).max(), 1e-30)
        assert np.abs(a - b).max() <= 1e-9 * scale + 1e-18, (what, np.abs(a - b).max(), scale)

    for l in range(p.L + 1):
        close(gr["dWy"][l], G["Wy"][l], f"dWy{l}")
        close(xg["dWyu"][l], G["Wyu"][l], f"dWyu{l}")
        close(xg["dbyu"][l], G["byu"][l], f"dbyu{l}")
        if l > 0:
            close(gr["dWz"][l], G["Wz"][l], f"dWz{l}")
            close(xg["dWzu"][l], G["Wzu"][l], f"dWzu{l}")
            close(xg["dbzu"][l], G["bzu"][l], f"dbzu{l}")
        # the additive gate d_l does not enter df/dy: its parameters get no gradient
        # (TF returns None, filtered out at multi-label-cls/icnn-back.py:137-138)
        assert G["Wzx"][l] is None or not np.any(G["Wzx"][l])
    for l in range(p.L):
        close(xg["dWu"][l], G["Wu"][l], f"dWu{l}")
        close(xg["dbu"][l], G["bu"][l], f"dbu{l}")


def test_kappa_recurrence():
    k = gd_grad_np.kappas(3, 0.1, 0.5)
    # c_3 = 1.5, c_2 = 1.75, c_1 = 1.875
    np.testing.assert_allclose(k, [-0.1875, -0.175, -0.15])


def test_oracle_matches_golden(golden_dir):
    """oracle/gd_grad_np.py vs the committed torch-autograd vectors (oracle/gen_golden_gd_grad.py)."""
    import os
    from oracle.gen_golden_gd_grad import CASES, true_labels
    gold = np.load(os.path.join(golden_dir, "gd_grad.npz"))
    for tag, (name, B, nIter, lr, mom) in CASES.items():
        p, x, y0 = synth.make_inputs(name, B=B)
        tY = true_labels(B, p.n)
        yN, gr = gd_grad_np.gd_backward(p, picnn_np.gates(p, x), y0, nIter, lr, mom,
                                        lambda y: 2.0 * (y - tY) / y.size)
        np.testing.assert_allclose(yN, gold[tag + "_yN"], rtol=1e-11, atol=1e-12)
        for l in range(p.L + 1):
            ref = gold["%s_Wy%d" % (tag, l)]
            assert np.abs(gr["dWy"][l] - ref).max() <= 1e-9 * np.abs(ref).max() + 1e-18
            if l > 0:
                ref = gold["%s_Wz%d" % (tag, l)]
                assert np.abs(gr["dWz"][l] - ref).max() <= 1e-9 * np.abs(ref).max() + 1e-18


def test_gd_backward_matches_finite_differences():
    """A third, autodiff-free pin: central differences of the unrolled-GD loss in single weights."""
    import copy
    p, x, y0 = synth.make_inputs("C1", B=6)
    tY = (np.random.RandomState(3).uniform(size=(6, p.n)) < 0.3).astype(np.float64)
    nIter, lr, mom = 5, 0.05, 0.3
    gts = picnn_np.gates(p, x)

    def loss(pp):
        yN, _ = picnn_np.momentum_gd(lambda y: picnn_np.fg_gated(pp, gts, y), y0, nIter, lr, mom)
        return float(((yN - tY) ** 2).mean())

    _, gr = gd_grad_np.gd_backward(p, gts, y0, nIter, lr, mom, lambda y: 2.0 * (y - tY) / y.size)
    rs = np.random.RandomState(4)
    eps = 1e-6
    checked = 0
    for name, key, layers in (("Wy", "dWy", range(p.L + 1)), ("Wz", "dWz", range(1, p.L + 1))):
        for l in layers:
            W = getattr(p, name)[l]
            for _ in range(3):
                i, j = rs.randint(W.shape[0]), rs.randint(W.shape[1])
                pp, pm = copy.deepcopy(p), copy.deepcopy(p)
                getattr(pp, name)[l][i, j] += eps
                getattr(pm, name)[l][i, j] -= eps
                fd = (loss(pp) - loss(pm)) / (2 * eps)
                an = gr[key][l][i, j]
                # a ReLU kink inside [W - eps, W + eps] would break the difference quotient; none at this seed
                assert abs(fd - an) <= 1e-5 * max(abs(an), np.abs(gr[key][l]).max()) + 1e-12, (name, l, i, j, fd, an)
                checked += 1
    assert checked == 3 * (2 * p.L + 1)


def test_host_xpath_backward_matches_oracle():
    """icnn_b200.gd_grad._xpath_backward (the host-side dense backprop of the gate adjoints into the
    x-path parameters; device-agnostic torch ops) against oracle/gd_grad_np.xpath_backward, on CPU."""
    import types
    import torch
    from icnn_b200.gd_grad import _xpath_backward
    p, x, y0 = synth.make_inputs("C1", B=9)
    rs = np.random.RandomState(8)
    dcy = [rs.randn(9, p.n) for _ in range(p.L + 1)]
    dcz = [None] + [rs.randn(9, p.hidden[l - 1]) for l in range(1, p.L + 1)]
    want = gd_grad_np.xpath_backward(p, x, dcy, dcz)
    t = lambda a: None if a is None else torch.tensor(np.asarray(a), dtype=torch.float64)  # noqa: E731
    net = types.SimpleNamespace(L=p.L, **{k: [t(a) for a in getattr(p, k)] for k in ("Wu", "bu", "Wzu", "bzu", "Wyu", "byu")})
    got = _xpath_backward(net, t(x), [t(a) for a in dcy], [t(a) for a in dcz])
    for k in ("Wu", "bu", "Wzu", "bzu", "Wyu", "byu"):
        for a, b in zip(got[k], want["d" + k]):
            if b is None:
                assert a is None
            else:
                np.testing.assert_allclose(a.numpy(), b, rtol=1e-11, atol=1e-13)
