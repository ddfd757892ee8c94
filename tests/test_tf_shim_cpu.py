"""The stand-in's PRIMITIVES (oracle/tf_shim.py) against independent implementations of the same published
semantics in torch.nn.functional -- the part of the reference-graph goldens that is restated rather than executed.
CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.tf_shim import BN_EPS, Shim


def _scope(sh, name):
    return sh.tf.variable_scope(name)


def test_fully_connected_bias_activation_and_flatten():
    rs = np.random.RandomState(0)
    W, b = rs.randn(12, 5), rs.randn(5)
    sh = Shim({"a/W": W, "a/b": b, "c/W": W})
    x = torch.tensor(rs.randn(3, 12))
    with _scope(sh, "a") as s:
        y = sh.tflearn.fully_connected(x, 5, scope=s, activation="relu")
    assert torch.allclose(y, torch.relu(F.linear(x, torch.tensor(W).t(), torch.tensor(b))), atol=1e-14)
    with _scope(sh, "c") as s:                       # bias-free layer on a 4-D input: tflearn flattens (NHWC order)
        y4 = sh.tflearn.fully_connected(x.reshape(3, 2, 3, 2), 5, scope=s, bias=False)
    assert torch.allclose(y4, x @ torch.tensor(W), atol=1e-14)
    with pytest.raises(KeyError):                    # a variable the store lacks
        with _scope(sh, "missing") as s:
            sh.tflearn.fully_connected(x, 5, scope=s)
    assert sh.unused_variables() == []


def test_batch_normalization_inference_mode_equals_torch_eval_batch_norm():
    rs = np.random.RandomState(1)
    C = 6
    v = {"u/bn/gamma": rs.uniform(0.5, 1.5, C), "u/bn/beta": rs.randn(C), "u/bn/moving_mean": rs.randn(C),
         "u/bn/moving_variance": rs.uniform(0.5, 2.0, C)}
    sh = Shim(v)
    t = lambda k: torch.tensor(v["u/bn/" + k])      # noqa: E731
    for shape in ((7, C), (2, 4, 3, C)):             # dense activations and an NHWC feature map
        x = torch.tensor(rs.randn(*shape))
        with _scope(sh, "u") as s:
            y = sh.tflearn.batch_normalization(x, scope=s, name="bn")
        xt = x if x.dim() == 2 else x.permute(0, 3, 1, 2)
        ref = F.batch_norm(xt, t("moving_mean"), t("moving_variance"), t("gamma"), t("beta"), training=False, eps=BN_EPS)
        ref = ref if x.dim() == 2 else ref.permute(0, 2, 3, 1)
        assert torch.allclose(y, ref, atol=1e-13)


def test_leaky_relu_and_gradients_sum_over_outputs():
    sh = Shim({})
    x = torch.tensor([-2.0, -0.5, 0.0, 0.3, 4.0], dtype=torch.float64, requires_grad=True)
    y = sh.tflearn.activations.leaky_relu(x, alpha=0.01)
    assert torch.allclose(y, F.leaky_relu(x, 0.01), atol=0)
    (g,) = sh.tf.gradients(y * y, x)                  # tf.gradients differentiates the SUM of ys
    (gr,) = torch.autograd.grad((F.leaky_relu(x, 0.01) ** 2).sum(), x)
    assert torch.allclose(g, gr, atol=1e-15)


@pytest.mark.parametrize("H,W,k,stride", [(16, 8, 8, 4), (4, 2, 4, 2), (7, 5, 3, 1), (6, 6, 4, 1), (9, 6, 4, 2), (5, 5, 8, 4)])
def test_conv_2d_same_padding(H, W, k, stride):
    """TensorFlow 'SAME': out = ceil(in / stride); for stride 1 it must equal torch's padding='same' (odd AND even
    kernels: both libraries put the extra pad at the end), for stride > 1 a hand-padded valid convolution."""
    rs = np.random.RandomState(2)
    cin, cout = 3, 4
    Wt = rs.randn(k, k, cin, cout)
    b = rs.randn(cout)
    sh = Shim({"c/W": Wt, "c/b": b})
    x = torch.tensor(rs.randn(2, H, W, cin))
    with _scope(sh, "c") as s:
        y = sh.tflearn.conv_2d(x, cout, k, strides=[1, stride, stride, 1], scope=s)
    assert tuple(y.shape) == (2, -(-H // stride), -(-W // stride), cout)
    xn, wn = x.permute(0, 3, 1, 2), torch.tensor(Wt).permute(3, 2, 0, 1)
    if stride == 1:
        ref = F.conv2d(xn, wn, torch.tensor(b), padding="same")
    else:                                             # explicit zero padding, then a VALID strided correlation
        def pads(size):
            out = -(-size // stride)
            tot = max((out - 1) * stride + k - size, 0)
            return tot // 2, tot - tot // 2
        (pt, pb), (pl, pr) = pads(H), pads(W)
        xp = torch.zeros(2, cin, H + pt + pb, W + pl + pr, dtype=torch.float64)
        xp[:, :, pt:pt + H, pl:pl + W] = xn
        ref = F.conv2d(xp, wn, torch.tensor(b), stride=stride)
    assert torch.allclose(y, ref.permute(0, 2, 3, 1), atol=1e-12)


def test_variable_scopes_nest_and_trainable_order_is_creation_order():
    sh = Shim({"q/u0/W": np.ones((2, 2)), "q/u0/b": np.zeros(2), "q/z0_u/W": np.ones((2, 1)), "q/z0_u/b": np.zeros(1)})
    x = torch.ones(1, 2, dtype=torch.float64)
    with sh.tf.variable_scope("q"):
        with sh.tf.variable_scope("z0_u") as s:
            assert s.name == "q/z0_u"
            sh.tflearn.fully_connected(x, 1, scope=s)
        with sh.tf.variable_scope("u0") as s:
            sh.tflearn.fully_connected(x, 2, scope=s)
    assert sh.created == ["q/z0_u/W", "q/z0_u/b", "q/u0/W", "q/u0/b"]
    assert sh.vars["q/u0/W"].name == "q/u0/W:0"
