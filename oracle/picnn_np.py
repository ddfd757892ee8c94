"""numpy float64 restatement of the fully-connected PICNN energy f(x, y; theta) and df/dy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference builds this function as a
TensorFlow graph, which cannot run here (TensorFlow/tflearn absent), so it is restated from the
cited lines; tflearn ``fully_connected`` is ``out = in @ W + b`` with W laid out ``[in, out]``.

Follows (paths relative to /root/reference):
  * multi-label-cls/icnn_ebundle.py:316-388  (Model.f: ReLU PICNN, ``nLabels`` appended to szs)
  * RL/src/icnn.py:325-404                   (Agent.negQ: leaky-ReLU PICNN, szs = [l1, l2])
  * multi-label-cls/icnn_ebundle.py:146      (dE_dy_ = tf.gradients(E_, y_))
  * RL/src/icnn.py:148-158                   (affine wrapper x in [0,1] -> a = 2x-1, grad *= 2)
  * multi-label-cls/icnn-back.py:116-131     (momentum gradient-descent inner loop)

Parity status of THIS file: "parity unpinned" by reference execution (the TF graph cannot be
run); it is pinned instead by (i) a finite-difference check of the analytic gradient and
(ii) an independently written torch-autograd forward (tests/test_oracle_picnn.py).

Index convention: z-layers i = 0..L, widths s_0..s_{L-1} = ``hidden``, s_L = 1.
"""
from __future__ import annotations

import numpy as np


class PicnnParams:
    """Weights of a fully-connected PICNN.

    hidden : list of z-layer widths s_0..s_{L-1} (the output layer of width 1 is implicit).
    Wy[i]  : [n, s_i]         'z{i}_yu/W'       (bias-free, unconstrained)
    Wz[i]  : [s_{i-1}, s_i]   'z{i}_zu_proj/W'  (bias-free, >= 0), i >= 1 (Wz[0] is None)
    x-path (evaluated once per solveBatch, outside the hot loop):
      Wu[i], bu[i]   : u_i = u_{i-1} @ Wu[i] + bu[i]  (relu for i < L-1), i = 0..L-1
      Wzu[i], bzu[i] : cz_i = relu(P_i @ Wzu[i] + bzu[i])   in R^{s_{i-1}}, i >= 1
      Wyu[i], byu[i] : cy_i =      P_i @ Wyu[i] + byu[i]    in R^{n}
      Wzx[i], bzx[i] : d_i  =      P_i @ Wzx[i] + bzx[i]    in R^{s_i}
      with P_0 = x, P_i = u_{i-1}.
    alpha : leaky-ReLU slope on the z path (0 -> ReLU, multi-label; 0.01 RL).
    """

    def __init__(self, m, n, hidden, alpha=0.0):
        self.m, self.n, self.hidden, self.alpha = int(m), int(n), [int(s) for s in hidden], float(alpha)
        self.L = len(self.hidden)
        self.sizes = self.hidden + [1]
        L = self.L
        self.Wy = [None] * (L + 1)
        self.Wz = [None] * (L + 1)
        self.Wu, self.bu = [None] * L, [None] * L
        self.Wzu, self.bzu = [None] * (L + 1), [None] * (L + 1)
        self.Wyu, self.byu = [None] * (L + 1), [None] * (L + 1)
        self.Wzx, self.bzx = [None] * (L + 1), [None] * (L + 1)

    def prev_width(self, i):
        """Width of P_i (the x-path activation feeding layer i's gates)."""
        return self.m if i == 0 else self.hidden[i - 1]


def synth_params(seed, m, n, hidden, alpha=0.0, gate_bias=0.0, dtype=np.float32):
    """Seeded synthetic weights: Wy ~ N(0,1/n), Wz = |N(0,1/s_prev)|,
    x-path/gate weights N(0, 1/fan_in), biases 0 (RL: gate biases 1, RL/src/icnn.py:364,375).
    Values are rounded to ``dtype`` (float32 = what the device stores) but returned as float64
    arrays so oracle and device see bit-identical parameters."""
    rs = np.random.RandomState(seed)
    p = PicnnParams(m, n, hidden, alpha)
    L = p.L

    def rnd(shape, fan_in):
        return (rs.randn(*shape) / np.sqrt(fan_in)).astype(dtype).astype(np.float64)

    for i in range(L):
        fin = p.prev_width(i)
        p.Wu[i] = rnd((fin, p.hidden[i]), fin)
        p.bu[i] = np.zeros(p.hidden[i])
    for i in range(L + 1):
        fin = p.prev_width(i)
        si = p.sizes[i]
        if i > 0:
            sp = p.sizes[i - 1]
            p.Wzu[i] = rnd((fin, sp), fin)
            p.bzu[i] = np.full(sp, gate_bias)
            p.Wz[i] = np.abs(rnd((sp, si), sp))
        p.Wyu[i] = rnd((fin, n), fin)
        p.byu[i] = np.full(n, gate_bias)
        p.Wy[i] = rnd((n, si), n)
        p.Wzx[i] = rnd((fin, si), fin)
        p.bzx[i] = np.zeros(si)
    return p


def gates(p: PicnnParams, x):
    """x-path: returns (cz, cy, d) lists indexed by z-layer (cz[0] is None).
    multi-label-cls/icnn_ebundle.py:339-347 (u path; batch-norm is treated as caller-supplied,
    i.e. identity here -- tflearn BN defaults are un-pinned),
    :354-356 (cz), :363-365 (cy), :372-373 (d)."""
    x = np.asarray(x, dtype=np.float64)
    L = p.L
    us = []
    prev = x
    for i in range(L):
        u = prev @ p.Wu[i] + p.bu[i]
        if i < L - 1:
            u = np.maximum(u, 0.0)
        us.append(u)
        prev = u
    cz, cy, d = [None] * (L + 1), [None] * (L + 1), [None] * (L + 1)
    for i in range(L + 1):
        P = x if i == 0 else us[i - 1]
        if i > 0:
            cz[i] = np.maximum(P @ p.Wzu[i] + p.bzu[i], 0.0)
        cy[i] = P @ p.Wyu[i] + p.byu[i]
        d[i] = P @ p.Wzx[i] + p.bzx[i]
    return cz, cy, d


def fg_gated(p: PicnnParams, gts, y, dtype=np.float64):
    """Energy f [B] and gradient df/dy [B, n] for iterate y [B, n] given precomputed gates.
    Forward  multi-label-cls/icnn_ebundle.py:349-387 / RL/src/icnn.py:356-404;
    gradient = what tf.gradients(E_, y_) (:146) evaluates, written out by hand."""
    cz, cy, d = gts
    y = np.asarray(y, dtype=dtype)
    L, a = p.L, p.alpha
    zs = []
    z = None
    for i in range(L + 1):
        pre = (y * cy[i].astype(dtype)) @ p.Wy[i].astype(dtype) + d[i].astype(dtype)
        if i > 0:
            pre = pre + (z * cz[i].astype(dtype)) @ p.Wz[i].astype(dtype)
        if i < L:
            z = np.where(pre > 0, pre, a * pre)
        else:
            z = pre
        zs.append(z)
    f = zs[L][:, 0]
    delta = np.ones_like(zs[L])
    g = np.zeros_like(y)
    for i in range(L, -1, -1):
        g += cy[i].astype(dtype) * (delta @ p.Wy[i].astype(dtype).T)
        if i > 0:
            dact = np.where(zs[i - 1] > 0, 1.0, a).astype(dtype)
            delta = dact * cz[i].astype(dtype) * (delta @ p.Wz[i].astype(dtype).T)
    return f, g


def make_fg(p: PicnnParams, x, dtype=np.float64, out_dtype=None, affine=False):
    """Closure fg(y) -> (f, g) with the reference's callback contract
    (multi-label-cls/icnn_ebundle.py:218-221).  ``dtype=np.float32`` mimics the TF graph's
    arithmetic, ``out_dtype=np.float32`` mimics the float32 fetch.  ``affine=True`` applies the
    RL wrapper (RL/src/icnn.py:148-153): the solver variable is x in [0,1], a = 2x-1, grad *= 2."""
    gts = gates(p, x)

    def fg(y):
        yy = 2.0 * np.asarray(y) - 1.0 if affine else y
        f, g = fg_gated(p, gts, yy, dtype=dtype)
        if affine:
            g = 2.0 * g
        if out_dtype is not None:
            f, g = f.astype(out_dtype), g.astype(out_dtype)
        return f, g

    return fg


def momentum_gd(fg, y0, nIter, lr, momentum):
    """Unrolled momentum gradient descent, multi-label-cls/icnn-back.py:120-131
    (= completion/icnn.back.py:133-147): v' = m v - lr g(y); y' = y - m v + (1+m) v'.
    No projection.  Returns (y_n, f(y_n))."""
    y = np.array(y0, dtype=np.float64)
    v = np.zeros_like(y)
    for _ in range(nIter):
        _, g = fg(y)
        v_new = momentum * v - lr * np.asarray(g, dtype=np.float64)
        y = y - momentum * v + (1.0 + momentum) * v_new
        v = v_new
    f, _ = fg(y)
    return y, np.asarray(f, dtype=np.float64)


def entr(y):
    """Sum_j -y log y - (1-y) log(1-y) with 0 log 0 = 0 (multi-label-cls/ebundle-vs-gd.py:38-41)."""
    y = np.asarray(y, dtype=np.float64)
    with np.errstate(all="ignore"):
        z = -y * np.log(y) - (1.0 - y) * np.log(1.0 - y)
    z[z != z] = 0.0
    return z.sum(axis=1)
