"""numpy float64 restatement of d loss / d theta through the unrolled momentum-GD inner loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What the reference computes: ``opt.compute_gradients(self.mse_, self.theta_)`` on the graph that
unrolls ``nIter`` momentum-GD steps on the energy (multi-label-cls/icnn-back.py:120-139 =
completion/icnn.back.py:133-156), i.e. TensorFlow double-backprop through
``tf.gradients(Ei_, yi_)``.  TensorFlow is absent here, so this file restates the result
analytically.  Pins (tests/test_oracle_gd_grad.py, tests/test_oracle_tfshim.py): (i) the reference's own
``Model.__init__`` / ``Model.f`` of multi-label-cls/icnn-back.py, unmodified, executed on the TF-primitive
stand-in oracle/tf_shim.py -- ``yn_``, ``mse_`` and every entry of ``opt.compute_gradients(mse_, theta_)`` agree
to 1e-9 (goldens tests/golden/picnn_tfshim.npz); (ii) an independently written torch autograd
(float64, ``create_graph=True``) of the same recurrence; (iii) central finite differences.  Not executable here:
the real TensorFlow kernels (the stand-in's autodiff is torch's).

Derivation (ReLU / leaky-ReLU energies are piecewise linear in y, so d2f/dy2 = 0 almost everywhere,
exactly what TF's ReluGrad-of-ReluGrad yields):

    v_{i+1} = m v_i - lr g_i,   y_{i+1} = y_i - m v_i + (1+m) v_{i+1},   g_i = df/dy(y_i; theta)

    adjoint of y_i:  a = dl/dy_N for every i (the Hessian term vanishes)
    adjoint of v_i:  c_i * a with c_N = 1+m, c_i = m c_{i+1} + 1
    adjoint of g_i:  kappa_i * a,  kappa_i = -lr c_{i+1}

    dl/dtheta = sum_i kappa_i * d/dtheta <g_i, a>  =  sum_i kappa_i * d/dtheta (Jf(y_i)[a])

With the activation pattern of iterate i fixed, the directional derivative Jf(y_i)[a] is the linear
"tangent network"  zt_l = act'(pre_l) o ((zt_{l-1} o cz_l) Wz_l + (a o cy_l) Wy_l)  whose backprop
multipliers are the primal delta_l.  Hence, per layer l (delta_L = 1):

    dWy_l = (a o cy_l)^T Delta_l            Delta_l = sum_i kappa_i delta_l^(i)
    dcy_l = a o (Delta_l Wy_l^T)
    dWz_l = sum_i kappa_i (zt_{l-1}^(i) o cz_l)^T delta_l^(i)
    dcz_l = sum_i kappa_i zt_{l-1}^(i) o (delta_l^(i) Wz_l^T)
    dd_l  = 0

`xpath_backward` chains (dcy, dcz) into the x-path parameters (plain dense-layer backprop of
multi-label-cls/icnn-back.py:255-262,269-291).
"""
from __future__ import annotations

import numpy as np


def kappas(nIter, lr, momentum):
    """kappa_i, i = 0..nIter-1 (adjoint weight of g_i relative to a = dl/dy_N)."""
    c = np.zeros(nIter + 1)
    if nIter > 0:
        c[nIter] = 1.0 + momentum
        for i in range(nIter - 1, 0, -1):
            c[i] = momentum * c[i + 1] + 1.0
    return np.array([-lr * c[i + 1] for i in range(nIter)])


def _forward(p, gts, y):
    cz, cy, d = gts
    L, al = p.L, p.alpha
    zs, z = [], None
    for i in range(L + 1):
        pre = (y * cy[i]) @ p.Wy[i].astype(np.float64) + d[i]
        if i > 0:
            pre = pre + (z * cz[i]) @ p.Wz[i].astype(np.float64)
        z = np.where(pre > 0, pre, al * pre) if i < L else pre
        zs.append(z)
    return zs


def gd_backward(p, gts, y0, nIter, lr, momentum, dl_dyn_fn):
    """Runs the GD loop from y0 and returns (y_N, grads) with
    grads = dict(dWy=[L+1], dWz=[L+1] (dWz[0] None), dcy=[L+1], dcz=[L+1] (dcz[0] None)).
    ``dl_dyn_fn(y_N) -> a`` is the loss gradient at the GD output, e.g. 2 (y_N - trueY) / (B n) for
    ``mse_ = reduce_mean(square(yn - trueY))`` (multi-label-cls/icnn-back.py:133)."""
    cz, cy, d = [[None if g is None else np.asarray(g, dtype=np.float64) for g in gs] for gs in gts]
    gts = (cz, cy, d)
    L, al = p.L, p.alpha
    Wy = [w.astype(np.float64) for w in p.Wy]
    Wz = [None] + [w.astype(np.float64) for w in p.Wz[1:]]
    y = np.array(y0, dtype=np.float64)
    v = np.zeros_like(y)
    traj = []
    for _ in range(nIter):
        zs = _forward(p, gts, y)
        delta = [None] * (L + 1)
        delta[L] = np.ones_like(zs[L])
        g = np.zeros_like(y)
        for i in range(L, -1, -1):
            g += cy[i] * (delta[i] @ Wy[i].T)
            if i > 0:
                dact = np.where(zs[i - 1] > 0, 1.0, al)
                delta[i - 1] = dact * cz[i] * (delta[i] @ Wz[i].T)
        traj.append((zs, delta))
        v_new = momentum * v - lr * g
        y = y - momentum * v + (1.0 + momentum) * v_new
        v = v_new
    a = np.asarray(dl_dyn_fn(y), dtype=np.float64)
    kap = kappas(nIter, lr, momentum)
    dWy = [np.zeros_like(w) for w in Wy]
    dWz = [None] + [np.zeros_like(w) for w in Wz[1:]]
    dcy = [np.zeros_like(c) for c in cy]
    dcz = [None] + [np.zeros_like(c) for c in cz[1:]]
    Delta = [np.zeros_like(dl) for dl in traj[0][1]] if nIter else None
    for i in range(nIter):
        zs, delta = traj[i]
        zt = None
        for l in range(L + 1):
            if l > 0:
                dWz[l] += kap[i] * (zt * cz[l]).T @ delta[l]
                dcz[l] += kap[i] * zt * (delta[l] @ Wz[l].T)
            Delta[l] += kap[i] * delta[l]
            if l < L:
                t = (a * cy[l]) @ Wy[l]
                if l > 0:
                    t = t + (zt * cz[l]) @ Wz[l]
                zt = np.where(zs[l] > 0, 1.0, al) * t
    for l in range(L + 1):
        if nIter:
            dWy[l] = (a * cy[l]).T @ Delta[l]
            dcy[l] = a * (Delta[l] @ Wy[l].T)
    return y, dict(dWy=dWy, dWz=dWz, dcy=dcy, dcz=dcz)


def xpath_backward(p, x, dcy, dcz):
    """Chain gate adjoints into the x-path parameters (dd = 0).  Returns a dict name -> list."""
    x = np.asarray(x, dtype=np.float64)
    L = p.L
    W = lambda ws: [None if w is None else np.asarray(w, dtype=np.float64) for w in ws]  # noqa: E731
    Wu, Wzu, Wyu = W(p.Wu), W(p.Wzu), W(p.Wyu)
    us, pres, prev = [], [], x
    for i in range(L):
        pre = prev @ Wu[i] + p.bu[i]
        u = np.maximum(pre, 0.0) if i < L - 1 else pre
        pres.append(pre); us.append(u); prev = u
    out = dict(dWu=[None] * L, dbu=[None] * L, dWzu=[None] * (L + 1), dbzu=[None] * (L + 1),
               dWyu=[None] * (L + 1), dbyu=[None] * (L + 1))
    dU = [np.zeros_like(u) for u in us]
    for i in range(L, -1, -1):
        P = x if i == 0 else us[i - 1]
        dP = dcy[i] @ Wyu[i].T
        out["dWyu"][i] = P.T @ dcy[i]
        out["dbyu"][i] = dcy[i].sum(0)
        if i > 0:
            pz = dcz[i] * ((P @ Wzu[i] + p.bzu[i]) > 0)
            out["dWzu"][i] = P.T @ pz
            out["dbzu"][i] = pz.sum(0)
            dP = dP + pz @ Wzu[i].T
            dU[i - 1] += dP
    for i in range(L - 1, -1, -1):
        du = dU[i] * (pres[i] > 0) if i < L - 1 else dU[i]
        P = x if i == 0 else us[i - 1]
        out["dWu"][i] = P.T @ du
        out["dbu"][i] = du.sum(0)
        if i > 0:
            dU[i - 1] += du @ Wu[i].T
    return out
