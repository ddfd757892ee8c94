"""numpy restatement of the reference's argmin differentiation.

TEST INFRASTRUCTURE ONLY.  Differentiates y* = argmin of the bundle model through its KKT system
and assembles the per-bundle-point (v, c) pairs the training step feeds back to the graph:
    crossEntrGrad   multi-label-cls/icnn_ebundle.py:390-417
    mseGrad         completion/icnn_ebundle.py:493-522
    train_step_fd   multi-label-cls/icnn_ebundle.py:296-314 / completion/icnn_ebundle.py:315-335
Pinned by tests/golden/argmin_grad.npz, produced by exec'ing the reference's own function
bodies and the reference's own ``Model.train_step_fd`` methods (extracted with ast; the scripts
themselves import TensorFlow at module level and cannot be imported) -- oracle/gen_golden_grad.py.
"""
import numpy as np


def argmin_grad(y, trueY, G, loss):
    """Returns (cy [n], clam [k], ct [1]).  loss = 'xent' (crossEntrGrad) or 'mse' (mseGrad)."""
    k, n = G.shape
    if loss == "xent":
        y_ = np.clip(y, 1e-8, 1.0 - 1e-8)                      # :393-395
        dl = trueY / y_ - (1.0 - trueY) / (1.0 - y_)           # :410
    else:
        y_ = y
        dl = -(y - trueY)                                      # completion :515
    zinv = 1.0 / (1.0 / y_ + 1.0 / (1.0 - y_))
    Gz = G * zinv
    H = np.zeros((k + 1, k + 1))
    H[:k, :k] = Gz.dot(G.T)
    H[:k, k] = 1.0
    H[k, :k] = 1.0
    b = np.concatenate([Gz.dot(dl), [0.0]])
    sol = np.linalg.solve(H, b)
    clam, ct = sol[:k], sol[k:]
    cy = zinv * dl - Gz.T.dot(clam)
    cy[(y == 0) | (y == 1)] = 0
    return cy, clam, ct


def train_step_pairs(yN, trueY, G, ys, lam, loss):
    """(fd_ys, fd_vs, fd_cs) of train_step_fd for one minibatch: one row per (sample, bundle point)."""
    fys, fvs, fcs = [], [], []
    for j in range(len(G)):
        if len(G[j]) == 0:
            continue
        cy, clam, _ = argmin_grad(yN[j], trueY[j], np.array(G[j], dtype=np.float64), loss)
        for i in range(len(G[j])):
            fys.append(ys[j][i])
            fvs.append(lam[j][i] * cy + clam[i] * (yN[j] - ys[j][i]))
            fcs.append(clam[i])
    return np.array(fys), np.array(fvs), np.array(fcs)
