"""numpy restatement of the RL agent's default inner optimiser.

TEST INFRASTRUCTURE ONLY.  Follows RL/src/icnn.py:
    adam(self, func, obs)      :160-215   batched Adam on the action, best-so-far tracking,
                                          rolling-average stop (a_diff < 1e-3 and i > 5), clip to (-1, 1)
    entropy(x)                 :455-458   the penalty subtracted from negQ (func = _fg_entr, :60-63,127-131)
Pinned by tests/golden/adam.npz, produced by exec'ing the reference's own `adam` body
(oracle/gen_golden_adam.py).
"""
import numpy as np

from . import picnn_np


def make_fg_entr(p, obs, dtype=np.float64):
    """func(obs, act) -> [negQ - entropy(act), d/dact]  (RL/src/icnn.py:60-63; entropy :455-458;
    tf.clip_by_value passes the gradient inside [1e-4, 1-1e-4] only)."""
    gts = picnn_np.gates(p, obs)

    def func(_obs, act):
        f, g = picnn_np.fg_gated(p, gts, act, dtype=dtype)
        xr_raw = (np.asarray(act, dtype=np.float64) + 1.0) / 2.0
        xr = np.clip(xr_raw, 0.0001, 0.9999)
        pen = xr * np.log(xr) + (1.0 - xr) * np.log(1.0 - xr)
        inside = (xr_raw >= 0.0001) & (xr_raw <= 0.9999)
        gpen = np.where(inside, 0.5 * (np.log(xr) - np.log(1.0 - xr)), 0.0)
        return [f + pen.sum(axis=1), g + gpen]

    return func


def adam(func, obs, dimA, max_iter=1000):
    """RL/src/icnn.py:160-215 (plotting removed).  Returns (act_best, iterations)."""
    b1, b2, lam, eps, alpha = 0.9, 0.999, 0.5, 1e-8, 0.01
    nBatch = obs.shape[0]
    act = np.zeros((nBatch, dimA))
    m = np.zeros_like(act)
    v = np.zeros_like(act)
    b1t, b2t = 1.0, 1.0
    act_best, a_diff, f_best = None, None, None
    for i in range(max_iter):
        f, g = func(obs, act)
        if i == 0:
            act_best = act.copy()
            f_best = np.array(f, dtype=np.float64).copy()
        else:
            prev = act_best.copy()
            I = f < f_best
            act_best[I] = act[I]
            f_best[I] = f[I]
            a_diff_i = np.mean(np.linalg.norm(act_best - prev, axis=1))
            a_diff = a_diff_i if a_diff is None else lam * a_diff + (1.0 - lam) * a_diff_i
            if a_diff < 1e-3 and i > 5:
                return act_best, i
        m = b1 * m + (1.0 - b1) * g
        v = b2 * v + (1.0 - b2) * (g * g)
        b1t *= b1
        b2t *= b2
        mhat = m / (1.0 - b1t)
        act = act - alpha * mhat / (np.sqrt(v) + eps)       # (the reference divides by sqrt(v), not vhat)
        act = np.clip(act, -1.0 + 1e-8, 1.0 - 1e-8)
    return act_best, max_iter
