"""Seeded synthetic workloads at the reference's dimensions (BASELINE.md
section 4) and the plain container for PICNN weights.

Pure numpy, no algorithm: the reference's data files are not in its repo and there is no
network, so every configuration is synthetic.  Everything is generated from
``np.random.RandomState(seed)``; values are float32-representable but returned as float64 so
the oracle and the device see bit-identical inputs.  Shared by bench.py, the tests and the
oracle (which imports it as a data generator only).
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
    bn[i] : optional inference-mode batch-norm after the ReLU of u_i, i < L-1
            (multi-label-cls/icnn_ebundle.py:343-345, RL/src/icnn.py:348-351 with --icnn_bn), given as the
            per-feature affine map it is at inference time: ``(scale, shift)`` with
            scale = gamma / sqrt(moving_var + eps), shift = beta - moving_mean * scale  (``bn_affine``).
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
        self.bn = [None] * L

    def prev_width(self, i):
        """Width of P_i (the x-path activation feeding layer i's gates)."""
        return self.m if i == 0 else self.hidden[i - 1]


def bn_affine(gamma, beta, moving_mean, moving_var, eps=1e-5):
    """Inference-mode batch normalisation as a per-feature affine map (tflearn batch_normalization with
    is_training = False: (x - mean) / sqrt(var + eps) * gamma + beta).  eps: tflearn's default is 1e-5."""
    scale = np.asarray(gamma, dtype=np.float64) / np.sqrt(np.asarray(moving_var, dtype=np.float64) + eps)
    return scale, np.asarray(beta, dtype=np.float64) - np.asarray(moving_mean, dtype=np.float64) * scale


def fold_batchnorm(p):
    """Fold the inference-mode batch-norm of the u-path into the weights of its consumers and return a copy
    WITHOUT bn:  u_i = s o r + t (r = relu(fc)) feeds u_{i+1} and the three gate layers of z-layer i+1 as
    P W + b, and (s o r + t) W + b = r (diag(s) W) + (b + t W).  Exact in real arithmetic; this is how the
    device path honours BN (the x-path GEMM epilogue stays bias + ReLU)."""
    import copy
    q = copy.deepcopy(p)
    for i in range(p.L - 1):
        if p.bn[i] is None:
            continue
        s_, t_ = (np.asarray(v, dtype=np.float64) for v in p.bn[i])
        for W, b in ((q.Wu, q.bu), (q.Wzu, q.bzu), (q.Wyu, q.byu), (q.Wzx, q.bzx)):
            W0 = np.asarray(W[i + 1], dtype=np.float64)
            b[i + 1] = np.asarray(b[i + 1], dtype=np.float64) + t_ @ W0
            W[i + 1] = s_[:, None] * W0
        q.bn[i] = None
    return q


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



# name -> dict(m, n, hidden, B, nIter, variant, alpha, gate_bias, affine, seed, y0)
CONFIGS = {
    # configs[0]: plumbing, the reference's CPU-runnable case
    "C1": dict(m=8, n=8, hidden=[16, 16], B=64, nIter=5, variant="lib", alpha=0.0, gate_bias=0.0,
               affine=False, seed=1, y0="half", xdist="normal"),
    # configs[1]: Olivetti dims, FC stand-in for the conv PICNN (completion/icnn_ebundle.py:345)
    "C2": dict(m=2048, n=2048, hidden=[512, 512], B=400, nIter=30, variant="lib", alpha=0.0,
               gate_bias=0.0, affine=False, seed=2, y0="meanvec", xdist="normal"),
    # configs[2]: Bibtex dims, layerSizes [600] -> [600, 159] (multi-label-cls/icnn_ebundle.py:321-323)
    "C3": dict(m=1836, n=159, hidden=[600, 159], B=4096, nIter=10, variant="lib", alpha=0.0,
               gate_bias=0.0, affine=False, seed=3, y0="half", xdist="normal"),
    # configs[3]: RL HalfCheetah dims (RL/src/agent.py: l1size=l2size=200, lrelu=0.01)
    "C4": dict(m=17, n=6, hidden=[200, 200], B=65536, nIter=5, variant="rl", alpha=0.01,
               gate_bias=1.0, affine=True, seed=4, y0="half", xdist="uniform"),
    # configs[4]: stress
    "C5": dict(m=512, n=4096, hidden=[1024, 1024, 1024, 1024], B=8192, nIter=50, variant="lib",
               alpha=0.0, gate_bias=0.0, affine=False, seed=5, y0="half", xdist="normal"),
    # north_star target shape (batch 4096 / n_y 512)
    "T": dict(m=512, n=512, hidden=[1024, 1024], B=4096, nIter=10, variant="lib", alpha=0.0,
              gate_bias=0.0, affine=False, seed=6, y0="half", xdist="normal"),
}


WY_SCALE = {"C1": 3.0, "C2": 3.0, "C3": 3.0, "C4": 1.0, "C5": 3.0, "T": 3.0}
for _k, _v in WY_SCALE.items():
    # W^y ~ N(0, wy_scale^2 / n).  SURVEY.md section 8d proposes N(0, 1/n); measured here, that makes
    # the ReLU PICNN so flat that every sample hits the duplicate-row stop after ~5 iterations
    # (C2: all 30-iteration solves end at t=5 with k<=6).  Scale 3 gives the bundle growth the
    # survey describes (C2: k up to 25, mean 17 iterations) -- see DESIGN.md "workloads".
    CONFIGS[_k]["wy_scale"] = _v


def make_inputs(cfg, B=None, seed=None):
    """Returns (params, x [B,m] float64, y0 [B,n] float64) for a config dict (or name)."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    B = cfg["B"] if B is None else B
    seed = cfg["seed"] if seed is None else seed
    p = synth_params(seed, cfg["m"], cfg["n"], cfg["hidden"], alpha=cfg["alpha"],
                              gate_bias=cfg["gate_bias"])
    sc = np.float32(cfg.get("wy_scale", 1.0))
    for i in range(len(p.Wy)):
        p.Wy[i] = (p.Wy[i].astype(np.float32) * sc).astype(np.float64)
    rs = np.random.RandomState(seed + 1000)
    if cfg["xdist"] == "uniform":
        x = rs.uniform(-1.0, 1.0, size=(B, cfg["m"]))
    else:
        x = rs.randn(B, cfg["m"])
    x = x.astype(np.float32).astype(np.float64)
    if cfg["y0"] == "meanvec":   # completion/icnn_ebundle.py:223 starts from the per-pixel mean
        v = rs.uniform(0.2, 0.8, size=(1, cfg["n"])).astype(np.float32).astype(np.float64)
        y0 = np.repeat(v, B, axis=0)
    else:
        y0 = np.full((B, cfg["n"]), 0.5)
    return p, x, y0
