"""Drop-in ``solveBatch`` of the reference's bundle-entropy library, running on a B200.

Reference surfaces reproduced (paths under /root/reference):
    lib/bundle_entropy.py:192        solveBatch(fg, initXs, nIter=10, callback=None, solver='pc')
    lib/bundle_entropy_dual.py:129   solveBatch(fg, initXs, nIter=10, callback=None)
    RL/src/bundle_entropy.py:85      solveBatch(fg, initXs, nIter=5,  callback=None)
returning ``(x, A, b, lam, xs, nIters)`` (callers name A, b as G, h:
multi-label-cls/icnn_ebundle.py:225).

Two ways to supply ``fg``:
  * fused mode   -- ``fg`` is a :class:`icnn_b200.BoundPICNN` (``PICNN.bind(x)``): the whole loop
                    (K1 PICNN f/grad kernel + K2 bundle step) runs on the device, no host round
                    trip per iteration;
  * callback mode -- ``fg`` is any Python callable ``fg(x ndarray[B,n]) -> (f[B], g[B,n])``
                    (e.g. a conv-PICNN in torch): one host hop per iteration like the reference,
                    the per-sample bundle work (K2) still runs on the device.
There is no CPU implementation here: without the native library / a CUDA device this raises.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
import torch

from . import _capi
from .picnn import BoundPICNN

__all__ = ["solveBatch", "solve", "BundleState", "VARIANT_DEFAULTS"]

# reference defaults per copy: (nIter, solver, line_search, prune_thr)
VARIANT_DEFAULTS = {
    "lib": dict(nIter=10, line_search=True, prune_thr=1e-8),
    "dual": dict(nIter=10, line_search=False, prune_thr=0.0),
    "rl": dict(nIter=5, line_search=True, prune_thr=0.0),
}
_EPS64 = float(np.finfo(np.float64).eps)


class _Rows:
    """Lazy list-of-lists view over the dense device bundle buffers.  ``rows[u]`` is a Python
    list (what the reference returns) built on first access; fetching everything at once would
    be a multi-GB device->host copy at the stress sizes (hard part 7)."""

    def __init__(self, state, kind):
        self._s, self._kind = state, kind
        self._cache = {}

    def __len__(self):
        return self._s.B

    def __getitem__(self, u):
        if isinstance(u, slice):
            return [self[i] for i in range(*u.indices(len(self)))]
        if u < 0:
            u += len(self)
        if not 0 <= u < len(self):
            raise IndexError(u)
        if u not in self._cache:
            self._cache[u] = self._s._fetch(self._kind, u)
        return self._cache[u]

    def __iter__(self):
        return (self[u] for u in range(len(self)))


class BundleState:
    """Device buffers of one solveBatch call (icnn_bundle_bufs in include/icnn_b200.h)."""

    def __init__(self, B, n, KS, device, keep_xs=True, nIter=10):
        self.B, self.n, self.KS, self.device = int(B), int(n), int(KS), device
        f32, f64, i32 = torch.float32, torch.float64, torch.int32
        e = lambda *shape, dtype: torch.empty(*shape, dtype=dtype, device=device)  # noqa: E731
        self.y = e(B, n, dtype=f64)
        self.y32 = e(B, n, dtype=f32)
        self.f = e(B, dtype=f32)
        self.G = e(B, KS, n, dtype=f32)
        self.ys = e(B, KS, n, dtype=f64) if keep_xs else None
        self.h = torch.zeros(B, KS, dtype=f64, device=device)
        self.lam = torch.zeros(B, KS, dtype=f64, device=device)
        self.rsum = torch.zeros(B, KS, dtype=f64, device=device)
        self.gram = torch.zeros(B, KS, KS, dtype=f64, device=device)
        self.perm = e(B, KS, dtype=i32)
        self.count = e(B, dtype=i32)
        self.status = e(B, dtype=i32)
        self.finished = e(B, dtype=i32)
        self.nIters = e(B, dtype=i32)
        self.nactive = e(nIter + 1, dtype=i32)
        self.newton_its = e(B, dtype=i32)
        self.ksum = e(B, dtype=i32)
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        self.c = _capi.BundleBufs(self.B, self.n, self.KS, p(self.y), p(self.y32), p(self.f), p(self.G),
                                  p(self.ys), p(self.h), p(self.lam), p(self.rsum), p(self.gram),
                                  p(self.perm), p(self.count), p(self.status), p(self.finished),
                                  p(self.nIters), p(self.nactive), p(self.newton_its), p(self.ksum))
        self._host = None

    # ---- ragged outputs ---------------------------------------------------------------------
    def _host_small(self):
        if self._host is None:
            self._host = dict(perm=self.perm.cpu().numpy(), count=self.count.cpu().numpy(),
                              h=self.h.cpu().numpy(), lam=self.lam.cpu().numpy())
        return self._host

    def _fetch(self, kind, u):
        hs = self._host_small()
        k = int(hs["count"][u])
        slots = hs["perm"][u, :k]
        if kind == "lam":
            return hs["lam"][u, slots].copy() if k else None
        if kind == "b":
            return [float(v) for v in hs["h"][u, slots]]
        src = self.G if kind == "A" else self.ys
        if src is None:
            raise RuntimeError("xs were not kept (keep_xs=False)")
        if k == 0:
            return []
        idx = torch.as_tensor(slots.astype(np.int64), device=self.device)
        rows = src[u].index_select(0, idx).cpu().numpy()
        return [rows[j] for j in range(k)]


def _make_cfg(variant, solver, nIter, line_search, rank_tol, max_inner, n, KS):
    dflt = VARIANT_DEFAULTS[variant]
    if variant == "lib":
        if solver == "pc":
            sol = _capi.SOLVER_PC
        elif solver in ("newton", "boyd"):
            # 'boyd' (lib/bundle_entropy.py:80-156) solves the same strictly convex subproblem on
            # the full KKT system; it is accepted and mapped to the converged dual Newton solve.
            sol = _capi.SOLVER_NEWTON
        else:
            raise RuntimeError("Solver unknown: " + str(solver))  # lib/bundle_entropy.py:232
    else:
        sol = _capi.SOLVER_NEWTON
    ls = dflt["line_search"] if line_search is None else bool(line_search)
    if rank_tol is None:
        # np.linalg.matrix_rank: tol = sigma_max * max(k, n) * eps (float64 rows); x16 margin for
        # the explicit-residual test that replaces the SVD ("rank stop")
        rank_tol = 16.0 * max(KS, n) * _EPS64
    return _capi.BundleCfg(_capi.VARIANT[variant], sol, int(ls), int(max_inner), float(dflt["prune_thr"]),
                           float(rank_tol), int(nIter), 0)


def _dtype_of(a):
    if isinstance(a, torch.Tensor):
        return {torch.float32: np.float32, torch.float64: np.float64}.get(a.dtype, None)
    return np.asarray(a).dtype.type


def _to_numpy(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def solveBatch(fg, initXs, nIter=None, callback=None, solver="pc", *, variant="lib", line_search=None,
               rank_tol=None, max_inner=0, keep_xs=True, device=None, strict=False, return_state=False):
    """argmin_y f(x, y) - H(y) over [0,1]^n by the bundle-entropy method, on the GPU.

    Positional/keyword arguments are the reference's; keyword-only extras select which of the
    reference's three copies is reproduced (``variant`` in {'lib', 'dual', 'rl'}) and tuning.
    ``solver``: 'pc' (Mehrotra predictor-corrector, the reference default), 'boyd' (accepted,
    mapped to the same optimum) or 'newton' (converged dual projected Newton).
    Returns ``(x, A, b, lam, xs, nIters)``; A/b/xs are lazy list-of-lists views, lam a lazy list
    of arrays.  ``initXs`` (numpy) is overwritten with the result like the reference does
    (lib/bundle_entropy.py:200,228).
    """
    if variant not in VARIANT_DEFAULTS:
        raise ValueError("variant must be one of %s" % sorted(VARIANT_DEFAULTS))
    if variant != "lib":
        solver = "newton"
    if nIter is None:
        nIter = VARIANT_DEFAULTS[variant]["nIter"]
    nIter = int(nIter)
    if not torch.cuda.is_available():
        raise RuntimeError("icnn_b200.solveBatch needs a CUDA device (no CPU fallback)")
    fused = isinstance(fg, BoundPICNN)
    dev = fg.net.device if fused else torch.device(device if device is not None else "cuda")
    x0 = initXs
    B, n = x0.shape
    if B == 0 or nIter < 1:
        return (_to_numpy(x0), [[] for _ in range(B)], [[] for _ in range(B)], [None] * B,
                [[] for _ in range(B)], [nIter] * B)
    # slot capacity: active rows <= min(nIter, n) for the rank-tested copies, + 1 free slot
    KS = (nIter if variant == "rl" else min(nIter, n)) + 1
    KS = max(KS, 2)
    cfg = _make_cfg(variant, solver, nIter, line_search, rank_tol, max_inner, n, KS)
    with torch.cuda.device(dev):
        st = BundleState(B, n, KS, dev, keep_xs=keep_xs, nIter=nIter)
        if isinstance(x0, torch.Tensor):
            st.y.copy_(x0.to(device=dev, dtype=torch.float64))
        else:
            st.y.copy_(torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float64)), non_blocking=False)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if fused and callback is None:
            if fg.B != B or fg.net.n != n:
                raise ValueError("fg is bound to a [%d, %d] problem, initXs is %s" % (fg.B, fg.net.n, (B, n)))
            _capi.check(_capi.lib.icnn_solve_batch_fused(fg.net._h, C.byref(fg.c_gates), C.byref(cfg),
                                                         C.byref(st.c), fg.ws.data_ptr(), stream))
        else:
            _capi.check(_capi.lib.icnn_bundle_init(C.byref(st.c), nIter, stream))
            for t in range(nIter):
                if fused:
                    _capi.check(_capi.lib.icnn_picnn_fg(
                        fg.net._h, C.byref(fg.c_gates), st.y32.data_ptr(), st.f.data_ptr(),
                        st.G.data_ptr(), 0, st.perm.data_ptr(), st.count.data_ptr(), KS,
                        fg.ws.data_ptr(), None, stream))
                    if callback is not None:
                        fi = st.f.cpu().numpy()
                        xi = st.y.cpu().numpy()
                        if variant == "rl":
                            callback(t, fi)          # RL/src/bundle_entropy.py:103-104
                        else:
                            callback(t, fi, xi)      # lib/bundle_entropy.py:208-209
                else:
                    xi = st.y.cpu().numpy()
                    fi, gi = fg(xi)
                    if t == 0 and rank_tol is None and variant != "rl" and _dtype_of(gi) == np.float32:
                        # np.linalg.matrix_rank scales its tolerance with the dtype of the rows: a
                        # float32 fg (the reference's TF fetch) stops samples at max(k, n) * eps32
                        cfg.rank_tol = float(max(KS, n) * np.finfo(np.float32).eps)
                    if callback is not None:
                        if variant == "rl":
                            callback(t, _to_numpy(fi))
                        else:
                            callback(t, _to_numpy(fi), xi)
                    fd = torch.as_tensor(fi, device=dev).to(torch.float32).contiguous().reshape(B)
                    gd = torch.as_tensor(gi, device=dev).to(torch.float32).contiguous().reshape(B, n)
                    _capi.check(_capi.lib.icnn_bundle_put_fg(C.byref(st.c), fd.data_ptr(), gd.data_ptr(), stream))
                _capi.check(_capi.lib.icnn_bundle_step(C.byref(cfg), C.byref(st.c), t, stream))
                if int(st.nactive[t + 1].item()) == 0:   # lib/bundle_entropy.py:239
                    break
        x = st.y.cpu().numpy()
        status = st.status.cpu().numpy()
    if np.any(status == _capi.ST_NONFINITE) or np.any(status == _capi.ST_SOLVE_FAIL):
        msg = "solveBatch: %d samples non-finite, %d with a failed inner solve" % (
            int(np.sum(status == _capi.ST_NONFINITE)), int(np.sum(status == _capi.ST_SOLVE_FAIL)))
        if strict:
            raise RuntimeError(msg)
        warnings.warn(msg)   # the reference runs under np.seterr(all='warn')
    if isinstance(x0, np.ndarray) and x0.dtype == np.float64:
        x0[...] = x
        x = x0
    nIters = [int(v) for v in st.nIters.cpu().numpy()]
    st.status_host = status
    out = (x, _Rows(st, "A"), _Rows(st, "b"), _Rows(st, "lam"), _Rows(st, "xs"), nIters)
    if return_state:
        return out + (st,)
    return out


def solve(fg, initX, nIter=10, callback=None, *, variant="dual", **kw):
    """Single-sample form of the reference (lib/bundle_entropy_dual.py:87-127 ``solve``; the
    ``solve`` of lib/bundle_entropy.py:168-190 is dead code there): ``fg(x [n]) -> (f, g [n])``,
    returns the final iterate.  Runs as a batch of one through :func:`solveBatch`."""
    x0 = np.array(initX, dtype=np.float64).reshape(1, -1)

    def fg1(xb):
        f, g = fg(xb[0])
        return np.atleast_1d(np.asarray(f, dtype=np.float64)), np.asarray(g, dtype=np.float64).reshape(1, -1)

    cb = None if callback is None else (lambda t, f, x=None: callback(t, f[0], None if x is None else x[0]))
    out = solveBatch(fg1, x0, nIter=nIter, callback=cb, variant=variant, **kw)
    return out[0][0]
