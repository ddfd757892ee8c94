"""Argmin differentiation on the device.

The training step of the reference differentiates y* = argmin of the final bundle model through
its KKT system and feeds per-bundle-point pairs (v, c) back to the graph:
    crossEntrGrad   multi-label-cls/icnn_ebundle.py:390-417
    mseGrad         completion/icnn_ebundle.py:493-522
    train_step_fd   multi-label-cls/icnn_ebundle.py:296-314 / completion/icnn_ebundle.py:315-335
Here it runs on the bundle state left on the device by ``solveBatch(..., return_state=True)``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi

LOSS = {"mse": 0, "xent": 1, "crossentr": 1}


def argmin_grad(state, trueY, loss="xent", assemble=True, return_device=False):
    """cy [B,n], clam (list of arrays, bundle order), ct [B] and, if ``assemble``, the stacked
    train_step_fd feeds (fd_ys, fd_vs, fd_cs) with one row per (sample, bundle point)."""
    if loss not in LOSS:
        raise ValueError("loss must be 'mse' or 'xent'")
    dev = state.device
    B, n, KS = state.B, state.n, state.KS
    with torch.cuda.device(dev):
        tY = torch.as_tensor(np.ascontiguousarray(trueY, dtype=np.float64), device=dev) \
            if not isinstance(trueY, torch.Tensor) else trueY.to(device=dev, dtype=torch.float64).contiguous()
        cy = torch.empty(B, n, dtype=torch.float64, device=dev)
        clam = torch.zeros(B, KS, dtype=torch.float64, device=dev)
        ct = torch.empty(B, dtype=torch.float64, device=dev)
        V = torch.empty(B, KS, n, dtype=torch.float64, device=dev) if (assemble and state.ys is not None) else None
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _capi.check(_capi.lib.icnn_argmin_grad(C.byref(state.c), LOSS[loss], tY.data_ptr(), cy.data_ptr(),
                                               clam.data_ptr(), ct.data_ptr(),
                                               None if V is None else V.data_ptr(), stream))
        if return_device:
            return cy, clam, ct, V
        count = state.count.cpu().numpy()
        cy_h, clam_h, ct_h = cy.cpu().numpy(), clam.cpu().numpy(), ct.cpu().numpy()
        out = (cy_h, [clam_h[u, :count[u]].copy() for u in range(B)], ct_h)
        if V is None:
            return out
        # gather the (sample, bundle point) rows in train_step_fd order
        idx_u = np.repeat(np.arange(B), count)
        idx_i = np.concatenate([np.arange(c) for c in count]) if count.sum() else np.zeros(0, dtype=np.int64)
        iu = torch.as_tensor(idx_u, device=dev)
        ii = torch.as_tensor(idx_i, device=dev)
        perm = state.perm.long()
        slots = perm[iu, ii]
        fd_vs = V[iu, ii].cpu().numpy()
        fd_ys = state.ys[iu, slots].cpu().numpy()
        fd_cs = clam[iu, ii].cpu().numpy()
        return out + ((fd_ys, fd_vs, fd_cs),)
