"""RL Adam argmin on the device.

Mirror of ``Agent.adam(func, obs)`` (RL/src/icnn.py:160-215) for func = ``_fg_entr``
(negQ - entropy(act), RL/src/icnn.py:60-63): returns the best action found per sample, in [-1, 1].
The RL agent's default optimiser (``--icnn_opt adam``, RL/src/agent.py:25-26)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi
from .picnn import BoundPICNN


def solve(fg: BoundPICNN, max_iter=1000, return_iters=False):
    if not isinstance(fg, BoundPICNN):
        raise TypeError("adam.solve needs a BoundPICNN (PICNN.bind(obs))")
    if fg.affine:
        raise ValueError("bind the observations without affine=True: Adam works on the action in [-1, 1]")
    net = fg.net
    dev = net.device
    B, n = fg.B, net.n
    with torch.cuda.device(dev):
        act_best = torch.empty(B, n, dtype=torch.float64, device=dev)
        f_best = torch.empty(B, dtype=torch.float64, device=dev)
        scratch = torch.empty(_capi.lib.icnn_adam_workspace_bytes(B, n), dtype=torch.uint8, device=dev)
        its = C.c_int32(0)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _capi.check(_capi.lib.icnn_adam_solve(net._h, C.byref(fg.c_gates), act_best.data_ptr(), f_best.data_ptr(),
                                              int(max_iter), C.byref(its), scratch.data_ptr(), fg.ws.data_ptr(), stream))
        out = act_best.cpu().numpy()
    return (out, int(its.value)) if return_iters else out
