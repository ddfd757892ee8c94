#!/usr/bin/env python
"""Experiment: split the minibatch into G independent row groups, one CUDA stream + one fused
solveBatch each (samples are independent, lib/bundle_entropy.py:211), so that a group whose samples
need few interior-point iterations does not wait at every outer iteration for the slowest sample of
the whole batch.  Prints ms per whole-batch solve for each G."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icnn_b200  # noqa: E402
from icnn_b200 import _capi, bundle_entropy, workloads  # noqa: E402
from icnn_b200.dist import shard_rows  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    Gs = [int(g) for g in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 16]
    cfg = workloads.CONFIGS[name]
    p, x, y0 = workloads.make_inputs(name)
    B, n, nIter = cfg["B"], cfg["n"], cfg["nIter"]
    dev = torch.device("cuda")
    net = icnn_b200.PICNN.from_params(p, device=dev)
    xd = torch.tensor(x, dtype=torch.float32, device=dev)
    y0d = torch.tensor(y0, device=dev)
    KS = min(nIter, n) + 1
    ccfg = bundle_entropy._make_cfg(cfg["variant"], "pc", nIter, None, None, 0, n, KS)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out = {}
    yref = None
    for G in Gs:
        groups = []
        for g in range(G):
            lo, hi = shard_rows(B, g, G)
            fg = net.bind(xd[lo:hi], affine=cfg["affine"])
            st = bundle_entropy.BundleState(hi - lo, n, KS, dev, keep_xs=True, nIter=nIter)
            groups.append((lo, hi, fg, st, torch.cuda.Stream()))
        torch.cuda.synchronize()

        def step():
            cur = torch.cuda.current_stream()
            for lo, hi, fg, st, s in groups:
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    st.y.copy_(y0d[lo:hi])
                    _capi.check(_capi.lib.icnn_solve_batch_fused(net._h, C.byref(fg.c_gates), C.byref(ccfg), C.byref(st.c),
                                                                 fg.ws.data_ptr(), C.c_void_p(s.cuda_stream)))
            for *_, s in groups:
                cur.wait_stream(s)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        y = torch.cat([st.y for _, _, _, st, _ in groups]).cpu().numpy()
        if yref is None:
            yref = y
        d = np.abs(y - yref).max(axis=1)
        out[G] = {"ms": round(float(np.mean(ts)), 3), "min_ms": round(float(np.min(ts)), 3),
                  "frac_rows_within_1e-4_of_G1": float(np.mean(d < 1e-4))}
        del groups
    print(json.dumps({"workload": name, "B": B, "groups": out}))


if __name__ == "__main__":
    main()
