#!/usr/bin/env python
"""Timing of icnn_gd_backward (C3 dims, the multi-label script's defaults) next to (i) the forward-only
GD loop and (ii) torch autograd double-backprop of the same unrolled graph on the same GPU (cuBLAS,
float32) -- the library baseline for this row.  Prints one JSON line."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icnn_b200  # noqa: E402
from icnn_b200 import workloads  # noqa: E402


def torch_arm(p, x, y0, tY, nIter, lr, mom, dev):
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=dev)  # noqa: E731
    names = ("Wy", "Wz", "Wu", "bu", "Wzu", "bzu", "Wyu", "byu", "Wzx", "bzx")
    T = {k: [None if a is None else t(a).requires_grad_() for a in getattr(p, k)] for k in names}
    x, tY = t(x), t(tY)
    L = p.L

    def step():
        us, prev = [], x
        for i in range(L):
            u = prev @ T["Wu"][i] + T["bu"][i]
            if i < L - 1:
                u = torch.relu(u)
            us.append(u); prev = u
        cz = [None] + [torch.relu((x if i == 0 else us[i - 1]) @ T["Wzu"][i] + T["bzu"][i]) for i in range(1, L + 1)]
        cy = [(x if i == 0 else us[i - 1]) @ T["Wyu"][i] + T["byu"][i] for i in range(L + 1)]
        d = [(x if i == 0 else us[i - 1]) @ T["Wzx"][i] + T["bzx"][i] for i in range(L + 1)]

        def energy(y):
            z = None
            for i in range(L + 1):
                pre = (y * cy[i]) @ T["Wy"][i] + d[i]
                if i > 0:
                    pre = pre + (z * cz[i]) @ T["Wz"][i]
                z = torch.relu(pre) if i < L else pre
            return z.reshape(-1)

        yi = t(y0).requires_grad_()
        vi = torch.zeros_like(yi)
        for _ in range(nIter):
            (gi,) = torch.autograd.grad(energy(yi).sum(), yi, create_graph=True)
            vn = mom * vi - lr * gi
            yi = yi - mom * vi + (1.0 + mom) * vn
            vi = vn
        loss = ((yi - tY) ** 2).mean()
        params = [w for k in names for w in T[k] if w is not None]
        return torch.autograd.grad(loss, params, allow_unused=True)

    return step


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else workloads.CONFIGS[name]["B"]
    nIter, lr, mom = 30, 0.01, 0.3
    p, x, y0 = workloads.make_inputs(name, B=B)
    tY = (np.random.RandomState(5).uniform(size=y0.shape) < 0.1).astype(np.float64)
    dev = torch.device("cuda")
    net = icnn_b200.PICNN.from_params(p)
    fg = net.bind(x)
    xd = torch.tensor(x, dtype=torch.float32, device=dev)
    y0d = torch.tensor(y0, dtype=torch.float32, device=dev)
    tYd = torch.tensor(tY, dtype=torch.float32, device=dev)
    ms_fwd = timeit(lambda: icnn_b200.gd.solve(fg, y0d, nIter=nIter, lr=lr, momentum=mom, return_device=True))
    ms_bwd = timeit(lambda: icnn_b200.gd_grad.gd_grad(fg, y0d, tYd, nIter=nIter, lr=lr, momentum=mom,
                                                      return_device=True))
    ms_full = timeit(lambda: icnn_b200.gd_grad.gd_grad(fg, y0d, tYd, nIter=nIter, lr=lr, momentum=mom, x=xd,
                                                       return_device=True))
    ms_torch = timeit(torch_arm(p, x, y0, tY, nIter, lr, mom, dev), reps=3, warm=1)
    print(json.dumps({"workload": name, "B": B, "nIter": nIter, "gd_solve_ms": round(ms_fwd, 3),
                      "gd_backward_ms": round(ms_bwd, 3), "gd_backward_with_xpath_ms": round(ms_full, 3),
                      "torch_autograd_f32_ms": round(ms_torch, 3)}))


if __name__ == "__main__":
    main()
