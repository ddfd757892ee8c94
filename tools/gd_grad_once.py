#!/usr/bin/env python
"""One icnn_gd_backward call on C3 dims (for an ncu launch list)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icnn_b200  # noqa: E402
from icnn_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
p, x, y0 = workloads.make_inputs(name)
tY = (np.random.RandomState(5).uniform(size=y0.shape) < 0.1).astype(np.float64)
fg = icnn_b200.PICNN.from_params(p).bind(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
yN, gr = icnn_b200.gd_grad.gd_grad(fg, y0, tY, nIter=30, lr=0.01, momentum=0.3, return_device=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(yN.mean()))
