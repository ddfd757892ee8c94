#!/usr/bin/env python
"""torchrun check of icnn_b200.dist.gd_grad_sharded: sample-sharded training backward + one NCCL
all-reduce must equal the single-GPU gradient of the whole minibatch (up to float32 summation order).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_gd_grad_check.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icnn_b200  # noqa: E402
from icnn_b200 import dist as idist, workloads  # noqa: E402


def main():
    rank, lrank = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl")
    ws = dist.get_world_size()
    B, nIter = 1024, 10
    p, x, y0 = workloads.make_inputs("C3", B=B)
    tY = (np.random.RandomState(5).uniform(size=y0.shape) < 0.1).astype(np.float64)
    net = icnn_b200.PICNN.from_params(p, device="cuda:%d" % lrank)
    yN, gr = idist.gd_grad_sharded(net, x, y0, tY, nIter=nIter)
    torch.cuda.synchronize()
    # timing of the sharded call (max over ranks)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        idist.gd_grad_sharded(net, x, y0, tY, nIter=nIter)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = None
    if rank == 0:
        yf, gf = icnn_b200.gd_grad.gd_grad(net.bind(x), y0, tY, nIter=nIter, x=x, return_device=True)
        errs = {}
        for k in idist.PARAM_KEYS:
            for i, (a, b) in enumerate(zip(gr[k], gf[k])):
                if a is not None:
                    errs["%s%d" % (k, i)] = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        dy = float((yN - yf).abs().max())
        out = {"world": ws, "B": B, "nIter": nIter, "ms_per_call_max_over_ranks": round(float(ms), 3),
               "yN_max_abs_diff": dy, "max_rel_err": max(errs.values()), "worst": max(errs, key=errs.get)}
        print(json.dumps(out))
        assert dy < 1e-5 and max(errs.values()) < 1e-4, out
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
