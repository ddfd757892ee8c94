#!/usr/bin/env python
"""torchrun check of the public multi-GPU entry icnn_b200.dist.solve_batch_sharded:
a RAGGED global batch (B not a multiple of the world size, and a second case with B < world size so that one
rank owns no rows) is sharded by contiguous row blocks, solved rank-locally and all-gathered; the gathered y*
must equal the unsharded single-GPU solve to the float32-summation-order tolerance of the fused loop (K1's
split-K factor follows the local batch size; K2 is bit-identical given the same rows).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_solve_check.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icnn_b200  # noqa: E402
from icnn_b200 import bundle_entropy, dist as idist, workloads  # noqa: E402


def main():
    rank, lrank = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    ws = dist.get_world_size()
    out = {"world": ws, "cases": []}
    for name, B, nIter in (("C3", 8 * ws + 5, 10), ("T", 64 * ws + 3, 10), ("C4", 1001, 5), ("C1", max(1, ws - 1), 5)):
        cfg = workloads.CONFIGS[name]
        p, x, y0 = workloads.make_inputs(name, B=B)
        net = icnn_b200.PICNN.from_params(p, device="cuda:%d" % lrank)
        y_all, loc = idist.solve_batch_sharded(net, x, y0, nIter=nIter, variant=cfg["variant"], affine=cfg["affine"])
        torch.cuda.synchronize()
        assert tuple(y_all.shape) == (B, cfg["n"])
        # every rank holds the same gathered result
        ref = y_all.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, y_all)
        if rank == 0:
            full = bundle_entropy.solveBatch(net.bind(x, affine=cfg["affine"]), y0.copy(), nIter=nIter,
                                             variant=cfg["variant"])
            d = np.abs(y_all.cpu().numpy() - full[0]).max(axis=1)
            rec = {"workload": name, "B": B, "rows_per_rank": idist.shard_sizes(B, ws), "nIter": nIter,
                   "max": float(d.max()), "median": float(np.median(d)), "frac_gt_1e-4": float(np.mean(d > 1e-4))}
            out["cases"].append(rec)
            # float32 summation order only: the same statement tests/test_gpu_bundle.py::test_shard_concat_equals_unsharded makes
            assert np.median(d) < 1e-5 and np.mean(d < 1e-4) >= 0.9, rec
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
