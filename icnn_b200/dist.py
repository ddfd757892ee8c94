"""Sample-sharded multi-GPU inner loop: one process per GPU, no data-path collective except a
single all-gather of the solved y* at the end.  The training backward of the
unrolled GD loop (``gd_grad_sharded``) has one real exchange step: the parameter gradients are sums
over samples, so they take one bucketed all-reduce; per-sample adjoints stay sharded.

Every sample's bundle, dual solve and iterate are independent (the reference loops
``for u in range(bsize)``, lib/bundle_entropy.py:211), so rows are split into contiguous blocks,
theta is replicated, and the only exchange is the final gather.  ``torch.distributed`` (NCCL on
GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(B, rank, world_size):
    """Contiguous row block [lo, hi) of rank ``rank``; the first B % ws ranks get one extra row."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(B, world_size):
    return [shard_rows(B, r, world_size)[1] - shard_rows(B, r, world_size)[0] for r in range(world_size)]


def allgather_rows(y_local, B, group=None):
    """All-gather row blocks [b_r, n] -> [B, n] on every rank (ragged last blocks allowed).
    Every rank must call it, including ranks whose samples all finished early."""
    ws = dist.get_world_size(group)
    sizes = shard_sizes(B, ws)
    n = y_local.shape[1]
    if len(set(sizes)) == 1:
        out = torch.empty(B, n, dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(out, y_local.contiguous(), group=group)
        return out
    mx = max(sizes)
    pad = torch.zeros(mx, n, dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def solve_batch_sharded(net, x, y0, nIter=None, solver="pc", variant="lib", affine=False, group=None,
                        **kw):
    """Rank-local fused solveBatch on this rank's row block of (x, y0), then one all-gather of
    y*.  ``net`` is this rank's PICNN replica.  Returns (y_all [B, n] float64 CUDA tensor,
    local result tuple)."""
    from . import bundle_entropy
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    B = x.shape[0]
    lo, hi = shard_rows(B, rank, ws)
    y0b = y0[lo:hi]
    if isinstance(y0b, torch.Tensor):
        # solveBatch overwrites its initXs in place: work on a copy of this rank's block, pinned when the caller's
        # batch is (the H2D of y0 and the D2H of y* then run at the pinned-memory rate)
        blk = None
        if y0b.is_pinned():
            try:
                blk = torch.empty(tuple(y0b.shape), dtype=y0b.dtype, pin_memory=True).copy_(y0b)
            except Exception:      # no pinned memory to be had: a pageable copy is still correct
                blk = None
        y0b = blk if blk is not None else y0b.clone()
    else:
        y0b = np.array(y0b, dtype=np.float64)
    if hi > lo:
        fg = net.bind(x[lo:hi], affine=affine)
        out = bundle_entropy.solveBatch(fg, y0b, nIter=nIter, solver=solver, variant=variant,
                                        return_state=True, **kw)
    else:   # B < world_size: this rank owns no rows but still takes part in the gather below
        out = bundle_entropy.solveBatch(None, y0b, nIter=nIter, solver=solver, variant=variant,
                                        return_state=True, device=net.device, **kw)
    st = out[-1]
    y_all = allgather_rows(st.y, B, group=group)
    return y_all, out[:-1]


PARAM_KEYS = ("Wy", "Wz", "Wu", "bu", "Wzu", "bzu", "Wyu", "byu")


def allreduce_grads(grads, keys=PARAM_KEYS, group=None):
    """Sum the parameter gradients over ranks with ONE all-reduce: every tensor of ``grads[k]``
    (k in keys, None entries skipped) is packed into a single flat bucket (NVSwitch collectives are
    latency- not link-bound, so one launch beats one per tensor), reduced, and unpacked in place."""
    ts = [t for k in keys if k in grads for t in grads[k] if t is not None]
    if not ts:
        return grads
    flat = torch.cat([t.reshape(-1) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in ts:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return grads


def gd_grad_sharded(net, x, y0, trueY, nIter=30, lr=0.01, momentum=0.3, group=None):
    """d mse / d theta through the unrolled GD loop (icnn_b200.gd_grad) for a global minibatch split
    into contiguous row blocks: rank-local ``icnn_gd_backward`` with the GLOBAL loss weight
    2/(B n), one all-reduce of the parameter gradients, one all-gather of y_N.  Returns
    (yN_all [B, n] CUDA tensor, grads) -- grads['dcy'/'dcz'] cover this rank's rows only."""
    from . import gd_grad as _gd
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    B = x.shape[0]
    lo, hi = shard_rows(B, rank, ws)
    fg = net.bind(x[lo:hi])
    yN, gr = _gd.gd_grad(fg, y0[lo:hi], trueY[lo:hi], nIter=nIter, lr=lr, momentum=momentum,
                         loss_scale=2.0 / (B * net.n), x=x[lo:hi], return_device=True)
    allreduce_grads(gr, group=group)
    return allgather_rows(yN, B, group=group), gr
