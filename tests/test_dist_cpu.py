"""Host-side logic of the sample-sharded multi-GPU path, on CPU with the gloo backend
(world_size 2 and 3): row partition, ragged all-gather, every rank reaches the collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icnn_b200 import dist as idist


def test_shard_rows_partition():
    for B in (1, 7, 64, 400, 65536):
        for ws in (1, 2, 3, 8):
            blocks = [idist.shard_rows(B, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = idist.shard_sizes(B, ws)
            assert sum(sizes) == B and max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, B, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        full = torch.arange(B * n, dtype=torch.float64).reshape(B, n)
        lo, hi = idist.shard_rows(B, rank, ws)
        # a rank whose samples "all finished early" still owns its rows and still calls the gather
        out = idist.allgather_rows(full[lo:hi].clone(), B)
        ok = bool(torch.equal(out, full))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,B", [(2, 8), (2, 7), (3, 10), (3, 2), (2, 1)])   # B < ws: a rank with NO rows still gathers
def test_allgather_rows_gloo(ws, B):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, B, 5, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(ws)) and all(ok for _, ok in res)


def _worker_grads(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        v = float(rank + 1)
        grads = {"Wy": [torch.full((3, 2), v), torch.full((3, 1), 2 * v)], "Wz": [None, torch.full((2, 1), -v)],
                 "bu": [torch.full((4,), 0.5 * v)], "dcy": [torch.full((5, 3), v)]}    # dcy: per-sample, not reduced
        idist.allreduce_grads(grads)
        tot = sum(range(1, ws + 1))
        ok = (torch.equal(grads["Wy"][0], torch.full((3, 2), float(tot))) and
              torch.equal(grads["Wy"][1], torch.full((3, 1), 2.0 * tot)) and
              torch.equal(grads["Wz"][1], torch.full((2, 1), -float(tot))) and
              torch.equal(grads["bu"][0], torch.full((4,), 0.5 * tot)) and
              torch.equal(grads["dcy"][0], torch.full((5, 3), v)) and grads["Wz"][0] is None)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2, 3])
def test_allreduce_grads_gloo(ws):
    """The one exchange step of the sharded training backward: a single bucketed SUM all-reduce of the
    parameter gradients; per-sample adjoints are left alone."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_grads, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(ws)) and all(ok for _, ok in res)


class _FakeNet:
    """Stands in for a PICNN replica: ``bind`` keeps the row block it was given."""
    device = torch.device("cpu")

    def bind(self, x, affine=False):
        return ("bound", np.asarray(x, dtype=np.float64), affine)


def _fake_solve_batch(fg, initXs, nIter=None, solver="pc", variant="lib", return_state=False, device=None, **kw):
    """The shape of bundle_entropy.solveBatch's result with a trivial 'solve': y = y0 + (global row id) + nIter."""
    import types
    y0 = initXs.numpy() if isinstance(initXs, torch.Tensor) else np.asarray(initXs)
    if fg is None:                       # a rank that owns no rows
        assert y0.shape[0] == 0 and device is not None
        y = np.zeros((0, y0.shape[1]))
    else:
        y = y0 + fg[1][:, :1] + float(nIter)
    st = types.SimpleNamespace(y=torch.from_numpy(np.ascontiguousarray(y)))
    return (y, [], [], [], [], [nIter] * y.shape[0], st)


def _worker_sharded(rank, ws, port, B, n, as_tensor, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from icnn_b200 import bundle_entropy
        bundle_entropy.solveBatch = _fake_solve_batch          # this process only
        x = np.arange(B, dtype=np.float64)[:, None] * np.ones((1, 3))     # x[i, :] = global row id
        y0 = np.random.RandomState(0).uniform(size=(B, n))
        y0_in = torch.from_numpy(y0.copy()) if as_tensor else y0.copy()
        y_all, local = idist.solve_batch_sharded(_FakeNet(), x, y0_in, nIter=4)
        lo, hi = idist.shard_rows(B, rank, ws)
        ok = (tuple(y_all.shape) == (B, n)
              and np.allclose(y_all.numpy(), y0 + np.arange(B)[:, None] + 4.0, rtol=0, atol=0)
              and len(local) == 6 and local[0].shape[0] == hi - lo
              and np.array_equal(np.asarray(y0_in), y0))                 # the caller's y0 is not touched (block copies)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,B,as_tensor", [(2, 5, False), (2, 5, True), (3, 2, False), (3, 1, True), (2, 8, False)])
def test_solve_batch_sharded_orchestration_gloo(ws, B, as_tensor):
    """icnn_b200.dist.solve_batch_sharded on CPU/gloo with the device solve replaced by a stub: contiguous row
    blocks reach the right rank, ragged and EMPTY shards (B < world size) still take part in the all-gather, numpy
    and torch y0 are both accepted, the gathered y* is in global row order on every rank."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sharded, args=(r, ws, port, B, 4, as_tensor, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(ws)) and all(ok for _, ok in res), res
