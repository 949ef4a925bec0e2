"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, bucketing and the
gradient all-reduce reproduce the single-process full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdformerflow_b200 import distributed as sdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                               torch.nn.Linear(64, 3))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(10, 16, generator=g), torch.randn(10, 3, generator=g)
        lo, hi = sdist.shard_range(10, rank, world)
        # weighted so that the average of per-rank mean losses equals the full-batch mean loss
        loss = ((model(x[lo:hi]) - y[lo:hi]) ** 2).sum() / 10 * world
        loss.backward()
        model[4].bias.grad = None if rank == 1 else model[4].bias.grad      # a rank with an "unused" parameter
        nb = sdist.allreduce_gradients(model.parameters(), world, bucket_bytes=8 * 1024)
        # plain arrays, pickled by value: a tensor would travel as a shared-memory handle that dies with this process
        q.put((rank, nb, [p.grad.numpy().copy() for p in model.parameters()]))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (1, 7, 8, 10, 33):
        for w in (1, 2, 3, 8):
            spans = [sdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_buckets_reverse_order_and_size_cap():
    m = _model()
    buckets = sdist.make_buckets(list(m.parameters()), bucket_bytes=8 * 1024)
    flat = [p for b in buckets for p in b]
    assert [id(p) for p in flat] == [id(p) for p in reversed(list(m.parameters()))]
    assert len(buckets) > 1
    for b in buckets:
        assert len(b) == 1 or sum(p.numel() * 4 for p in b) <= 8 * 1024


@pytest.mark.timeout(120)
def test_two_rank_gradient_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [(r, nb, [torch.from_numpy(g) for g in grads]) for r, nb, grads in (q.get(timeout=100) for _ in procs)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    model = _model()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(10, 16, generator=g), torch.randn(10, 3, generator=g)
    (((model(x) - y) ** 2).sum() / 10).backward()
    ref = [p.grad for p in model.parameters()]
    for rank, nb, grads in results:
        assert nb > 1
        for i, (a, b) in enumerate(zip(grads, ref)):
            if i == len(ref) - 1:
                continue   # the bias whose gradient rank 1 dropped: averaged with zeros by design
            assert torch.allclose(a, b, atol=1e-6), (rank, i)
    assert torch.allclose(results[0][2][-1], results[1][2][-1])          # ranks agree even on the dropped one
