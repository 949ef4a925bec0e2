"""Data-parallel plumbing: one process per GPU under torchrun, batch sharded by rank,
one exchange per step — the gradient all-reduce over NCCL / NVLink.  The reference itself only has
single-process nn.DataParallel (train_mdr_supervised_SNN.py:125-128); BatchNorm statistics stay per
replica exactly like there.

Two ways to reduce gradients:
  * wrap(model, local_rank): torch DDP — 25 MB buckets launched from backward hooks, overlapped with
    the backward kernels (what bench.py uses);
  * allreduce_gradients(params): explicit bucketed flat all-reduce after backward (no overlap), backend
    agnostic — used by the gloo CPU tests and usable under CUDA graphs.
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world ranks get one extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def make_buckets(params, bucket_bytes=25 * 1024 * 1024):
    """Greedy buckets in REVERSE parameter order (gradients become ready back to front)."""
    buckets, cur, size = [], [], 0
    for p in reversed([p for p in params if p.requires_grad]):
        nbytes = p.numel() * p.element_size()
        if cur and size + nbytes > bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(p)
        size += nbytes
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(params, world=None, bucket_bytes=25 * 1024 * 1024, average=True):
    """Flat bucketed all-reduce (sum, then / world) of .grad; parameters without a gradient contribute zeros
    (so ranks with unused parameters, e.g. PSN's dead attn_sn, stay in lock-step)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return 0
    n = 0
    for bucket in make_buckets(list(params), bucket_bytes):
        for p in bucket:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in bucket]
        flat = torch.cat([g.reshape(-1) for g in grads])          # one launch in, one multi-tensor launch out
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(world)
        torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])
        n += 1
    return n


def wrap(model, local_rank, find_unused_parameters=False):
    """DistributedDataParallel with NCCL over NVLink; identity for a single process."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    return torch.nn.parallel.DistributedDataParallel(
        model, device_ids=[local_rank], bucket_cap_mb=25, gradient_as_bucket_view=True,
        find_unused_parameters=find_unused_parameters)
