"""Replacement for tfoptflow/multi_gpus.py in a one-process-per-GPU world.

The reference does single-process in-graph replication: per-tower gradients are stacked and
`tf.reduce_mean`-ed on a controller device (`average_gradients`, tfoptflow/multi_gpus.py:68-86),
variables live on the controller and are re-copied to every GPU each step (:6-7, `assign_to_device`
:35-55).  Here every rank owns one GPU and a full weight replica; the only exchange step of the
training configuration is ONE all-reduce(mean) over flattened gradient buckets on NCCL/NVLink
(gloo on CPU for tests).  The hot-path operator itself has no parameters and needs no collective:
batch elements are independent (tower slicing, tfoptflow/model_pwcnet.py:285-286)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def get_available_gpus():
    """tfoptflow/multi_gpus.py:58 - names of the visible GPUs."""
    return [f"/device:GPU:{i}" for i in range(torch.cuda.device_count())]


def init_from_env(backend: str | None = None):
    """Initialises torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_batch(batch_size: int, rank: int, world: int):
    """Contiguous batch slice of this rank, exactly the reference's tower slicing
    (`x_tnsr[n * batch_size:(n + 1) * batch_size]`, tfoptflow/model_pwcnet.py:285-286).
    Remainders go to the lowest ranks."""
    base, rem = divmod(batch_size, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def average_gradients(grads, group=None, bucket_bytes: int = 32 << 20):
    """All-reduce(mean) of a list of gradient tensors in place (None entries are skipped, like the
    reference's `if g is not None`, tfoptflow/multi_gpus.py:77).  Gradients are packed into flat fp32
    buckets so the collective count does not depend on the number of variables (the reference moves every
    variable separately: 4.7M / 14.1M parameters, README.md:55)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return grads
    world = dist.get_world_size(group)
    live = [g for g in grads if g is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1).to(torch.float32) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
        bucket, size = [], 0

    for g in live:
        bucket.append(g)
        size += g.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return grads


class GradientBucket:
    """Persistent flat fp32 bucket for the training configuration's one exchange step.

    The reference computes gradients per tower, and in mixed precision un-scales them by the loss scaler and clips
    them to global norm 5.0 PER TOWER, BEFORE averaging (tfoptflow/model_pwcnet.py:308-315); the towers' gradients are
    then `tf.reduce_mean`-ed per variable on the controller (`average_gradients`, tfoptflow/multi_gpus.py:68-86).
    Here: one multi-tensor pack into a bucket allocated once, the same per-rank unscale + clip on the flat view, ONE
    all-reduce(avg) over NCCL / NVLink (sum + divide on gloo, which has no AVG), one multi-tensor unpack."""

    def __init__(self, grads):
        live = [g for g in grads if g is not None]
        if not live:
            raise ValueError("no gradients")
        self.sizes = [g.numel() for g in live]
        self.flat = torch.empty(sum(self.sizes), dtype=torch.float32, device=live[0].device)
        self.views, off = [], 0
        for g, n in zip(live, self.sizes):
            self.views.append(self.flat[off:off + n].view(g.shape))
            off += n

    def average(self, grads, loss_scale: float | None = None, clip_norm: float | None = None, group=None):
        live = [g for g in grads if g is not None]  # `if g is not None`, tfoptflow/multi_gpus.py:77
        if [g.numel() for g in live] != self.sizes:
            raise ValueError("gradient list does not match the bucket")
        torch._foreach_copy_(self.views, live)
        if loss_scale is not None:  # grads / loss_scaler, model_pwcnet.py:313
            self.flat.mul_(1.0 / float(loss_scale))
        if clip_norm is not None:  # tf.clip_by_global_norm: t * clip / max(global_norm, clip), model_pwcnet.py:314
            norm = torch.linalg.vector_norm(self.flat)
            self.flat.mul_(float(clip_norm) / torch.clamp(norm, min=float(clip_norm)))
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        torch._foreach_copy_(live, self.views)
        return grads

