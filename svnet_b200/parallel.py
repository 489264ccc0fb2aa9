"""Multi-GPU inference: shard by point-cloud batch, one process per GPU, one all-gather of the
outputs per batch.  Replaces the reference's nn.DataParallel
(main_cls_dgcnn.py:125): in eval mode every cloud's output depends only on that cloud (BatchNorm uses
running stats, the gate mean is per cloud, kNN is per cloud), so no collective is needed inside the
forward.  torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world, rank):
    """Contiguous split of ``batch`` clouds over ``world`` ranks; earlier ranks get the remainder."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, world=None, rank=None):
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(x.shape[0], world, rank)
    return x[lo:hi]


def gather_outputs(y_local, batch, world=None):
    """All-gather per-rank outputs (B_r, ...) back into (batch, ...) on every rank.  Uneven shards are
    padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return y_local
    per = (batch + world - 1) // world
    tail = tuple(y_local.shape[1:])
    buf = y_local
    if y_local.shape[0] != per:
        buf = torch.zeros((per,) + tail, dtype=y_local.dtype, device=y_local.device)
        buf[:y_local.shape[0]].copy_(y_local)
    out = torch.empty((world * per,) + tail, dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(out, buf.contiguous())
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(batch, world, r)
        parts.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(parts, dim=0)


class ShardedInference:
    """model(x) over a batch sharded across the ranks of the default process group.

    Every rank takes the same decisions from ``batch`` alone, so no rank can be left alone in the
    collective: when the batch is smaller than the world (the tail batch of an eval epoch: ModelNet40's
    2468 test clouds at batch 32 end with 4 clouds), ranks without a cloud run the model on cloud 0 to learn
    the output shape and contribute zero rows to the all-gather."""

    def __init__(self, model):
        self.model = model

    @torch.no_grad()
    def __call__(self, x, *extra):
        batch = x.shape[0]
        if batch == 0:
            raise ValueError("ShardedInference: empty batch")          # identical on every rank
        xs = shard_batch(x).contiguous()
        es = [shard_batch(e).contiguous() for e in extra]
        if xs.shape[0] > 0:
            y = self.model(xs, *es)
        else:
            y = self.model(x[:1].contiguous(), *[e[:1].contiguous() for e in extra])[:0]
        return gather_outputs(y, batch)
