"""Data parallelism of the hot path: clouds are independent, so a batch shards over ranks by the batch
dimension and the ONLY exchange is the all-reduce of parameter gradients (reference: DistributedSampler +
DistributedDataParallel(broadcast_buffers=False), /root/reference/pytorch/function/train_modelnet_dist.py:117,206;
BatchNorm statistics are per GPU, not synchronised).  One process per GPU, NCCL over NVLink/NVSwitch
(gloo on CPU in the tests)."""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """contiguous [lo, hi) of `total` clouds owned by `rank` (earlier ranks take the remainder)"""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, average=True):
    """one flat all-reduce over every parameter gradient (a LocalAggregation has <= 42 KB of them)"""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    if average:
        flat /= dist.get_world_size()
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n
