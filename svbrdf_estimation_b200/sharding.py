"""Batch sharding of the rendering loss across the GPUs of one box (one process per GPU).

The path shards by independent units: every batch element has its own scene records and pixels
(losses.py:34-44) and the only cross-sample operation is the final mean (losses.py:50).  Each rank
therefore evaluates the fused kernel on its slice with no data-path collective; what crosses
NVLink is (a) the scalar loss for logging and (b) the CNN gradient all-reduce that
``DistributedDataParallel`` already does.  With equal slices the mean of the ranks' local losses is
the global loss and DDP's gradient averaging yields exactly the gradient of that global mean.
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world_size):
    """Contiguous [start, stop) slice of the batch owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    base, extra = divmod(int(global_batch), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def shard_seed(seed, rank):
    """Per-rank seed for the scene sampler, so ranks draw different light/view configurations.  Seed and rank go
    through a 64-bit mixer (the splitmix64 finaliser the library's stateless sampler uses): (s, r+1) and (s+1, r)
    do not collide, as plain ``seed + rank`` would.  63 bits, usable with ``torch.manual_seed``."""
    return _mix64(_mix64(int(seed) & 0xFFFFFFFFFFFFFFFF) ^ ((int(rank) + 1) * 0xD1342543DE82EF95 & 0xFFFFFFFFFFFFFFFF)) >> 1


def global_mean_loss(local_loss, local_batch, group=None):
    """All-reduce of the scalar loss: sum_r(local_loss_r * B_r) / sum_r(B_r) (every element of the
    loss tensor has the same weight because N, H and W are equal on all ranks).  Detached - logging
    only; gradients come from the local loss."""
    buf = torch.stack((local_loss.detach().to(torch.float64) * float(local_batch),
                       torch.tensor(float(local_batch), dtype=torch.float64, device=local_loss.device)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[0] / buf[1]).to(local_loss.dtype)


def local_grad_to_global(local_grad, local_batch, global_batch):
    """d(global mean loss)/d(local input) from d(local mean loss)/d(local input)."""
    return local_grad * (float(local_batch) / float(global_batch))


__all__ = ["shard_range", "shard_seed", "global_mean_loss", "local_grad_to_global"]
