"""Multi-GPU plumbing for the hot path: one process per GPU, sequences (B) sharded across ranks, T frames of a
sequence stay on one rank (the reference's temporal ops need the whole sequence, modules/mesh_encoder.py:161,467-476).
The forward path has no data-path collective (every op is per-cloud); training adds one flat gradient all-reduce.
Mirrors the partition of torch's DistributedSampler used at train_temporal.py:86 / utils/train_utils.py:12-31
(contiguous, non-shuffled shards as in the reference's evaluation sampler)."""
import torch
import torch.distributed as dist


def shard_sequences(num_sequences: int, rank: int, world: int):
    """Contiguous [lo, hi) range of sequence indices owned by `rank`; the first (num % world) ranks get one extra."""
    base, extra = divmod(num_sequences, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Device/step time as the max over ranks (never wall clock on one rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_flat_gradients(params, group=None):
    """ONE all-reduce (sum, then divide) over a flat fp32 buffer of every gradient -- the B200-native shape for the
    1.3-5.5 MB payload of this model (latency-bound on NVLink 5), instead of DDP's bucketed graph walk
    (train_temporal.py:186-187)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(dist.get_world_size())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
