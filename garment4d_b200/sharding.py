"""Multi-GPU plumbing for the hot path: one process per GPU, sequences (B) sharded across ranks, the T frames of a
sequence stay on one rank (the reference's temporal ops need the whole sequence, modules/mesh_encoder.py:161,467-476).
The forward path has no data-path collective (every op is per cloud); training adds ONE flat gradient all-reduce."""
import torch
import torch.distributed as dist


def shard_sequences(num_sequences: int, rank: int, world: int):
    """Indices of the sequences owned by `rank`: the partition of the reference's non-shuffling DistributedSampler
    (utils/train_utils.py:12-31, and torch's sampler used at train_temporal.py:86 with shuffle off): the index list is
    padded by wrapping around to a multiple of `world`, then dealt out strided -- ``indices[rank:total:world]`` -- so
    every rank owns exactly ceil(num_sequences / world) sequences and all ranks run the same number of steps."""
    if num_sequences <= 0:
        return []
    per_rank = (num_sequences + world - 1) // world
    total = per_rank * world
    indices = list(range(num_sequences))
    while len(indices) < total:                       # wrap-around padding (train_utils.py:24-26)
        indices += indices[: total - len(indices)]
    return indices[rank:total:world]


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Device/step time as the max over ranks (never wall clock on one rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class FlatGradientReducer:
    """ONE all-reduce (sum, then divide by the group size) over a flat fp32 buffer holding every gradient -- the
    B200-native shape for the 1.3-5.5 MB payload of this model (latency-bound on NVLink 5) instead of DDP's bucketed
    graph walk with find_unused_parameters (train_temporal.py:186-187).

    The buffer covers EVERY parameter that requires grad, in a fixed order, whether or not it received a gradient on
    this rank in this step (missing gradients contribute zeros), so all ranks always reduce the same number of elements.
    Parameters' ``.grad`` are views into the buffer: backward accumulates straight into it and the optimizer reads the
    reduced values from it -- no gather/scatter copies around the collective."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """Re-attach the views (an optimizer or zero_grad(set_to_none=True) may have dropped them); grads that were
        replaced by fresh tensors are copied back into the buffer."""
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            off += p.numel()

    def all_reduce(self, async_op=False):
        """Sum over the group, then average.  Returns the work handle when async_op (call .wait() before reading grads;
        the division is then the caller's: ``reducer.flat.div_(reducer.world)``)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return None
        self.rebind()
        if async_op:
            return dist.all_reduce(self.flat, group=self.group, async_op=True)
        dist.all_reduce(self.flat, group=self.group)
        self.flat.div_(dist.get_world_size(self.group))
        return None

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_flat_gradients(params, group=None):
    """One-shot form: averages the gradients of `params` over the group with a single all-reduce.  Every parameter that
    requires grad takes part (a missing gradient counts as zeros and is materialised), so buffer sizes match on all ranks."""
    params = [p for p in params if p.requires_grad]
    if not params or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in params])
    dist.all_reduce(flat, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
