"""Ray sharding across the GPUs of one box: one process per GPU, weights replicated, rays
partitioned in contiguous ranges, ONE NCCL all-reduce of the flat fp32 parameter gradients per step.

The reference has no distributed code; rays are independent, so the only exchange step training needs
is the sum of the MLP weight gradients (3.19 MB for the DTU model).  Random tensors are drawn for the
GLOBAL batch on every rank from the same seed and sliced, so a sharded step equals the unsharded one
ray for ray.
"""
import torch
import torch.distributed as dist

from .model.ray_sampler import RefRng


def shard_range(n, rank, world, align=1):
    """Contiguous [lo, hi) of `n` items for `rank`; the first ranks get the extra items.  `align` > 1 shards whole blocks
    of `align` items (rendering: convergence groups of 512 rays must not straddle ranks, or the sharded frame would group
    rays differently from the unsharded one)."""
    if align > 1:
        nb = (n + align - 1) // align
        lo_b, hi_b = shard_range(nb, rank, world)
        return min(lo_b * align, n), min(hi_b * align, n)
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_input(model_input, rank, world):
    """Slices the reference's model_input dict ('uv' (1,R,2) [+ 'rgb']) to this rank's rays."""
    R = model_input['uv'].shape[1]
    lo, hi = shard_range(R, rank, world)
    out = dict(model_input)
    out['uv'] = model_input['uv'][:, lo:hi].contiguous()
    for k in ('rgb', 'object_mask'):
        if k in out and out[k] is not None and out[k].dim() >= 2 and out[k].shape[1] == R:
            out[k] = out[k][:, lo:hi].contiguous()
    return out


class ShardedRng(RefRng):
    """Draws every per-ray random tensor for the global batch (same CPU generator state on all ranks) and
    keeps this rank's rows; shared draws (the extras permutation) are identical everywhere."""

    def __init__(self, device, n_global, lo, hi):
        super().__init__(device)
        self.n_global, self.lo, self.hi = n_global, lo, hi

    def rand(self, *shape):
        full = torch.rand((self.n_global,) + tuple(shape[1:]))
        return self._up(full[self.lo:self.hi].contiguous().pin_memory())

    def randint(self, high, shape):
        full = torch.randint(high, (self.n_global,) + tuple(shape[1:]))
        return self._up(full[self.lo:self.hi].contiguous().pin_memory())

    def uniform(self, shape, lo, hi):
        full = torch.empty((self.n_global,) + tuple(shape[1:])).uniform_(lo, hi)
        return self._up(full[self.lo:self.hi].contiguous().pin_memory())


class GradAllReducer(object):
    """Flat fp32 gradient buffer + one all-reduce(sum) per step; gradients come back as the global mean
    (each rank's loss is a mean over its own equally sized shard)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def nbytes(self):
        return self.flat.numel() * 4

    def allreduce_(self, world=None):
        world = world or dist.get_world_size(self.group)
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat


def all_converged(flag, group=None):
    """Optional batch-global convergence (ray_sampler.py:136) across ranks: max of the per-rank flags."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return flag
