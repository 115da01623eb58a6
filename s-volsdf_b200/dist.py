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


class PeerGradBuffer(object):
    """The flat gradient buffer of `GradAllReducer` placed in SYMMETRIC memory (torch.distributed._symmetric_memory: every
    rank's allocation is mapped into every other rank's address space over NVLink) plus the flag words of the device-side
    barrier.  `FusedAdam(..., peer=PeerGradBuffer(params))` then replaces [copy-in, NCCL all-reduce, divide, copy-out, norm
    launch, Adam launch] by [copy-in, ONE kernel] (`svs_adam_step_allreduce`, csrc/optim.cu): the kernel reads the
    gradients of all ranks straight from peer memory.  world == 1: ordinary device memory, same kernel."""

    MAX_PEERS = 8

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.world > self.MAX_PEERS:
            raise ValueError('PeerGradBuffer supports up to %d ranks of one NVLink domain' % self.MAX_PEERS)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4        # every tensor starts on a 16-byte boundary (float4 loads over NVLink)
        self.n_flat = off
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            name = pg.group_name
            if hasattr(symm, 'enable_symm_mem_for_group'):
                try:
                    symm.enable_symm_mem_for_group(name)
                except Exception:
                    pass
            self.flat = symm.empty(self.n_flat, dtype=torch.float32, device=dev)
            self.flags = symm.empty(2 * self.MAX_PEERS, dtype=torch.int32, device=dev)
            self.flat.zero_()
            self.flags.zero_()
            self._h_flat = symm.rendezvous(self.flat, name)
            self._h_flags = symm.rendezvous(self.flags, name)
            self.peer_grads = [int(x) for x in self._h_flat.buffer_ptrs]
            self.peer_flags = [int(x) for x in self._h_flags.buffer_ptrs]
            torch.cuda.synchronize()
            dist.barrier(group)                     # nobody signals before every flag word is zero
        else:
            self.flat = torch.zeros(self.n_flat, dtype=torch.float32, device=dev)
            self.flags = None
            self.peer_grads, self.peer_flags = [self.flat.data_ptr()], None
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self.gmean = torch.zeros(self.n_flat, dtype=torch.float32, device=dev)
        self.mean_views = [self.gmean[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self.ctrl = torch.zeros(4 + 1024, dtype=torch.int32, device=dev)   # barrier words + per-block partial norms

    def nbytes(self):
        return self.n_flat * 4

    def load_grads(self):
        """this rank's gradients -> the symmetric buffer (one multi-tensor copy)"""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)


def all_converged(flag, group=None):
    """Optional batch-global convergence (ray_sampler.py:136) across ranks: max of the per-rank flags."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return flag
