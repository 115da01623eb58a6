"""Host-side multi-GPU logic on CPU: ray sharding, RNG slicing and the flat gradient all-reduce (gloo, 2 ranks)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svolsdf_b200 import dist as sdist
import svolsdf_b200.scene as S


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 1024, 1025, 65536):
        for w in (1, 2, 3, 8):
            spans = [sdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_input_slices_rays():
    inp = S.make_input('dtu', 10)
    inp['rgb'] = S.gt_rgb(10)
    parts = [sdist.shard_input(inp, r, 3) for r in range(3)]
    assert torch.equal(torch.cat([p['uv'] for p in parts], 1), inp['uv'])
    assert torch.equal(torch.cat([p['rgb'] for p in parts], 1), inp['rgb'])
    assert parts[0]['pose'] is inp['pose']


class _CpuShardedRng(sdist.ShardedRng):
    def _up(self, t):   # no CUDA in this container: keep the slice on the host
        self.h2d_bytes += t.numel() * t.element_size()
        return t

    def rand(self, *shape):
        return self._up(torch.rand((self.n_global,) + tuple(shape[1:]))[self.lo:self.hi].contiguous())

    def randint(self, high, shape):
        return self._up(torch.randint(high, (self.n_global,) + tuple(shape[1:]))[self.lo:self.hi].contiguous())


def test_sharded_rng_equals_global_draws():
    """Concatenating every rank's slices reproduces the unsharded draws in the reference's order."""
    R = 11
    torch.manual_seed(5)
    g_rand, g_u, g_perm, g_idx = torch.rand(R, 128), torch.rand(R, 64), torch.randperm(128), torch.randint(98, (R,))
    got = []
    for r in range(3):
        lo, hi = sdist.shard_range(R, r, 3)
        torch.manual_seed(5)
        rng = _CpuShardedRng('cpu', R, lo, hi)
        a, b = rng.rand(hi - lo, 128), rng.rand(hi - lo, 64)
        p = torch.randperm(128)
        c = rng.randint(98, (hi - lo,))
        got.append((a, b, p, c))
    assert torch.equal(torch.cat([g[0] for g in got]), g_rand)
    assert torch.equal(torch.cat([g[1] for g in got]), g_u)
    assert all(torch.equal(g[2], g_perm) for g in got)
    assert torch.equal(torch.cat([g[3] for g in got]), g_idx)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
    x = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10.0
    lo, hi = sdist.shard_range(8, rank, world)
    red = sdist.GradAllReducer(lin.parameters())
    lin(x[lo:hi]).pow(2).mean().backward()       # local mean over an equal shard
    red.allreduce_(world)
    flag = torch.tensor([rank], dtype=torch.int32)
    sdist.all_converged(flag)
    if rank == 0:
        torch.save({'grads': [p.grad.clone() for p in lin.parameters()], 'flag': flag, 'nbytes': red.nbytes()}, out)
    dist.destroy_process_group()


def test_grad_allreduce_matches_unsharded(tmp_path):
    out = str(tmp_path / 'g.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
    x = torch.arange(40, dtype=torch.float32).reshape(8, 5) / 10.0
    lin(x).pow(2).mean().backward()
    for a, p in zip(got['grads'], lin.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6)
    assert int(got['flag']) == 1 and got['nbytes'] == sum(p.numel() for p in lin.parameters()) * 4
