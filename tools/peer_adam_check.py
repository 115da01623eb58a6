"""svs_adam_step_allreduce (all-reduce over NVLink peer memory + clip + guard + Adam in one kernel) against NCCL
all-reduce + FusedAdam.  Run with torchrun on >= 2 GPUs (also works on 1):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_adam_check.py"""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svolsdf_b200.optim import FusedAdam
from svolsdf_b200.dist import PeerGradBuffer

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)


def params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(256, 39), (256,), (256, 1), (217, 256), (257, 256), (3, 256), (), (5001,), (1,)] + [(256, 256)] * 8
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]


a, b = params(0), params(0)
ref = FusedAdam(a, lr=5e-4, max_grad_norm=1.0)
peer = PeerGradBuffer(b)
fus = FusedAdam(b, lr=5e-4, max_grad_norm=1.0, peer=peer)
g = torch.Generator().manual_seed(100 + rank)
for it in range(5):
    scale = 10.0 if it % 2 == 0 else 1e-3
    for pa, pb in zip(a, b):
        gr = (torch.randn(pa.shape, generator=g) * scale).to(dev)
        if it == 3 and rank == world - 1 and pa.dim() == 2 and pa.shape[0] == 217:
            gr[5, 7] = float('nan')          # the guard must fire on every rank
        pa.grad, pb.grad = gr.clone(), gr.clone()
    if world > 1:
        for pa in a:
            dist.all_reduce(pa.grad)
            pa.grad.div_(world)
    ref.step()
    fus.step()
    torch.cuda.synchronize()
    worst = max(float((pa - pb).abs().max()) for pa, pb in zip(a, b))
    fin = all(bool(torch.isfinite(pb).all()) for pb in b)
    if rank == 0:
        print('step %d: max |param diff| vs NCCL + FusedAdam %.3e, finite %s, norm^2 %.4e vs %.4e' % (
            it, worst, fin, float(fus.last_grad_norm_sq), float(ref.last_grad_norm_sq)), flush=True)
    assert worst < 1e-6 and fin
# replicas bit-identical
if world > 1:
    flat = torch.cat([p.detach().reshape(-1) for p in b])
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    same = all(torch.equal(others[0], o) for o in others)
    if rank == 0:
        print('replicas bit-identical:', same, flush=True)
    assert same
# latency
for fn, name in ((lambda: fus.step(), 'peer kernel (copy-in + 1 launch)'),):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print('%s: %.1f us per step' % (name, e0.elapsed_time(e1) / 50 * 1e3), flush=True)


def nccl_step():
    grads = [p.grad for p in a]
    flatg = torch.cat([x.reshape(-1) for x in grads])
    if world > 1:
        dist.all_reduce(flatg)
        flatg.div_(world)
    ref.step()


for _ in range(5):
    nccl_step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    nccl_step()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print('cat + NCCL all-reduce + div + FusedAdam (2 launches): %.1f us per step' % (e0.elapsed_time(e1) / 50 * 1e3), flush=True)
if world > 1:
    dist.destroy_process_group()
