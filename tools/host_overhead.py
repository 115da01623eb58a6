"""How long does the HOST take to issue one train step (no synchronisation) vs the device time?  (diagnostic)"""
import os, sys, time, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import svolsdf_b200._lib as L
import svolsdf_b200.conf as C
import svolsdf_b200.scene as S
from svolsdf_b200.model.network import VolSDFNetwork

dev = 'cuda'
R = int(os.environ.get('R', '1024'))
torch.manual_seed(0)
model = VolSDFNetwork(C.dtu_model_conf()).to(dev).train().set_engine(L.ENGINE_TC)
opt = torch.optim.Adam(model.parameters(), lr=5e-4)
inp = {k: v.to(dev) for k, v in S.make_input('dtu', R).items()}
gt = S.gt_rgb(R).to(dev).reshape(-1, 3)


def step():
    out = model(inp, fast=1)
    loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('host issue %.3f ms/step, wall incl. drain %.3f ms/step' % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
if os.environ.get('PROFILE'):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
if os.environ.get('KPROF'):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        ct = getattr(e, 'device_time_total', None)
        if ct is None:
            ct = getattr(e, 'cuda_time_total', 0)
        if ct > 0 and e.device_type.name == 'CUDA':
            rows.append((ct / 5.0, e.count / 5.0, e.key[:90]))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print('total device us/step %.1f' % tot)
    for r in rows[:45]:
        print('%9.1f us  x%5.1f  %s' % r)
