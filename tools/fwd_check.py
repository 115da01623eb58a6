"""SDF forward chain of the tcgen05 engine: correctness against the fp32 engine + device timing.
Run on the GPU box (always under `timeout`):  timeout 120 python tools/fwd_check.py [tag]
The library is taken from SVS_LIB_PATH when set (A/B runs of build variants).  Not part of the product or tests."""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model  # noqa: E402
import svolsdf_b200._lib as L  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'lib'
print('== %s : %s' % (tag, L.LIB_PATH), flush=True)


def pts(P, seed=0, radius=3.6, d=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(P, d, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * (torch.rand(P, 1, generator=g) * radius)
    return x.cuda()


for kind in ('dtu', 'bmvs'):
    a = build_model(kind, perturb=True, beta=0.05, device='cuda')
    b = build_model(kind, perturb=True, beta=0.05, device='cuda').set_engine(L.ENGINE_TC)
    for P in (1, 128, 129, 1000, 20000, 148 * 128 * 2 + 77):
        x = pts(P, seed=P)
        with torch.no_grad():
            s1 = b.implicit_network.get_sdf_vals(x)
            torch.cuda.synchronize()
            s0 = a.implicit_network.get_sdf_vals(x)
            y1 = b.implicit_network(x)
            torch.cuda.synchronize()
            y0 = a.implicit_network(x)
        print('%s P %6d  sdf max|d| %.3e   y max|d| %.3e  (|y| max %.2f)' % (
            kind, P, float((s1 - s0).abs().max()), float((y1 - y0).abs().max()), float(y0.abs().max())), flush=True)
    if kind == 'bmvs' and hasattr(b, 'bg_implicit_network'):
        x4 = pts(5000, seed=3, radius=1.0, d=4)
        with torch.no_grad():
            y1 = b.bg_implicit_network(x4)
            torch.cuda.synchronize()
            y0 = a.bg_implicit_network(x4)
        print('bmvs bg net (d_in 4, 10 freqs)  y max|d| %.3e (|y| max %.2f)' % (float((y1 - y0).abs().max()), float(y0.abs().max())), flush=True)

# timing: 131072 points (1024 rays x 128 samples), no-grad sdf
b = build_model('dtu', perturb=True, beta=0.05, device='cuda').set_engine(L.ENGINE_TC)
for P in (131072, 1 << 20):
    x = pts(P, seed=1)
    with torch.no_grad():
        for _ in range(3):
            b.implicit_network.get_sdf_vals(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            b.implicit_network.get_sdf_vals(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    fl = 1049088.0 * P
    L.prof_enable(True)
    with torch.no_grad():
        for _ in range(5):
            b.implicit_network.get_sdf_vals(x)
    torch.cuda.synchronize()
    pr = L.prof_collect()
    L.prof_enable(False)
    for k, v in pr.items():
        print('   kernel %-22s %.4f ms/launch  %.1f TFLOP/s' % (k, v['ms'] / v['launches'], v['flops'] / max(v['ms'], 1e-9) / 1e9), flush=True)
    print('TIME %s sdf fwd P=%d : %.4f ms  = %.1f TFLOP/s (incl. weight pack + launch overheads)' % (tag, P, ms, fl / ms / 1e9), flush=True)
