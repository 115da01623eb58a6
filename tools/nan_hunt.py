"""Uninitialised-read detector: poison the caching allocator's free blocks with NaN, then run a train step of each
model on each engine and report which outputs / parameter gradients come back non-finite.
    timeout 120 python tools/nan_hunt.py      (GPU box; not part of the product or the tests)"""
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
import svolsdf_b200.scene as S  # noqa: E402


def poison(gb=6):
    xs = [torch.full((1 << 28,), float('nan'), device='cuda') for _ in range(gb)]
    torch.cuda.synchronize()
    del xs


for kind in ('bmvs', 'dtu'):
    for eng, ename in ((L.ENGINE_TC, 'tc'), (L.ENGINE_FP32, 'fp32')):
        R = 48
        model = build_model(kind, perturb=True, beta=0.05, device='cuda').train().set_engine(eng)
        inp = {k: v.cuda() for k, v in S.make_input(kind, R).items()}
        gt = S.gt_rgb(R).reshape(-1, 3).cuda()
        for trial in range(2):
            poison()
            torch.manual_seed(321)
            out = model(inp, fast=1)
            bad_out = [k for k, v in out.items() if torch.is_tensor(v) and v.is_floating_point() and not bool(torch.isfinite(v).all())]
            dep = out['depth_values_all'] if kind == 'bmvs' else out['depth_values']
            loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean() + \
                0.05 * out['weights'].pow(2).sum(1).mean() + 0.1 * dep.mean()
            model.zero_grad()
            poison()
            loss.backward()
            bad_g = [n for n, p in model.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all())]
            print('%s %s trial %d: loss %.5f  non-finite outputs %s  non-finite grads %s' % (
                kind, ename, trial, float(loss), bad_out, bad_g[:4] + (['... %d total' % len(bad_g)] if len(bad_g) > 4 else [])), flush=True)
