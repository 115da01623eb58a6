"""Bring-up check of the tcgen05 (bf16) engine against the fp32 SIMT engine, kernel family by kernel family.
Run on the GPU box:  timeout 200 python tools/tc_check.py [stage ...]   (not part of the product or the tests)."""
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')

from helpers import build_model  # noqa: E402
import svolsdf_b200._lib as L  # noqa: E402
from svolsdf_b200 import functional as F  # noqa: E402

DEV = 'cuda'


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def mx(a, b):
    return float((a.double() - b.double()).abs().max())


def points(P, seed=0, radius=3.6):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(P, 3, generator=g)
    x = x / x.norm(dim=1, keepdim=True) * (torch.rand(P, 1, generator=g) * radius)
    return x.to(DEV)


def models(kind='dtu'):
    a = build_model(kind, perturb=True, beta=0.05, device=DEV)
    b = build_model(kind, perturb=True, beta=0.05, device=DEV)
    b.set_engine(L.ENGINE_TC)
    return a, b


def stage_fwd(P=1000):
    a, b = models()
    x = points(P)
    with torch.no_grad():
        s0 = a.implicit_network.get_sdf_vals(x)
        s1 = b.implicit_network.get_sdf_vals(x)
        torch.cuda.synchronize()
        print('get_sdf_vals  max|d| %.3e  (|sdf| max %.2f)' % (mx(s1, s0), float(s0.abs().max())), flush=True)
        y0 = a.implicit_network(x)
        y1 = b.implicit_network(x)
        torch.cuda.synchronize()
        print('forward y     max|d| %.3e  rel %.3e   sdf col max|d| %.3e' % (mx(y1, y0), rel(y1, y0), mx(y1[:, 0], y0[:, 0])), flush=True)


def stage_outputs(P=1000):
    a, b = models()
    x = points(P)
    with torch.no_grad():
        sa, fa, ga = a.implicit_network.get_outputs(x)
        sb, fb, gb = b.implicit_network.get_outputs(x)
        torch.cuda.synchronize()
    print('get_outputs   sdf max|d| %.3e  feat rel %.3e  grad rel %.3e max|d| %.3e' % (mx(sb, sa), rel(fb, fa), rel(gb, ga), mx(gb, ga)), flush=True)
    with torch.no_grad():
        g0 = a.implicit_network.gradient(x)
        g1 = b.implicit_network.gradient(x)
    print('gradient      rel %.3e' % rel(g1, g0), flush=True)


def _grads(m):
    return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}


def stage_backward(P=int(os.environ.get('TC_P', '1000'))):
    a, b = models()
    x = points(P)
    g = torch.Generator().manual_seed(3)
    wy = torch.randn(P, 256, generator=g).to(DEV) * 0.01
    ws = torch.randn(P, 1, generator=g).to(DEV)
    wg = torch.randn(P, 3, generator=g).to(DEV)
    res = []
    for m in (a, b):
        m.zero_grad()
        sdf, feat, grad = m.implicit_network.get_outputs(x)
        loss = (feat * wy).sum() + (sdf * ws).sum() + (grad * wg).sum()
        loss.backward()
        torch.cuda.synchronize()
        res.append(_grads(m))
    worst = 0
    for n in res[0]:
        if n not in res[1]:
            print('  missing grad', n)
            continue
        r = rel(res[1][n], res[0][n])
        worst = max(worst, r)
        if r > 1e-2:
            print('  %-40s rel %.3e  |g| %.3e' % (n, r, float(res[0][n].norm())), flush=True)
    print('sdf backward (top+tangent) worst rel %.3e' % worst, flush=True)
    # eikonal-style: gradient() only
    res = []
    for m in (a, b):
        m.zero_grad()
        grad = m.implicit_network.gradient(x)
        ((grad.norm(dim=1) - 1) ** 2).mean().backward()
        torch.cuda.synchronize()
        res.append(_grads(m))
    worst = max(rel(res[1][n], res[0][n]) for n in res[0])
    print('eikonal backward worst rel %.3e' % worst, flush=True)


def stage_render(P=1000):
    a, b = models()
    x = points(P)
    g = torch.Generator().manual_seed(4)
    nrm = torch.randn(P, 3, generator=g).to(DEV)
    view = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1).to(DEV)
    feat = (torch.randn(P, 256, generator=g) * 0.3).to(DEV)
    wr = torch.randn(P, 3, generator=g).to(DEV)
    res, outs, din = [], [], []
    for m in (a, b):
        m.zero_grad()
        f = feat.clone().requires_grad_(True)
        n = nrm.clone().requires_grad_(True)
        rgb = m.rendering_network(x, n, view, f)
        (rgb * wr).sum().backward()
        torch.cuda.synchronize()
        outs.append(rgb.detach())
        din.append((f.grad, n.grad))
        res.append(_grads(m))
    print('render fwd    max|d| %.3e' % mx(outs[1], outs[0]), flush=True)
    print('render d_feat rel %.3e  d_normals rel %.3e' % (rel(din[1][0], din[0][0]), rel(din[1][1], din[0][1])), flush=True)
    for n in res[0]:
        print('  %-40s rel %.3e' % (n, rel(res[1][n], res[0][n])), flush=True)


def stage_speed(P=131072):
    a, b = models()
    x = points(P)
    for name, m in (('fp32', a), ('bf16', b)):
        with torch.no_grad():
            for _ in range(2):
                m.implicit_network.get_sdf_vals(x)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                m.implicit_network.get_sdf_vals(x)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
        print('get_sdf_vals %s: %.3f ms for %d points -> %.1f TFLOP/s' % (name, dt * 1e3, P, P * 1049088 / dt / 1e12), flush=True)


def stage_model(R=512):
    """full train step: fp32 engine vs tensor-core engine on the SAME sample positions"""
    import svolsdf_b200.scene as S
    a, b = models()
    a.train()
    b.train()
    inp = {k: v.to(DEV) for k, v in S.make_input('dtu', R).items()}
    gt = S.gt_rgb(R).reshape(-1, 3).to(DEV)
    outs, grads = [], []
    zs = None
    for m in (a, b):
        m.zero_grad()
        if zs is not None:
            orig = m.ray_sampler.get_z_vals

            def patched(*args, _o=orig, **kw):   # same RNG consumption, fp32-engine sample positions
                _o(*args, **kw)
                return zs
            m.ray_sampler.get_z_vals = patched
        torch.manual_seed(123)
        out = m(inp, fast=1)
        if zs is None:
            zs = m.last_z
        loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
        loss.backward()
        torch.cuda.synchronize()
        outs.append({k: v.detach() for k, v in out.items()})
        grads.append(_grads(m))
        print('loss %.6f' % float(loss), flush=True)
    for k in ('rgb_values', 'depth_values', 'weights', 'grad_theta'):
        print('  %-14s max|d| %.3e rel %.3e' % (k, mx(outs[1][k], outs[0][k]), rel(outs[1][k], outs[0][k])), flush=True)
    worst = 0
    for n in grads[0]:
        r = rel(grads[1][n], grads[0][n])
        worst = max(worst, r)
        if r > 5e-3:
            print('  %-40s rel %.3e |g| %.3e' % (n, r, float(grads[0][n].norm())), flush=True)
    print('model train step worst param-grad rel %.3e' % worst, flush=True)
    # sampler with the tensor-core SDF: how far do the sample positions move?
    b2 = models()[1].train()
    torch.manual_seed(123)
    b2(inp, fast=1)
    print('sampler z (tc sdf vs fp32 sdf) max|d| %.3e' % mx(b2.last_z[0], zs[0]), flush=True)


def stage_bmvs(R=int(os.environ.get('TC_R', '256'))):
    import svolsdf_b200.scene as S
    a, b = models('bmvs')
    a.train()
    b.train()
    inp = {k: v.to(DEV) for k, v in S.make_input('bmvs', R).items()}
    gt = S.gt_rgb(R).reshape(-1, 3).to(DEV)
    outs, grads = [], []
    zs = None
    for m in (a, b):
        m.zero_grad()
        if zs is not None:
            orig = m.ray_sampler.get_z_vals

            def patched(*args, _o=orig, **kw):
                _o(*args, **kw)
                return zs
            m.ray_sampler.get_z_vals = patched
        torch.manual_seed(123)
        out = m(inp, fast=1)
        if zs is None:
            zs = m.last_z
        loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
        loss.backward()
        torch.cuda.synchronize()
        outs.append({k: v.detach() for k, v in out.items()})
        grads.append(_grads(m))
        print('loss %.6f' % float(loss), flush=True)
    for k in ('rgb_values', 'depth_values', 'weights', 'grad_theta'):
        print('  %-14s max|d| %.3e rel %.3e' % (k, mx(outs[1][k], outs[0][k]), rel(outs[1][k], outs[0][k])), flush=True)
    worst = 0
    for n in grads[0]:
        r = rel(grads[1][n], grads[0][n])
        worst = max(worst, r)
        if r > 1e-2 or 'bg_rendering' in n:
            print('  %-40s rel %.3e |g| %.3e' % (n, r, float(grads[0][n].norm())), flush=True)
    print('bmvs train step worst param-grad rel %.3e' % worst, flush=True)


def stage_det(R=int(os.environ.get('TC_R', '64'))):
    """run-to-run determinism of the bmvs train step per engine (races show up as large differences)"""
    import svolsdf_b200.scene as S
    inp = {k: v.to(DEV) for k, v in S.make_input('bmvs', R).items()}
    gt = S.gt_rgb(R).reshape(-1, 3).to(DEV)
    for name, eng in (('fp32', L.ENGINE_FP32), ('tc', L.ENGINE_TC)):
        m = build_model('bmvs', perturb=True, beta=0.05, device=DEV).train().set_engine(eng)
        runs = []
        for it in range(3):
            m.zero_grad()
            torch.manual_seed(321)
            out = m(inp, fast=1)
            loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean() + \
                0.05 * out['weights'].pow(2).sum(1).mean() + 0.1 * out['depth_values_all'].mean()
            loss.backward()
            torch.cuda.synchronize()
            runs.append(_grads(m))
        for it in (1, 2):
            worst = max((rel(runs[it][n], runs[0][n]), n) for n in runs[0])
            print('%s run %d vs run 0: worst rel %.3e (%s)' % (name, it, worst[0], worst[1]), flush=True)


def stage_tails():
    a, b = models()
    for P in (1, 5, 127, 128, 129, 300):
        x = points(P, seed=P)
        with torch.no_grad():
            s0, f0, g0 = a.implicit_network.get_outputs(x)
            s1, f1, g1 = b.implicit_network.get_outputs(x)
        print('P=%d sdf %.2e feat %.2e grad %.2e' % (P, mx(s1, s0), rel(f1, f0), rel(g1, g0)), flush=True)


if __name__ == '__main__':
    stages = sys.argv[1:] or ['fwd', 'outputs', 'backward', 'render', 'speed']
    for s in stages:
        print('== stage', s, flush=True)
        globals()['stage_' + s]()
