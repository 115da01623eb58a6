"""Bring-up check + timing of SVS_ENGINE_TC_SPLIT against the fp32 engine (GPU box; not part of the product/tests).
    timeout 300 python tools/split_check.py"""
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
import svolsdf_b200.scene as S  # noqa: E402

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


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    engines = {'fp32': L.ENGINE_FP32, 'tc': L.ENGINE_TC, 'split': L.ENGINE_TC_SPLIT}
    ms = {k: build_model('dtu', perturb=True, beta=0.05, device=DEV).set_engine(e) for k, e in engines.items()}
    for P in (1000, 131072):
        x = points(P, seed=P)
        with torch.no_grad():
            ref = ms['fp32'].implicit_network.get_sdf_vals(x)
            for k in ('tc', 'split'):
                s = ms[k].implicit_network.get_sdf_vals(x)
                torch.cuda.synchronize()
                print('P=%d get_sdf_vals[%s] max|d| %.3e' % (P, k, mx(s, ref)), flush=True)
            yr = ms['fp32'].implicit_network(x)
            for k in ('tc', 'split'):
                y = ms[k].implicit_network(x)
                torch.cuda.synchronize()
                print('P=%d forward y[%s] max|d| %.3e rel %.3e' % (P, k, mx(y, yr), rel(y, yr)), flush=True)
            sr, fr, gr = ms['fp32'].implicit_network.get_outputs(x)
            for k in ('tc', 'split'):
                s, f, g = ms[k].implicit_network.get_outputs(x)
                torch.cuda.synchronize()
                print('P=%d get_outputs[%s] sdf %.3e feat %.3e grad rel %.3e' % (P, k, mx(s, sr), mx(f, fr), rel(g, gr)), flush=True)
    # rendering net
    P = 100352
    x = points(P, seed=5)
    g = torch.Generator().manual_seed(4)
    nrm = torch.randn(P, 3, generator=g).to(DEV)
    view = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1).to(DEV)
    feat = (torch.randn(P, 256, generator=g) * 0.3).to(DEV)
    with torch.no_grad():
        rr = ms['fp32'].rendering_network(x, nrm, view, feat)
        for k in ('tc', 'split'):
            r = ms[k].rendering_network(x, nrm, view, feat)
            torch.cuda.synchronize()
            print('render fwd[%s] max|d| %.3e' % (k, mx(r, rr)), flush=True)
    # timings of the chains
    L.load().svs_prof_enable(1)
    x = points(131072, seed=1)
    xm = points(100352 + 2048, seed=2)
    for k in ('tc', 'split'):
        with torch.no_grad():
            for _ in range(5):
                ms[k].implicit_network.get_sdf_vals(x)
        ms[k].train()
        for _ in range(3):
            y, s, gq = ms[k].implicit_network.outputs_fused(xm, clamp=100352)
            r = ms[k].rendering_network(xm[:100352], gq[:100352], view, y[:100352], _feat_col=1)
        torch.cuda.synchronize()
        import ctypes
        buf = ctypes.create_string_buffer(1 << 16)
        n = L.load().svs_prof_collect(buf, len(buf))
        print('--- engine', k)
        for line in buf.raw[:n].decode().strip().split('\n'):
            name, cnt, ms_tot, fl, by = line.split('\t')
            cnt, ms_tot, fl = int(cnt), float(ms_tot), float(fl)
            print('  %-22s x%-3d %.3f ms/launch  %.0f TFLOP/s (algorithmic)' % (name, cnt, ms_tot / cnt, fl / cnt / (ms_tot / cnt) / 1e9))
    # whole train step
    inp = {k: v.to(DEV) for k, v in S.make_input('dtu', 1024).items()}
    gt = S.gt_rgb(1024).reshape(-1, 3).to(DEV)
    for k in ('tc', 'split'):
        m = ms[k].train()

        def step():
            out = m(inp, fast=1)
            loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
            m.zero_grad()
            loss.backward()
        print('eager train step [%s]: %.3f ms' % (k, timeit(step, 10)))


if __name__ == '__main__':
    main()
