"""Parity table: engines x beta at 1024 rays vs the fp64 oracle on the model's own sample positions (GPU box)."""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model, conf_of, max_abs, rel_err, state_dict_cpu
from oracle import volsdf_oracle as O
import svolsdf_b200._lib as L
import svolsdf_b200.scene as S

DEV = 'cuda'
kind = sys.argv[1] if len(sys.argv) > 1 else 'dtu'
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
betas = [float(b) for b in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0.05, 0.01, 0.003, 0.001]
for beta in betas:
    for ename, engine in (('fp32', L.ENGINE_FP32), ('tc', L.ENGINE_TC), ('tc_split', L.ENGINE_TC_SPLIT)):
        model = build_model(kind, perturb=True, beta=beta, device=DEV).train().set_engine(engine)
        sd = state_dict_cpu(model)
        inp = S.make_input(kind, R)
        gt = S.gt_rgb(R)
        torch.manual_seed(321)
        out = model({k: v.to(DEV) for k, v in inp.items()}, fast=1)
        loss = (out['rgb_values'] - gt.reshape(-1, 3).to(DEV)).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
        model.zero_grad()
        loss.backward()
        ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
        z, z_eik = model.last_z
        torch.manual_seed(321)
        rng = O.draw_rng(R, True, bg=(kind == 'bmvs'))
        if kind == 'dtu':
            o = O.volsdf_forward(ref, conf_of(kind), inp, True, fast=1, rng=rng, dtype=torch.float64, z_override=(z.cpu(), z_eik.cpu(), None))
        else:
            o = O.volsdf_bg_forward(ref, conf_of(kind), inp, True, fast=1, rng=rng, dtype=torch.float64,
                                    z_override=((z[0].cpu(), z[1].cpu()), z_eik.cpu(), None))
        O.volsdf_loss(o, gt).backward()
        hit = o['weights'].detach().sum(1, keepdim=True) > 1e-2
        keys = ['rgb_values', 'depth_values', 'weights', 'grad_theta'] + (['depth_values_all'] if kind == 'bmvs' else [])
        errs = {}
        for k in keys:
            a, b = out[k].detach().cpu(), o[k].detach()
            errs[k] = max_abs(a, b)
            if k == 'depth_values':
                errs['depth(hit)'] = max_abs(a[hit], b[hit])
                d = (a.double() - b).abs().flatten()
                errs['depth p99'] = float(d.kthvalue(int(0.99 * d.numel()))[0])
        rows = sorted(((rel_err(p.grad.cpu(), ref[n].grad), n) for n, p in model.named_parameters()
                       if ref[n].grad is not None and float(ref[n].grad.norm()) > 1e-10), reverse=True)
        print('%s R=%d beta=%g %-8s %s | worst grad %.2e %s, 2nd %.2e %s' % (
            kind, R, beta, ename, ' '.join('%s %.2e' % (k, v) for k, v in errs.items()), rows[0][0], rows[0][1], rows[1][0], rows[1][1]), flush=True)
