"""Scratch diagnostics run on the GPU box while bringing kernels up (not part of the product or tests)."""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')

from helpers import build_model, conf_of, state_dict_cpu  # noqa: E402
from oracle import volsdf_oracle as O  # noqa: E402
import svolsdf_b200.scene as S  # noqa: E402

DEV = 'cuda'


def dbg_depth2pts():
    mb = build_model('bmvs', perturb=True, device=DEV)
    inp = S.make_input('bmvs', 128)
    rd, cl = O.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
    depth = torch.rand(128, 32, generator=torch.Generator().manual_seed(2))
    o = cl.unsqueeze(1).repeat(128, 32, 1)
    dd = rd[0].unsqueeze(1).repeat(1, 32, 1)
    pts, dreal = O.depth2pts_outside(o, dd, depth, 3.0)
    p2, d2 = mb.depth2pts_outside(o.to(DEV), dd.to(DEV), depth.to(DEV))
    p2, d2 = p2.cpu(), d2.cpu()
    print('oracle nan', int(torch.isnan(pts).sum()), 'cuda nan', int(torch.isnan(p2).sum()), int(torch.isnan(d2).sum()))
    bad = torch.isnan(p2).any(-1).nonzero()
    print('bad idx', bad[:5].tolist())
    for r, s in bad[:3].tolist():
        print(' o', o[r, s].tolist(), 'd', dd[r, s].tolist(), 'depth', depth[r, s].item(), 'oracle', pts[r, s].tolist())
    ok = ~torch.isnan(p2).any(-1)
    print('max err on finite', float((p2[ok] - pts[ok]).abs().max()))


def dbg_sampler(kind, training, beta):
    from test_gpu_model import _sampler_case
    (z_ref, z_eik_ref, tr), got, trace = _sampler_case(kind, training, beta)
    print('== sampler', kind, training, beta, 'iters', len(trace), len(tr.iters))
    for i, (a, b) in enumerate(zip(trace, tr.iters)):
        for k in ('z', 'sdf', 'beta', 'inds', 'samples', 'samples_idx'):
            if k not in b or a.get(k) is None:
                continue
            x, y = a[k].cpu(), b[k]
            if x.dtype != y.dtype:
                x = x.to(y.dtype)
            ne = (x != y)
            if ne.any():
                d = (x.double() - y.double()).abs()
                idx = ne.nonzero()
                print(' it', i, k, 'mismatch', int(ne.sum()), 'of', ne.numel(), 'maxdiff', float(d.max()),
                      'first', idx[:4].tolist(), 'cols', sorted(set(idx[:, -1].tolist()))[:12])
                if k == 'samples':
                    r, j = idx[0].tolist()
                    print('   got', x[r, j].item(), 'ref', y[r, j].item(), 'ind', int(b['inds'][r, j]), 'n', b['n'],
                          'cdf around', b['cdf'][r, max(0, int(b['inds'][r, j]) - 2):int(b['inds'][r, j]) + 2].tolist())
            else:
                print(' it', i, k, 'OK')
    zg = got[0][0] if kind == 'bmvs' else got[0]
    zr = z_ref[0] if kind == 'bmvs' else z_ref
    print(' final z equal', torch.equal(zg.cpu(), zr), 'z_eik equal', torch.equal(got[1].cpu(), z_eik_ref))
    if kind == 'bmvs':
        print(' z_bg equal', torch.equal(got[0][1].cpu(), z_ref[1]))


def dbg_far():
    from svolsdf_b200 import functional as F
    inp = S.make_input('bmvs', 96)
    rd, cl = O.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
    dirs, cam = rd[0].contiguous(), cl.expand(96, 3).contiguous()
    ref = O.get_sphere_intersections(cam, dirs, 3.0)
    nf, bad = F.sphere_intersections(cam.to(DEV), dirs.to(DEV), 3.0)
    ne = (nf.cpu() != ref)
    print('far mismatch', int(ne.sum()), 'of', ne.numel(), float((nf.cpu() - ref).abs().max()))
    zr = O.uniform_z(torch.zeros(96, 1), ref[:, 1:], 128)
    from svolsdf_b200.model.ray_sampler import UniformSampler

    class M(object):
        training = False
    us = UniformSampler(3.0, 0.0, 128, take_sphere_intersection=True)
    zg = us.get_z_vals(dirs.to(DEV), cam.to(DEV), M(), _far_ray=ref[:, 1].contiguous().to(DEV))
    ne = zg.cpu() != zr
    print('uniform z (same far) mismatch', int(ne.sum()), float((zg.cpu() - zr).abs().max()), ne.nonzero()[:4].tolist())


if __name__ == '__main__':
    which = sys.argv[1:] or ['d2p', 'far', 'samp']
    if 'd2p' in which:
        dbg_depth2pts()
    if 'far' in which:
        dbg_far()
    if 'samp' in which:
        dbg_sampler('dtu', False, None)
        dbg_sampler('dtu', False, 0.01)
        dbg_sampler('bmvs', True, None)
