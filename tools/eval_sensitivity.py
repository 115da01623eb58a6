"""How sensitive is the eval render (5 sampler iterations) to 1e-6-level sdf perturbations?  fp32 engine vs itself with noise
injected into the sampler's sdf, and vs the split engine (GPU box)."""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model
import svolsdf_b200._lib as L
import svolsdf_b200.scene as S
DEV = 'cuda'
R = 700
for beta in (0.05, 0.01):
    a = build_model('dtu', perturb=True, beta=beta, device=DEV).eval()
    inp = {k: v.to(DEV) for k, v in S.make_input('dtu', R).items()}
    torch.manual_seed(5)
    oa = a(inp)
    def stats(name, ob):
        d = (ob['depth_values'] - oa['depth_values']).abs().flatten()
        r = (ob['rgb_values'] - oa['rgb_values']).abs().max(1)[0]
        q = lambda t, p: float(t.kthvalue(max(1, int(p * t.numel())))[0])
        print('beta %g %-28s depth p50 %.2e p90 %.2e p98 %.2e max %.2e | rgb p98 %.2e max %.2e | iters %s' % (
            beta, name, q(d, .5), q(d, .9), q(d, .98), float(d.max()), q(r, .98), float(r.max()), ob.get('it')), flush=True)
    for noise in (1e-7, 1e-6, 1e-5):
        b = build_model('dtu', perturb=True, beta=beta, device=DEV).eval()
        g = torch.Generator(device=DEV).manual_seed(1)
        orig = b.implicit_network.get_sdf_vals
        def noisy(x, orig=orig, noise=noise):
            s = orig(x)
            return s + noise * torch.randn(s.shape, device=DEV, generator=g)
        b.implicit_network.get_sdf_vals = noisy
        torch.manual_seed(5)
        ob = b(inp)
        ob['it'] = b.ray_sampler.last_iters
        stats('fp32 + sampler noise %.0e' % noise, ob)
    for ename, e in (('tc_split', L.ENGINE_TC_SPLIT), ('tc', L.ENGINE_TC)):
        b = build_model('dtu', perturb=True, beta=beta, device=DEV).eval().set_engine(e)
        torch.manual_seed(5)
        ob = b(inp)
        ob['it'] = b.ray_sampler.last_iters
        stats(ename, ob)
