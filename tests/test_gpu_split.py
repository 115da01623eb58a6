"""SVS_ENGINE_TC_SPLIT — the benchmarked tcgen05 engine (split fp16 hi + lo operands in every forward chain, single fp16
operands in the backward chains) — against the fp64 oracle at the north_star tolerances: rgb / depth max-abs <= 1e-3,
parameter gradients relative error <= 1e-2.  Sizes: the bench's own (1024 rays = 800 tiles > 148 persistent CTAs, every
chain — forward, reverse sweep, tangent sweep, backward, weight gradients — walks 5-6 tiles per CTA) and trained-like
beta (0.01, 0.001), where a 1e-3 sdf error would be 1 beta.  End-to-end tests against the reference's goldens (own
sampler, no injected sample positions) are in test_gpu_model.py, parametrised over this engine."""
import pytest
import torch

from helpers import build_model, conf_of, max_abs, rel_err, state_dict_cpu
from oracle import volsdf_oracle as O
import svolsdf_b200._lib as L
import svolsdf_b200.scene as S

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _points(P, seed=0, radius=3.6):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(P, 3, generator=g)
    return x / x.norm(dim=1, keepdim=True) * (torch.rand(P, 1, generator=g) * radius)


@pytest.mark.parametrize('P', [1, 127, 129, 5000, 131072])
def test_sdf_forward_vs_fp64_oracle(P):
    """get_sdf_vals / forward / get_outputs on the split engine vs the fp64 oracle; 131072 points = 1024 tiles (the
    sampler launch of a 1024-ray step), both sphere-clamp branches (points reach |x| = 3.6 > 3)"""
    m = build_model('dtu', perturb=True, beta=0.05, device=DEV).set_engine(L.ENGINE_TC_SPLIT)
    sd = {k: v.double() for k, v in state_dict_cpu(m).items()}
    x = _points(P, seed=P)
    with torch.no_grad():
        y_ref = O.sdf_net(sd, 'implicit_network', x.double(), 6)
        s_ref = O.sphere_clamp(y_ref[:, :1], x.double(), 3.0, float(conf_of('dtu').get_config('implicit_network').get('sphere_scale', 1.0)))
        xd = x.to(DEV)
        s = m.implicit_network.get_sdf_vals(xd).cpu()
        y = m.implicit_network(xd).cpu()
        s2, f2, g2 = m.implicit_network.get_outputs(xd)
    # |sdf| reaches ~15 in the clamp region: 5e-5 absolute is 3e-6 relative there
    assert max_abs(s, s_ref) < 5e-5, max_abs(s, s_ref)
    assert max_abs(y, y_ref) < 5e-5, max_abs(y, y_ref)
    assert max_abs(s2.cpu(), s_ref) < 5e-5 and max_abs(f2.cpu(), y_ref[:, 1:]) < 5e-5
    if P <= 5000:
        xg = x.double().requires_grad_(True)
        yy = O.sphere_clamp(O.sdf_net(sd, 'implicit_network', xg, 6)[:, :1], xg, 3.0,
                            float(conf_of('dtu').get_config('implicit_network').get('sphere_scale', 1.0)))
        g_ref = torch.autograd.grad(yy.sum(), xg)[0]
        # the reverse sweep keeps single fp16 operands: normals to ~1e-3 relative
        assert rel_err(g2.cpu(), g_ref) < 2e-3, rel_err(g2.cpu(), g_ref)


@pytest.mark.parametrize('engine', [pytest.param(L.ENGINE_TC_SPLIT, id='tc_split'), pytest.param(L.ENGINE_FP32, id='fp32')])
@pytest.mark.parametrize('beta', [0.05, 0.01, 0.001])
def test_train_step_1024_rays_vs_fp64_oracle(beta, engine):
    """The bench's step (1024 rays, L1 + 0.1 eikonal) on the bench's engine: outputs and ALL parameter gradients against
    fp64 autograd on the oracle, on the sample positions the model drew with its own sampler."""
    R = 1024
    model = build_model('dtu', perturb=True, beta=beta, device=DEV).train().set_engine(engine)
    sd = state_dict_cpu(model)
    inp = S.make_input('dtu', R)
    gt = S.gt_rgb(R)
    torch.manual_seed(321)
    out = model({k: v.to(DEV) for k, v in inp.items()}, fast=1)
    loss = (out['rgb_values'] - gt.reshape(-1, 3).to(DEV)).abs().mean() + \
        0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
    model.zero_grad()
    loss.backward()
    ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    z, z_eik = model.last_z
    torch.manual_seed(321)
    rng = O.draw_rng(R, True)
    o = O.volsdf_forward(ref, conf_of('dtu'), inp, True, fast=1, rng=rng, dtype=torch.float64,
                         z_override=(z.cpu(), z_eik.cpu(), None))
    rl = O.volsdf_loss(o, gt)
    rl.backward()
    errs = {k: max_abs(out[k].detach().cpu(), o[k].detach()) for k in ('rgb_values', 'depth_values', 'weights', 'grad_theta')}
    rows = []
    for name, p in model.named_parameters():
        rg = ref[name].grad
        if rg is None or float(rg.norm()) < 1e-10:
            continue
        rows.append((rel_err(p.grad.cpu(), rg), name))
    rows.sort(reverse=True)
    print('beta %g: %s | loss %.7f vs %.7f | worst grads %s' % (beta, errs, float(loss), float(rl), rows[:4]))
    assert errs['rgb_values'] < 1e-3 and errs['depth_values'] < 1e-3, errs
    assert errs['grad_theta'] < 5e-3 and errs['weights'] < 1e-3, errs      # normals: single-fp16 reverse sweep
    assert abs(float(loss) - float(rl)) < 1e-4
    assert len(rows) == 43 and rows[0][0] < 1e-2, rows[:6]


def test_bmvs_train_step_vs_fp64_oracle():
    """BlendedMVS model (background SDF net with d_in = 4 / 10 frequencies, 'nerf' rendering net), 256 rays"""
    R = 256
    model = build_model('bmvs', perturb=True, beta=0.02, device=DEV).train().set_engine(L.ENGINE_TC_SPLIT)
    sd = state_dict_cpu(model)
    inp = S.make_input('bmvs', R)
    gt = S.gt_rgb(R)
    torch.manual_seed(321)
    out = model({k: v.to(DEV) for k, v in inp.items()}, fast=1)
    loss = (out['rgb_values'] - gt.reshape(-1, 3).to(DEV)).abs().mean() + \
        0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
    model.zero_grad()
    loss.backward()
    ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    z, z_eik = model.last_z
    torch.manual_seed(321)
    rng = O.draw_rng(R, True, bg=True)
    o = O.volsdf_bg_forward(ref, conf_of('bmvs'), inp, True, fast=1, rng=rng, dtype=torch.float64,
                            z_override=((z[0].cpu(), z[1].cpu()), z_eik.cpu(), None))
    rl = O.volsdf_loss(o, gt)
    rl.backward()
    assert max_abs(out['rgb_values'].detach().cpu(), o['rgb_values'].detach()) < 1e-3
    assert max_abs(out['weights'].detach().cpu(), o['weights'].detach()) < 1e-3
    # depth = sum(w z) / (sum(w) + 1e-8) is ill-conditioned on rays that hit nothing (the BlendedMVS model has no sphere
    # clamp; same convention as test_gpu_model.py); depth_values_all inherits that through 1 / depth of the background
    hit = o['weights'].detach().sum(1, keepdim=True) > 1e-2
    assert max_abs(out['depth_values'].detach().cpu()[hit], o['depth_values'].detach()[hit]) < 1e-3
    rows = []
    for name, p in model.named_parameters():
        rg = ref[name].grad
        if rg is None or float(rg.norm()) < 1e-10:
            continue
        rows.append((rel_err(p.grad.cpu(), rg), name))
    rows.sort(reverse=True)
    print('bmvs worst grads', rows[:4])
    assert rows[0][0] < 1e-2, rows[:6]


def test_eval_render_matches_fp32_engine():
    """eval forward at beta = 0.01 (5 sampler iterations, 700 rays = 5 x 700 x 128 sampler points).  Both engines run
    their own sampler (same iteration count required); the maps are compared on the fp32 engine's sample positions:
    with 700 rays some graze the surface, where 1e-6 of sdf moves a bisection decision and with it the samples — the
    end-to-end (own positions) comparison against the reference's recorded outputs is test_gpu_model.py's golden test."""
    a = build_model('dtu', perturb=True, beta=0.01, device=DEV).eval()
    b = build_model('dtu', perturb=True, beta=0.01, device=DEV).eval().set_engine(L.ENGINE_TC_SPLIT)
    inp = {k: v.to(DEV) for k, v in S.make_input('dtu', 700).items()}
    torch.manual_seed(5)
    oa = a(inp)
    zs = a.last_z
    torch.manual_seed(5)
    ob_own = b(inp)
    assert a.ray_sampler.last_iters == b.ray_sampler.last_iters == 5
    # own positions: measured with tools/eval_sensitivity.py, 1e-7 of noise on the fp32 engine's OWN sampler sdf already
    # moves the depth of the worst of 700 rays by 7e-3 and the 98th percentile by 3e-5 (1e-6: 1.2e-2 / 5e-4) — a few
    # grazing rays are ill-conditioned in the reference algorithm itself, so the bulk is what can be compared
    d = (ob_own['depth_values'] - oa['depth_values']).abs().flatten()
    assert float(d.median()) < 1e-5 and float(d.kthvalue(int(0.9 * d.numel()))[0]) < 2e-4
    assert max_abs(ob_own['rgb_values'], oa['rgb_values']) < 1e-3
    orig = b.ray_sampler.get_z_vals

    def patched(*args, **kw):
        orig(*args, **kw)
        return zs
    b.ray_sampler.get_z_vals = patched
    torch.manual_seed(5)
    ob = b(inp)
    assert max_abs(ob['rgb_values'], oa['rgb_values']) < 1e-3
    assert max_abs(ob['depth_values'], oa['depth_values']) < 1e-3
    assert max_abs(ob['normal_map'], oa['normal_map']) < 5e-3


@pytest.mark.parametrize('engine', [pytest.param(L.ENGINE_FP32, id='fp32'), pytest.param(L.ENGINE_TC_SPLIT, id='tc_split')])
def test_grouped_render_equals_the_reference_chunk_loop(engine):
    """one launch over many rays with 512-ray convergence groups INSIDE it (ErrorBoundSampler.group_size)
    must equal the reference's render loop, which calls the model on 512 rays at a time (vsdf.py:246-262,
    eval_vsdf.py:216-228): same sampler iterations per group, same maps ray for ray.  The ray set mixes a group that
    misses the object (image corner), groups on it, and a ragged tail."""
    from svolsdf_b200.render import render_rays
    m = build_model('dtu', perturb=True, beta=0.03, device=DEV).eval().set_engine(engine)
    inp = S.make_input('dtu', 16)
    xs = torch.arange(512, dtype=torch.float32)
    corner = torch.stack([xs % 64, torch.div(xs, 64, rounding_mode='floor')], -1)                    # 64 x 8 block at (0, 0)
    centre = torch.stack([800 + xs % 32 * 3, 600 + torch.div(xs, 32, rounding_mode='floor') * 3], -1)
    g = torch.Generator().manual_seed(3)
    rnd = torch.stack([torch.randint(0, 1600, (812,), generator=g), torch.randint(0, 1200, (812,), generator=g)], -1).float()
    uv = torch.cat([corner, centre, rnd], 0)[None].to(DEV)                                            # 1836 rays = 3 groups + 300
    K_, pose = inp['intrinsics'].to(DEV), inp['pose'].to(DEV)
    torch.manual_seed(3)
    got = render_rays(m, K_, pose, uv, chunk=4096, group=512)
    parts, iters = [], []
    for lo in range(0, uv.shape[1], 512):
        torch.manual_seed(3)
        parts.append(m({'intrinsics': K_, 'pose': pose, 'uv': uv[:, lo:lo + 512].contiguous()}))
        iters.append(m.ray_sampler.last_iters)
    assert got['sampler_iters'] == iters, (got['sampler_iters'], iters)
    assert len(set(iters)) > 1, iters          # the groups really stop at different iterations
    for k in ('rgb_values', 'depth_values', 'normal_map'):
        assert torch.equal(got[k], torch.cat([p[k] for p in parts], 0)), k


def test_cta_pair_mode_and_dy_register_path_match_the_default(tmp_path):
    """SVS_F3_PAIR=1 (cta_group::2: clusters of two CTAs, one M = 256 instruction per weight block for both tiles) computes
    the same arithmetic in the same order as independent CTAs: sdf / y bit-identical, including an odd tile count (the
    pair's padding tile); saved activations: the gradients of a small step agree to the rounding order of the atomics.  The switch is read once
    per process, so each mode runs in its own interpreter."""
    import os
    import subprocess
    import sys
    script = r'''
import sys, torch, warnings
warnings.filterwarnings('ignore')
sys.path.insert(0, %r); sys.path.insert(0, %r)
from helpers import build_model
import svolsdf_b200._lib as L
m = build_model('dtu', perturb=True, beta=0.05, device='cuda').set_engine(L.ENGINE_TC_SPLIT).train()
g = torch.Generator().manual_seed(5)
out = {}
for P in (129, 5000, 12345):          # 2, 40, 97 tiles (odd: padding tile in the last pair)
    x = ((torch.rand(P, 3, generator=g) * 2 - 1) * 1.2).cuda()
    with torch.no_grad():
        out['s%%d' %% P] = m.implicit_network.get_sdf_vals(x).cpu()
        out['y%%d' %% P] = m.implicit_network(x).cpu()
x = ((torch.rand(3001, 3, generator=g) * 2 - 1) * 1.2).cuda()   # 57 rows in the last tile: its dy bytes are not a multiple of 16
sdf, feat, grad = m.implicit_network.get_outputs(x)
(sdf.sum() + feat.square().mean() + grad.square().sum()).backward()
for n, p in m.implicit_network.named_parameters():
    out['g_' + n] = p.grad.cpu()
torch.save(out, sys.argv[1])
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode, env_add in (('0', {'SVS_F3_PAIR': '0'}), ('1', {'SVS_F3_PAIR': '1'}), ('nobulk', {'SVS_DY_BULK': '0'})):
        env = dict(os.environ, **env_add)
        path = str(tmp_path / ('out%s.pt' % mode))
        subprocess.run([sys.executable, '-c', script, path], check=True, env=env, timeout=300)
        res[mode] = torch.load(path)
    assert res['0'].keys() == res['1'].keys()
    # SVS_DY_BULK=0: the backward chain converts dy with register loads instead of staging the rows through the aux ring
    # (the path taken when dy is not 16-byte aligned): same fp16 operand tile, same gradients
    for k in res['0']:
        if k.startswith('g_'):
            assert rel_err(res['nobulk'][k], res['0'][k]) < 1e-4, (k, rel_err(res['nobulk'][k], res['0'][k]))
    # forward results bit for bit; weight gradients are sums of fp32 atomics (run-to-run rounding order) over identical tiles
    bad = {k: float((res['0'][k] - res['1'][k]).abs().max()) for k in res['0'] if not k.startswith('g_') and not torch.equal(res['0'][k], res['1'][k])}
    assert not bad, bad
    for k in res['0']:
        if k.startswith('g_'):
            assert rel_err(res['1'][k], res['0'][k]) < 1e-4, (k, rel_err(res['1'][k], res['0'][k]))


def test_mesh_grid_and_point_queries_vs_fp64_oracle():
    """svolsdf_b200.mesh.sdf_grid / sdf_points (the reference's meshing queries `model.implicit_network(x)[:, 0]`,
    utils/plots.py:61,69-76) on the benchmarked engine against the fp64 oracle — ragged chunks, an odd
    number of tiles per chunk, points outside the bounding sphere (raw network output: no clamp in this query)."""
    from svolsdf_b200.mesh import sdf_grid, sdf_points
    m = build_model('dtu', perturb=True, device=DEV).eval().set_engine(L.ENGINE_TC_SPLIT)
    sd = {k: v.double() for k, v in state_dict_cpu(m).items()}
    grid, axes = sdf_grid(m, resolution=33, bound=(-1.2, 1.2, -0.9, 1.5, -3.5, 3.5), chunk=7001)
    assert grid.shape == (33, 33, 33)
    pts = torch.stack(torch.meshgrid(*[a.cpu() for a in axes], indexing='ij'), -1).reshape(-1, 3)
    with torch.no_grad():
        ref = O.sdf_net(sd, 'implicit_network', pts.double(), 6)[:, 0]
    assert max_abs(grid.reshape(-1).cpu(), ref) < 5e-5, max_abs(grid.reshape(-1).cpu(), ref)
    assert float(grid.min()) < 0 < float(grid.max())       # the surface crosses the box
    x = _points(5003, seed=4)
    with torch.no_grad():
        ref_p = O.sdf_net(sd, 'implicit_network', x.double(), 6)[:, 0]
    assert max_abs(sdf_points(m, x.to(DEV), chunk=999).cpu(), ref_p) < 5e-5
