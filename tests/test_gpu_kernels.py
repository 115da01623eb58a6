"""GPU parity: every CUDA entry point against the CPU oracle on the same seeded inputs (through the C ABI).

Tolerances (fp32 engine): element-wise kernels <= 1e-5, MLP outputs <= 5e-5 absolute (fp32 accumulation
order), parameter gradients <= 2e-3 relative per tensor against the fp64 oracle; the sampler in exact
mode is BIT-EXACT (indices, counts and positions).
"""
import pytest
import torch

from helpers import build_model, conf_of, load_golden, max_abs, rel_err, state_dict_cpu
from oracle import volsdf_oracle as O
import svolsdf_b200._lib as L
import svolsdf_b200.scene as S

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def dtu():
    m = build_model('dtu', perturb=True, device=DEV)
    return m, state_dict_cpu(m)


@pytest.fixture(scope='module')
def bmvs():
    m = build_model('bmvs', perturb=True, device=DEV)
    return m, state_dict_cpu(m)


def _points(n, scale=2.4, seed=5, d=3):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, d, generator=g) * 2 - 1) * scale


def test_library_loaded_is_ours():
    import svolsdf_b200._lib as L
    assert L.load().svs_abi_version() == 6
    assert L.LIB_PATH.endswith('libsvolsdf_b200.so')


def test_pinned_rng_matches_reference_draws():
    from svolsdf_b200.model.ray_sampler import RefRng
    torch.manual_seed(9)
    r = RefRng(DEV)
    a, b, c, d = r.rand(7, 5), r.randperm(11), r.randint(98, (6,)), r.uniform((4, 3), -3.0, 3.0)
    torch.manual_seed(9)
    a0, b0, c0 = torch.rand(7, 5), torch.randperm(11), torch.randint(98, (6,))
    d0 = torch.empty(4, 3).uniform_(-3.0, 3.0)
    assert torch.equal(a.cpu(), a0) and torch.equal(b.cpu().long(), b0) and torch.equal(c.cpu(), c0)
    assert torch.equal(d.cpu(), d0)


def test_raygen_and_sphere():
    from svolsdf_b200 import functional as F
    from svolsdf_b200.utils import rend_util
    for kind in ('dtu', 'bmvs'):
        inp = S.make_input(kind, 500)
        rd, cl = O.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
        tmp, _ = O.get_camera_params(inp['uv'], torch.eye(4)[None], inp['intrinsics'])
        d, c, ds = F.raygen(inp['uv'][0].to(DEV), inp['pose'][0].to(DEV), inp['intrinsics'][0].to(DEV))
        assert max_abs(d.cpu(), rd[0]) < 1e-6
        assert max_abs(c.cpu(), cl.expand(500, 3)) == 0
        assert max_abs(ds.cpu(), tmp[0, :, 2:]) < 1e-6
        d2, c2 = rend_util.get_camera_params(inp['uv'].to(DEV), inp['pose'].to(DEV), inp['intrinsics'].to(DEV))
        assert max_abs(d2.cpu(), rd) < 1e-6 and max_abs(c2.cpu(), cl) == 0
        nf = rend_util.get_sphere_intersections(c, d, r=3.0)
        assert max_abs(nf.cpu(), O.get_sphere_intersections(cl.expand(500, 3), rd[0], 3.0)) < 1e-5
    # a ray that misses the sphere must raise where the reference exit()s: at once outside training ...
    miss = (torch.tensor([[10., 0, 0]], device=DEV), torch.tensor([[0., 1, 0]], device=DEV))
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            rend_util.get_sphere_intersections(*miss, r=1.0)
    # ... and without a host synchronisation inside a training step: the condition sticks to a device flag
    with torch.enable_grad():
        rend_util.get_sphere_intersections(*miss, r=1.0)
        rend_util.get_sphere_intersections(c, d, r=3.0)
        with pytest.raises(RuntimeError):
            rend_util.check_bounding_sphere()
        rend_util.check_bounding_sphere()      # reading clears it


def test_embed_and_density(dtu):
    from svolsdf_b200.model.embedder import get_embedder
    x = _points(300)
    for L_, d_in in ((6, 3), (1, 3), (10, 4), (4, 3)):
        xx = _points(300, d=d_in)
        fn, w = get_embedder(L_, input_dims=d_in)
        out = fn(xx.to(DEV))
        assert out.shape == (300, w)
        assert max_abs(out.cpu(), O.embed(xx, L_)) < 2e-6
    m, sd = dtu
    s = torch.randn(64, 98, generator=torch.Generator().manual_seed(3)) * 0.3
    beta = O.get_beta(sd['density.beta'], 0.0001)
    assert rel_err(m.density(s.to(DEV)).cpu(), O.laplace_density(s, beta)) < 1e-6
    rows = torch.rand(64, 1, generator=torch.Generator().manual_seed(4)) * 0.1 + 0.01
    assert rel_err(m.density(s.to(DEV), beta=rows.to(DEV)).cpu(), O.laplace_density(s, rows)) < 1e-6
    assert abs(float(m.density.get_beta()) - float(beta)) < 1e-9
    # autograd through the stand-alone density
    sg = s.clone().requires_grad_(True)
    bp = sd['density.beta'].clone().requires_grad_(True)
    O.laplace_density(sg, O.get_beta(bp, 0.0001)).mul(s.cos()).sum().backward()
    s_dev = s.to(DEV).requires_grad_(True)
    m.density.beta.grad = None
    m.density(s_dev).mul(s.cos().to(DEV)).sum().backward()
    assert rel_err(s_dev.grad.cpu(), sg.grad) < 1e-5
    assert rel_err(m.density.beta.grad.cpu(), bp.grad) < 1e-4
    m.density.beta.grad = None


def test_sdf_forward_paths(dtu, bmvs):
    m, sd = dtu
    x = _points(1000)                 # |x| up to 4.1: both branches of the sphere clamp
    with torch.no_grad():
        y = m.implicit_network(x.to(DEV))
        s = m.implicit_network.get_sdf_vals(x.to(DEV))
    ref = O.sdf_net(sd, 'implicit_network', x, 6)
    assert y.shape == (1000, 257)
    assert max_abs(y.cpu(), ref) < 5e-5
    ref_s = O.sdf_vals(sd, 'implicit_network', x, 6, 3.0, 20.0)
    assert max_abs(s.cpu(), ref_s) < 5e-5
    assert (ref_s < ref[:, :1] - 1e-3).any() and (ref_s == ref[:, :1]).any()
    mb, sdb = bmvs
    xb = torch.cat([torch.nn.functional.normalize(_points(700), dim=1), torch.rand(700, 1)], 1)
    with torch.no_grad():
        yb = mb.bg_implicit_network(xb.to(DEV))
    assert max_abs(yb.cpu(), O.sdf_net(sdb, 'bg_implicit_network', xb, 10)) < 1e-4
    # ragged sizes: not a multiple of any tile
    for n in (1, 31, 129):
        with torch.no_grad():
            yy = m.implicit_network(x[:n].to(DEV))
        assert max_abs(yy.cpu(), ref[:n]) < 5e-5
    with torch.no_grad():
        assert m.implicit_network(x[:0].to(DEV)).shape == (0, 257)


def test_sdf_outputs_and_gradient(dtu):
    m, sd = dtu
    sd64 = {k: v.double() for k, v in sd.items()}
    x = _points(777)
    with torch.no_grad():
        sdf, feat, grad = m.implicit_network.get_outputs(x.to(DEV))
        g2 = m.implicit_network.gradient(x.to(DEV))
    rs, rf, rg = O.sdf_outputs(sd64, 'implicit_network', x.double(), 6, 3.0, 20.0, create_graph=False)
    assert max_abs(sdf.cpu(), rs) < 5e-5 and max_abs(feat.cpu(), rf) < 5e-5
    assert max_abs(grad.cpu(), rg) < 2e-4, max_abs(grad.cpu(), rg)
    assert rel_err(grad.cpu(), rg) < 1e-5
    rg2 = O.sdf_gradient(sd64, 'implicit_network', x.double(), 6, create_graph=False)
    assert rel_err(g2.cpu(), rg2) < 1e-5
    clamped = (rs < O.sdf_net(sd64, 'implicit_network', x.double(), 6)[:, :1])
    assert clamped.any() and (~clamped).any()


def test_render_forward(dtu, bmvs):
    m, sd = dtu
    g = torch.Generator().manual_seed(8)
    x, n = _points(555), torch.randn(555, 3, generator=g)
    d = torch.nn.functional.normalize(torch.randn(555, 3, generator=g), dim=1)
    f = torch.randn(555, 256, generator=g) * 0.3
    with torch.no_grad():
        rgb = m.rendering_network(x.to(DEV), n.to(DEV), d.to(DEV), f.to(DEV))
    assert max_abs(rgb.cpu(), O.render_net(sd, 'rendering_network', x, n, d, f, 'idr', 1)) < 2e-5
    mb, sdb = bmvs
    with torch.no_grad():
        rgbb = mb.bg_rendering_network(None, None, d.to(DEV), f.to(DEV))
    assert max_abs(rgbb.cpu(), O.render_net(sdb, 'bg_rendering_network', None, None, d, f, 'nerf', 4)) < 2e-5


def _grad_report(model, prefix, ref_sd):
    worst, rows = 0.0, []
    for name, p in model.named_parameters():
        if not name.startswith(prefix) or name not in ref_sd:
            continue
        rg = ref_sd[name].grad
        rg = torch.zeros_like(ref_sd[name]) if rg is None else rg
        og = torch.zeros_like(p) if p.grad is None else p.grad
        e = rel_err(og.cpu(), rg) if float(rg.norm()) > 1e-12 else float(og.norm())
        rows.append((name, e, float(rg.norm())))
        worst = max(worst, e)
    return worst, rows


def test_sdf_backward_and_double_backward(dtu):
    """Random upstream gradients on y, clamped sdf and d sdf/dx; parameter gradients vs fp64 autograd."""
    m, sd = dtu
    x = _points(600)
    g = torch.Generator().manual_seed(21)
    dy = torch.randn(600, 257, generator=g) * 0.1
    dsdf = torch.randn(600, 1, generator=g)
    dgr = torch.randn(600, 3, generator=g)
    ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    rs, rf, rg = O.sdf_outputs(ref, 'implicit_network', x.double(), 6, 3.0, 20.0, create_graph=True)
    ry = O.sdf_net(ref, 'implicit_network', x.double(), 6)
    ((ry * dy.double()).sum() + (rs * dsdf.double()).sum() + (rg * dgr.double()).sum()).backward()
    m.zero_grad()
    m.train()
    y, sdf, grad = m.implicit_network.outputs_fused(x.to(DEV), clamp=True)
    ((y[:, :257] * dy.to(DEV)).sum() + (sdf * dsdf.to(DEV)).sum() + (grad * dgr.to(DEV)).sum()).backward()
    worst, rows = _grad_report(m, 'implicit_network', ref)
    assert worst < 2e-3, sorted(rows, key=lambda r: -r[1])[:5]
    # eikonal form: only d/dx, no clamp
    ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    rg = O.sdf_gradient(ref, 'implicit_network', x.double(), 6, create_graph=True)
    ((rg.norm(2, dim=1) - 1) ** 2).mean().backward()
    m.zero_grad()
    gt = m.implicit_network.gradient(x.to(DEV))
    ((gt.norm(2, dim=1) - 1) ** 2).mean().backward()
    worst, rows = _grad_report(m, 'implicit_network', ref)
    assert worst < 2e-3, sorted(rows, key=lambda r: -r[1])[:5]
    m.zero_grad()


def test_render_backward(dtu, bmvs):
    for (m, sd), net, mode, mv in ((dtu, 'rendering_network', 'idr', 1), (bmvs, 'bg_rendering_network', 'nerf', 4)):
        g = torch.Generator().manual_seed(8)
        P = 444
        x, n = _points(P), torch.randn(P, 3, generator=g)
        d = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=1)
        f = torch.randn(P, 256, generator=g) * 0.3
        up = torch.randn(P, 3, generator=g)
        ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
        n64, f64 = n.double().requires_grad_(True), f.double().requires_grad_(True)
        idr = mode == 'idr'
        r = O.render_net(ref, net, x.double() if idr else None, n64 if idr else None, d.double(), f64, mode, mv)
        (r * up.double()).sum().backward()
        m.zero_grad()
        m.train()
        nd, fd = n.to(DEV).requires_grad_(True), f.to(DEV).requires_grad_(True)
        rgb = getattr(m, net)(x.to(DEV) if idr else None, nd if idr else None, d.to(DEV), fd)
        (rgb * up.to(DEV)).sum().backward()
        worst, rows = _grad_report(m, net, ref)
        assert worst < 2e-3, sorted(rows, key=lambda r_: -r_[1])[:5]
        assert rel_err(fd.grad.cpu(), f64.grad) < 1e-4
        if idr:
            assert rel_err(nd.grad.cpu(), n64.grad) < 1e-4
        m.zero_grad()


@pytest.mark.parametrize('fast', [False, True])
@pytest.mark.parametrize('variant', ['fg', 'fg_tail', 'bg'])
def test_composite_forward_backward(variant, fast):
    """fast=False: canonical arithmetic (fp32 parity engine); fast=True: SVS_COMP_FAST (MUFU exp, fp32 scans — what the
    tcgen05 engine composites with).  Tolerances of the fast mode are 5x looser and written below."""
    from svolsdf_b200 import functional as F
    import svolsdf_b200._lib as L
    g = torch.Generator().manual_seed(31)
    R, S_ = 200, {'fg': 98, 'fg_tail': 97, 'bg': 32}[variant]
    z = torch.sort(torch.rand(R, S_, generator=g) * 5.5, -1)[0]
    sdf = torch.randn(R, S_, generator=g) * 0.3 + 0.2
    rgb = torch.rand(R, S_, 3, generator=g)
    ds = torch.rand(R, 1, generator=g) * 0.1 + 0.9
    nrm = torch.randn(R, S_, 3, generator=g)
    bp = torch.tensor(0.07)
    zmax = z[:, -1] + torch.rand(R, generator=g)
    up_rgb, up_dep, up_w, up_bt = (torch.randn(R, 3, generator=g), torch.randn(R, 1, generator=g),
                                   torch.randn(R, S_, generator=g), torch.randn(R, generator=g))
    # oracle (fp64 autograd)
    z64, s64, c64 = z.double(), sdf.double().requires_grad_(True), rgb.double().requires_grad_(True)
    b64 = bp.double().requires_grad_(True)
    beta = O.get_beta(b64, 0.0001)
    if variant == 'fg':
        w = O.volume_rendering(z64, s64.reshape(-1, 1), beta)
        bt = None
    elif variant == 'fg_tail':
        w, bt = O.volume_rendering_fg_bg(z64, zmax.double(), s64.reshape(-1, 1), beta)
    else:
        zf = torch.flip(z64, dims=[-1])
        w = O.bg_volume_rendering(zf, s64)
        bt = None
    zz = torch.flip(z64, dims=[-1]) if variant == 'bg' else z64
    rv, dv, nm = O.composite(w, c64, zz, ds.double(), nrm.double())
    loss = (rv * up_rgb.double()).sum() + (dv * up_dep.double()).sum() + (w * up_w.double()).sum()
    if bt is not None:
        loss = loss + (bt * up_bt.double()).sum()
    loss.backward()
    # CUDA
    flags = {'fg': 0, 'fg_tail': L.COMP_ZMAX_TAIL, 'bg': L.COMP_ABS_DENSITY | L.COMP_REVERSED}[variant]
    if fast:
        flags |= L.COMP_FAST
    k = 5.0 if fast else 1.0
    zd = (torch.flip(z, dims=[-1]) if variant == 'bg' else z).contiguous().to(DEV)
    sd_, cd = sdf.to(DEV).requires_grad_(True), rgb.to(DEV).requires_grad_(True)
    bd = bp.to(DEV).requires_grad_(True)
    wd, rvd, dvd, nmd, btd = F.composite(zd, sd_, cd, None if variant == 'bg' else bd, 0.0001, ds.to(DEV),
                                         normals=nrm.to(DEV), z_max=zmax.to(DEV) if variant == 'fg_tail' else None,
                                         flags=flags)
    assert max_abs(wd.cpu(), w) < 2e-6 * k and max_abs(rvd.cpu(), rv) < 5e-6 * k and max_abs(nmd.cpu(), nm) < 5e-6 * k
    assert max_abs(dvd.cpu(), dv) < 2e-5 * k
    l2 = (rvd * up_rgb.to(DEV)).sum() + (dvd * up_dep.to(DEV)).sum() + (wd * up_w.to(DEV)).sum()
    if bt is not None:
        assert max_abs(btd.cpu(), bt) < 2e-6 * k
        l2 = l2 + (btd * up_bt.to(DEV)).sum()
    l2.backward()
    assert rel_err(sd_.grad.cpu(), s64.grad) < 2e-4 * k, rel_err(sd_.grad.cpu(), s64.grad)
    assert rel_err(cd.grad.cpu(), c64.grad) < 1e-5 * k
    if variant != 'bg':
        assert rel_err(bd.grad.cpu(), b64.grad) < 2e-4 * k, (float(bd.grad), float(b64.grad))


def test_depth2pts_outside(bmvs):
    mb, _ = bmvs
    inp = S.make_input('bmvs', 128)
    rd, cl = O.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
    depth = torch.rand(128, 32, generator=torch.Generator().manual_seed(2)) / 3.0   # 1/r with r >= R_s = 3
    o = cl.unsqueeze(1).repeat(128, 32, 1)
    dd = rd[0].unsqueeze(1).repeat(1, 32, 1)
    pts, dreal = O.depth2pts_outside(o, dd, depth, 3.0)
    p2, d2 = mb.depth2pts_outside(o.to(DEV), dd.to(DEV), depth.to(DEV))
    assert max_abs(p2.cpu(), pts) < 2e-5 and rel_err(d2.cpu(), dreal) < 1e-5


# ---- MVS cost lookup (VolOpt.cost_mapping, vsdf.py:382-452) ---------------------------------------

def _mapper(views, img_res, inverse_depth):
    from svolsdf_b200.mvs import CostMapper
    ids = [25, 22, 28][:len(views)]
    return ids, CostMapper([v['cost'][None] for v in views], [v['z_mvs'][None] for v in views], [v['K'] for v in views],
                           [v['c2w'] for v in views], ids, img_res, inverse_depth=inverse_depth)


@pytest.mark.parametrize('tag,own,inv', [('own1_inv', 22, True), ('none_inv', 99, True), ('own0_lin', 25, False)])
def test_cost_mapping_matches_reference_golden(tag, own, inv):
    """CUDA kernel vs the outputs of the reference's own cost_mapping (golden made by oracle/make_golden.py)"""
    g = load_golden('mvs_cost_mapping')
    views = S.mvs_views(img_res=(72, 96))
    xyz = torch.from_numpy(g['xyz']).to(DEV)
    _, cm = _mapper(views, (72, 96), inv)
    cj, cmv, va = cm(torch.zeros(xyz.shape[:2], device=DEV), torch.tensor([own]), xyz)
    assert va.dtype == torch.bool and tuple(cj.shape) == tuple(xyz.shape[:2])
    # the validity tests compare fp32 coordinates with thresholds: identical except where a coordinate is within
    # rounding of a threshold (the world -> camera product is a matmul in the reference, three FMAs here)
    mism = (va.cpu().numpy() != g[tag + '_valid'])
    assert mism.mean() < 2e-3, mism.mean()
    keep = torch.from_numpy(~mism)
    assert max_abs(cj.cpu()[keep], torch.from_numpy(g[tag + '_cost_j'])[keep]) < 2e-5
    assert max_abs(cmv.cpu()[keep], torch.from_numpy(g[tag + '_cost_mvs'])[keep]) < 2e-5


@pytest.mark.parametrize('n_rays,n_samples', [(1, 1), (1024, 98), (777, 33)])
def test_cost_mapping_matches_oracle(n_rays, n_samples):
    """larger / ragged shapes against the CPU oracle (itself pinned to the reference in test_oracle_vs_golden.py)"""
    views = S.mvs_views(n_views=3, dz=48, h=72, w=96, img_res=(288, 384), seed=9)
    xyz = S.mvs_points(n_rays, n_samples, seed=10)
    ids, cm = _mapper(views, (288, 384), True)
    cj, cmv, va = cm(torch.zeros(n_rays, n_samples, device=DEV), torch.tensor([ids[2]]), xyz.to(DEV))
    rj, rm, rv = O.cost_mapping(xyz, views, (288, 384), 2, inverse_depth=True)
    mism = va.cpu() != rv
    assert float(mism.float().mean()) < 2e-3
    keep = ~mism
    # the synthetic volume is white noise (neighbouring voxels differ by O(0.1)) sampled at 384 x 288 x 48: one ulp of
    # the normalised coordinate (the reference's world -> camera product is a matmul, here three separately rounded
    # multiply-adds) moves the lookup by 2e-3 of a voxel, i.e. up to ~2e-4 in the interpolated value
    assert max_abs(cj.cpu()[keep], rj[keep]) < 5e-4 and max_abs(cmv.cpu()[keep], rm[keep]) < 5e-4
    assert float((cj.cpu()[keep] - rj[keep]).abs().mean()) < 2e-6
    if n_rays > 1:
        assert float(rv.float().mean()) > 0.1


def test_cost_mapping_rejects_bad_arguments():
    from svolsdf_b200.mvs import CostMapper
    views = S.mvs_views(n_views=1, dz=4, h=6, w=8, img_res=(12, 16))
    with pytest.raises(L.SvsError):
        CostMapper([views[0]['cost'][None]] * 9, [views[0]['z_mvs'][None]] * 9, [views[0]['K']] * 9, [views[0]['c2w']] * 9,
                   list(range(9)), (12, 16))
    ids, cm = _mapper(views, (1, 16), True)
    with pytest.raises(L.SvsError):
        cm(torch.zeros(2, 2, device=DEV), torch.tensor([0]), torch.zeros(2, 2, 3, device=DEV))


def test_cost_mapping_own_view_on_the_device():
    """`ts` as a CUDA int32 tensor: the own view is chosen inside the kernel (CUDA-graph replays) — same results"""
    views = S.mvs_views(n_views=3, dz=48, h=72, w=96, img_res=(288, 384), seed=9)
    xyz = S.mvs_points(300, 40, seed=11).to(DEV)
    ids, cm = _mapper(views, (288, 384), True)
    z = torch.zeros(300, 40, device=DEV)
    for own in (ids[0], ids[2], 999):
        a = cm(z, torch.tensor([own]), xyz)
        b = cm(z, torch.tensor([own], dtype=torch.int32, device=DEV), xyz)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


@pytest.mark.parametrize('gce,confi', [(1, 0.0), (0, 0.0), (0.5, 0.0), (0.5, 0.02)])
def test_fused_mvs_loss_matches_lookup_plus_loss_module(gce, confi):
    """`CostMapper.mvs_loss` (svs_mvs_loss: lookup + GCE term + d/d weights in one kernel, p_i p_j in registers) against
    the two-step path — `CostMapper.cost_mapping` (checked against the reference's golden above) followed by
    `VolSDFLoss.get_mvs_loss` (pinned to the reference's loss in test_oracle_vs_golden.py) with torch autograd."""
    from svolsdf_b200.model.loss import VolSDFLoss
    views = S.mvs_views(n_views=3, dz=48, h=72, w=96, img_res=(288, 384), seed=9)
    n_rays, n_samples = 777, 98
    xyz = S.mvs_points(n_rays, n_samples, seed=12).to(DEV)
    ids, cm = _mapper(views, (288, 384), True)
    g = torch.Generator().manual_seed(3)
    w = torch.softmax(torch.randn(n_rays, n_samples, generator=g) * 2, dim=1).to(DEV).requires_grad_(True)
    own = torch.tensor([ids[1]])
    # two-step reference path
    pj, pi, _ = cm(torch.zeros(n_rays, n_samples, device=DEV), own, xyz)
    loss_mod = VolSDFLoss(rgb_loss='torch.nn.L1Loss', eikonal_weight=0.1, mvs_weight=1.0, gce=gce, confi=confi)
    ref = loss_mod.get_mvs_loss({'pi': pi, 'pj': pj, 'weights': w})
    (g_ref,) = torch.autograd.grad(ref, w)
    conf_ref = (pi * pj).sum(-1)
    # fused
    w2 = w.detach().clone().requires_grad_(True)
    val, conf = cm.mvs_loss(w2, own, xyz, gce=gce, confi=confi)
    (g_fused,) = torch.autograd.grad(val * 3.0, w2)          # upstream factor: the backward scales the stored gradient
    assert not conf.requires_grad
    assert rel_err(conf.cpu(), conf_ref.cpu()) < 1e-6
    assert abs(float(val) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    assert rel_err(g_fused.cpu() / 3.0, g_ref.cpu()) < 1e-5
    assert float((conf_ref > confi).float().mean()) > 0.05   # the confidence test keeps some rays and drops others
    # the loss module takes the fused value and the ray confidences in place of pi / pj
    out = {'mvs_loss_fused': val, 'conf_ray': conf, 'weights': w2}
    assert loss_mod.get_mvs_loss(out) is val
    assert torch.equal(loss_mod._conf_ray(out), conf)


def test_fused_mvs_loss_rejects_bad_arguments():
    views = S.mvs_views(n_views=1, dz=4, h=6, w=8, img_res=(12, 16))
    ids, cm = _mapper(views, (12, 16), True)
    with pytest.raises(L.SvsError):   # weights do not match the samples
        cm.mvs_loss(torch.zeros(2, 3, device=DEV), torch.tensor([0]), torch.zeros(2, 2, 3, device=DEV))
    with pytest.raises(L.SvsError):   # more samples per ray than a warp covers
        cm.mvs_loss(torch.zeros(2, 300, device=DEV), torch.tensor([0]), torch.zeros(2, 300, 3, device=DEV))
