"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container (where /root/reference exists):  python -m oracle.make_golden
The reference has no golden vectors of its own, so these files ARE the pin: seeded
synthetic scenes of SURVEY.md §8d pushed through volsdf/model/network.py::VolSDFNetwork.forward and
network_bg.py::VolSDFNetworkBG.forward, with torch.searchsorted / torch.sort wrapped (not modified) to
record the sampler's per-iteration indices.  Weights are not stored: they are re-created on the test
side from torch.manual_seed(0) + scene.perturb_, and a checksum in each file guards that.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
import svolsdf_b200.conf as C  # noqa: E402
import svolsdf_b200.scene as S  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


class Capture(object):
    """Wraps torch.searchsorted / torch.sort while the reference runs and records their index outputs."""

    def __init__(self, model=None):
        self.inds, self.sort_idx, self.cdf, self.sort_in = [], [], [], []
        self.model = model
        self.dens_beta, self.bounds = [], []

    def __enter__(self):
        self._ss, self._sort = torch.searchsorted, torch.sort

        def ss(*a, **k):
            r = self._ss(*a, **k)
            self.inds.append(r.clone())
            self.cdf.append(a[0].clone())
            return r

        def srt(*a, **k):
            r = self._sort(*a, **k)
            self.sort_idx.append(r[1].clone())
            self.sort_in.append(a[0].clone())
            return r

        torch.searchsorted, torch.sort = ss, srt
        if self.model is not None:   # per-iteration sampler state, recorded through hooks (no source change)
            def dens_hook(mod, args, kwargs):
                if kwargs.get('beta', None) is not None:
                    self.dens_beta.append(kwargs['beta'].detach().clone())
            self._h = self.model.density.register_forward_pre_hook(dens_hook, with_kwargs=True)
            smp = self.model.ray_sampler
            self._orig_bound = smp.get_error_bound

            def bound(beta, mdl, sdf, z_vals, dists, d_star):
                self.bounds.append((z_vals.detach().clone(), d_star.detach().clone(),
                                    sdf.detach().reshape(z_vals.shape).clone()))
                return self._orig_bound(beta, mdl, sdf, z_vals, dists, d_star)
            smp.get_error_bound = bound
        return self

    def __exit__(self, *exc):
        torch.searchsorted, torch.sort = self._ss, self._sort
        if self.model is not None:
            self._h.remove()
            del self.model.ray_sampler.get_error_bound


def param_checksum(model):
    return float(sum(p.detach().double().sum() for p in model.parameters())), \
        float(sum(p.detach().double().abs().sum() for p in model.parameters()))


def grad_fingerprint(model, n_pick=32, seed=11):
    """Per-parameter (L2 norm, sum, n_pick sampled entries at seeded positions)."""
    g = torch.Generator().manual_seed(seed)
    fp = {}
    for name, p in sorted(model.named_parameters(), key=lambda kv: kv[0]):
        gr = p.grad.detach().reshape(-1).double() if p.grad is not None else torch.zeros(p.numel(), dtype=torch.double)
        pick = torch.randint(0, gr.numel(), (n_pick,), generator=g)
        fp['grad_norm/' + name] = np.array(float(gr.norm()))
        fp['grad_sum/' + name] = np.array(float(gr.sum()))
        fp['grad_pick_idx/' + name] = pick.numpy()
        fp['grad_pick/' + name] = gr[pick].numpy()
    return fp


def build(ns, kind, perturb, beta):
    torch.manual_seed(0)
    if kind == 'dtu':
        model = ns.network.VolSDFNetwork(C.dtu_model_conf())
    else:
        model = ns.network_bg.VolSDFNetworkBG(C.bmvs_model_conf())
    if perturb or beta is not None:
        S.perturb_(model, seed=7, w_std=S.PERTURB_W if perturb else 0.0, b_std=S.PERTURB_B if perturb else 0.0, beta=beta)
    return model


def run_case(ns, name, kind, n_rays, training, perturb=False, beta=None, fast=None):
    model = build(ns, kind, perturb, beta)
    inp = S.make_input(kind, n_rays)
    rec = {'meta/kind': kind, 'meta/n_rays': n_rays, 'meta/training': training, 'meta/perturb': perturb,
           'meta/beta': -1.0 if beta is None else beta}
    cs = param_checksum(model)
    rec['meta/param_sum'], rec['meta/param_abs_sum'] = cs
    torch.manual_seed(123)
    with Capture(model) as cap:
        if training:
            model.train()
            out = model(inp, fast=1)
        else:
            model.eval()
            out = model(inp) if fast is None else model(inp, fast=fast)
    for k, v in out.items():
        rec['out/' + k] = v.detach().numpy()
    for i, t in enumerate(cap.inds):
        rec['sampler/inds_%d' % i] = t.numpy().astype(np.int32)
    for i, t in enumerate(cap.sort_idx):
        rec['sampler/sort_idx_%d' % i] = t.numpy().astype(np.int32)
    rec['sampler/n_searchsorted'] = len(cap.inds)
    for i, t in enumerate(cap.cdf):
        rec['sampler/cdf_%d' % i] = t.numpy()
    for i, t in enumerate(cap.sort_in):
        rec['sampler/sort_in_%d' % i] = t.numpy()
    n_it = len(cap.inds)
    per = len(cap.bounds) // max(n_it, 1)          # 1 + beta_iters error-bound calls per iteration
    assert per * n_it == len(cap.bounds) and len(cap.dens_beta) == (per + 1) * n_it
    for i in range(n_it):
        z_i, dstar_i, sdf_i = cap.bounds[per * i]
        rec['sampler/z_%d' % i], rec['sampler/d_star_%d' % i], rec['sampler/sdf_%d' % i] = \
            z_i.numpy(), dstar_i.numpy(), sdf_i.numpy()
        rec['sampler/beta_%d' % i] = cap.dens_beta[(per + 1) * i + per].reshape(-1).numpy()
    if training:
        gt = S.gt_rgb(n_rays)
        rgb_loss = (out['rgb_values'] - gt.reshape(-1, 3)).abs().mean()
        eik = ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
        loss = rgb_loss + 0.1 * eik
        if ns.loss is not None:  # cross-check against the reference's own VolSDFLoss
            lf = ns.loss.VolSDFLoss(rgb_loss='torch.nn.L1Loss', eikonal_weight=0.1)
            lf.iter_step = 0
            ref_loss = lf(out, {'rgb': gt})['loss']
            assert abs(float(ref_loss) - float(loss)) < 1e-7, (float(ref_loss), float(loss))
        model.zero_grad()
        loss.backward()
        rec['loss'] = np.array(float(loss))
        rec.update(grad_fingerprint(model))
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **rec)
    print('%-28s %8.1f KB  iters=%d' % (name, os.path.getsize(path) / 1024.0, len(cap.inds)))
    return rec


def unit_vectors(ns):
    """Sub-module level vectors: PE, SDF MLP forward / gradient, rendering MLP, density, camera, bg lift."""
    rec = {}
    model = build(ns, 'dtu', True, None)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(256, 3, generator=g) * 2 - 1) * 2.4   # some points beyond |x|=3 -> sphere branch
    rec['x'] = x.numpy()
    rec['pe6'] = ns.embedder.get_embedder(6)[0](x).numpy()
    imp = model.implicit_network
    with torch.no_grad():
        rec['sdf_forward'] = imp(x.clone()).numpy()
        rec['sdf_vals'] = imp.get_sdf_vals(x.clone()).numpy()
    sdf, feat, grad = imp.get_outputs(x.clone())
    rec['out_sdf'], rec['out_feat'], rec['out_grad'] = sdf.detach().numpy(), feat.detach().numpy(), grad.detach().numpy()
    rec['gradient'] = imp.gradient(x.clone()).detach().numpy()
    d = torch.nn.functional.normalize(torch.randn(256, 3, generator=g), dim=1)
    rec['view_dirs'] = d.numpy()
    rec['rgb'] = model.rendering_network(x, grad.detach(), d, feat.detach()).detach().numpy()
    s = torch.randn(64, 98, generator=g) * 0.3
    rec['dens_sdf'] = s.numpy()
    rec['dens'] = model.density(s).detach().numpy()
    z = torch.sort(torch.rand(64, 98, generator=g) * 6, -1)[0]
    rec['vr_z'] = z.numpy()
    w, _ = model.volume_rendering(z, s.reshape(-1, 1))
    rec['vr_weights'] = w.detach().numpy()
    inp = S.make_input('dtu', 128)
    rd, cl = ns.rend_util.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
    rec['cam_uv'], rec['cam_dirs'], rec['cam_loc'] = inp['uv'].numpy(), rd.numpy(), cl.numpy()
    rec['sphere_isect'] = ns.rend_util.get_sphere_intersections(cl.repeat(128, 1), rd[0], r=3.0).numpy()
    bgm = build(ns, 'bmvs', True, None)
    depth = torch.rand(128, 32, generator=g)
    o = cl.unsqueeze(1).repeat(128, 32, 1)
    dd = rd[0].unsqueeze(1).repeat(1, 32, 1)
    pts, dreal = bgm.depth2pts_outside(o, dd, depth)
    rec['bg_depth'], rec['bg_pts'], rec['bg_depth_real'] = depth.numpy(), pts.numpy(), dreal.numpy()
    with torch.no_grad():
        bo = bgm.bg_implicit_network(pts.reshape(-1, 4)[:256])
        rec['bg_sdf_forward'] = bo.numpy()
        rec['bg_rgb'] = bgm.bg_rendering_network(None, None, dd.reshape(-1, 3)[:256], bo[:, 1:]).numpy()
    rec['meta/param_sum_dtu'], _ = param_checksum(model)
    rec['meta/param_sum_bmvs'], _ = param_checksum(bgm)
    path = os.path.join(OUT, 'units.npz')
    np.savez_compressed(path, **rec)
    print('%-28s %8.1f KB' % ('units', os.path.getsize(path) / 1024.0))


def reference_cost_mapping():
    """VolOpt.cost_mapping as the reference wrote it: volsdf/vsdf.py cannot be imported here (pyhocon, tensorboard and
    a dataset on disk), so the method's source is cut out of the file with `ast` at run time and compiled on its own —
    executed verbatim, never copied into this repo."""
    import ast
    path = os.path.join(ref_import.REF_ROOT, 'volsdf', 'vsdf.py')
    tree = ast.parse(open(path).read())
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == 'cost_mapping':
            fn = node
    assert fn is not None
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    scope = {'torch': torch, 'grid_sample': torch.nn.functional.grid_sample}
    exec(compile(mod, path, 'exec'), scope)
    return scope['cost_mapping']


def mvs_case(ns):
    """cost_mapping + VolSDFLoss with the MVS terms on, on the synthetic MVS scene of svolsdf_b200.scene.mvs_views."""
    import types
    import svolsdf_b200.scene as S
    img_res = (72, 96)
    views = S.mvs_views(img_res=img_res)
    xyz = S.mvs_points()
    ids = [25, 22, 28]
    fake = types.SimpleNamespace(
        trains_i=ids, costs={i: v['cost'][None] for i, v in enumerate(views)},
        z_mvs={i: v['z_mvs'][None] for i, v in enumerate(views)},
        train_dataset=types.SimpleNamespace(img_res=list(img_res), intrinsics_all={k: views[i]['K'] for i, k in enumerate(ids)},
                                            pose_all={k: views[i]['c2w'] for i, k in enumerate(ids)}),
        hparams=types.SimpleNamespace(inverse_depth=True), stg=0)
    fn = reference_cost_mapping()
    # the volumes are not stored: the test side re-creates them from the seed; a checksum guards that
    out = {'xyz': xyz.numpy(), 'views_checksum': np.asarray(sum(float(v[k].double().sum()) for v in views for k in ('cost', 'z_mvs', 'K', 'c2w')))}
    for tag, same, inv in (('own1_inv', 22, True), ('none_inv', 99, True), ('own0_lin', 25, False)):
        fake.hparams.inverse_depth = inv
        with torch.no_grad():
            cj, cm, va = fn(fake, torch.zeros(xyz.shape[:2]), torch.tensor([same]), xyz.clone())
        out[tag + '_cost_j'], out[tag + '_cost_mvs'], out[tag + '_valid'] = cj.numpy(), cm.numpy().copy(), va.numpy()
    # the loss with its MVS / sparsity terms (loss.py:80-115), three generalised-cross-entropy settings
    if ns.loss is not None:
        g = torch.Generator().manual_seed(8)
        N, D = xyz.shape[:2]
        mo = {'rgb_values': torch.rand(N, 3, generator=g), 'grad_theta': torch.randn(2 * N, 3, generator=g),
              'weights': torch.softmax(torch.randn(N, D, generator=g), -1) * 0.9, 'depth_values': torch.rand(N, 1, generator=g) + 1.5,
              'pj': torch.from_numpy(out['own1_inv_cost_j']), 'pi': torch.from_numpy(out['own1_inv_cost_mvs'])}
        gt = {'rgb': torch.rand(1, N, 3, generator=g), 'rgb_smooth': None}
        gt['rgb_smooth'] = gt['rgb']
        for k in ('rgb_values', 'grad_theta', 'weights', 'depth_values'):
            out['loss_in_' + k] = mo[k].numpy()
        out['loss_in_rgb'] = gt['rgb'].numpy()
        for gce in (1, 0, 0.5):
            for sparse in (0.0, 0.3):
                L_ = ns.loss.VolSDFLoss('torch.nn.L1Loss', eikonal_weight=0.1, mvs_weight=0.5, sparse_weight=sparse, anneal_rgb=100 if sparse else 0,
                                        gce=gce, confi=0.02)
                L_.iter_step = 25
                r = L_(mo, gt)
                for k, v_ in r.items():
                    out['loss_gce%s_sp%s_%s' % (gce, sparse, k)] = np.asarray(float(v_))
    np.savez_compressed(os.path.join(OUT, 'mvs_cost_mapping.npz'), **out)
    print('mvs_cost_mapping.npz', {k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items() if 'loss_gce' in k or 'valid' in k})


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_import.load()
    torch.set_num_threads(8)
    unit_vectors(ns)
    run_case(ns, 'dtu_eval_r64', 'dtu', 64, False)
    run_case(ns, 'dtu_eval_r32_beta001', 'dtu', 32, False, perturb=True, beta=0.01)
    run_case(ns, 'dtu_train_r64', 'dtu', 64, True)
    run_case(ns, 'dtu_train_r64_pert', 'dtu', 64, True, perturb=True, beta=0.02)
    run_case(ns, 'bmvs_eval_r32', 'bmvs', 32, False, perturb=True, beta=0.02)
    run_case(ns, 'bmvs_train_r32', 'bmvs', 32, True, perturb=True)
    mvs_case(ns)


if __name__ == '__main__':
    main()
