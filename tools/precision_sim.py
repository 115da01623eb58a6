"""CPU model of the tcgen05 engine's operand rounding (design tool, not part of the product or the tests).

Runs one VolSDF train step with the hand-derived chains of SURVEY.md Appendix F (forward, reverse sweep, tangent sweep,
backward, weight gradients) in fp64, rounding every MMA operand the way the kernels store it: 'h' = one fp16 value,
'hh' = fp16 hi + fp16 lo (split operand, 3 MMAs), 'x' = exact.  Compares outputs and parameter gradients with fp64
autograd on the oracle.  Used to decide which chains need split operands to meet rgb/depth <= 1e-3, grads <= 1e-2.

    python tools/precision_sim.py --rays 256 --beta 0.05 --cfg fwd=hh,rev=hh,rfwd=hh
"""
import argparse
import math
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')

from oracle import volsdf_oracle as O  # noqa: E402
import svolsdf_b200.conf as C  # noqa: E402
import svolsdf_b200.scene as S  # noqa: E402

D = torch.float64


def q16(x):
    return x.float().clamp(-65504, 65504).half().double()


def qmode(x, m):
    if m == 'x':
        return x
    if m == 'h':
        return q16(x)
    if m == 'hh':
        hi = q16(x)
        return hi + q16(x - hi)
    if m == 'b':
        return x.float().bfloat16().double()
    raise ValueError(m)


def softplus(z):
    return torch.nn.functional.softplus(z, beta=100)


def dsoft_from_h(h):
    return 1.0 - torch.exp(-100.0 * h)


def pe(x, L):
    return O.embed(x, L)


def pe_jac(x, L):
    """(P, 3+6L, 3) Jacobian of PE"""
    P = x.shape[0]
    rows = [torch.eye(3, dtype=D).expand(P, 3, 3)]
    for k in range(L):
        f = float(2 ** k)
        rows.append(torch.diag_embed(f * torch.cos(x * f)))
        rows.append(torch.diag_embed(-f * torch.sin(x * f)))
    return torch.cat(rows, 1)


class SdfNet(object):
    def __init__(self, sd, net, multires, skip=4):
        self.W, self.b = [], []
        n = O.count_layers(sd, net)
        for l in range(n):
            self.W.append(O.effective_weight(sd, '%s.lin%d' % (net, l)).double())
            self.b.append(sd['%s.lin%d.bias' % (net, l)].double())
        self.L = n
        self.multires, self.skip = multires, skip

    def forward(self, x, m, m_save='h'):
        """returns y, saved (a_l as stored for the backward: fp16 images)"""
        h0 = pe(x, self.multires)
        a = qmode(h0, m)
        acts_m, acts_s = [a], [qmode(h0, m_save)]
        for l in range(self.L):
            z = a @ qmode(self.W[l], m).t() + self.b[l]
            if l == self.L - 1:
                return z, acts_s, acts_m
            h = softplus(z)
            if l + 1 == self.skip:
                h = torch.cat([h, h0], 1) / math.sqrt(2)
            a = qmode(h, m)
            acts_m.append(a)
            acts_s.append(qmode(h, m_save))

    def reverse(self, x, acts, m, m_save='h'):
        """acts: stored a_l (l = 0..L-1).  returns g (P,3), U list (stored)"""
        Lr = self.L
        U = [None] * (Lr - 1)
        hs = 1.0
        a8 = acts[Lr - 1]
        s = dsoft_from_h(a8 * (math.sqrt(2) if Lr - 1 == self.skip else 1.0))
        u = s * self.W[Lr - 1][0, :]
        um = qmode(u, m)
        U[Lr - 2] = qmode(u, m_save)
        e = None
        for l in range(Lr - 2, 0, -1):
            p = um @ qmode(self.W[l], m)          # (P, in_l)
            if l == self.skip:
                n_h = self.W[l - 1].shape[0]
                e = p[:, n_h:] / math.sqrt(2)
                p = p[:, :n_h] / math.sqrt(2)
                s = dsoft_from_h(acts[l][:, :n_h] * math.sqrt(2))
            else:
                s = dsoft_from_h(acts[l])
            u = s * p
            um = qmode(u, m)
            U[l - 1] = qmode(u, m_save)
        p0 = um @ qmode(self.W[0], m)
        if e is not None:
            p0 = p0 + e
        J = pe_jac(x, self.multires)
        g = torch.einsum('pc,pcd->pd', p0, J)
        return g, U

    def tangent(self, x, acts, U, vbar, m, gs=1.0):
        """vbar: (P,3) dL/dg (already multiplied by the clamp weight).  returns Q (stored), ZETA (stored), q_last"""
        J = pe_jac(x, self.multires)
        q0 = torch.einsum('pcd,pd->pc', J, vbar) * gs
        q = qmode(q0, m)
        Q, Z = [q], []
        for l in range(self.L - 1):
            r = q @ qmode(self.W[l], m).t()
            hn = acts[l + 1]
            if l + 1 == self.skip:
                n_h = self.W[l].shape[0]
                s = dsoft_from_h(hn[:, :n_h] * math.sqrt(2))
            else:
                s = dsoft_from_h(hn)
            zeta = 100.0 * (1 - s) * U[l] * r
            Z.append(qmode(zeta, m))
            qn = s * r
            if l + 1 == self.skip:
                qn = torch.cat([qn, q0], 1) / math.sqrt(2)
            q = qmode(qn, m)
            Q.append(q)
        return Q, Z

    def backward(self, acts, Z, dy, m, gs=1.0):
        """dy (P, 257).  returns DZ list (stored) for l = 0..L-1"""
        dz = qmode(dy * gs, m)
        DZ = [None] * self.L
        DZ[self.L - 1] = dz
        for l in range(self.L - 1, 0, -1):
            da = dz @ qmode(self.W[l], m)
            if l == self.skip:
                n_h = self.W[l - 1].shape[0]
                da = da[:, :n_h] / math.sqrt(2)
                s = dsoft_from_h(acts[l][:, :n_h] * math.sqrt(2))
            else:
                s = dsoft_from_h(acts[l])
            v = s * da
            if Z is not None:
                v = v + Z[l - 1]
            dz = qmode(v, m)
            DZ[l - 1] = dz
        return DZ

    def dw(self, acts, U, Q, DZ, gs=1.0):
        dW, db = [], []
        for l in range(self.L):
            g = DZ[l].t() @ acts[l]
            if Q is not None and l < self.L - 1:
                g = g + U[l].t() @ Q[l]
            dW.append(g / gs)
            db.append(DZ[l].sum(0) / gs)
        if Q is not None:
            dW[self.L - 1][0, :] += Q[self.L - 1].sum(0) / gs
        return dW, db


class RenderNet(object):
    def __init__(self, sd, net):
        self.W, self.b = [], []
        n = O.count_layers(sd, net)
        for l in range(n):
            self.W.append(O.effective_weight(sd, '%s.lin%d' % (net, l)).double())
            self.b.append(sd['%s.lin%d.bias' % (net, l)].double())
        self.L = n

    def forward(self, inp, m, m_save='h'):
        a = qmode(inp, m)
        acts = [qmode(inp, m_save)]
        pre = []
        for l in range(self.L):
            z = a @ qmode(self.W[l], m).t() + self.b[l]
            if l == self.L - 1:
                return torch.sigmoid(z), acts
            h = torch.relu(z)
            a = qmode(h, m)
            acts.append(qmode(h, m_save))

    def backward(self, acts, rgb, d_rgb, m, gs=1.0):
        dz = qmode(d_rgb * rgb * (1 - rgb) * gs, m)
        DZ = [None] * self.L
        DZ[self.L - 1] = dz
        for l in range(self.L - 1, 0, -1):
            da = dz @ qmode(self.W[l], m)
            dz = qmode(torch.where(acts[l] > 0, da, torch.zeros_like(da)), m)
            DZ[l - 1] = dz
        d_in = DZ[0] @ qmode(self.W[0], m) / gs
        dW = [DZ[l].t() @ acts[l] / gs for l in range(self.L)]
        db = [DZ[l].sum(0) / gs for l in range(self.L)]
        return d_in, dW, db


def wn_grads(sd, prefix, dW, db, out):
    if prefix + '.weight_g' in sd:
        g, v = sd[prefix + '.weight_g'].double(), sd[prefix + '.weight_v'].double()
        nrm = v.norm(2, dim=1, keepdim=True)
        vh = v / nrm
        dg = (dW * vh).sum(1, keepdim=True)
        out[prefix + '.weight_g'] = dg
        out[prefix + '.weight_v'] = (g / nrm) * (dW - dg * vh)
    else:
        out[prefix + '.weight'] = dW
    out[prefix + '.bias'] = db


def grad_scale(amax, target):
    if not (amax > 0):
        return 1.0
    return 2.0 ** max(-40, min(40, math.floor(math.log2(target / amax))))


def simulate(sd, conf, inp, gt, z, z_eik, eik_pts, cfg, loss_kind='l1'):
    imp = conf.get_config('implicit_network')
    radius = conf.get_float('scene_bounding_sphere', default=1.0)
    scale = float(imp.get('sphere_scale', 1.0))
    multires = int(imp['multires'])
    beta_min = float(conf.get_config('density')['beta_min'])
    sdd = {k: v.detach().double() for k, v in sd.items()}
    ray_dirs, cam_loc = O.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
    tmp, _ = O.get_camera_params(inp['uv'], torch.eye(4)[None], inp['intrinsics'])
    depth_scale = tmp[0, :, 2:].double()
    R = ray_dirs.shape[1]
    cam = cam_loc.unsqueeze(1).repeat(1, R, 1).reshape(-1, 3).double()
    dirs = ray_dirs.reshape(-1, 3).double()
    zc = z.double()
    Sn = zc.shape[1]
    pts = (cam.unsqueeze(1) + zc.unsqueeze(2) * dirs.unsqueeze(1)).reshape(-1, 3)
    df = dirs.unsqueeze(1).repeat(1, Sn, 1).reshape(-1, 3)
    eik_near = (cam.unsqueeze(1) + z_eik.double().unsqueeze(2) * dirs.unsqueeze(1)).reshape(-1, 3)
    allp = torch.cat([pts, eik_pts.double(), eik_near], 0)
    n_main = pts.shape[0]

    net = SdfNet(sdd, 'implicit_network', multires)
    rnet = RenderNet(sdd, 'rendering_network')
    y, acts, acts_m = net.forward(allp, cfg['fwd'], cfg.get('save', 'h'))
    g_all, U = net.reverse(allp, acts_m if cfg.get('rev_hact') else acts, cfg['rev'], cfg.get('save', 'h'))
    # sphere clamp on the main points
    nrm = pts.norm(dim=1, keepdim=True)
    sphere = scale * (radius - nrm)
    y0 = y[:n_main, :1]
    cw = (y0 < sphere).double() + 0.5 * (y0 == sphere).double()
    sdf = torch.minimum(y0, sphere)
    g_main = cw * g_all[:n_main] + (1 - cw) * (-scale * pts / nrm)
    feat = y[:n_main, 1:]
    vd = O.embed(df, int(conf.get_config('rendering_network')['multires_view']))
    rin = torch.cat([pts, vd, g_main, feat], -1)
    rgb, racts = rnet.forward(rin, cfg['rfwd'], cfg.get('save', 'h'))

    # compositor + loss in fp64 autograd on leaves
    sdf_l = sdf.detach().clone().requires_grad_(True)
    rgb_l = rgb.detach().clone().requires_grad_(True)
    gth_l = g_all[n_main:].detach().clone().requires_grad_(True)
    beta_p = sdd['density.beta'].clone().requires_grad_(True)
    beta = O.get_beta(beta_p, beta_min)
    weights = O.volume_rendering(zc, sdf_l, beta)
    rgb_values, depth_values, _ = O.composite(weights, rgb_l.reshape(-1, Sn, 3), zc, depth_scale)
    if loss_kind == 'l1':
        rgb_loss = (rgb_values - gt.reshape(-1, 3).double()).abs().mean()
    else:
        rgb_loss = ((rgb_values - gt.reshape(-1, 3).double()) ** 2).mean()
    loss = rgb_loss + 0.1 * ((gth_l.norm(2, dim=1) - 1) ** 2).mean()
    loss.backward()
    d_sdf, d_rgb, d_gth = sdf_l.grad, rgb_l.grad, gth_l.grad

    grads = {'density.beta': beta_p.grad}
    # rendering net backward
    gs_r = grad_scale(float(d_rgb.abs().max()), 64.0)
    d_in, dWr, dbr = rnet.backward(racts, rgb, d_rgb, cfg['rbwd'], gs_r)
    d_normals = d_in[:, 12:15]
    d_feat = d_in[:, 15:]
    for l in range(rnet.L):
        wn_grads(sdd, 'rendering_network.lin%d' % l, dWr[l], dbr[l], grads)
    # sdf net backward
    vbar = torch.cat([cw * d_normals, d_gth], 0)
    dy = torch.zeros_like(y)
    dy[:n_main, :1] = cw * d_sdf
    dy[:n_main, 1:] = d_feat
    amax = max(float(dy.abs().max()), float(vbar.abs().max()) * 2 ** (multires - 1))
    gs = grad_scale(amax, 8.0)
    Q, Z = net.tangent(allp, acts, U, vbar, cfg['tan'], gs)
    DZ = net.backward(acts, Z, dy, cfg['bwd'], gs)
    dW, db = net.dw(acts, U, Q, DZ, gs)
    for l in range(net.L):
        wn_grads(sdd, 'implicit_network.lin%d' % l, dW[l], db[l], grads)
    return {'rgb_values': rgb_values.detach(), 'depth_values': depth_values.detach(), 'loss': float(loss),
            'grad_theta': g_all[n_main:].detach(), 'sdf': sdf.detach(), 'grads': grads, 'weights': weights.detach()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rays', type=int, default=256)
    ap.add_argument('--beta', type=float, default=0.05)
    ap.add_argument('--cfg', default='')
    ap.add_argument('--loss', default='l1')
    ap.add_argument('--seed', type=int, default=123)
    ap.add_argument('--sampler-mode', default='x')
    ap.add_argument('--sampler-noise', type=float, default=0.0, help='perturb the sampler sdf (models fp16 sampler chain)')
    args = ap.parse_args()
    from helpers import build_model, state_dict_cpu
    cfg = {k: 'h' for k in ('fwd', 'rev', 'rfwd', 'rbwd', 'tan', 'bwd')}
    for kv in filter(None, args.cfg.split(',')):
        k, v = kv.split('=')
        if k == 'all':
            for kk in list(cfg):
                cfg[kk] = v
            if v == 'x':
                cfg['save'] = 'x'
        elif k == 'rev_hact':
            cfg[k] = int(v)
        else:
            cfg[k] = v
    model = build_model('dtu', perturb=True, beta=args.beta)
    sd = state_dict_cpu(model)
    conf = C.dtu_model_conf()
    R = args.rays
    inp = S.make_input('dtu', R)
    gt = S.gt_rgb(R)
    torch.manual_seed(args.seed)
    rng = O.draw_rng(R, True)
    # fp64 oracle (truth) on the oracle's own sample positions
    ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    o = O.volsdf_forward(ref, conf, inp, True, fast=1, rng=rng, dtype=D)
    z, trace = o['z_vals'], o['trace']
    # z_eik is not returned: recompute through the override path
    torch.manual_seed(args.seed)
    rng2 = O.draw_rng(R, True)
    sd32 = {k: v.detach().float() for k, v in sd.items()}
    beta0 = O.get_beta(sd32['density.beta'], 1e-4)
    smp = conf.get_config('ray_sampler')
    ray_dirs, cam_loc = O.get_camera_params(inp['uv'], inp['pose'], inp['intrinsics'])
    cam = cam_loc.unsqueeze(1).repeat(1, R, 1).reshape(-1, 3)
    dirs = ray_dirs.reshape(-1, 3)

    def sdf_fn(p):
        with torch.no_grad():
            s = O.sdf_vals(sd32, 'implicit_network', p, 6, 3.0, float(conf.get_config('implicit_network').get('sphere_scale', 1.0)))
            if args.sampler_noise > 0:
                g = torch.Generator().manual_seed(99)
                s = s + args.sampler_noise * torch.randn(s.shape, generator=g) * s.abs().clamp(max=1.0)
            if args.sampler_mode != 'x':
                netq = SdfNet({k: v.double() for k, v in sd32.items()}, 'implicit_network', 6)
                yq = netq.forward(p.double(), args.sampler_mode)[0][:, :1]
                s = O.sphere_clamp(yq, p.double(), 3.0, float(conf.get_config('implicit_network').get('sphere_scale', 1.0))).float()
            return s
    zz, z_eik, _ = O.sampler_get_z_vals(
        dirs, cam, sdf_fn, beta0, training=True, near=float(smp['near']), scene_radius=3.0,
        n_samples=int(smp['N_samples']), n_samples_eval=int(smp['N_samples_eval']),
        n_samples_extra=int(smp['N_samples_extra']), eps=float(smp['eps']),
        beta_iters=int(smp['beta_iters']), max_total_iters=int(smp['max_total_iters']), fast=1, rng=rng2)
    if args.sampler_noise == 0 and args.sampler_mode == 'x':
        assert torch.equal(zz, z)
    if args.loss == 'l1':
        rl = O.volsdf_loss(o, gt)
    else:
        rl = ((o['rgb_values'] - gt.reshape(-1, 3).double()) ** 2).mean() + 0.1 * ((o['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
    rl.backward()
    sim = simulate(sd, conf, inp, gt, zz, z_eik, rng['eik_pts'], cfg, args.loss)
    print('cfg', cfg, 'rays', R, 'beta', args.beta, 'loss', args.loss)
    print('loss %.8f  (oracle %.8f)' % (sim['loss'], float(rl)))
    hit = o['weights'].detach().sum(1, keepdim=True) > 1e-2
    for k in ('rgb_values', 'depth_values', 'grad_theta', 'sdf', 'weights'):
        a, b = sim[k], o[k].detach()
        print('  %-14s max|d| %.3e' % (k, float((a - b.reshape(a.shape)).abs().max())))
    a, b = sim['depth_values'], o['depth_values'].detach()
    print('  depth (hit rays, %d of %d) max|d| %.3e' % (int(hit.sum()), hit.numel(), float((a - b)[hit].abs().max())))
    rows = []
    for name, p in ref.items():
        if p.grad is None or float(p.grad.norm()) < 1e-10:
            continue
        g = sim['grads'][name].reshape(p.grad.shape)
        rows.append((float((g - p.grad).norm() / p.grad.norm()), name))
    rows.sort(reverse=True)
    print('  worst grads:', ', '.join('%s %.2e' % (n.replace('implicit_network', 'sdf').replace('rendering_network', 'rnd'), e) for e, n in rows[:6]))
    print('  median grad rel err %.2e' % rows[len(rows) // 2][0])


if __name__ == '__main__':
    main()
