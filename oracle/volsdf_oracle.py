"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's VolSDF hot path (the parity oracle).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module, and only as the checker.  The product (`s-volsdf_b200/`) never imports it.

PARITY PINNING.  The reference ships no tests or golden vectors for this path, so
this oracle is pinned against the reference ITSELF: `oracle/make_golden.py` imports the unmodified
`/root/reference/volsdf/model/*` in the build container, runs it on the seeded synthetic scenes of
SURVEY.md §8d and commits the outputs under `tests/golden/`; `tests/test_oracle_vs_golden.py` checks
this restatement against those files (and against the live reference when it is present).

Written against plain torch CPU tensors (the reference's arithmetic IS torch's); the neural-network
parts use autograd in fp32 or fp64 as an independent check of the hand-derived CUDA backward, the
sampler/compositor parts spell out a CANONICAL fp32 arithmetic that the CUDA "exact" mode reproduces
bit-for-bit:
  * +,-,*,/ are IEEE fp32, evaluated in the reference's operator order, never fused;
  * sqrt is the correctly rounded fp32 root (evaluated in fp64 and rounded: torch's vectorised CPU sqrtf is
    1 ulp off for ~0.7% of inputs — measured — so it cannot serve as the canonical definition);
  * exp / expm1 are evaluated in fp64 and rounded once to fp32;
  * cumsum accumulates in fp64 and rounds every prefix to fp32 (what torch's CPU cumsum does);
  * row sums used for normalisation are the fp64 sum rounded to fp32 (== last cumsum element);
  * sort is stable (ties keep concatenation order: old samples before new ones).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# canonical helpers
# --------------------------------------------------------------------------------------------


def _exp(x):
    return torch.exp(x.double()).to(x.dtype)


def _expm1(x):
    return torch.expm1(x.double()).to(x.dtype)


def _sqrt(x):
    return torch.sqrt(x.double()).to(x.dtype)


def _cumsum(x):
    return torch.cumsum(x.double(), -1).to(x.dtype)


def _rowsum(x):
    return x.double().sum(-1, keepdim=True).to(x.dtype)


# --------------------------------------------------------------------------------------------
# camera / rays                                     volsdf/utils/rend_util.py:60-95,143-156,200-216
# --------------------------------------------------------------------------------------------


def lift(x, y, z, intrinsics):
    fx, fy = intrinsics[:, 0, 0:1], intrinsics[:, 1, 1:2]
    cx, cy = intrinsics[:, 0, 2:3], intrinsics[:, 1, 2:3]
    sk = intrinsics[:, 0, 1:2]
    x_lift = (x - cx + cy * sk / fy - sk * y / fy) / fx * z
    y_lift = (y - cy) / fy * z
    return torch.stack((x_lift, y_lift, z), dim=-1)


def get_camera_params(uv, pose, intrinsics):
    """uv (B,N,2), pose (B,4,4), intrinsics (B,4,4) -> ray_dirs (B,N,3), cam_loc (B,3)."""
    cam_loc = pose[:, :3, 3]
    x_cam, y_cam = uv[:, :, 0], uv[:, :, 1]
    z_cam = torch.ones_like(x_cam)
    pts = lift(x_cam, y_cam, z_cam, intrinsics)  # (B,N,3)
    world = torch.bmm(pose[:, :3, :3], pts.permute(0, 2, 1)) + pose[:, :3, 3:]
    world = world.permute(0, 2, 1)
    d = world - cam_loc[:, None, :]
    d = torch.nn.functional.normalize(d, dim=2)
    return d, cam_loc


def get_sphere_intersections(cam_loc, ray_dirs, r):
    """(R,3),(R,3) -> (R,2) near/far; raises where the reference prints and exit()s (rend_util.py:209-211)."""
    dot = (ray_dirs * cam_loc).sum(-1, keepdim=True)
    on = _sqrt((cam_loc * cam_loc).sum(-1, keepdim=True))
    under = dot ** 2 - (on ** 2 - r ** 2)
    if (under <= 0).sum() > 0:
        raise ValueError('BOUNDING SPHERE PROBLEM!')
    out = _sqrt(under) * torch.tensor([-1.0, 1.0], dtype=under.dtype) - dot
    return out.clamp_min(0.0)


# --------------------------------------------------------------------------------------------
# networks                                               volsdf/model/network.py:10-190, embedder.py
# --------------------------------------------------------------------------------------------


def embed(x, n_freqs):
    """[x, sin(2^0 x), cos(2^0 x), ...]  (embedder.py:10-36; frequencies are exact powers of two)."""
    if n_freqs <= 0:
        return x
    outs = [x]
    for k in range(n_freqs):
        f = float(2 ** k)
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, -1)


def effective_weight(sd, prefix):
    """weight-norm: W = g * v / ||v||_row (network.py:64-65) or the plain weight (bg nets)."""
    if prefix + '.weight_g' in sd:
        g, v = sd[prefix + '.weight_g'], sd[prefix + '.weight_v']
        return g * v / v.norm(2, dim=1, keepdim=True)
    return sd[prefix + '.weight']


def count_layers(sd, net):
    n = 0
    while (net + '.lin%d.bias' % n) in sd:
        n += 1
    return n


def sdf_net(sd, net, x, multires, skip_in=(4,)):
    """ImplicitNetwork.forward (network.py:71-88): (P,d_in) -> (P,1+feature)."""
    inp = embed(x, multires)
    h = inp
    nl = count_layers(sd, net)
    for l in range(nl):
        if l in skip_in:
            h = torch.cat([h, inp], 1) / math.sqrt(2)
        W = effective_weight(sd, '%s.lin%d' % (net, l))
        h = torch.nn.functional.linear(h, W, sd['%s.lin%d.bias' % (net, l)])
        if l < nl - 1:
            h = torch.nn.functional.softplus(h, beta=100)
    return h


def sphere_clamp(sdf, x, radius, scale):
    """min(sdf, scale*(R - |x|)) (network.py:108-112,127-130); no-op when radius <= 0."""
    if radius > 0.0:
        return torch.minimum(sdf, scale * (radius - x.norm(2, 1, keepdim=True)))
    return sdf


def sdf_vals(sd, net, x, multires, radius, scale):
    """ImplicitNetwork.get_sdf_vals (network.py:125-131)."""
    return sphere_clamp(sdf_net(sd, net, x, multires)[:, :1], x, radius, scale)


def sdf_outputs(sd, net, x, multires, radius, scale, create_graph):
    """ImplicitNetwork.get_outputs (network.py:105-123): sdf (clamped), features, d sdf/dx by autograd."""
    x = x.detach().requires_grad_(True)
    out = sdf_net(sd, net, x, multires)
    sdf = sphere_clamp(out[:, :1], x, radius, scale)
    grad = torch.autograd.grad(sdf, x, torch.ones_like(sdf), create_graph=create_graph, retain_graph=True)[0]
    return sdf, out[:, 1:], grad


def sdf_gradient(sd, net, x, multires, create_graph):
    """ImplicitNetwork.gradient (network.py:90-103): d forward[:,0]/dx, no sphere clamp."""
    x = x.detach().requires_grad_(True)
    y = sdf_net(sd, net, x, multires)[:, :1]
    return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph, retain_graph=True)[0]


def render_net(sd, net, points, normals, view_dirs, feats, mode, multires_view):
    """RenderingNetwork.forward (network.py:170-190)."""
    vd = embed(view_dirs, multires_view)
    if mode == 'idr':
        h = torch.cat([points, vd, normals, feats], -1)
    else:
        h = torch.cat([vd, feats], -1)
    nl = count_layers(sd, net)
    for l in range(nl):
        W = effective_weight(sd, '%s.lin%d' % (net, l))
        h = torch.nn.functional.linear(h, W, sd['%s.lin%d.bias' % (net, l)])
        if l < nl - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)


# --------------------------------------------------------------------------------------------
# density + compositing                         volsdf/model/density.py:21-35, network.py:281-295
# --------------------------------------------------------------------------------------------


def get_beta(beta_param, beta_min):
    return beta_param.abs() + beta_min


def laplace_density(sdf, beta, canonical=False):
    """alpha*(0.5 + 0.5*sign(s)*expm1(-|s|/beta)) (density.py:21-26); beta scalar or (R,1)."""
    alpha = 1 / beta
    em = _expm1(-sdf.abs() / beta) if canonical else torch.expm1(-sdf.abs() / beta)
    return alpha * (0.5 + 0.5 * sdf.sign() * em)


def volume_rendering(z_vals, sdf, beta, canonical=False):
    """VolSDFNetwork.volume_rendering (network.py:281-295) -> weights (R,S)."""
    R, S = z_vals.shape
    density = laplace_density(sdf.reshape(R, S), beta, canonical)
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    dists = torch.cat([dists, torch.full((R, 1), 1e10, dtype=z_vals.dtype)], -1)
    fe = dists * density
    sfe = torch.cat([torch.zeros(R, 1, dtype=z_vals.dtype), fe[:, :-1]], -1)
    if canonical:
        alpha = 1 - _exp(-fe)
        trans = _exp(-_cumsum(sfe))
    else:
        alpha = 1 - torch.exp(-fe)
        trans = torch.exp(-torch.cumsum(sfe, -1))
    return alpha * trans


def composite(weights, rgb, z_vals, depth_scale, normals=None):
    """rgb/depth/normal maps (network.py:239-248,270-276)."""
    rgb_values = torch.sum(weights.unsqueeze(-1) * rgb, 1)
    depth = torch.sum(weights * z_vals, 1, keepdim=True) / (weights.sum(1, keepdim=True) + 1e-8)
    depth = depth_scale * depth
    nm = None
    if normals is not None:
        n = normals / normals.norm(2, -1, keepdim=True)
        nm = torch.sum(weights.unsqueeze(-1) * n, 1)
    return rgb_values, depth, nm


def volume_rendering_fg_bg(z_vals, z_max, sdf, beta):
    """VolSDFNetworkBG.volume_rendering (network_bg.py:147-164) -> weights (R,S), bg_transmittance (R,)."""
    R, S = z_vals.shape
    density = laplace_density(sdf.reshape(R, S), beta)
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], z_max.unsqueeze(-1) - z_vals[:, -1:]], -1)
    fe = dists * density
    sfe = torch.cat([torch.zeros(R, 1, dtype=z_vals.dtype), fe], -1)
    alpha = 1 - torch.exp(-fe)
    trans = torch.exp(-torch.cumsum(sfe, -1))
    return alpha * trans[:, :-1], trans[:, -1]


def bg_volume_rendering(z_vals_bg, bg_sdf):
    """VolSDFNetworkBG.bg_volume_rendering (network_bg.py:166-180), AbsDensity (density.py:33-35)."""
    R, S = z_vals_bg.shape
    density = bg_sdf.reshape(R, S).abs()
    dists = z_vals_bg[:, :-1] - z_vals_bg[:, 1:]
    dists = torch.cat([dists, torch.full((R, 1), 1e10, dtype=z_vals_bg.dtype)], -1)
    fe = dists * density
    sfe = torch.cat([torch.zeros(R, 1, dtype=z_vals_bg.dtype), fe[:, :-1]], -1)
    alpha = 1 - torch.exp(-fe)
    trans = torch.exp(-torch.cumsum(sfe, -1))
    return alpha * trans


def depth2pts_outside(ray_o, ray_d, depth, radius):
    """NeRF++ inverted-sphere lift (network_bg.py:182-214): (...,3),(...,3),(...) -> (...,4), real depth."""
    o_dot_d = torch.sum(ray_d * ray_o, dim=-1)
    under = o_dot_d ** 2 - ((ray_o ** 2).sum(-1) - radius ** 2)
    d_sphere = torch.sqrt(under) - o_dot_d
    p_sphere = ray_o + d_sphere.unsqueeze(-1) * ray_d
    p_mid = ray_o - o_dot_d.unsqueeze(-1) * ray_d
    p_mid_norm = torch.norm(p_mid, dim=-1)
    axis = torch.cross(ray_o, p_sphere, dim=-1)
    axis = axis / torch.norm(axis, dim=-1, keepdim=True)
    phi = torch.asin(p_mid_norm / radius)
    theta = torch.asin(p_mid_norm * depth)
    ang = (phi - theta).unsqueeze(-1)
    p_new = p_sphere * torch.cos(ang) + torch.cross(axis, p_sphere, dim=-1) * torch.sin(ang) + \
        axis * torch.sum(axis * p_sphere, dim=-1, keepdim=True) * (1. - torch.cos(ang))
    p_new = p_new / torch.norm(p_new, dim=-1, keepdim=True)
    pts = torch.cat((p_new, depth.unsqueeze(-1)), dim=-1)
    d1 = -o_dot_d / torch.sum(ray_d * ray_d, dim=-1)
    ray_d_cos = 1. / torch.norm(ray_d, dim=-1)
    depth_real = 1. / (depth + 1e-6) * torch.cos(theta) * ray_d_cos + d1
    return pts, depth_real


# --------------------------------------------------------------------------------------------
# ErrorBoundSampler, canonical fp32 arithmetic                 volsdf/model/ray_sampler.py:15-229
# --------------------------------------------------------------------------------------------


def uniform_z(near, far, n, t_rand=None):
    """UniformSampler.get_z_vals (ray_sampler.py:22-43). near/far (R,1); t_rand (R,n) or None (eval)."""
    t = torch.linspace(0., 1., steps=n)
    z = near * (1. - t) + far * t
    if t_rand is not None:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def beta_upper_bound(z, eps):
    """Lemma-2 start value (ray_sampler.py:76-78); the row sum is canonical (fp64, rounded)."""
    dists = z[:, 1:] - z[:, :-1]
    coef = 1.0 / (4.0 * torch.log(torch.tensor(eps + 1.0)))
    bound = coef * _rowsum(dists * dists).squeeze(-1)
    return _sqrt(bound)


def d_star_bound(z, d):
    """Theorem-1 interval bound (ray_sampler.py:96-111). z,d (R,n) -> (R,n-1)."""
    dists = z[:, 1:] - z[:, :-1]
    a, b, c = dists, d[:, :-1].abs(), d[:, 1:].abs()
    first = a * a + b * b <= c * c
    second = a * a + c * c <= b * b
    ds = torch.zeros_like(a)
    ds = torch.where(first, b, ds)
    ds = torch.where(second, c, ds)
    s = (a + b + c) / 2.0
    area = s * (s - a) * (s - b) * (s - c)
    mask = ~first & ~second & (b + c - a > 0)
    tri = (2.0 * _sqrt(area)) / a
    ds = torch.where(mask, tri, ds)
    same = (d[:, 1:].sign() * d[:, :-1].sign() == 1)
    return same.to(ds.dtype) * ds


def error_bound(beta, sdf, dists, d_star):
    """ErrorBoundSampler.get_error_bound (ray_sampler.py:221-229); beta 0-dim or (R,1)."""
    R = sdf.shape[0]
    density = laplace_density(sdf, beta, canonical=True)
    sfe = torch.cat([torch.zeros(R, 1, dtype=sdf.dtype), dists * density[:, :-1]], -1)
    integral = _cumsum(sfe)
    eps_sec = _exp(-d_star / beta) * (dists * dists) / (4 * beta * beta)
    eint = _cumsum(eps_sec)
    bo = (torch.clamp(_exp(eint), max=1.e6) - 1.0) * _exp(-integral[:, :-1])
    return bo.max(-1)[0]


def stable_merge(z_old, samples):
    """sort(cat[z, samples]) with ties kept in concatenation order (ray_sampler.py:189-190)."""
    cat = torch.cat([z_old, samples], -1)
    z, idx = torch.sort(cat, dim=-1, stable=True)
    return z, idx


class SamplerTrace(object):
    """Everything the CUDA sampler is checked against, per iteration."""

    def __init__(self):
        self.iters = []          # dict per iteration: n, beta, inds, samples, samples_idx, not_converge
        self.z_final = None
        self.z_eik = None
        self.z_bg = None
        self.sdf_evals = []      # number of rows x new samples evaluated per iteration


def sampler_bound_step(z, sdf, beta_in, beta0, eps, beta_iters):
    """d* and the beta line search of one iteration (ray_sampler.py:96-123) -> beta (R,), d_star (R,n-1)."""
    R = z.shape[0]
    dists = z[:, 1:] - z[:, :-1]
    d_star = d_star_bound(z, sdf)
    err = error_bound(beta0, sdf, dists, d_star)
    beta = torch.where(err <= eps, beta0.expand_as(beta_in), beta_in)
    bmin, bmax = beta0.expand(R).clone(), beta.clone()
    for _ in range(beta_iters):
        mid = (bmin + bmax) / 2.
        err = error_bound(mid.unsqueeze(-1), sdf, dists, d_star)
        bmax = torch.where(err <= eps, mid, bmax)
        bmin = torch.where(err > eps, mid, bmin)
    return bmax, d_star


def sampler_resample_step(z, sdf, beta, d_star, cont, u, add_tiny):
    """weights -> pdf -> cdf -> inverse CDF of one iteration (ray_sampler.py:126-185).
    Returns cdf (R,n), inds (R,N), samples (R,N)."""
    R, n = z.shape
    f32 = z.dtype
    dists = z[:, 1:] - z[:, :-1]
    density = laplace_density(sdf, beta.unsqueeze(-1), canonical=True)
    dists_p = torch.cat([dists, torch.full((R, 1), 1e10, dtype=f32)], -1)
    fe = dists_p * density
    sfe = torch.cat([torch.zeros(R, 1, dtype=f32), fe[:, :-1]], -1)
    alpha = 1 - _exp(-fe)
    trans = _exp(-_cumsum(sfe))
    weights = alpha * trans
    if cont:
        b = beta.unsqueeze(-1)
        eps_sec = _exp(-d_star / b) * (dists * dists) / (4 * b * b)
        bo = (torch.clamp(_exp(_cumsum(eps_sec)), max=1.e6) - 1.0) * trans[:, :-1]
        pdf = bo + add_tiny
    else:
        pdf = weights[:, :-1] + 1e-5
    pdf = pdf / _rowsum(pdf)
    cdf = torch.cat([torch.zeros(R, 1, dtype=f32), _cumsum(pdf)], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=n - 1)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    zb, za = torch.gather(z, 1, below), torch.gather(z, 1, above)
    denom = ca - cb
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cb) / denom
    samples = zb + t * (za - zb)
    return cdf, inds, samples


def sampler_get_z_vals(ray_dirs, cam_loc, sdf_fn, beta0, *, training, near, scene_radius, n_samples,
                       n_samples_eval, n_samples_extra, eps, beta_iters, max_total_iters, fast=-1,
                       inverse_sphere_bg=False, n_samples_inverse_sphere=0, add_tiny=0.0, rng=None):
    """ErrorBoundSampler.get_z_vals (ray_sampler.py:67-219).

    sdf_fn(points (P,3)) -> (P,1) clamped sdf (get_sdf_vals).  `rng` supplies the reference's CPU draws
    in Appendix-C order: dict with 't_rand' (R,n_eval), 'u_final' (R,n_samples), 'perm' (n_eval,),
    'eik_idx' (R,), 't_rand_bg' (R,n_bg); eval uses only 'eik_idx'.
    Returns (z_final | (z_final, z_bg), z_eik, trace).
    """
    tr = SamplerTrace()
    R = ray_dirs.shape[0]
    f32 = torch.float32
    max_iters = fast if fast >= 0 else max_total_iters
    far_default = 2.0 * scene_radius
    near_t = near * torch.ones(R, 1, dtype=f32)
    if inverse_sphere_bg:
        far_t = get_sphere_intersections(cam_loc, ray_dirs, scene_radius)[:, 1:]
    else:
        far_t = far_default * torch.ones(R, 1, dtype=f32)
    z = uniform_z(near_t, far_t, n_samples_eval, rng['t_rand'] if training else None)
    samples, samples_idx = z, None
    beta = beta_upper_bound(z, eps)
    total, not_conv = 0, True
    sdf = None
    while not_conv and total < max_iters:
        pts = cam_loc.unsqueeze(1) + samples.unsqueeze(2) * ray_dirs.unsqueeze(1)
        s_new = sdf_fn(pts.reshape(-1, 3)).reshape(R, -1).to(f32)
        tr.sdf_evals.append(int(s_new.numel()))
        if samples_idx is not None:
            sdf = torch.gather(torch.cat([sdf, s_new], -1), 1, samples_idx)
        else:
            sdf = s_new
        n = z.shape[1]
        beta, d_star = sampler_bound_step(z, sdf, beta, beta0, eps, beta_iters)
        total += 1
        not_conv = bool(beta.max() > beta0)
        cont = not_conv and total < max_iters
        N = n_samples_eval if cont else n_samples
        if cont or not training:
            u = torch.linspace(0., 1., steps=N).unsqueeze(0).repeat(R, 1)
        else:
            u = rng['u_final']
        cdf, inds, samples = sampler_resample_step(z, sdf, beta, d_star, cont, u, add_tiny)
        it = {'n': n, 'beta': beta.clone(), 'inds': inds.clone(), 'samples': samples.clone(),
              'not_converge': not_conv, 'cont': cont, 'z': z.clone(), 'sdf': sdf.clone(), 'cdf': cdf.clone()}
        if cont:
            z, samples_idx = stable_merge(z, samples)
            it['samples_idx'] = samples_idx.clone()
        tr.iters.append(it)
    z_samples = samples
    n = z.shape[1]
    if n_samples_extra > 0:
        if training:
            sidx = rng['perm'][:n_samples_extra]
        else:
            sidx = torch.linspace(0, n - 1, n_samples_extra).long()
        extra = torch.cat([near_t, far_t, z[:, sidx]], -1)
    else:
        extra = torch.cat([near_t, far_t], -1)
    z_final, _ = torch.sort(torch.cat([z_samples, extra], -1), -1)
    z_eik = torch.gather(z_final, 1, rng['eik_idx'].unsqueeze(-1))
    tr.z_final, tr.z_eik = z_final, z_eik
    if inverse_sphere_bg:
        zb = uniform_z(torch.zeros(R, 1), torch.ones(R, 1), n_samples_inverse_sphere,
                       rng['t_rand_bg'] if training else None)
        zb = zb * (1. / scene_radius)
        tr.z_bg = zb
        return (z_final, zb), z_eik, tr
    return z_final, z_eik, tr


def draw_rng(R, training, n_eval=128, n_samples=64, n_final=98, bg=False, n_bg=32, radius=3.0):
    """The reference's CPU default-generator draws in its order."""
    rng = {}
    if training:
        rng['t_rand'] = torch.rand((R, n_eval))
        rng['u_final'] = torch.rand((R, n_samples))
        rng['perm'] = torch.randperm(n_eval)
    rng['eik_idx'] = torch.randint(n_final, (R,))
    if bg and training:
        rng['t_rand_bg'] = torch.rand((R, n_bg))
    if training:
        rng['eik_pts'] = torch.empty(R, 3).uniform_(-radius, radius)
    return rng


# --------------------------------------------------------------------------------------------
# full model forward                                     volsdf/model/network.py:206-279 (DTU)
# --------------------------------------------------------------------------------------------


def volsdf_forward(sd, conf, inp, training, fast=-1, rng=None, dtype=torch.float32, z_override=None):
    """VolSDFNetwork.forward restated.  `sd` = state_dict-like mapping of (detached or grad-requiring)
    tensors; `conf` = model ConfTree.  The MLP math runs in `dtype` (fp64 for tight gradient checks);
    the sampler always runs its canonical fp32 arithmetic on fp32 SDF values."""
    imp = conf.get_config('implicit_network')
    rnd = conf.get_config('rendering_network')
    smp = conf.get_config('ray_sampler')
    radius = conf.get_float('scene_bounding_sphere', default=1.0)
    white = conf.get_bool('white_bkgd', default=False)
    sdf_radius = 0.0 if white else radius
    scale = float(imp.get('sphere_scale', 1.0))
    multires = int(imp['multires'])
    beta_min = float(conf.get_config('density')['beta_min'])

    uv, pose, K = inp['uv'], inp['pose'], inp['intrinsics']
    ray_dirs, cam_loc = get_camera_params(uv, pose, K)
    tmp, _ = get_camera_params(uv, torch.eye(4)[None], K)
    depth_scale = tmp[0, :, 2:]
    R = ray_dirs.shape[1]
    cam = cam_loc.unsqueeze(1).repeat(1, R, 1).reshape(-1, 3)
    dirs = ray_dirs.reshape(-1, 3)

    sd32 = {k: v.detach().float() for k, v in sd.items()}
    beta0 = get_beta(sd32['density.beta'], beta_min)
    if rng is None:
        rng = draw_rng(R, training, radius=radius)

    def sdf_fn(p):
        with torch.no_grad():
            return sdf_vals(sd32, 'implicit_network', p, multires, sdf_radius, scale)

    if z_override is None:
        z, z_eik, trace = sampler_get_z_vals(
            dirs, cam, sdf_fn, beta0, training=training, near=float(smp['near']), scene_radius=radius,
            n_samples=int(smp['N_samples']), n_samples_eval=int(smp['N_samples_eval']),
            n_samples_extra=int(smp['N_samples_extra']), eps=float(smp['eps']),
            beta_iters=int(smp['beta_iters']), max_total_iters=int(smp['max_total_iters']), fast=fast, rng=rng)
    else:
        z, z_eik, trace = z_override
    S = z.shape[1]
    sdd = {k: v.to(dtype) for k, v in sd.items()}
    zc, cam_c, dirs_c = z.to(dtype), cam.to(dtype), dirs.to(dtype)
    pts = cam_c.unsqueeze(1) + zc.unsqueeze(2) * dirs_c.unsqueeze(1)
    pf = pts.reshape(-1, 3)
    df = dirs_c.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
    sdf, feat, grad = sdf_outputs(sdd, 'implicit_network', pf, multires, sdf_radius, scale, create_graph=training)
    rgb = render_net(sdd, 'rendering_network', pf, grad, df, feat, rnd['mode'], int(rnd['multires_view']))
    rgb = rgb.reshape(-1, S, 3)
    beta = get_beta(sdd['density.beta'], beta_min)
    weights = volume_rendering(zc, sdf, beta)
    rgb_values, depth_values, _ = composite(weights, rgb, zc, depth_scale.to(dtype))
    if white:   # white background assumption (network.py:244-247)
        bg = torch.tensor(conf.get_list('bg_color', default=[1.0, 1.0, 1.0]), dtype=dtype)
        rgb_values = rgb_values + (1.0 - weights.sum(-1, keepdim=True)) * bg.unsqueeze(0)
    out = {'rgb_values': rgb_values, 'depth_values': depth_values, 'depth_vals': zc * depth_scale.to(dtype),
           'weights': weights, 'xyz': pts, 'z_vals': z, 'sdf': sdf, 'gradients': grad, 'rgb': rgb,
           'trace': trace}
    if training:
        eik_near = (cam_c.unsqueeze(1) + z_eik.to(dtype).unsqueeze(2) * dirs_c.unsqueeze(1)).reshape(-1, 3)
        ep = torch.cat([rng['eik_pts'].to(dtype), eik_near], 0)
        out['grad_theta'] = sdf_gradient(sdd, 'implicit_network', ep, multires, create_graph=True)
    else:
        g = grad.detach()
        n = (g / g.norm(2, -1, keepdim=True)).reshape(-1, S, 3)
        out['normal_map'] = torch.sum(weights.unsqueeze(-1) * n, 1)
    return out


def volsdf_loss(out, rgb_gt, eikonal_weight=0.1):
    """VolSDFLoss with rgb_loss=L1Loss(mean), mvs/sparse off (loss.py:80-115 at config/vol/dtu.yaml:16-20)."""
    rgb_loss = (out['rgb_values'] - rgb_gt.reshape(-1, 3).to(out['rgb_values'].dtype)).abs().mean()
    eik = ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean() if 'grad_theta' in out else 0.0
    return rgb_loss + eikonal_weight * eik


# --------------------------------------------------------------------------------------------
# MVS cost lookup + full loss of the training loop        volsdf/vsdf.py:382-452, volsdf/model/loss.py:80-115
# --------------------------------------------------------------------------------------------


def _lerp_lookup(vol, coords):
    """grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) of ONE channel, spelled out.
    vol: (H, W) with coords (..., 2) = (x, y), or (Dz, H, W) with coords (..., 3) = (x, y, z); coords in [-1, 1]."""
    dims = list(vol.shape)[::-1]                       # sizes along x, y(, z)
    pos = [(coords[..., k] + 1.0) / 2.0 * (dims[k] - 1) for k in range(len(dims))]
    lo = [torch.floor(p_) for p_ in pos]
    out = torch.zeros_like(pos[0])
    for corner in range(1 << len(dims)):
        w = torch.ones_like(pos[0])
        idx, ok = [], torch.ones_like(pos[0], dtype=torch.bool)
        for k in range(len(dims)):
            hi = (corner >> k) & 1
            c = lo[k] + hi
            w = w * ((pos[k] - lo[k]) if hi else (lo[k] + 1.0 - pos[k]))
            ok = ok & (c >= 0) & (c <= dims[k] - 1)
            idx.append(c.clamp(0, dims[k] - 1).long())
        val = vol[idx[1], idx[0]] if len(dims) == 2 else vol[idx[2], idx[1], idx[0]]
        out = out + torch.where(ok, w * val, torch.zeros_like(w))
    return out


def cost_mapping(xyz, views, img_res, same_view_index, inverse_depth=True):
    """VolOpt.cost_mapping (vsdf.py:382-452).  xyz (N, D, 3) world points; views: list of dicts with cost (Dz,H,W),
    z_mvs (Dz,H,W), K (4,4), c2w (4,4); same_view_index: position in `views` of the batch's own image or -1.
    Returns results_cost_j (N,D), results_cost_mvs (N,D), valid_mask (N,D) bool."""
    h, w = img_res
    N, D, _ = xyz.shape
    cost_sum = torch.zeros(N, D, dtype=xyz.dtype)
    cost_own = torch.zeros(N, D, dtype=xyz.dtype)
    valid = torch.zeros(N, D, dtype=torch.bool)
    for i, v in enumerate(views):
        K, c2w = v['K'], v['c2w'][:3]
        fx, fy, cx, cy, sk = K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]          # :395-399
        p = (xyz - c2w[:, 3].view(1, 1, 3)) @ c2w[:, :3]                           # :401-402 world -> camera
        z = p[..., 2]
        x, y = p[..., 0] / z, p[..., 1] / z                                        # :406
        y = y * fy + cy                                                            # :407
        x = x * fx + cx + (y - cy) * sk / fy                                       # :408
        x = x / ((w - 1) / 2) - 1                                                  # :410-411
        y = y / ((h - 1) / 2) - 1
        bad = (z < 1e-5) | (x > 1.001) | (x < -1.001) | (y > 1.001) | (y < -1.001)  # :418
        x, y, z = [torch.where(bad, torch.full_like(t, -99.0), t) for t in (x, y, z)]   # :419
        xy = torch.stack([x, y], -1)
        near = _lerp_lookup(v['z_mvs'][0], xy)                                     # :420,424
        far = _lerp_lookup(v['z_mvs'][-1], xy)                                     # :425
        if inverse_depth:                                                          # :426-428
            far = torch.where(bad, torch.full_like(far, 1e-8), far)
            zn = 2 * (1.0 - near / z) / (1.0 - near / far) - 1
        else:                                                                      # :432
            zn = 2 * (z - near) / (far - near) - 1
        bad2 = (near < 1e-5) | (far < 1e-5) | (zn > 1.01) | (zn < -1.01) | bad      # :434
        x, y, zn = [torch.where(bad2, torch.full_like(t, -99.0), t) for t in (x, y, zn)]   # :435
        c = _lerp_lookup(v['cost'], torch.stack([x, y, zn], -1))                   # :440
        if i == same_view_index:                                                   # :443-444
            cost_own = c
        else:                                                                      # :446-448
            cost_sum = cost_sum + c
            valid = valid | ~bad2
    cost_own = torch.where(valid, cost_own, torch.zeros_like(cost_own))            # :450
    return cost_sum, cost_own, valid


def volsdf_full_loss(out, rgb_gt, eikonal_weight=0.1, rgb_weight=1.0, mvs_weight=0.0, sparse_weight=0.0, gce=1.0,
                     confi=0.0, anneal_sparse=0.0):
    """VolSDFLoss.forward (loss.py:80-115) with rgb_loss=L1Loss(mean); `out` may carry 'pi'/'pj' from cost_mapping.
    anneal_sparse is the linear annealing factor of loss.py:101-104 (0 outside the annealing window: then the rgb term
    is the plain L1 and the sparsity term is off)."""
    dt = out['rgb_values'].dtype
    res = {}
    diff = (out['rgb_values'] - rgb_gt.reshape(-1, 3).to(dt)).abs()
    res['rgb_loss'] = diff.mean()                                                  # :38-47
    res['eikonal_loss'] = ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean() if 'grad_theta' in out else torch.zeros((), dtype=dt)
    res['mvs_loss'] = torch.zeros((), dtype=dt)
    res['sparse_loss'] = torch.zeros((), dtype=dt)
    if 'pi' in out and mvs_weight > 0:                                             # :53-68
        pw = out['pi'] * out['pj']
        wgt = out['weights']
        if gce == 1:
            l = -pw * wgt
        elif gce == 0:
            l = -pw * torch.log(wgt + 1e-8)
        else:
            l = -pw * wgt.detach() ** gce * torch.log(wgt + 1e-8)
        res['mvs_loss'] = ((pw.sum(1) > confi).to(dt) * l.sum(1)).mean()
    if 'pi' in out and sparse_weight > 0 and anneal_sparse > 0:                    # :70-78, 96-104
        conf_ray = (out['pi'] * out['pj']).sum(-1)
        dep = (out['depth_values_all'] if 'depth_values_all' in out else out['depth_values']).squeeze()
        res['sparse_loss'] = ((1.0 / (dep + 1e-3)) * (conf_ray < confi)).mean()
        res['rgb_loss'] = (diff.mean(-1) * (conf_ray < 1e-8)).mean()               # :40-45 with t=1e-8 (rgb_smooth ground truth)
    res['loss'] = rgb_weight * res['rgb_loss'] + eikonal_weight * res['eikonal_loss'] + mvs_weight * res['mvs_loss'] + \
        sparse_weight * anneal_sparse * res['sparse_loss']                         # :107-110
    return res


# --------------------------------------------------------------------------------------------
# BlendedMVS model forward                                volsdf/model/network_bg.py:37-145
# --------------------------------------------------------------------------------------------


def volsdf_bg_forward(sd, conf, inp, training, fast=-1, rng=None, dtype=torch.float32, z_override=None):
    imp = conf.get_config('implicit_network')
    rnd = conf.get_config('rendering_network')
    smp = conf.get_config('ray_sampler')
    bgc = conf.get_config('bg_network')
    radius = conf.get_float('scene_bounding_sphere', default=1.0)
    multires = int(imp['multires'])
    beta_min = float(conf.get_config('density')['beta_min'])
    uv, pose, K = inp['uv'], inp['pose'], inp['intrinsics']
    ray_dirs, cam_loc = get_camera_params(uv, pose, K)
    tmp, _ = get_camera_params(uv, torch.eye(4)[None], K)
    depth_scale = tmp[0, :, 2:]
    R = ray_dirs.shape[1]
    cam = cam_loc.unsqueeze(1).repeat(1, R, 1).reshape(-1, 3)
    dirs = ray_dirs.reshape(-1, 3)
    sd32 = {k: v.detach().float() for k, v in sd.items()}
    beta0 = get_beta(sd32['density.beta'], beta_min)
    if rng is None:
        rng = draw_rng(R, training, bg=True, radius=radius)

    def sdf_fn(p):
        with torch.no_grad():
            return sdf_vals(sd32, 'implicit_network', p, multires, 0.0, 1.0)

    if z_override is not None:
        (z_all, z_bg), z_eik, trace = z_override
    else:
        (z_all, z_bg), z_eik, trace = sampler_get_z_vals(
            dirs, cam, sdf_fn, beta0, training=training, near=float(smp['near']), scene_radius=radius,
            n_samples=int(smp['N_samples']), n_samples_eval=int(smp['N_samples_eval']),
            n_samples_extra=int(smp['N_samples_extra']), eps=float(smp['eps']),
            beta_iters=int(smp['beta_iters']), max_total_iters=int(smp['max_total_iters']), fast=fast, rng=rng,
            inverse_sphere_bg=True, n_samples_inverse_sphere=int(smp['N_samples_inverse_sphere']),
            add_tiny=float(smp.get('add_tiny', 0.0)))
    sdd = {k: v.to(dtype) for k, v in sd.items()}
    cam_c, dirs_c = cam.to(dtype), dirs.to(dtype)
    z_max = z_all[:, -1].to(dtype)
    z = z_all[:, :-1].to(dtype)
    S = z.shape[1]
    pts = cam_c.unsqueeze(1) + z.unsqueeze(2) * dirs_c.unsqueeze(1)
    pf = pts.reshape(-1, 3)
    df = dirs_c.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
    sdf, feat, grad = sdf_outputs(sdd, 'implicit_network', pf, multires, 0.0, 1.0, create_graph=training)
    if not training:
        near_dirs, _ = get_camera_params(uv, inp['near_pose'], K)
        near_dirs = near_dirs.reshape(-1, 3).to(dtype)
        df = near_dirs.unsqueeze(1).repeat(1, S, 1).reshape(-1, 3)
    rgb = render_net(sdd, 'rendering_network', pf, grad, df, feat, rnd['mode'], int(rnd['multires_view'])).reshape(-1, S, 3)
    beta = get_beta(sdd['density.beta'], beta_min)
    weights, bg_trans = volume_rendering_fg_bg(z, z_max, sdf, beta)
    fg_rgb = torch.sum(weights.unsqueeze(-1) * rgb, 1)
    Sb = z_bg.shape[1]
    zb = torch.flip(z_bg, dims=[-1]).to(dtype)
    bg_dirs = dirs_c.unsqueeze(1).repeat(1, Sb, 1)
    bg_locs = cam_c.unsqueeze(1).repeat(1, Sb, 1)
    bg_pts, bg_depth = depth2pts_outside(bg_locs, bg_dirs, zb, radius)
    bimp, brnd = bgc.get_config('implicit_network'), bgc.get_config('rendering_network')
    bo = sdf_net(sdd, 'bg_implicit_network', bg_pts.reshape(-1, 4), int(bimp['multires']))
    bg_sdf, bg_feat = bo[:, :1], bo[:, 1:]
    bdf = bg_dirs.reshape(-1, 3)
    if not training:
        bdf = near_dirs.unsqueeze(1).repeat(1, Sb, 1).reshape(-1, 3)
    bg_rgb = render_net(sdd, 'bg_rendering_network', None, None, bdf, bg_feat, brnd['mode'],
                        int(brnd['multires_view'])).reshape(-1, Sb, 3)
    bg_w = bg_volume_rendering(zb, bg_sdf)
    bg_rgb_values = torch.sum(bg_w.unsqueeze(-1) * bg_rgb, 1)
    ds = depth_scale.to(dtype)
    w_all = torch.cat([weights, bg_trans[:, None] * bg_w], 1)
    dv_all = ds * torch.cat([z, bg_depth], 1)
    depth_all = torch.sum(w_all * dv_all, 1, keepdim=True) / (w_all.sum(1, keepdim=True) + 1e-8)
    depth_vals = z * ds
    depth_values = torch.sum(weights * depth_vals, 1, keepdim=True) / (weights.sum(1, keepdim=True) + 1e-8)
    rgb_values = fg_rgb + bg_trans.unsqueeze(-1) * bg_rgb_values
    out = {'rgb_values': rgb_values, 'depth_values_all': depth_all, 'depth_values': depth_values,
           'depth_vals': depth_vals, 'weights': weights, 'xyz': pts.detach(), 'trace': trace,
           'z_vals': z_all, 'z_bg': z_bg}
    if training:
        eik_near = (cam_c.unsqueeze(1) + z_eik.to(dtype).unsqueeze(2) * dirs_c.unsqueeze(1)).reshape(-1, 3)
        ep = torch.cat([rng['eik_pts'].to(dtype), eik_near], 0)
        out['grad_theta'] = sdf_gradient(sdd, 'implicit_network', ep, multires, create_graph=True)
    else:
        g = grad.detach()
        n = (g / g.norm(2, -1, keepdim=True)).reshape(-1, S, 3)
        out['normal_map'] = torch.sum(weights.unsqueeze(-1) * n, 1)
    return out
