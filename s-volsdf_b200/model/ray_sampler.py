"""UniformSampler / ErrorBoundSampler with the reference's interface (volsdf/model/ray_sampler.py).

The Python side only owns the outer loop of Algorithm 1 (each iteration needs the SDF network) and the
reference's CPU random draws; every per-ray computation runs in the warp-per-ray kernels of
csrc/sampler.cu.  Training (`fast=1`) has no host synchronisation; evaluation reads one 4-byte
convergence flag per iteration (the reference synchronises ~47 times per iteration).
"""
import abc

import torch

from .. import _lib as L
from .. import functional as F
from .._lib import ptr


class RefRng(object):
    """The reference's random draws: CPU default generator, same shapes, same order,
    staged through pinned memory so the host->device copies never block the stream."""

    def __init__(self, device):
        self.device = device
        self.h2d_bytes = 0

    def _up(self, t):
        self.h2d_bytes += t.numel() * t.element_size()
        return t.to(self.device, non_blocking=True)

    def rand(self, *shape):
        return self._up(torch.rand(shape, pin_memory=True))

    def randperm(self, n):
        return self._up(torch.randperm(n, pin_memory=True).to(torch.int32))

    def randint(self, high, shape):
        return self._up(torch.randint(high, shape, pin_memory=True))

    def uniform(self, shape, lo, hi):
        return self._up(torch.empty(shape, pin_memory=True).uniform_(lo, hi))


class RecordedRng(RefRng):
    """Replays draws that were made (in the reference's order) and uploaded ahead of time: used when the
    caller wants the step's inputs resident in HBM before it starts (bench.py's device-timed `value`)."""

    def __init__(self, device, tape):
        super().__init__(device)
        self.tape = list(tape)
        self.pos = 0

    def _next(self):
        t = self.tape[self.pos]
        self.pos += 1
        return t

    def rand(self, *shape):
        return self._next()

    def randperm(self, n):
        return self._next()

    def randint(self, high, shape):
        return self._next()

    def uniform(self, shape, lo, hi):
        return self._next()


class TapeRng(object):
    """Wraps a RefRng-like source and records the device tensors it hands out (to build RecordedRng tapes)."""

    def __init__(self, inner):
        self.inner, self.tape, self.calls = inner, [], []

    @property
    def h2d_bytes(self):
        return self.inner.h2d_bytes

    def _rec(self, t):
        self.tape.append(t)
        return t

    def rand(self, *shape):
        self.calls.append(('rand', shape))
        return self._rec(self.inner.rand(*shape))

    def randperm(self, n):
        self.calls.append(('randperm', (n,)))
        return self._rec(self.inner.randperm(n))

    def randint(self, high, shape):
        self.calls.append(('randint', (high, shape)))
        return self._rec(self.inner.randint(high, shape))

    def uniform(self, shape, lo, hi):
        self.calls.append(('uniform', (shape, lo, hi)))
        return self._rec(self.inner.uniform(shape, lo, hi))


_const_cache = {}


def _linspace(n, device):
    key = ('lin', n, str(device))
    if key not in _const_cache:
        _const_cache[key] = torch.linspace(0., 1., steps=n).to(device)   # CPU linspace, as the reference
    return _const_cache[key]


def _extra_idx_eval(n, k, device):
    key = ('extra', n, k, str(device))
    if key not in _const_cache:
        _const_cache[key] = torch.linspace(0, n - 1, k).long().to(torch.int32).to(device)
    return _const_cache[key]


class RaySampler(metaclass=abc.ABCMeta):
    def __init__(self, near, far):
        self.near = near
        self.far = far

    @abc.abstractmethod
    def get_z_vals(self, ray_dirs, cam_loc, model):
        pass


def _cfg(near, far, eps=0.1, add_tiny=0.0, beta_iters=10, exact=True):
    c = L.SamplerCfg()
    c.near, c.far, c.eps, c.add_tiny = float(near), float(far), float(eps), float(add_tiny)
    # 1/(4*log(1+eps)) exactly as ray_sampler.py:77 evaluates it (fp32 tensor ops on the host)
    c.inv4logeps = float(1.0 / (4.0 * torch.log(torch.tensor(eps + 1.0))))
    c.beta_iters, c.exact = int(beta_iters), 1 if exact else 0
    return c


class UniformSampler(RaySampler):
    """ray_sampler.py:15-43"""

    def __init__(self, scene_bounding_sphere, near, N_samples, take_sphere_intersection=False, far=-1):
        super().__init__(near, 2.0 * scene_bounding_sphere if far == -1 else far)
        self.N_samples = N_samples
        self.scene_bounding_sphere = scene_bounding_sphere
        self.take_sphere_intersection = take_sphere_intersection

    def get_z_vals(self, ray_dirs, cam_loc, model, iter_step=None, _rng=None, _far_ray=None, _want_beta=False):
        R = ray_dirs.shape[0]
        dev = ray_dirs.device
        far_ray = _far_ray
        if self.take_sphere_intersection and far_ray is None:
            from ..utils import rend_util
            far_ray = rend_util.get_sphere_intersections(cam_loc, ray_dirs, r=self.scene_bounding_sphere)[:, 1].contiguous()
        cfg = _cfg(self.near, -1.0 if self.take_sphere_intersection else self.far)
        n = self.N_samples
        t_rand = None
        if model.training:
            rng = _rng or RefRng(dev)
            t_rand = rng.rand(R, n)
        z = torch.empty(R, n, dtype=torch.float32, device=dev)
        beta = torch.empty(R, dtype=torch.float32, device=dev)
        L.call('svs_sampler_init', cfg, R, n, ptr(_linspace(n, dev)), ptr(t_rand), ptr(far_ray), ptr(z), ptr(beta),
               L.stream())
        return (z, beta) if _want_beta else z


class ErrorBoundSampler(RaySampler):
    """VolSDF Algorithm 1 (ray_sampler.py:46-229)."""

    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra,
                 eps, beta_iters, max_total_iters,
                 inverse_sphere_bg=False, N_samples_inverse_sphere=0, add_tiny=0.0):
        super().__init__(near, 2.0 * scene_bounding_sphere)
        self.N_samples = N_samples
        self.N_samples_eval = N_samples_eval
        self.uniform_sampler = UniformSampler(scene_bounding_sphere, near, N_samples_eval,
                                              take_sphere_intersection=inverse_sphere_bg)
        self.N_samples_extra = N_samples_extra
        self.eps = eps
        self.beta_iters = beta_iters
        self.max_total_iters = max_total_iters
        self.scene_bounding_sphere = scene_bounding_sphere
        self.add_tiny = add_tiny
        self.inverse_sphere_bg = inverse_sphere_bg
        if inverse_sphere_bg:
            self.inverse_sphere_sampler = UniformSampler(1.0, 0.0, N_samples_inverse_sphere, False, far=1.0)
        self.exact = True          # fp64 transcendentals + fp64 scans (bit-exact vs the oracle)
        self.trace = None          # set to a list to record per-iteration tensors (tests)
        self.last_iters = 0
        # eval only: rays [k * group_size, (k + 1) * group_size) of ONE call form a convergence group of their own, i.e.
        # the call computes what the reference computes when its render loop feeds the model `split_n_pixels` rays at a
        # time (vsdf.py:246-262, utils/general.py:24-39; 500 at train-time renders, 512 in eval_vsdf.py).  None: the whole
        # call is one group (the reference's semantics for a single model call).
        self.group_size = None
        self.last_group_iters = None

    def get_z_vals(self, ray_dirs, cam_loc, model, fast=-1, iter_step=None, _rng=None, _sdf_fn=None):
        R = ray_dirs.shape[0]
        dev = ray_dirs.device
        ray_dirs = ray_dirs.detach().contiguous().float()
        cam_loc = cam_loc.detach().contiguous().float()
        max_total_iters = fast if fast >= 0 else self.max_total_iters
        rng = _rng or RefRng(dev)
        training = model.training
        beta_param = model.density.beta.detach().reshape(1)
        beta_min = float(model.density.beta_min)
        sdf_fn = _sdf_fn or model.implicit_network.get_sdf_vals

        far_ray = None
        if self.inverse_sphere_bg:
            from ..utils import rend_util
            far_ray = rend_util.get_sphere_intersections(cam_loc, ray_dirs, r=self.scene_bounding_sphere)[:, 1].contiguous()
        cfg = _cfg(self.near, -1.0 if self.inverse_sphere_bg else self.far, self.eps, self.add_tiny,
                   self.beta_iters, self.exact)

        if (self.group_size and not training and R > self.group_size and self.trace is None and max_total_iters > 0):
            return self._get_z_vals_grouped(ray_dirs, cam_loc, model, rng, sdf_fn, far_ray, cfg, int(self.group_size),
                                            max_total_iters, beta_param, beta_min, iter_step)
        self.last_group_iters = None

        # uniform start + Lemma-2 beta (ray_sampler.py:72-78)
        z, beta = self.uniform_sampler.get_z_vals(ray_dirs, cam_loc, model, iter_step=iter_step, _rng=rng,
                                                  _far_ray=far_ray, _want_beta=True)
        samples, samples_idx, sdf = z, None, None
        total_iters, not_converge = 0, True
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        st = L.stream()
        while not_converge and total_iters < max_total_iters:
            n, n_new = z.shape[1], samples.shape[1]
            pts = F.ray_points(cam_loc, ray_dirs, samples).reshape(-1, 3)
            with torch.no_grad():
                sdf_new = sdf_fn(pts).reshape(R, n_new).contiguous()
            sdf_m = torch.empty(R, n, dtype=torch.float32, device=dev)
            flag.zero_()
            L.call('svs_sampler_bound', cfg, R, n, n_new, ptr(z), ptr(sdf), ptr(sdf_new), ptr(samples_idx), ptr(sdf_m),
                   ptr(beta_param), beta_min, ptr(beta), ptr(flag), st)
            sdf = sdf_m
            total_iters += 1
            if total_iters < max_total_iters:
                not_converge = bool(flag.item())   # the only host sync; never reached when fast=1 (training)
            cont = not_converge and total_iters < max_total_iters
            n_u = self.N_samples_eval if cont else self.N_samples
            if cont or not training:
                u, per_ray = _linspace(n_u, dev), 0
            else:
                u, per_ray = rng.rand(R, n_u), 1
            samples = torch.empty(R, n_u, dtype=torch.float32, device=dev)
            inds = torch.empty(R, n_u, dtype=torch.int32, device=dev) if self.trace is not None else None
            z_m = torch.empty(R, n + n_u, dtype=torch.float32, device=dev) if cont else None
            sidx = torch.empty(R, n + n_u, dtype=torch.int32, device=dev) if cont else None
            L.call('svs_sampler_resample', cfg, R, n, n_u, 1 if cont else 0, ptr(z), ptr(sdf), ptr(beta), ptr(u), per_ray,
                   ptr(samples), ptr(inds), ptr(z_m), ptr(sidx), st)
            if self.trace is not None:
                self.trace.append({'n': n, 'z': z, 'sdf': sdf, 'beta': beta.clone(), 'inds': inds, 'samples': samples,
                                   'samples_idx': sidx, 'cont': cont})
            if cont:
                z, samples_idx = z_m, sidx
        self.last_iters = total_iters

        # final sample set (ray_sampler.py:193-212)
        n = z.shape[1]
        n_extra = self.N_samples_extra
        extra_idx = None
        if n_extra > 0:
            extra_idx = rng.randperm(n)[:n_extra].contiguous() if training else _extra_idx_eval(n, n_extra, dev)
        n_s = samples.shape[1]
        m = n_s + 2 + n_extra
        z_final = torch.empty(R, m, dtype=torch.float32, device=dev)
        z_eik = torch.empty(R, 1, dtype=torch.float32, device=dev)
        eik_idx = rng.randint(m, (R,))
        L.call('svs_sampler_finalize', cfg, R, n, n_s, ptr(z), ptr(samples.contiguous()), ptr(extra_idx), n_extra,
               ptr(far_ray), ptr(eik_idx), ptr(z_final), ptr(z_eik), st)
        if self.inverse_sphere_bg:
            z_bg = self.inverse_sphere_sampler.get_z_vals(ray_dirs, cam_loc, model, _rng=rng)
            z_bg = z_bg * (1. / self.scene_bounding_sphere)
            return (z_final, z_bg), z_eik
        return z_final, z_eik


def _grouped(self, ray_dirs, cam_loc, model, rng, sdf_fn, far_ray, cfg, G, max_total_iters, beta_param, beta_min, iter_step):
    """Eval sampling of R rays as ceil(R / G) independent convergence groups inside one call: every
    group runs the reference's loop (ray_sampler.py:83-190) until ITS rays have converged (`beta_ray.max() > beta0` is
    taken per group, :136); groups that stop are finalised and leave the working set, the others get 128 more samples.
    All active groups share the sample count (128 per iteration), so one launch per kernel serves them all.  One host
    synchronisation per iteration (the reference has one per group and iteration)."""
    R = ray_dirs.shape[0]
    dev = ray_dirs.device
    st = L.stream()
    n_extra = self.N_samples_extra
    m = self.N_samples + 2 + n_extra
    z_final = torch.empty(R, m, dtype=torch.float32, device=dev)
    z_eik = torch.empty(R, 1, dtype=torch.float32, device=dev)
    eik_idx = rng.randint(m, (R,))
    z, beta = self.uniform_sampler.get_z_vals(ray_dirs, cam_loc, model, iter_step=iter_step, _rng=rng,
                                              _far_ray=far_ray, _want_beta=True)
    n_groups = (R + G - 1) // G
    idx = torch.arange(R, device=dev)
    gid = torch.div(idx, G, rounding_mode='floor')
    beta0 = beta_param.abs() + beta_min
    dirs, cams, far = ray_dirs, cam_loc, far_ray
    samples, samples_idx, sdf = z, None, None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    group_iters = [0] * n_groups
    it = 0
    while True:
        Ra, n, n_new = z.shape[0], z.shape[1], samples.shape[1]
        pts = F.ray_points(cams, dirs, samples).reshape(-1, 3)
        with torch.no_grad():
            sdf_new = sdf_fn(pts).reshape(Ra, n_new).contiguous()
        sdf_m = torch.empty(Ra, n, dtype=torch.float32, device=dev)
        L.call('svs_sampler_bound', cfg, Ra, n, n_new, ptr(z), ptr(sdf), ptr(sdf_new), ptr(samples_idx), ptr(sdf_m),
               ptr(beta_param), beta_min, ptr(beta), ptr(flag), st)
        sdf = sdf_m
        it += 1
        if it < max_total_iters:
            nc_group = torch.zeros(n_groups, dtype=torch.int32, device=dev)
            nc_group.index_add_(0, gid, (beta > beta0).to(torch.int32))
            cont_row = nc_group[gid] > 0
            sel_c = cont_row.nonzero().squeeze(1)          # host sync: how many rays go on
            n_cont = int(sel_c.numel())
        else:
            cont_row, sel_c, n_cont = None, None, 0
        if n_cont < Ra:
            if n_cont == 0:
                zf, sf, bf, idx_f, far_f, gid_f = z, sdf, beta, idx, far, gid
            else:
                sel_f = (~cont_row).nonzero().squeeze(1)
                zf, sf, bf, idx_f, gid_f = z[sel_f].contiguous(), sdf[sel_f].contiguous(), beta[sel_f].contiguous(), idx[sel_f], gid[sel_f]
                far_f = far[sel_f].contiguous() if far is not None else None
            Rf = zf.shape[0]
            n_u = self.N_samples
            smp = torch.empty(Rf, n_u, dtype=torch.float32, device=dev)
            L.call('svs_sampler_resample', cfg, Rf, n, n_u, 0, ptr(zf), ptr(sf), ptr(bf), ptr(_linspace(n_u, dev)), 0,
                   ptr(smp), None, None, None, st)
            extra_idx = _extra_idx_eval(n, n_extra, dev) if n_extra > 0 else None
            zfin = torch.empty(Rf, m, dtype=torch.float32, device=dev)
            zeik = torch.empty(Rf, 1, dtype=torch.float32, device=dev)
            L.call('svs_sampler_finalize', cfg, Rf, n, n_u, ptr(zf), ptr(smp), ptr(extra_idx), n_extra, ptr(far_f),
                   ptr(eik_idx[idx_f].contiguous()), ptr(zfin), ptr(zeik), st)
            if n_cont == 0 and Rf == R:
                z_final, z_eik = zfin, zeik
            else:
                z_final.index_copy_(0, idx_f, zfin)
                z_eik.index_copy_(0, idx_f, zeik)
            for g in torch.unique(gid_f).tolist():
                group_iters[g] = it
        if n_cont == 0:
            break
        if n_cont < Ra:
            z, sdf, beta = z[sel_c].contiguous(), sdf[sel_c].contiguous(), beta[sel_c].contiguous()
            dirs, cams, idx, gid = dirs[sel_c].contiguous(), cams[sel_c].contiguous(), idx[sel_c], gid[sel_c]
            far = far[sel_c].contiguous() if far is not None else None
        Rc = z.shape[0]
        n_u = self.N_samples_eval
        samples = torch.empty(Rc, n_u, dtype=torch.float32, device=dev)
        z_m = torch.empty(Rc, n + n_u, dtype=torch.float32, device=dev)
        sidx = torch.empty(Rc, n + n_u, dtype=torch.int32, device=dev)
        L.call('svs_sampler_resample', cfg, Rc, n, n_u, 1, ptr(z), ptr(sdf), ptr(beta), ptr(_linspace(n_u, dev)), 0,
               ptr(samples), None, ptr(z_m), ptr(sidx), st)
        z, samples_idx = z_m, sidx
    self.last_iters = it
    self.last_group_iters = group_iters
    if self.inverse_sphere_bg:
        z_bg = self.inverse_sphere_sampler.get_z_vals(ray_dirs, cam_loc, model, _rng=rng)
        z_bg = z_bg * (1. / self.scene_bounding_sphere)
        return (z_final, z_bg), z_eik
    return z_final, z_eik


ErrorBoundSampler._get_z_vals_grouped = _grouped
